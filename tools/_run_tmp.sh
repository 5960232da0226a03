O=gpurun_out
( time python -m pytest tests -m gpu -x -q -k "menger or carve or probe or far_field or sponge" ) > $O/pytest_r2y.log 2>&1; tail -4 $O/pytest_r2y.log
python tools/carve_gain.py menger-sponge guide 2>&1 | tee $O/carve_gain_r2y.txt
