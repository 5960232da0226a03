O=gpurun_out
( time python -m pytest tests -m gpu -x -q -k "group or variant or auto_special or row_tile or generic_program" ) > $O/pytest_r2l.log 2>&1; tail -30 $O/pytest_r2l.log
