O=gpurun_out
( time python -m pytest tests -m gpu -x -q -k "not (full_one_light or config3 or config5 or config4)" ) > $O/pytest_r2k.log 2>&1; tail -5 $O/pytest_r2k.log
python bench.py --quick --steps 10 --warmup 3 > $O/bench_r2k_quick.json 2>$O/bench_r2k_quick.err
python -c "import json; d=json.load(open('$O/bench_r2k_quick.json')); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frame_ms',round(d['extra']['ms_per_frame'],4),'frac',round(d['roofline']['frac'],4))" || tail -20 $O/bench_r2k_quick.err
