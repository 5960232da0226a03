O=gpurun_out
( time python -m pytest tests -m gpu -x -q -k "not (full_one_light or config3 or config5 or config4 or group or variant)" ) > $O/pytest_r2q.log 2>&1; tail -4 $O/pytest_r2q.log
python bench.py --quick --steps 10 --warmup 3 > $O/bench_r2q_quick.json 2>$O/bench_r2q_quick.err
python -c "import json; d=json.load(open('$O/bench_r2q_quick.json')); r=d['roofline']; print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frame_ms',round(d['extra']['ms_per_frame'],4),'frac',round(r['frac'],4),'kernel_ms/frame',round(r['kernel_ms_per_frame'],4))" || tail -20 $O/bench_r2q_quick.err
python bench.py --quick --steps 4 --warmup 2 --width 3840 --height 2160 --frames-per-step 8 > $O/bench_r2q_4k.json 2>$O/bench_r2q_4k.err
python -c "import json; d=json.load(open('$O/bench_r2q_4k.json')); r=d['roofline']; print('4K value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frame_ms',round(d['extra']['ms_per_frame'],4),'frac',round(r['frac'],4),'kernel_ms/frame',round(r['kernel_ms_per_frame'],4))" || tail -20 $O/bench_r2q_4k.err
