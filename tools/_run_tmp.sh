O=gpurun_out
( time python -m pytest tests -m gpu -x -q -k "not (config3 or config5 or config4 or group or variant or 1080p_preview)" ) > $O/pytest_r2t.log 2>&1; tail -4 $O/pytest_r2t.log
python bench.py --quick --mode full --steps 3 --warmup 2 --frames-per-step 4 > $O/bench_r2t_full.json 2>$O/bench_r2t_full.err
python -c "import json; d=json.load(open('$O/bench_r2t_full.json')); r=d['roofline']; print('full value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frame_ms',round(d['extra']['ms_per_frame'],4),'frac',round(r['frac'],4),'kernel_ms/frame',round(r['kernel_ms_per_frame'],4))" || tail -20 $O/bench_r2t_full.err
python bench.py --quick --steps 10 --warmup 3 > $O/bench_r2t_quick.json 2>$O/bench_r2t_quick.err
python -c "import json; d=json.load(open('$O/bench_r2t_quick.json')); r=d['roofline']; print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frame_ms',round(d['extra']['ms_per_frame'],4),'frac',round(r['frac'],4),'kernel_ms/frame',round(r['kernel_ms_per_frame'],4))" || tail -20 $O/bench_r2t_quick.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file $O/launches_r2t_full.csv python bench.py --quick --mode full --steps 2 --warmup 1 --frames-per-step 4 --contexts 1 > $O/launches_r2t_full.log 2>&1
python tools/launch_summary.py $O/launches_r2t_full.csv | tee $O/launches_r2t_full_summary.txt
