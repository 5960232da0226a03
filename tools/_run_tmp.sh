O=gpurun_out
( time python -m pytest tests -m gpu -x -q -k "mandelbulb or config4" ) > $O/pytest_r2n.log 2>&1; tail -5 $O/pytest_r2n.log
fl=exact; TAG=r2n
python bench.py --scene mandelbulb --step-counts 512 --width 3840 --height 2160 --flavour $fl --steps 3 --warmup 1 --frames-per-step 4 \
      --no-second-flavour --config3-steps 0 --cpu-band-rows 24 > $O/bench_${TAG}_config4_$fl.json 2> $O/bench_${TAG}_config4_$fl.err
python -c "
import json
d=json.load(open('$O/bench_${TAG}_config4_$fl.json')); r=d['roofline']
print('config4 $fl', round(d['value'],1), 'Mpx/s e2e', round(d['e2e']['value'],1), 'fp32 frac', round(r['frac'],4), 'regs', r['registers_per_thread'])"
ncu --set full --clock-control none --import-source on -k regex:rm_wf_march_preview_kernel -s 2 -c 1 -o $O/prof_config4_${fl}_${TAG} \
      python bench.py --quick --scene mandelbulb --step-counts 512 --width 1920 --height 1080 --flavour $fl --steps 1 --warmup 1 --frames-per-step 1 --contexts 1 > $O/ncu_config4_${fl}_${TAG}.log 2>&1
