#!/bin/bash
# usage: bash tools/gpu_config4.sh <tag> : BASELINE.json config 4 (Mandelbulb DE, 3840x2160, steps [512]) in both flavours with
# the config-4 roofline model (bench.py header), and ncu --set full of the march kernel of each
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
for fl in fast exact; do
  python bench.py --scene mandelbulb --step-counts 512 --width 3840 --height 2160 --flavour $fl --steps 3 --warmup 1 --frames-per-step 4 \
      --no-second-flavour --config3-steps 0 --cpu-band-rows 24 > $O/bench_${TAG}_config4_$fl.json 2> $O/bench_${TAG}_config4_$fl.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_${TAG}_config4_$fl.json")); r=d["roofline"]
    print("config4 $fl", round(d["value"],1), "Mpx/s e2e", round(d["e2e"]["value"],1), "fp32 frac", round(r["frac"],4), "xu frac", round(r["config4"]["xu_bound"]["frac"],4),
          "trips/eval", round(r["config4"]["de_trips_per_eval"],3), "evals/px", round(r["executed_steps_per_px"],2), "regs", r["registers_per_thread"], "cpu", d.get("cpu_baseline",{}).get("value"))
except Exception as e:
    print("config4 $fl FAILED", e); print(open("$O/bench_${TAG}_config4_$fl.err").read()[-2000:])
PY
  ncu --set full --clock-control none --import-source on -k regex:rm_wf_march_preview_kernel -s 2 -c 1 -o $O/prof_config4_${fl}_${TAG} \
      python bench.py --quick --scene mandelbulb --step-counts 512 --width 1920 --height 1080 --flavour $fl --steps 1 --warmup 1 --frames-per-step 1 --contexts 1 > $O/ncu_config4_${fl}_${TAG}.log 2>&1
done
