#!/bin/bash
# BASELINE.json configs 2-5 on one GPU (config 5 with a bounded number of poses).  usage: bash tools/gpu_configs.sh <tag>
TAG=${1:-cfg}
O=gpurun_out
mkdir -p $O
run() { name=$1; shift
  python bench.py --no-cpu-baseline --no-second-flavour "$@" > $O/cfg_${TAG}_${name}.json 2> $O/cfg_${TAG}_${name}.err
  python - <<PY
import json
try:
    d=json.load(open("$O/cfg_${TAG}_${name}.json"))
    r=d["roofline"]
    print("${name}:", round(d["value"],1), "Mpx/s  e2e", round(d["e2e"]["value"],1), "| Msteps/s executed", round(d["extra"]["msteps_per_s_executed"]), "ref-equiv", round(d["extra"]["msteps_per_s_ref_equiv"]), "| march frac", r["frac"] and round(r["frac"],4), "steps/px", round(r["executed_steps_per_px"],1), "ms/step", round(d["ms_per_step"],3))
except Exception as e:
    print("${name} FAILED", e); print(open("$O/cfg_${TAG}_${name}.err").read()[-1500:])
PY
}
for fl in exact fast; do
run c2_1080p_preview_$fl --steps 20 --warmup 5 --flavour $fl
run c2_1080p_full_$fl --steps 8 --warmup 3 --mode full --flavour $fl
run c3_4k_preview_$fl --steps 12 --warmup 3 --width 3840 --height 2160 --flavour $fl
run c3_4k_full_$fl --steps 4 --warmup 2 --width 3840 --height 2160 --mode full --flavour $fl
run c4_mandelbulb_4k_512_preview_$fl --steps 6 --warmup 2 --width 3840 --height 2160 --scene mandelbulb --step-counts 512 --flavour $fl
run c4_mandelbulb_4k_512_full_$fl --steps 3 --warmup 2 --width 3840 --height 2160 --scene mandelbulb --step-counts 512 --mode full --flavour $fl
run c5_8k_16spp_preview_$fl --steps 4 --warmup 2 --width 7680 --height 4320 --spp 16 --flavour $fl
done
