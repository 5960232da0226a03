#!/bin/bash
# usage: bash tools/gpu_r2_iter.sh <tag> [pytest -k expression]
# one development iteration on the GPU box: the small parity suite (+ the 1080p preview oracle comparisons),
# the default bench line, a launch list and one ncu --set full of the march kernel
TAG=${1:-it}
KEXPR=${2:-"not (full_one_light or config3 or config5 or config4)"}
O=gpurun_out
mkdir -p $O
( time python -m pytest tests -m gpu -x -q -k "$KEXPR" ) > $O/pytest_${TAG}.log 2>&1; tail -15 $O/pytest_${TAG}.log
python bench.py --steps 30 --warmup 5 --no-second-flavour --no-cpu-baseline > $O/bench_${TAG}_preview.json 2> $O/bench_${TAG}_preview.err
python - <<PY
import json
try:
    d=json.load(open("$O/bench_${TAG}_preview.json")); r=d["roofline"]
    print("preview exact", round(d["value"],1), "Mpx/s  e2e", round(d["e2e"]["value"],1), " frac", round(r["frac"],4), "whole", round(r["whole_step_frac"],4), " kernel_ms/step", round(r["kernel_ms_per_step"],4), "step_ms", round(d["ms_per_step"],4), "launches", d["gpu_launches"], "evals/step", r["executed_sdf_evals_per_step"], "far share", round(r["far_field_evals_share"],4))
except Exception as e:
    print("bench FAILED", e); print(open("$O/bench_${TAG}_preview.err").read()[-3000:])
PY
if [ -z "$NOPROF" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $O/launches_${TAG}.csv python bench.py --steps 6 --warmup 5 --contexts 1 --no-cpu-baseline --no-second-flavour > $O/launches_${TAG}.log 2>&1
python tools/launch_summary.py $O/launches_${TAG}.csv > $O/launches_${TAG}_summary.txt 2>&1; cat $O/launches_${TAG}_summary.txt
ncu --set full --clock-control none --import-source on -k regex:rm_wf_march_preview_kernel -s 12 -c 2 -o $O/prof_march_${TAG} python bench.py --steps 2 --warmup 3 --contexts 1 --no-cpu-baseline --no-second-flavour > $O/ncu_march_${TAG}.log 2>&1
fi
