#!/bin/bash
# One GPU-box visit: tests, bench lines, tolerance report, ncu passes.  usage: bash tools/gpu_round.sh <tag> [quick]
TAG=${1:-r1c}
QUICK=$2
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/env_${TAG}.txt; nproc >> $O/env_${TAG}.txt
python -m pytest tests -m gpu -x -q > $O/pytest_${TAG}.log 2>&1; tail -5 $O/pytest_${TAG}.log
for cfg in "preview exact" "preview fast" "full exact" "full fast"; do
  set -- $cfg
  python bench.py --steps 20 --warmup 5 --mode $1 --flavour $2 --no-cpu-baseline --no-second-flavour > $O/bench_${TAG}_$1_$2.json 2> $O/bench_${TAG}_$1_$2.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_${TAG}_$1_$2.json"))
    print("$1 $2", round(d["value"],1), "Mpx/s  e2e", round(d["e2e"]["value"],1), " frac", round(d["roofline"]["frac"],4), " kernel_ms/step", round(d["roofline"]["kernel_ms_per_step"],4), "step_ms", round(d["ms_per_step"],4), "launches", d["gpu_launches"], "share", round(d["roofline"]["kernel_share_of_step"],3), "steps/px", round(d["roofline"]["executed_steps_per_px"],2), "regs", d["roofline"]["registers_per_thread"])
except Exception as e:
    print("$1 $2 FAILED", e); print(open("$O/bench_${TAG}_$1_$2.err").read()[-2000:])
PY
done
python tools/flavour_report.py --poses 0 40 --modes preview full > $O/flavour_${TAG}.json 2> $O/flavour_${TAG}.err; cat $O/flavour_${TAG}.json | tr -d '\n ' ; echo
if [ "$QUICK" != "quick" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-second-flavour > $O/ncu_launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rm_wf_march -s 3 -c 1 -o $O/prof_preview_fast_${TAG} \
    python bench.py --steps 2 --warmup 3 --flavour fast --no-cpu-baseline --no-second-flavour > $O/ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rm_wf_march -s 12 -c 1 -o $O/prof_full_fast_${TAG} \
    python bench.py --steps 2 --warmup 3 --mode full --flavour fast --no-cpu-baseline --no-second-flavour >> $O/ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rm_wf_march -s 12 -c 1 -o $O/prof_full_exact_${TAG} \
    python bench.py --steps 2 --warmup 3 --mode full --no-cpu-baseline --no-second-flavour >> $O/ncu_${TAG}.log 2>&1
fi
