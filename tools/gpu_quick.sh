#!/bin/bash
# usage: bash tools/gpu_quick.sh <tag> : tests + the four bench lines with 1 and 2 contexts
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
if [ -z "$NOTEST" ]; then python -m pytest tests -m gpu -x -q > $O/pytest_${TAG}.log 2>&1; tail -3 $O/pytest_${TAG}.log; fi
for nc in ${CTXS:-1 2}; do
for cfg in "preview exact" "preview fast" "full exact" "full fast"; do
  set -- $cfg
  python bench.py --steps 24 --warmup 6 --mode $1 --flavour $2 --contexts $nc --no-cpu-baseline --no-second-flavour > $O/bench_${TAG}_$1_$2_c$nc.json 2> $O/bench_${TAG}_$1_$2_c$nc.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_${TAG}_$1_$2_c$nc.json"))
    print("ctx=$nc $1 $2", round(d["value"],1), "Mpx/s  e2e", round(d["e2e"]["value"],1), " frac", round(d["roofline"]["frac"],4), "whole", round(d["roofline"]["whole_step_frac"],4), " kernel_ms/step", round(d["roofline"]["kernel_ms_per_step"],4), "step_ms", round(d["ms_per_step"],4), "launches", d["gpu_launches"])
except Exception as e:
    print("ctx=$nc $1 $2 FAILED", e); print(open("$O/bench_${TAG}_$1_$2_c$nc.err").read()[-2000:])
PY
done
done
