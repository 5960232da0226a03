#!/bin/bash
# Sweep of the march kernel's tunables (env overrides) on the GPU box.  usage: bash tools/sweep_march.sh <tag>
TAG=${1:-sweep}
O=gpurun_out
mkdir -p $O
: > $O/sweep_${TAG}.txt
for res in "1920 1080" "3840 2160"; do
  set -- $res; W=$1; H=$2
  for fl in fast exact; do
    for chunk in 32 128; do
      for bps in 3 4 6 8; do
        RMB_WF_CHUNK=$chunk RMB_MARCH_BLOCKS_PER_SM=$bps python bench.py --steps 10 --warmup 3 --width $W --height $H --flavour $fl --no-cpu-baseline --no-second-flavour > $O/tmp_sweep.json 2> $O/tmp_sweep.err
        python - <<PY >> $O/sweep_${TAG}.txt
import json
try:
    d=json.load(open("$O/tmp_sweep.json"))
    print("$W x $H $fl chunk=$chunk blocks/SM=$bps value", round(d["value"],1), "frac", round(d["roofline"]["frac"],4), "march_ms", round(d["roofline"]["kernel_ms_per_step"],4), "step_ms", round(d["ms_per_step"],4))
except Exception as e:
    print("$W x $H $fl chunk=$chunk blocks/SM=$bps FAILED", e)
PY
      done
    done
  done
done
cat $O/sweep_${TAG}.txt
ncu --set full --clock-control none --import-source on -k regex:rm_wf_setup -s 3 -c 1 -o $O/prof_setup_${TAG} python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-second-flavour > $O/ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rm_wf_bounce -s 3 -c 1 -o $O/prof_bounce_${TAG} python bench.py --steps 2 --warmup 3 --mode full --no-cpu-baseline --no-second-flavour >> $O/ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rm_display -s 3 -c 1 -o $O/prof_display_${TAG} python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-second-flavour >> $O/ncu_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_full_${TAG}.csv python bench.py --steps 3 --warmup 3 --mode full --no-cpu-baseline --no-second-flavour > $O/ncu_launches_full_${TAG}.log 2>&1
