#!/bin/bash
TAG=${1:-l}
O=gpurun_out
mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $O/launches_preview_${TAG}.csv python bench.py --steps 3 --warmup 3 --contexts 1 --no-cpu-baseline --no-second-flavour > $O/ncu_l_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_full_${TAG}.csv python bench.py --steps 3 --warmup 3 --contexts 1 --mode full --no-cpu-baseline --no-second-flavour >> $O/ncu_l_${TAG}.log 2>&1
python tools/launch_summary.py $O/launches_preview_${TAG}.csv
python tools/launch_summary.py $O/launches_full_${TAG}.csv
