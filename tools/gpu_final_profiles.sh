#!/bin/bash
# Final round profile set: ncu --set full of the march kernel (exact + fast, preview + castRay), launch lists.
TAG=${1:-r1z}
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --contexts 1 --no-cpu-baseline --no-second-flavour"
ncu --set full --clock-control none --import-source on -k regex:rm_wf_march_preview -s 3 -c 1 -o $O/prof_march_preview_exact_${TAG} $B > $O/ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rm_wf_march_preview -s 3 -c 1 -o $O/prof_march_preview_fast_${TAG} $B --flavour fast >> $O/ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rm_wf_march_cast -s 12 -c 1 -o $O/prof_march_cast_exact_${TAG} $B --mode full >> $O/ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rm_wf_march_cast -s 12 -c 1 -o $O/prof_march_cast_fast_${TAG} $B --mode full --flavour fast >> $O/ncu_${TAG}.log 2>&1
ncu --set full --clock-control none -k regex:rm_wf_setup -s 3 -c 1 -o $O/prof_setup_${TAG} $B >> $O/ncu_${TAG}.log 2>&1
ncu --set full --clock-control none -k regex:rm_wf_bounce -s 3 -c 1 -o $O/prof_bounce_${TAG} $B --mode full >> $O/ncu_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $O/launches_preview_exact_${TAG}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-second-flavour > $O/ncu_launches_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_full_exact_${TAG}.csv python bench.py --steps 3 --warmup 3 --mode full --no-cpu-baseline --no-second-flavour >> $O/ncu_launches_${TAG}.log 2>&1
ls -la $O/*${TAG}*
