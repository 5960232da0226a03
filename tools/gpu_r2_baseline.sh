#!/bin/bash
# usage: bash tools/gpu_r2_baseline.sh <tag> : the whole -m gpu suite (incl. the BASELINE-size oracle comparisons),
# then the default bench lines and a fresh ncu --set full of the setup kernel (VERDICT r1 Weak 3)
TAG=${1:-r2a}
O=gpurun_out
mkdir -p $O
nproc > $O/nproc_${TAG}.txt
( time python -m pytest tests -m gpu -x -q --durations=25 ) > $O/pytest_${TAG}.log 2>&1; tail -40 $O/pytest_${TAG}.log
python bench.py --steps 30 --warmup 5 --no-second-flavour > $O/bench_${TAG}_preview.json 2> $O/bench_${TAG}_preview.err; tail -c 600 $O/bench_${TAG}_preview.json
python bench.py --steps 30 --warmup 5 --no-second-flavour --no-cpu-baseline --width 3840 --height 2160 > $O/bench_${TAG}_4k.json 2> $O/bench_${TAG}_4k.err
ncu --set full --clock-control none --import-source on -k regex:rm_wf_setup_kernel -s 6 -c 1 -o $O/prof_setup_${TAG} python bench.py --steps 2 --warmup 3 --contexts 1 --no-cpu-baseline --no-second-flavour > $O/ncu_setup_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $O/launches_${TAG}.csv python bench.py --steps 6 --warmup 5 --contexts 1 --no-cpu-baseline --no-second-flavour > $O/launches_${TAG}.log 2>&1
python tools/launch_summary.py $O/launches_${TAG}.csv > $O/launches_${TAG}_summary.txt 2>&1; cat $O/launches_${TAG}_summary.txt
