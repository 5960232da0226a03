#!/bin/bash
# Round profile recipe (B200_PROFILING.md): bench lines, ncu launch list, ncu --set full of the top kernel.
# usage: bash tools/gpu_profile.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_preview_exact.json 2> gpurun_out/bench_${TAG}_preview_exact.err
python bench.py --steps 20 --warmup 5 --flavour fast --no-cpu-baseline --no-second-flavour > gpurun_out/bench_${TAG}_preview_fast.json 2> gpurun_out/bench_${TAG}_preview_fast.err
python bench.py --steps 5 --warmup 3 --mode full --no-cpu-baseline --no-second-flavour > gpurun_out/bench_${TAG}_full_exact.json 2> gpurun_out/bench_${TAG}_full_exact.err
python bench.py --steps 5 --warmup 3 --mode full --flavour fast --no-cpu-baseline --no-second-flavour > gpurun_out/bench_${TAG}_full_fast.json 2> gpurun_out/bench_${TAG}_full_fast.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-second-flavour > gpurun_out/ncu_launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rm_preview -s 3 -c 2 -o gpurun_out/prof_preview_exact_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-second-flavour > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rm_preview -s 3 -c 2 -o gpurun_out/prof_preview_fast_${TAG} \
    python bench.py --steps 2 --warmup 3 --flavour fast --no-cpu-baseline --no-second-flavour >> gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/*.err
cat gpurun_out/bench_${TAG}_*.json
