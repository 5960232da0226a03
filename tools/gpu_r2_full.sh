#!/bin/bash
# usage: bash tools/gpu_r2_full.sh <tag> : the whole -m gpu suite, the default bench (all arms), the reference arm,
# a launch list and ncu --set full captures of the march / setup / display kernels
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke_${TAG}.log 2>&1; tail -2 $O/smoke_${TAG}.log
( time python -m pytest tests -m gpu -x -q --durations=8 ) > $O/pytest_${TAG}.log 2>&1; tail -16 $O/pytest_${TAG}.log
python bench.py > $O/bench_${TAG}_default.json 2> $O/bench_${TAG}_default.err; tail -c 1500 $O/bench_${TAG}_default.json; tail -5 $O/bench_${TAG}_default.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_${TAG}_reference.json 2> $O/bench_${TAG}_reference.err; tail -c 400 $O/bench_${TAG}_reference.json
python bench.py --quick --width 3840 --height 2160 --steps 6 --frames-per-step 8 > $O/bench_${TAG}_4k.json 2> $O/bench_${TAG}_4k.err
python bench.py --quick --mode full --steps 4 --frames-per-step 4 > $O/bench_${TAG}_full.json 2> $O/bench_${TAG}_full.err
python - <<PY
import json
for n in ("default","4k","full"):
    try:
        d=json.load(open("$O/bench_${TAG}_%s.json"%n)); r=d["roofline"]
        print(n, round(d["value"],1), "Mpx/s  e2e", round(d["e2e"]["value"],1), " frac", round(r["frac"],4), "whole", round(r["whole_step_frac"],4), " kernel_ms/frame", round(r["kernel_ms_per_frame"],4), "frame_ms", round(d["extra"]["ms_per_frame"],4), "launches", d["gpu_launches"])
    except Exception as e:
        print(n, "FAILED", e); print(open("$O/bench_${TAG}_%s.err"%n).read()[-2000:])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 160 --csv --log-file $O/launches_${TAG}.csv python bench.py --quick --steps 2 --warmup 1 --frames-per-step 8 --contexts 1 > $O/launches_${TAG}.log 2>&1
python tools/launch_summary.py $O/launches_${TAG}.csv > $O/launches_${TAG}_summary.txt 2>&1; cat $O/launches_${TAG}_summary.txt
ncu --set full --clock-control none --import-source on -k regex:rm_wf_march_preview_kernel -s 12 -c 2 -o $O/prof_march_${TAG} python bench.py --quick --steps 1 --warmup 1 --frames-per-step 4 --contexts 1 > $O/ncu_march_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rm_wf_setup_kernel|rm_display_kernel|rm_wf_shade_preview_kernel|rm_wf_far_preview_kernel" -s 24 -c 4 -o $O/prof_stages_${TAG} python bench.py --quick --steps 1 --warmup 1 --frames-per-step 4 --contexts 1 > $O/ncu_stages_${TAG}.log 2>&1
