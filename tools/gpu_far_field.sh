#!/bin/bash
# far-field pipeline: launch-shape sweep, then the whole GPU suite and the default bench under the best setting
O=gpurun_out; mkdir -p $O; rm -f $O/best_env.sh
timeout 100 python tools/sweep_carve.py > $O/far_field_sweep.txt 2>&1; cat $O/far_field_sweep.txt | grep -v Warning | tail -12
[ -f $O/best_env.sh ] && . $O/best_env.sh
env | grep RMB_ > $O/far_field_env.txt
timeout 150 python -m pytest tests -m gpu -q --maxfail=6 > $O/far_field_tests.log 2>&1; echo "tests rc=$?" >> $O/far_field_tests.log; tail -4 $O/far_field_tests.log
timeout 100 python bench.py > $O/far_field_bench.json 2> $O/far_field_bench.err
timeout 60 python bench.py --mode full --no-cpu-baseline > $O/far_field_bench_full.json 2> $O/far_field_bench_full.err
python - <<'PY'
import json
for n in ("bench", "bench_full"):
    try:
        j = json.loads(open(f"gpurun_out/far_field_{n}.json").read().strip().splitlines()[-1])
        r = j["roofline"]
        print(n, "value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "ms", round(j["ms_per_step"], 4), "frac", round(r["frac"], 4),
              "far", round(r.get("far_field_evals_share", 0), 3), "kernel_ms", round(r["kernel_ms_per_step"], 4), "launches", j["gpu_launches"],
              "fast", (j.get("extra", {}).get("fast_flavour") or {}).get("value"))
    except Exception as e:
        print(n, "failed", e)
PY
