#!/bin/bash
# one GPU call: the GPU test-suite, then the default bench with and without the far-field pipeline
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q --maxfail=6 > gpurun_out/carve_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/carve_tests.log
tail -25 gpurun_out/carve_tests.log
timeout 120 python bench.py > gpurun_out/carve_bench_on.json 2> gpurun_out/carve_bench_on.err
RMB_CARVE=0 timeout 60 python bench.py --no-cpu-baseline --no-second-flavour > gpurun_out/carve_bench_off.json 2> gpurun_out/carve_bench_off.err
python - <<'PY'
import json
for n in ("on", "off"):
    try:
        j = json.loads(open(f"gpurun_out/carve_bench_{n}.json").read().strip().splitlines()[-1])
        r = j["roofline"]
        print(n, "value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "ms", round(j["ms_per_step"], 4), "frac", round(r["frac"], 4),
              "far", round(r.get("far_field_evals_share", 0), 3), "kernel_ms", round(r["kernel_ms_per_step"], 4), "launches", j["gpu_launches"],
              "fast", (j.get("extra", {}).get("fast_flavour") or {}).get("value"))
    except Exception as e:
        print(n, "failed", e)
PY
