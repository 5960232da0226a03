#!/bin/bash
# second GPU call of the far-field work: the new tests, a few launch-shape variants, full mode, launch list, one full ncu capture
O=gpurun_out; mkdir -p $O
B="python bench.py --no-cpu-baseline --no-second-flavour"
timeout 90 python -m pytest tests -m gpu -q -k "carve" > $O/carve2_tests.log 2>&1; tail -2 $O/carve2_tests.log
timeout 50 $B > $O/carve2_base.json 2> $O/carve2_base.err
RMB_PAUSE_LANES=12 timeout 50 $B > $O/carve2_pause12.json 2> $O/carve2_pause12.err
RMB_MARCH_BLOCKS_PER_SM=4 timeout 50 $B > $O/carve2_bps4.json 2> $O/carve2_bps4.err
timeout 70 $B --mode full > $O/carve2_full.json 2> $O/carve2_full.err
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_preview_exact_r1carve.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-second-flavour > $O/ncu_launches_r1carve.log 2>&1
timeout 90 ncu --set full --clock-control none --import-source on -k regex:rm_wf_march_preview -s 6 -c 1 -o $O/prof_march_preview_exact_r1carve $B --steps 2 --warmup 3 --contexts 1 > $O/ncu_r1carve.log 2>&1
python - <<'PY'
import json
for n in ("base", "pause12", "bps4", "full"):
    try:
        j = json.loads(open(f"gpurun_out/carve2_{n}.json").read().strip().splitlines()[-1])
        r = j["roofline"]
        print(n, "value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "ms", round(j["ms_per_step"], 4), "frac", round(r["frac"], 4),
              "far", round(r.get("far_field_evals_share", 0), 3), "kernel_ms", round(r["kernel_ms_per_step"], 4), "launches", j["gpu_launches"])
    except Exception as e:
        print(n, "failed", e)
PY
ls -la $O/*r1carve*
