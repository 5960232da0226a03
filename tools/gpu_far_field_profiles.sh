#!/bin/bash
# far-field pipeline: shipped defaults sanity bench, launch list, 4K, one full ncu capture of the march kernel
O=gpurun_out; mkdir -p $O
B="python bench.py --no-cpu-baseline --no-second-flavour"
timeout 40 $B > $O/far_field_default.json 2> $O/far_field_default.err
timeout 50 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_preview_exact_r1final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-second-flavour > $O/ncu_launches_r1final.log 2>&1
timeout 40 $B --width 3840 --height 2160 --steps 12 > $O/far_field_4k.json 2> $O/far_field_4k.err
timeout 60 ncu --set full --clock-control none --import-source on -k regex:rm_wf_march_preview -s 6 -c 1 -o $O/prof_march_preview_exact_r1final $B --steps 2 --warmup 3 --contexts 1 > $O/ncu_r1final.log 2>&1
python - <<'PY'
import json
for n in ("default", "4k"):
    try:
        j = json.loads(open(f"gpurun_out/far_field_{n}.json").read().strip().splitlines()[-1])
        r = j["roofline"]
        print(n, "value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "ms", round(j["ms_per_step"], 4), "frac", round(r["frac"], 4),
              "far", round(r.get("far_field_evals_share", 0), 3), "kernel_ms", round(r["kernel_ms_per_step"], 4), "launches", j["gpu_launches"])
    except Exception as e:
        print(n, "failed", e)
PY
