#!/usr/bin/env python3
"""Launch-shape sweep of the far-field pipeline in ONE process (the cubin cache is process-wide, the knobs are read at
every draw): RMB_PAUSE_LANES x RMB_MARCH_BLOCKS_PER_SM x RMB_RETURN_BLOCKS_PER_SM on the default bench workload.
Prints a table and writes the best setting as shell exports to gpurun_out/best_env.sh.
usage: python tools/sweep_carve.py [extra bench.py flags]"""
import contextlib
import io
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

VARIANTS = [(0, None, 2), (12, None, 2), (20, None, 2), (12, 4, 2), (20, 4, 2), (12, 3, 2), (20, 3, 3), (24, 4, 4), (16, 2, 2)]
KEYS = ("RMB_PAUSE_LANES", "RMB_MARCH_BLOCKS_PER_SM", "RMB_RETURN_BLOCKS_PER_SM")


def main():
    extra = sys.argv[1:]
    rows = []
    for v in VARIANTS:
        for k, x in zip(KEYS, v):
            if x is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(x)
        sys.argv = ["bench.py", "--no-cpu-baseline", "--no-second-flavour", "--steps", "20", "--warmup", "3"] + extra
        buf = io.StringIO()
        try:
            with contextlib.redirect_stdout(buf):
                bench.main()
            j = json.loads(buf.getvalue().strip().splitlines()[-1])
            rows.append((j["value"], v, j["e2e"]["value"], j["roofline"]["kernel_ms_per_step"], j["roofline"]["frac"]))
            print("pause=%s bps=%s ret=%s  value %.1f  e2e %.1f  march_ms %.4f  frac %.4f" % (*v, j["value"], j["e2e"]["value"],
                  j["roofline"]["kernel_ms_per_step"], j["roofline"]["frac"]), flush=True)
        except Exception as e:  # noqa: BLE001
            print("variant", v, "failed:", str(e)[:200], flush=True)
    if rows:
        best = max(rows)
        out = ROOT / "gpurun_out"
        out.mkdir(exist_ok=True)
        lines = ["export %s=%s" % (k, x) for k, x in zip(KEYS, best[1]) if x is not None]
        (out / "best_env.sh").write_text("\n".join(lines) + "\n")
        print("best:", best[1], "value %.1f" % best[0], flush=True)


if __name__ == "__main__":
    main()
