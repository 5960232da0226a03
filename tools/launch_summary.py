#!/usr/bin/env python3
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv)."""
import csv
import sys
from collections import defaultdict

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
d = defaultdict(list)
for r in rows:
    d[r["Kernel Name"].split("(")[0][:48]].append(float(r["Metric Value"]) / 1e3)
# rm_fp32_peak_kernel is bench.py's live measurement of the roofline denominator, not part of a step
probe = d.pop("rm_fp32_peak_kernel", [])
tot = sum(sum(v) for v in d.values())
print(f"# {sys.argv[1]}: {len(rows)} launches; step kernels {tot / 1e3:.3f} ms (per-launch times under ncu are serialised and cold-cache); "
      f"excluded: {len(probe)} launches of the FFMA peak probe")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:50s} n={len(v):4d} avg={sum(v) / len(v):9.1f} us  total={sum(v) / 1e3:9.3f} ms  share={100 * sum(v) / tot:5.1f} %")
