#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`): launches, total time and share per kernel.
Usage: launch_summary.py launches.csv "command that was profiled" > profiles/<round>_launches_summary.txt"""
import csv, sys
from collections import defaultdict
path = sys.argv[1]
cmd = sys.argv[2] if len(sys.argv) > 2 else "?"
lines = [l for l in open(path) if l.startswith('"')]
agg = defaultdict(lambda: [0, 0.0])
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    a = agg[r["Kernel Name"]]
    a[0] += 1
    a[1] += float(r["Metric Value"].replace(",", "")) * scale.get(r["Metric Unit"], 1.0)
tot = sum(v[1] for v in agg.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none, command: %s" % cmd)
print("# (per-launch times under ncu are cold-cache and serialised: shares, not absolutes)")
print("# launches  total_us  share  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6d %10.1f %6.1f%%  %s" % (v[0], v[1], 100 * v[1] / tot, k[:110]))
