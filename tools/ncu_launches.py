#!/usr/bin/env python
"""Sums an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`) per kernel.
usage: ncu_launches.py launches.csv "comment for the header line" > profiles/<round>_launches_summary.txt"""
import collections
import csv
import sys

rows = [ln for ln in open(sys.argv[1]) if ln.startswith('"')]
agg = collections.defaultdict(lambda: [0, 0.0])
for d in csv.DictReader(rows):
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", ""))
    unit = d.get("Metric Unit", "ns")
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    k = d["Kernel Name"][:80]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("# total ms, launches, share, kernel")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{v[1] / 1e6:10.3f} {v[0]:4d} {100 * v[1] / tot:5.1f}% {k}")
