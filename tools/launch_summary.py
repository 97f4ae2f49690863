#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total / average device time
and share.  python tools/launch_summary.py launches.csv [skip_fraction]  (skip_fraction of the launches from the start are
ignored: warm-up)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
body = rows[1:]
body = body[int(len(body) * skip):]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in body:
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    k = r[ki].split("(")[0]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{len(body)} launches, {tot:.1f} us in kernels")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:58s} n={v[0]:5d} total={v[1]:10.1f} us avg={v[1] / v[0]:9.2f} us share={v[1] / tot:.3f}")
