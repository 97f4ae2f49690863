#!/usr/bin/env python
"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` export (needs -lineinfo)."""
import csv, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
lines = {}
for row in csv.reader(open(path)):
    if not row: continue
    if row[0] == "File Path": cur_file = row[1].split("/")[-1]; continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No": hdr = row; continue
    if hdr is None or row[2] != "-": continue      # only source-line rows (Address == "-")
    try:
        ln = int(row[0])
    except ValueError:
        continue
    isamp = hdr.index("# Samples"); iinst = hdr.index("Instructions Executed"); ith = hdr.index("Thread Instructions Executed")
    k = (cur_file, ln)
    s, i, t, src = lines.get(k, (0, 0, 0, row[1]))
    lines[k] = (s + int(row[isamp]), i + int(row[iinst]), t + int(row[ith]), row[1])
ts = sum(v[0] for v in lines.values()); ti = sum(v[1] for v in lines.values()); tt = sum(v[2] for v in lines.values())
print(f"samples {ts}  warp inst {ti}  avg threads {tt/max(ti,1):.2f}")
for (f, ln), (s, i, t, src) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f:22s}:{ln:4d} inst {100*i/ti:5.1f}%  samples {100*s/ts:5.1f}%  thr {t/max(i,1):5.1f}  {src.strip()[:100]}")
