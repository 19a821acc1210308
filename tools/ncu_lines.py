#!/usr/bin/env python
"""Hot CUDA source lines of a capture: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv; ncu_lines.py x.csv [top]`.
Aggregates the stall samples of the SASS rows under each (file, line) row."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg, fname, hdr, cur = {}, None, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_smp, i_ex = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0]:                                  # a source line row (its own columns are the totals of the SASS rows below)
        cur = (fname, int(r[0]), r[1].strip()[:110])
        s, e = float(r[i_smp] or 0), float(r[i_ex] or 0)
        a = agg.setdefault(cur, [0.0, 0.0])
        a[0] += s; a[1] += e
tot = sum(v[0] for v in agg.values()) or 1
print(f"{tot:.0f} samples")
for (f, ln, src), (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * s / tot:5.1f}%  {e:10.0f} instr  {f}:{ln}  {src}")
