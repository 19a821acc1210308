#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, mean, share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) != len(hdr):
        continue
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] in ("ns", "nsecond") else v * (1e3 if r[iu] in ("ms", "msecond") else 1.0)   # -> us
    a = agg.setdefault(r[ik].split("(")[0][:48], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':50s} {'launches':>8s} {'total us':>10s} {'mean us':>9s} {'share':>6s}")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:50s} {n:8d} {t:10.1f} {t / n:9.1f} {t / tot * 100:5.1f}%")
print(f"{'total':50s} {sum(a[0] for a in agg.values()):8d} {tot:10.1f}")
