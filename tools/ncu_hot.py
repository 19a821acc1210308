#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` output: hot SASS instructions by executed count and stall samples."""
import csv
import sys

path = sys.argv[1]
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
rows = list(csv.reader(open(path)))
k = 0
while k < len(rows):
    if rows[k] and rows[k][0] == "Kernel Name":
        name = rows[k][1]
        hdr = rows[k + 1]
        j = k + 2
        data = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr):
                data.append(rows[j])
            j += 1
        ia, isrc, ismp, iex, ith = (hdr.index(x) for x in ("Address", "Source", "# Samples", "Instructions Executed", "Avg. Threads Executed"))

        def num(x):
            try:
                return float(x)
            except ValueError:
                return 0.0
        tot = sum(num(r[ismp]) for r in data) or 1
        totex = sum(num(r[iex]) for r in data) or 1
        print(f"== {name}: {len(data)} SASS, {totex:.0f} warp-instr, {tot:.0f} samples")
        for i, r in enumerate(data):
            ex, sm = num(r[iex]), num(r[ismp])
            if ex > thresh * totex or sm > thresh * tot:
                print(f"{i:5d} {ex / totex * 100:5.2f}%ex {sm / tot * 100:5.2f}%smp thr={r[ith]:>5} {r[isrc][:100]}")
        k = j
    else:
        k += 1
