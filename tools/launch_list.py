#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel and grid: usage launch_list.py launches.csv"""
import collections
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]
ki, vi, gi, bi = H.index('Kernel Name'), H.index('Metric Value'), H.index('Grid Size'), H.index('Block Size')
acc = collections.OrderedDict()
for r in rows[hdr + 2:]:
    if len(r) <= vi:
        continue
    acc.setdefault((r[ki].split('(')[0].replace('void ', '').replace('gvs::', ''), r[gi], r[bi]), []).append(float(r[vi].replace(',', '')))
tot = sum(sum(v) for v in acc.values())
for g, v in acc.items():
    print(f"{g[0][:34]:34s} {g[1]:16s} {g[2]:13s} n={len(v):3d} avg={sum(v) / len(v) / 1000:9.1f} us  share={100 * sum(v) / tot:5.1f}%")
