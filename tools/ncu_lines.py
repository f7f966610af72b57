#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export by CUDA source line.
usage: ncu_lines.py export.csv <function-substring> [top]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cur = None; hdr = None
agg = defaultdict(lambda: [0, 0, ''])
curline = None
for r in rows:
    if not r: continue
    if r[0] == 'Function Name': cur = r[1]; hdr = None; continue
    if r[0] == 'File Path': continue
    if r[0] == 'Line No':
        hdr = r if 'Instructions Executed' in r and '# Samples' in r else None
        if hdr: ii = hdr.index('Instructions Executed'); si = hdr.index('# Samples')
        continue
    if hdr is None or cur is None or want not in cur: continue
    if r[0] != '-' and r[0] != '':
        curline = r[0]; agg[curline][2] = r[1].strip()[:120]
    try:
        if r[2] != '-':
            agg[curline][0] += int(r[ii]); agg[curline][1] += int(r[si])
    except Exception: pass
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print('total instr', tot, 'samples', ts)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{100*v[0]/max(tot,1):5.1f}% instr {100*v[1]/max(ts,1):5.1f}% smp  L{k}: {v[2]}')
