#!/usr/bin/env python
"""Per-CUDA-line instruction and stall-sample shares from `ncu -i X --page source --csv --print-source cuda,sass`.
usage: ncu_lines.py file.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
hdr = None
agg = {}
fname = ''
cur = None
for r in rows:
    if r and r[0] == 'File Name':
        fname = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No':
        hdr = r; iI = hdr.index('Instructions Executed'); iS = hdr.index('# Samples'); continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0]:                       # a CUDA line row
        cur = (fname, int(r[0]), r[1]); agg.setdefault(cur, [0, 0, 0]); 
        continue
    if cur is None or not r[iI].isdigit(): continue
    agg[cur][0] += int(r[iI]); agg[cur][1] += int(r[iS]); agg[cur][2] += 1
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("total warp instructions %d, samples %d" % (ti, ts))
for k in sorted(agg, key=lambda k: (k[0], k[1])):
    v = agg[k]
    if v[0] * 100 >= minpct * ti or v[1] * 100 >= minpct * ts:
        print("%s:%-5d inst %5.2f%%  samples %5.2f%%  sass %3d | %s" % (k[0], k[1], 100 * v[0] / ti, 100 * v[1] / max(ts, 1), v[2], k[2].strip()[:120]))
