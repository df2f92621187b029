#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from an .ncu-rep (source page).  usage: ncu_hot.py file.ncu-rep [n]"""
import csv, io, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'rows': []}
        blocks.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = r
    elif cur is not None:
        cur['rows'].append(r)
for b in blocks:
    h = b['hdr']
    ci, si = h.index('Warp Stall Sampling (All Samples)'), h.index('Source')
    st = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
    tot = sum(float(r[ci] or 0) for r in b['rows']) or 1
    print('=' * 20, b['name'][:80], 'samples', tot)
    for k, r in enumerate(b['rows']):
        r.append(k)
    for r in sorted(b['rows'], key=lambda r: -float(r[ci] or 0))[:n]:
        why = sorted(((float(r[i] or 0), h[i][6:]) for i in st), reverse=True)[:2]
        print(f"{float(r[ci]) / tot * 100:5.1f}%  #{r[-1]:5d} {r[si].strip()[:70]:70s} {why[0][1]}:{why[0][0]:.0f} {why[1][1]}:{why[1][0]:.0f}")
