#!/usr/bin/env python
"""Per-SASS-line executed-instruction histogram from an .ncu-rep (source page): prints contiguous address regions
with their share of executed warp instructions.  usage: ncu_inst.py file.ncu-rep [bucket]"""
import csv, io, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
ci, si, wi = h.index('Instructions Executed'), h.index('Source'), h.index('Warp Stall Sampling (All Samples)')
body = rows[2:]
tot = sum(float(r[ci] or 0) for r in body) or 1
stot = sum(float(r[wi] or 0) for r in body) or 1
print('total warp instructions', tot, 'lines', len(body))
for k in range(0, len(body), bucket):
    blk = body[k:k + bucket]
    n = sum(float(r[ci] or 0) for r in blk)
    s = sum(float(r[wi] or 0) for r in blk)
    ops = {}
    for r in blk:
        op = r[si].strip().split()[0] if not r[si].strip().startswith('@') else r[si].strip().split()[1]
        ops[op] = ops.get(op, 0) + float(r[ci] or 0)
    top = sorted(ops.items(), key=lambda x: -x[1])[:5]
    print(f'#{k:5d}-{k + len(blk) - 1:5d} inst {n / tot * 100:5.1f}%  stall {s / stot * 100:5.1f}%  per-line {n / len(blk) / 1e3:8.1f}k  ' +
          ' '.join(f'{o}:{v / tot * 100:.1f}' for o, v in top))
