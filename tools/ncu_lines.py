#!/usr/bin/env python
"""Warp-stall samples and executed instructions per CUDA source line from an .ncu-rep captured with --import-source on.
usage: ncu_lines.py file.ncu-rep kernel-regex [n]"""
import csv, io, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'cuda,sass', '-k', 'regex:' + sys.argv[2]],
                     capture_output=True, text=True).stdout
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, agg = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        hdr = None
    elif r[0] == 'Line No':
        hdr = r
    elif hdr is not None and r[0].isdigit() and len(r) > 8 and r[2] == '-':  # a CUDA line (SASS rows carry an address)
        wi, ii = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
        st = {h[6:]: float(r[i] or 0) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}
        key = (fname, int(r[0]))
        a = agg.setdefault(key, dict(src=r[1].strip(), s=0.0, i=0.0, st={}))
        a['s'] += float(r[wi] or 0)
        a['i'] += float(r[ii] or 0)
        for k, v in st.items():
            a['st'][k] = a['st'].get(k, 0) + v
ts = sum(a['s'] for a in agg.values()) or 1
ti = sum(a['i'] for a in agg.values()) or 1
print(f'samples {ts:.0f}  warp instructions {ti:.0f}')
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]['s'])[:n]:
    why = sorted(a['st'].items(), key=lambda kv: -kv[1])[:2]
    print(f"{a['s'] / ts * 100:5.1f}% st {a['i'] / ti * 100:5.1f}% in  {f}:{ln:<4d} {a['src'][:80]:80s} " + ' '.join(f'{k}:{v:.0f}' for k, v in why))
