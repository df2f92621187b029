#!/usr/bin/env python
"""Opcode histogram per kernel of liblgs.so (cuobjdump -sass): the evidence that the hot kernels are Blackwell-native --
UTCHMMA / UTCBAR (tcgen05.mma / commit), LDTM (tcgen05.ld), UBLKCP (TMA bulk copy), LDGSTS (cp.async), LDGMC / multimem
(NVSwitch multicast), FFMA2 (packed fp32), MATCH (warp match in the radix sort), RED / ATOM.
    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "leg_slam_b200", "liblgs.so")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "LDGSTS", "LDGMC", "STGMC", "REDG", "RED", "ATOMG", "ATOMS", "FFMA2", "FFMA",
         "MUFU", "MATCH", "SHFL", "HMMA", "LDS", "STS", "LDG", "STG", "BAR", "SYNCS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass leg_slam_b200/liblgs.so : opcode counts per kernel (static instruction counts, sm_100a)")
    print("# columns: total |", " ".join(WATCH))
    tot = collections.Counter()
    for (name, c), dn in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dn).replace("lgs::", "").replace("void ", "")
        row = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
        print(f"{short[:70]:70s} total={sum(c.values()):6d} | {row}")
        tot.update(c)
    print("# library totals:", " ".join(f"{w}={tot[w]}" for w in WATCH if tot[w]))


if __name__ == "__main__":
    main()
