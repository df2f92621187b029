#!/usr/bin/env python
"""Time lgs_adam_multi alone on the cfgB parameter set (7 tensors, 123 floats per Gaussian).  python tools/bench_adam.py [P]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leg_slam_b200 import FusedAdam  # noqa: E402

if __name__ == "__main__":
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
    dev = torch.device("cuda:0")
    rows = [3, 3, 45, 64, 1, 3, 4]
    ps = [torch.nn.Parameter(torch.randn(P, r, device=dev)) for r in rows]
    for p in ps:
        p.grad = torch.randn_like(p)
    opt = FusedAdam([dict(params=[p], lr=1e-3) for p in ps], eps=1e-15)
    for _ in range(5):
        opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    n = 50
    e0.record()
    for _ in range(n):
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"P={P} variant={os.environ.get('LGS_ADAM_VARIANT', '0')}: {ms:.4f} ms  {P * 123 * 28 / ms / 1e6:.0f} GB/s")
