#!/usr/bin/env python
"""Profiling target: N kernel-path mapping iterations of cfgB (the same KernelPath bench.py times),
nothing else, so ncu captures stay short.   python tools/prof_step.py [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    sc, cam, up = bench.make_workload(0, 1, dev)
    kp = bench.KernelPath(sc, cam, up, dev, 1)
    for i in range(steps):
        kp.step(i)
    torch.cuda.synchronize()
    print("ok R=%d" % kp.R)
