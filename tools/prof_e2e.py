#!/usr/bin/env python
"""Where the end-to-end step goes: host enqueue time per step (no synchronisation) against device time per step, for the
mapper path of bench.py (ours).   python tools/prof_e2e.py [--steps 200]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--tiny", action="store_true", help="2000 Gaussians at 64x48: the device work vanishes, what is left is the host's cost per step")
    args = ap.parse_args()
    if args.tiny:
        bench.P_GAUSS, bench.WIDTH, bench.HEIGHT = 2000, 64, 48
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    sc, cam, up = bench.make_workload(0, 1, dev)
    e2e = bench.E2EPath(sc, cam, dev, 1, "ours")
    for i in range(20):
        e2e.step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        e2e.step(i)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    out = dict(host_enqueue_ms_per_step=(t1 - t0) * 1e3 / args.steps, wall_ms_per_step=(t2 - t0) * 1e3 / args.steps,
               device_ms_per_step=e0.elapsed_time(e1) / args.steps, last_num_rendered=e2e.mapper.last_num_rendered,
               overflow_steps=e2e.mapper.overflow_steps)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
