#!/usr/bin/env python
"""Where does an e2e mapping iteration (bench.E2EPath.step) spend its time?  torch.profiler table."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == "__main__":
    impl = sys.argv[1] if len(sys.argv) > 1 else "ours"
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    sc, cam, up = bench.make_workload(0, 1, dev)
    e = bench.E2EPath(sc, cam, dev, 1, impl)
    for i in range(5):
        e.step(i)
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for i in range(10):
        e.step(i)
    torch.cuda.synchronize()
    print("wall ms/step", (time.perf_counter() - t0) * 100)
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for i in range(5):
            e.step(i)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
