#!/usr/bin/env python
"""Distribution of the render backward's work items at cfgB: half-records per (tile, pixel-warp half), read from the backward
scratch buffer after one iteration of bench.KernelPath (layout: render_bwd.cu launch_render_bwd)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
sc, cam, up = bench.make_workload(int(os.environ.get("VIEW", "0")), 1, dev)
kp = bench.KernelPath(sc, cam, up, dev, 1)
for _ in range(2):
    kp.step()
torch.cuda.synchronize()
R = kp.R
tiles = ((bench.WIDTH + 7) // 8) * ((bench.HEIGHT + 7) // 8)
base = (kp.scratch_bwd.data_ptr() + 255) & ~255
off = base - kp.scratch_bwd.data_ptr() + 2 * R * 272
cnt = kp.scratch_bwd[off:off + 8 * tiles].view(torch.int32).cpu().numpy().astype(np.int64)
print("items", cnt.size, "records", int(cnt.sum()), "max", int(cnt.max()), "mean", float(cnt.mean()))
print("percentiles 50/90/99/99.9:", [int(np.percentile(cnt, p)) for p in (50, 90, 99, 99.9)])
b = (cnt + 63) // 64
print("batches", int(b.sum()), "max per item", int(b.max()), "items >= 8 batches:", int((b >= 8).sum()), "holding",
      int(b[b >= 8].sum()), "batches")
# list-scheduling simulation: 296 CTAs take items in queue order; cost of an item = 1 + batches (arbitrary units)
import heapq
def simulate(order, n_cta=296):
    h = [0.0] * n_cta
    heapq.heapify(h)
    for i in order:
        t = heapq.heappop(h)
        heapq.heappush(h, t + 0.3 + b[i])
    return max(h)
ideal = (0.3 * cnt.size + b.sum()) / 296
print("makespan / ideal: queue order %.3f, heavy first %.3f, sorted descending %.3f" % (
    simulate(range(cnt.size)) / ideal, simulate(np.argsort(-(b >= 8).astype(int), kind="stable")) / ideal,
    simulate(np.argsort(-b, kind="stable")) / ideal))
print("last 32 items of the queue:", b[-32:].tolist())
