#!/usr/bin/env python
"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from leg_slam_b200 import cosine_query, mapper as M, rasterize_points as rp, relevance_scores, synthetic  # noqa: E402

dev = torch.device("cuda:0")
for name in ("ragged_sh1", "precomp_nolf", "dense_opaque"):
    cs = cases.make_case(name, dev)
    R, color, lf, depth, radii, geom, binning, img = rp.rasterize_gaussians(*cases.fwd_args(cs))
    rp.rasterize_gaussians_backward(*cases.bwd_args(cs, radii, geom, R, binning, img))
    rp.mark_visible(cs["means3D"], cs["viewmatrix"], cs["projmatrix"])
sc = synthetic.make_scene(3000, seed=1, mean_scale=0.06, device=dev)
cams = synthetic.make_cameras(2, 93, 61, seed=1)
g = torch.Generator().manual_seed(2)
win = [M.Keyframe(c.to(dev), torch.rand(3, 61, 93, generator=g).to(dev), torch.randn(64, 37, 37, generator=g).to(dev),
                  (torch.rand(1, 61, 93, generator=g) * 3).to(dev)) for c in cams]
m = M.Mapper(sc, sh_degree=3)
for _ in range(2):
    m.train_step(win)
f = torch.randn(1000, 64, generator=g).to(dev)
t = torch.randn(20, 64, generator=g).to(dev)
cosine_query(f, t)
cosine_query(f, t, simt=True)
relevance_scores(f, t[0])
torch.cuda.synchronize()
print("sanitize target ok")
