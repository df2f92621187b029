#!/usr/bin/env python
"""Times the geometry operators next to the mapping path (leg_slam_b200.ingest) against the compiled, unmodified reference
operators (oracle/_ref/ref_geometry.so, ref_simple_knn.so) on one GPU: CUDA events, warm-up, median of `reps`.
    python tools/bench_geometry.py [--out gpurun_out/geometry_bench.json]
The loop-closure operator works in place and changes its inputs (flags are cleared), so every repetition runs on a fresh
copy made outside the timed region; both arms include the host read-back of the count the reference's interface has."""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def timed(fn, setup, reps=20, warm=3):
    ts = []
    for i in range(warm + reps):
        args = setup()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(*args)
        b.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import build_ref
    import test_ingest as TI
    from leg_slam_b200 import ingest
    ref = build_ref.load_geometry()
    dev = torch.device("cuda:0")
    res = {}
    g = torch.Generator().manual_seed(1)
    # loop closure over a 2 M Gaussian map
    P = 2_000_000
    pts, rots, nt, un = TI._loop_closure_case(P, 5)
    T, view, proj = TI._pose(g).to(dev), TI._pose(g, t=(0.1, 0.2, 0.5)).to(dev), torch.eye(4, device=dev)
    base = [t.to(dev) for t in (pts, rots, nt, un)]
    setup = lambda: [t.clone() for t in base]
    res["loop_closure_P2M_ms"] = {
        "ours": timed(lambda p, r, n, u: ingest.scaleAndTransformThenMarkVisiblePoints(p, r, n, u, T, view, proj, 0, 1.0), setup),
        "reference": timed(lambda p, r, n, u: ref.scale_and_transform_then_mark_visible(p, r, n, u, T, view, proj, 0, 1.0), setup)}
    # inactive-geometry densification over the keypoints of one frame
    for N in (2000, 8000):
        px, has, p3, colors = [t.to(dev) for t in TI._keypoint_case(N, 1296, 968, 9)]
        intr = [1169.7, 1169.7, 647.5, 483.5]
        f = ingest.monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints
        res[f"inactive_geo_N{N}_ms"] = {"ours": timed(lambda: f(px, has, p3, colors, 900.0, intr, 1296), lambda: []),
                                        "reference": timed(lambda: ref.inactive_geo_densify(px, has, p3, colors, 900.0, intr, 1296), lambda: [])}
    # depth image -> points -> world, one 1296x968 keyframe
    W, H = 1296, 968
    depth = (torch.rand(W * H, generator=g) * 5 + 0.1).to(dev)
    mask = (torch.rand(W * H, generator=g) > 0.3).to(dev)
    intr = [1169.7, 1169.7, 647.5, 483.5]
    res["reproject_1296x968_ms"] = {"ours": timed(lambda: ingest.reprojectDepthPinhole(depth, mask, intr, W), lambda: []),
                                    "reference": timed(lambda: ref.reproject_depth_pinhole(depth, mask, intr, W), lambda: [])}
    cam = ingest.reprojectDepthPinhole(depth, mask, intr, W)
    res["transform_1.25M_ms"] = {"ours": timed(lambda: ingest.transformPoints(cam, T), lambda: []),
                                 "reference": timed(lambda: ref.transform_points(cam.clone(), T), lambda: [])}
    res["gpu"] = torch.cuda.get_device_name(0)
    print(json.dumps(res, indent=1))
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
