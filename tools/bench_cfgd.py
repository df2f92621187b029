#!/usr/bin/env python
"""BASELINE.json configs[3] ("cfgD"): ScanNet-shaped 2M Gaussians at 1296x968, one keyframe per iteration out of an
8-keyframe window, densify / prune every 100 iterations (reference schedule src/gaussian_mapper.cpp:737-761), through
`Mapper.train_step` + `Mapper.densify_and_prune` (all liblgs launches).  Beside it: the unmodified reference rasterizer +
eager torch loss + torch.optim.Adam on the same scene and views (its density control cannot be built here -- Eigen / OpenCV --
so that arm times iterations only; the torch restatement of densifyAndPrune is timed by tools/bench_extra.py).
Prints one JSON object (not a bench.py line).    python tools/bench_cfgd.py [--iters 300] [--P 2000000]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench  # noqa: E402
from leg_slam_b200 import mapper as M, synthetic  # noqa: E402

ROOM = (8.0, 6.0, 3.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--P", type=int, default=2_000_000)
    ap.add_argument("--ref-iters", type=int, default=20)
    ap.add_argument("--grad-threshold", type=float, default=0.0,
                    help="> 0: fixed densify_grad_threshold; 0: per round, the value that selects --grow-frac of the Gaussians")
    ap.add_argument("--grow-frac", type=float, default=0.03)
    ap.add_argument("--no-roundup", action="store_true")
    ap.add_argument("--profile-densify", action="store_true", help="cProfile of every densify_and_prune call -> stderr")
    ap.add_argument("--skip-ref", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    if not args.no_roundup:
        # sizes rounded to 1/8 of a power of two: the buffers of the slightly larger P after a densification round fit the blocks
        # the previous round released, so the caching allocator serves them without cudaMalloc (20-40 ms per GB on this box)
        torch.cuda.memory._set_allocator_settings("roundup_power2_divisions:8")
    W, H = 1296, 968
    sc = synthetic.make_scene(args.P, seed=4, room=ROOM, device=dev)
    cams_host = synthetic.make_cameras(8, W, H, fx=1169.7, fy=1169.7, room=ROOM, seed=4)
    cams = [c.to(dev) for c in cams_host]
    g = torch.Generator().manual_seed(9)
    win = [M.Keyframe(c, torch.rand(3, H, W, generator=g).to(dev), torch.randn(64, 37, 37, generator=g).to(dev),
                      (torch.rand(1, H, W, generator=g) * 3).to(dev)) for c in cams]
    lrs = {k: v * bench.LR_SCALE for k, v in M.DEFAULT_LRS.items()}
    out = dict(config=f"cfgD: {args.P} Gaussians (ScanNet-shaped room {ROOM}), {W}x{H}, fx=fy=1169.7, SH degree 3, 64-D feature, "
                      f"8-keyframe window, 1 keyframe per iteration, densify/prune every 100 iterations "
                      f"(grad threshold {args.grad_threshold if args.grad_threshold > 0 else f'selecting {args.grow_frac:.0%} per round'}, "
                      f"min opacity 0.02, size threshold 20, extent 5.0), allocator roundup_power2_divisions:8 = {not args.no_roundup}")

    m = M.Mapper(sc, lrs=lrs, sh_degree=3, track_densify_stats=True)
    for i in range(5):
        m.train_step([win[i % 8]], presharded=True)
    torch.cuda.synchronize()
    gen = torch.Generator(device=dev).manual_seed(3)
    seg, dens, t_seg = [], [], time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    Rs = []
    for it in range(1, args.iters + 1):
        m.train_step([win[it % 8]], presharded=True)
        Rs.append(m.last_num_rendered)
        if it % 100 == 0:
            torch.cuda.synchronize()
            now = time.perf_counter()
            seg.append(dict(iters=f"{it - 99}..{it}", P=int(m.params["xyz"].shape[0]), ms_per_iter=(now - t_seg) * 1e3 / 100))
            thr = args.grad_threshold
            if thr <= 0:  # the synthetic loss has no meaningful gradient scale: pick the threshold by the fraction it selects
                avg = (m.stats.xyz_gradient_accum / m.stats.denom.clamp_min(1)).reshape(-1)
                k = max(1, int(avg.numel() * (1.0 - args.grow_frac)))
                thr = float(torch.kthvalue(avg, k).values)
                torch.cuda.synchronize()
                now = time.perf_counter()
            if args.profile_densify:
                import cProfile
                import pstats
                pr = cProfile.Profile()
                pr.enable()
            info = m.densify_and_prune(thr, 0.02, 5.0, 20, generator=gen)
            if args.profile_densify:
                torch.cuda.synchronize()
                pr.disable()
                print(f"---- densify at iteration {it}", file=sys.stderr)
                pstats.Stats(pr, stream=sys.stderr).sort_stats("tottime").print_stats(8)
            info["grad_threshold"] = thr
            info["reserved_GB"] = round(torch.cuda.memory_reserved() / 1e9, 2)
            info["allocated_GB"] = round(torch.cuda.memory_allocated() / 1e9, 2)
            torch.cuda.synchronize()
            t_seg = time.perf_counter()
            dens.append(dict(at_iter=it, ms=(t_seg - now) * 1e3, **info))
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    out["ours"] = dict(iters=args.iters, wall_s=total, iters_per_s=args.iters / total, segments=seg, densify=dens,
                       mean_R=sum(Rs) / len(Rs), final_P=int(m.params["xyz"].shape[0]),
                       api="Mapper.train_step + Mapper.densify_and_prune (device-resident keyframes)")
    del m
    torch.cuda.empty_cache()

    # reference arm: iterations only, at the initial P
    if args.skip_ref:
        print(json.dumps(out))
        return
    try:
        bench.WIDTH, bench.HEIGHT = W, H
        e2e = bench.E2EPath(sc, cams_host[0], dev, 1, "reference", cams=cams_host[:1])
        mp = e2e.mapper
        for i in range(3):
            mp.train_step([win[i % 8]], presharded=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(args.ref_iters):
            mp.train_step([win[i % 8]], presharded=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["reference"] = dict(iters=args.ref_iters, ms_per_iter=dt * 1e3 / args.ref_iters, iters_per_s=args.ref_iters / dt, P=args.P,
                                path="unmodified reference rasterizer (oracle/_ref) + eager torch activations/loss + torch.optim.Adam; "
                                     "no density control in this arm")
    except Exception as ex:  # reference .so not shipped
        out["reference"] = dict(unavailable=f"{type(ex).__name__}: {ex}")
    del sc, win
    torch.cuda.empty_cache()
    out["cfgC_1gpu"] = cfgc(dev)
    print(json.dumps(out))


def cfgc(dev, iters=20):
    """BASELINE.json configs[2] on ONE GPU: Replica-shaped 1M Gaussians, 640x480, 8-keyframe window per iteration (gradients
    summed over the 8 views, one Adam step) -- the single-GPU end of the data-parallel scaling line."""
    W, H, P = 640, 480, 1_000_000
    sc = synthetic.make_scene(P, seed=3, device=dev)
    cams_host = synthetic.make_cameras(8, W, H, seed=3)
    cams = [c.to(dev) for c in cams_host]
    g = torch.Generator().manual_seed(10)
    win = [M.Keyframe(c, torch.rand(3, H, W, generator=g).to(dev), torch.randn(64, 37, 37, generator=g).to(dev),
                      (torch.rand(1, H, W, generator=g) * 3).to(dev)) for c in cams]
    lrs = {k: v * bench.LR_SCALE for k, v in M.DEFAULT_LRS.items()}
    res = dict(config=f"cfgC on 1 GPU: {P} Gaussians, {W}x{H}, 8 views per iteration, fwd+bwd per view + one Adam step")

    def run(mp, n):
        for _ in range(3):
            mp.train_step(win, presharded=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            mp.train_step(win, presharded=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return dict(ms_per_iteration=dt * 1e3 / n, views_per_s=8 * n / dt)
    m = M.Mapper(sc, lrs=lrs, sh_degree=3)
    res["ours"] = run(m, iters)
    del m
    torch.cuda.empty_cache()
    try:
        bench.WIDTH, bench.HEIGHT = W, H
        e2e = bench.E2EPath(sc, cams_host[0], dev, 1, "reference", cams=cams_host[:1])
        res["reference"] = run(e2e.mapper, 5)
    except Exception as ex:
        res["reference"] = dict(unavailable=f"{type(ex).__name__}: {ex}")
    return res


if __name__ == "__main__":
    main()
