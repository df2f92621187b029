#!/usr/bin/env python
"""Development diagnostic (GPU box): our path vs the compiled reference on the same inputs.
Prints every parity metric instead of stopping at the first failure."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from leg_slam_b200 import synthetic, debug  # noqa: E402
from leg_slam_b200 import rasterize_points as ours  # noqa: E402
import build_ref  # noqa: E402
import refbuf  # noqa: E402

ref = build_ref.load()
dev = torch.device("cuda:0")


def rel(a, b):
    d = (a.double() - b.double()).abs().max().item()
    n = b.double().abs().max().item()
    return d / max(n, 1e-30)


def run(P, W, H, seed=0, include_lf=True, timing=False):
    print(f"==== P={P} {W}x{H} lf={include_lf}")
    sc = synthetic.make_scene(P, seed=seed, device=dev)
    cam = synthetic.make_cameras(1, W, H, seed=seed)[0].to(dev)
    a = synthetic.activate(sc)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    empty = torch.empty(0, device=dev)
    args = (bg, a["means3D"], empty, a["lang_feats"], a["opacities"], a["scales"], a["rotations"], 1.0, empty,
            cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, H, W, a["shs"], 3, cam.campos, False, include_lf)
    Rr, cr, lr, dr, radr, gr, br, ir = ref.rasterize_gaussians(*args)
    Ro, co, lo, do, rado, go, bo, io = ours.rasterize_gaussians(*args)
    torch.cuda.synchronize()
    print("num_rendered ref/ours", Rr, Ro, "visible", int((radr > 0).sum()))
    print("radii equal:", bool((radr == rado).all()), "mismatches", int((radr != rado).sum()))
    vg = refbuf.ref_geom_view(gr, P)
    og = debug.geom_view(go, P)
    vis = radr > 0
    rec = og["records"]
    print("depth bits equal:", bool((vg["depths"][vis].view(torch.int32) == rec[vis, 2].view(torch.int32)).all()))
    print("means2D bits equal:", bool((vg["means2D"][vis].view(torch.int32) == rec[vis, 0:2].contiguous().view(torch.int32)).all()))
    print("cov3D bits equal:", bool((vg["cov3D"][vis].view(torch.int32) == og["cov3D"][vis].view(torch.int32)).all()),
          "rel", rel(og["cov3D"][vis], vg["cov3D"][vis]))
    cmism = (vg["conic_opacity"][vis].view(torch.int32) != rec[vis, 4:8].contiguous().view(torch.int32)).sum().item()
    print("conic_opacity bit mismatches:", cmism, "rel", rel(rec[vis, 4:8], vg["conic_opacity"][vis]))
    print("rgb rel:", rel(rec[vis, 8:11], vg["rgb"][vis]))
    print("tiles_touched equal:", bool((vg["tiles_touched"] == og["tiles_touched"]).all()))
    if Rr == Ro and Rr > 0:
        vb = refbuf.ref_binning_view(br, Rr)
        ob = debug.reference_keys(go, bo, io, P, Ro, W, H)
        for k_ref, k_our in (("keys_unsorted", "keys_unsorted"), ("point_list_unsorted", "values_unsorted"),
                             ("keys_sorted", "keys_sorted"), ("point_list", "point_list")):
            print(f"{k_our} equal:", bool((vb[k_ref] == ob[k_our]).all()))
        vi = refbuf.ref_image_view(ir, W, H)
        oi = debug.image_view(io, W, H)
        print("ranges equal:", bool((vi["ranges"] == oi["ranges"]).all()))
        print("n_contrib equal:", bool((vi["n_contrib"] == oi["n_contrib"]).all()),
              "mismatch", int((vi["n_contrib"] != oi["n_contrib"]).sum()))
        print("final_T rel:", rel(oi["final_T"], vi["final_T"]))
    print("color rel", rel(co, cr), "lf rel", rel(lo, lr), "depth rel", rel(do, dr))

    # ---- backward
    g = torch.Generator(device="cpu").manual_seed(seed + 7)
    dc = (torch.randn(3, H, W, generator=g) / (H * W)).to(dev)
    dl = (torch.randn(64, H, W, generator=g) / (H * W)).to(dev)
    dd = (torch.randn(1, H, W, generator=g) / (H * W)).to(dev)
    bargs = lambda rad, gb, R, bb, ib: (bg, a["means3D"], rad, empty, a["lang_feats"], a["scales"], a["rotations"], 1.0,  # noqa: E731
                                        empty, cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, dc, dl, dd,
                                        a["shs"], 3, cam.campos, gb, R, bb, ib, include_lf)
    gr_ = ref.rasterize_gaussians_backward(*bargs(radr, gr, Rr, br, ir))
    go_ = ours.rasterize_gaussians_backward(*bargs(rado, go, Ro, bo, io))
    torch.cuda.synchronize()
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dlang_feats", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh",
             "dL_dscales", "dL_drotations"]
    for n, x, y in zip(names, go_, gr_):
        print(f"  {n:16s} rel {rel(x, y):.3e}   |ref|max {y.abs().max().item():.3e}")

    if timing:
        def t(fn, n=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        print("  time fwd ref %.3f ms  ours %.3f ms" % (t(lambda: ref.rasterize_gaussians(*args)),
                                                        t(lambda: ours.rasterize_gaussians(*args))))
        print("  time bwd ref %.3f ms  ours %.3f ms" % (t(lambda: ref.rasterize_gaussians_backward(*bargs(radr, gr, Rr, br, ir))),
                                                        t(lambda: ours.rasterize_gaussians_backward(*bargs(rado, go, Ro, bo, io)))))


if __name__ == "__main__":
    t0 = time.time()
    run(10_000, 320, 240, timing=True)
    run(10_000, 317, 235, seed=3)
    run(100_000, 640, 480, seed=1, timing=True)
    run(500_000, 640, 480, seed=2, timing=True)
    run(10_000, 320, 240, include_lf=False, seed=4)
    print("done in %.1fs" % (time.time() - t0))
