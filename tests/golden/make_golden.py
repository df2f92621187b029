#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference, on a GPU box.

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'    # then copy *.npz here

Runs the reference's own rasterizer (oracle/_ref/ref_rasterizer.so = the reference sources
compiled for sm_100 by oracle/build_ref.py, nothing of ours on the path) on the seeded cases
of tests/cases.py and stores its outputs; torch.optim.Adam (unfused, eps=1e-15, the reference's
7 parameter groups) and torch's F.normalize+matmul likewise.  These files are what pins the
CPU oracle (tests/test_oracle_golden.py) and, through it and directly, the CUDA path.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import cases  # noqa: E402
import refbuf  # noqa: E402
import build_ref  # noqa: E402


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    ref = build_ref.load()
    dev = torch.device("cuda:0")
    for name in cases.CASES:
        cs = cases.make_case(name, dev)
        R, color, lf, depth, radii, geom, binning, img = ref.rasterize_gaussians(*cases.fwd_args(cs))
        grads = ref.rasterize_gaussians_backward(*cases.bwd_args(cs, radii, geom, R, binning, img))
        torch.cuda.synchronize()
        P, W, H = cs["P"], cs["W"], cs["H"]
        gv, bv, iv = refbuf.ref_geom_view(geom, P), refbuf.ref_binning_view(binning, R), refbuf.ref_image_view(img, W, H)
        vis = (radii > 0)
        n = lambda t: t.detach().cpu().numpy()  # noqa: E731
        d = dict(num_rendered=np.int64(R), radii=n(radii), visible=n(vis),
                 depths=n(gv["depths"] * vis), means2D=n(gv["means2D"] * vis[:, None]),
                 conic_opacity=n(gv["conic_opacity"] * vis[:, None]), tiles_touched=n(gv["tiles_touched"]),
                 keys_sorted=n(bv["keys_sorted"]), point_list=n(bv["point_list"]),
                 keys_unsorted=n(bv["keys_unsorted"]), values_unsorted=n(bv["point_list_unsorted"]),
                 ranges=n(iv["ranges"]), n_contrib=n(iv["n_contrib"]), final_T=n(iv["final_T"]),
                 out_color=n(color), out_depth=n(depth), out_lf_sub=n(lf[cases.LF_GOLDEN_CH]))
        if cs["shs"].numel():
            d["rgb"] = n(gv["rgb"] * vis[:, None])
            d["cov3D"] = n(gv["cov3D"] * vis[:, None])
        for gname, gt in zip(cases.GRAD_NAMES, grads):
            gt = n(gt)
            d[gname] = gt[:, cases.LF_GOLDEN_CH] if gname == "dL_dlang_feats" else gt
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **d)
        print(name, "R", R, "visible", int(vis.sum()), "bytes", os.path.getsize(os.path.join(out_dir, name + ".npz")))

    # ---- Adam: torch.optim.Adam on CUDA, single-tensor (unfused) implementation
    params, grads = cases.adam_case()
    tp = {k: torch.nn.Parameter(v.clone().to(dev)) for k, v in params.items()}
    opt = torch.optim.Adam([dict(params=[tp[k]], lr=cases.ADAM_LRS[k], name=k) for k in tp], lr=0.0, eps=1e-15,
                           foreach=False, fused=False)
    d = {}
    for step, gr in enumerate(grads):
        for k in tp:
            tp[k].grad = gr[k].to(dev)
        opt.step()
        for k in tp:
            d[f"p{step}_{k}"] = tp[k].detach().cpu().numpy()
    for k in tp:
        d[f"m_{k}"] = opt.state[tp[k]]["exp_avg"].cpu().numpy()
        d[f"v_{k}"] = opt.state[tp[k]]["exp_avg_sq"].cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "adam.npz"), **d)

    # ---- cosine query: the reference's torch sequence (eval/find_objects_gaussians.py:160-175)
    feats, text = cases.cosine_case()
    f = torch.nn.functional.normalize(feats.to(dev).double(), dim=1)
    t = torch.nn.functional.normalize(text.to(dev).double(), dim=1)
    sim = (f @ t.t())
    s0 = sim[:, 0]
    rel = 1 - (s0 - s0.min()) / (s0.max() - s0.min())
    sim32 = torch.nn.functional.normalize(feats.to(dev), dim=1) @ torch.nn.functional.normalize(text.to(dev), dim=1).t()
    np.savez_compressed(os.path.join(out_dir, "cosine.npz"), sim=sim.float().cpu().numpy(),
                        sim_fp32_torch=sim32.cpu().numpy(), relevance0=rel.float().cpu().numpy())
    print("golden written to", out_dir)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
