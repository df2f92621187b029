#!/usr/bin/env python
"""Secondary configurations of BASELINE.json (not bench.py lines): cfgD render sizes and cfgE semantic query.
Prints one JSON object; used for profiles/ and DESIGN.md."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from leg_slam_b200 import cosine_image, cosine_query, heatmap_render, rasterize_points as rp, synthetic  # noqa: E402


def t(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda:0")
    out = {}
    # cfgE: 2M x 64 against 256 text embeddings
    g = torch.Generator().manual_seed(1)
    feats = torch.randn(2_000_000, 64, generator=g).to(dev)
    text = torch.randn(256, 64, generator=g).to(dev)
    ms = t(lambda: cosine_query(feats, text))
    ref_ms = t(lambda: torch.nn.functional.normalize(feats, dim=1) @ torch.nn.functional.normalize(text, dim=1).t())
    a = cosine_query(feats, text)
    b = torch.nn.functional.normalize(feats.double(), dim=1) @ torch.nn.functional.normalize(text.double(), dim=1).t()
    out["cfgE_cosine_2M_x_256"] = dict(ms=ms, torch_normalize_matmul_ms=ref_ms, out_GBps=2e6 * 256 * 4 / ms / 1e6,
                                       algorithmic_GBps=(2e6 * 256 + 4 * 256 * 2e6) / ms / 1e6,
                                       max_abs_err_vs_fp64=float((a.double() - b).abs().max()))
    del a, b, feats, text
    # cfgD: ScanNet-shaped 2M Gaussians, 1296x968
    W, H, P = 1296, 968, 2_000_000
    sc = synthetic.make_scene(P, seed=4, room=(8.0, 6.0, 3.0), device=dev)
    cam = synthetic.make_cameras(1, W, H, fx=1169.7, fy=1169.7, room=(8.0, 6.0, 3.0), seed=4)[0].to(dev)
    act = synthetic.activate(sc)
    e = torch.empty(0, device=dev)
    bg = torch.zeros(3, device=dev)
    args = (bg, act["means3D"], e, act["lang_feats"], act["opacities"], act["scales"], act["rotations"], 1.0, e, cam.viewmatrix,
            cam.projmatrix, cam.tanfovx, cam.tanfovy, H, W, act["shs"], 3, cam.campos, False, True)
    R, color, lf, depth, radii, geom, binning, img = rp.rasterize_gaussians(*args)
    gg = torch.Generator().manual_seed(2)
    dc = (torch.randn(3, H, W, generator=gg) / (H * W)).to(dev)
    dl = (torch.randn(64, H, W, generator=gg) / (H * W)).to(dev)
    dd = (torch.randn(1, H, W, generator=gg) / (H * W)).to(dev)
    bargs = (bg, act["means3D"], radii, e, act["lang_feats"], act["scales"], act["rotations"], 1.0, e, cam.viewmatrix,
             cam.projmatrix, cam.tanfovx, cam.tanfovy, dc, dl, dd, act["shs"], 3, cam.campos, geom, R, binning, img, True)
    out["cfgD_2M_1296x968"] = dict(R=R, visible=int((radii > 0).sum()), fwd_ms=t(lambda: rp.rasterize_gaussians(*args), 5, 2),
                                   bwd_ms=t(lambda: rp.rasterize_gaussians_backward(*bargs), 5, 2))
    try:
        import build_ref
        ref = build_ref.load()
        Rr, *_r, gr, br, ir = ref.rasterize_gaussians(*args)
        rb = (bg, act["means3D"], _r[3], e, act["lang_feats"], act["scales"], act["rotations"], 1.0, e, cam.viewmatrix,
              cam.projmatrix, cam.tanfovx, cam.tanfovy, dc, dl, dd, act["shs"], 3, cam.campos, gr, Rr, br, ir, True)
        out["cfgD_2M_1296x968"].update(ref_fwd_ms=t(lambda: ref.rasterize_gaussians(*args), 3, 1),
                                       ref_bwd_ms=t(lambda: ref.rasterize_gaussians_backward(*rb), 3, 1), ref_R=Rr)
    except Exception as ex:  # reference .so not shipped
        out["cfgD_2M_1296x968"]["ref"] = str(ex)
    # cfgE end to end on the same 2M-Gaussian scene: query (one text embedding of the batch) -> min-max inversion -> heat colours
    # -> forward through the colors_precomp path; and the per-pixel query on the rendered [64,H,W] feature image
    text1 = torch.randn(64, generator=torch.Generator().manual_seed(3)).to(dev)
    hm = lambda: heatmap_render(act["means3D"], act["opacities"], act["scales"], act["rotations"], act["lang_feats"], text1, cam)  # noqa: E731
    out["cfgE_query_then_heatmap_render_2M_1296x968"] = dict(ms=t(hm, 5, 2))
    t8 = torch.randn(8, 64, generator=torch.Generator().manual_seed(4)).to(dev)
    ms_px = t(lambda: cosine_image(lf, t8), 10, 3)
    ref_px = t(lambda: torch.stack([torch.nn.functional.cosine_similarity(lf, q[:, None, None], dim=0) for q in t8]), 5, 2)
    out["per_pixel_cosine_1296x968_x8_queries"] = dict(ms=ms_px, torch_cosine_similarity_ms=ref_px,
                                                       algorithmic_GBps=(256 + 32) * H * W / ms_px / 1e6)
    del sc, act, color, lf, depth, geom, binning, img
    torch.cuda.empty_cache()
    out.update(extra_rows(dev))
    print(json.dumps(out))


def extra_rows(dev):
    """fused densify/prune vs the torch restatement of the reference sequence, k-NN vs the
    compiled reference simple-knn."""
    import densify_ref as DR
    from leg_slam_b200 import densify as D, ingest
    out = {}
    for P in (500_000, 2_000_000):
        sc = synthetic.make_scene(P, seed=7, device=dev)
        g = torch.Generator().manual_seed(8)
        m = DR.Model(sc)
        for k in DR.PARAMS:
            m.m[k] = torch.randn(m.p[k].shape, generator=g).to(dev) * 0.01
            m.v[k] = torch.rand(m.p[k].shape, generator=g).to(dev) * 1e-4
        m.denom = torch.randint(0, 6, (P, 1), generator=g).float().to(dev)
        m.xyz_gradient_accum = m.denom * (torch.rand(P, 1, generator=g) * 4e-4).to(dev)
        z = {}

        def normal01(n):
            if z.get("n") != n:
                z["n"], z["z"] = n, torch.randn(n, 3, device=dev)
            return z["z"]
        st = D.DensifyStats(P, dev)
        st.xyz_gradient_accum, st.denom = m.xyz_gradient_accum, m.denom
        args = (2e-4, 0.005, 6.0, 20)
        info = D.densify_and_prune(m.p, m.m, m.v, st, *args, normal01=normal01)[4]

        def ref_once():
            c = DR.Model(m.p)
            c.m, c.v = dict(m.m), dict(m.v)
            c.xyz_gradient_accum, c.denom = m.xyz_gradient_accum.clone(), m.denom.clone()
            c.densify_and_prune(*args, normal01)
        out[f"densify_prune_P{P}"] = dict(info=info, fused_ms=t(lambda: D.densify_and_prune(m.p, m.m, m.v, st, *args, normal01=normal01), 5, 2),
                                          torch_restatement_ms=t(ref_once, 3, 1),
                                          algorithmic_GB=round(info["new_P"] * 123 * 4 * 3 * 2 / 1e9, 3))
        pts = sc["xyz"].contiguous()
        row = dict(ours_ms=t(lambda: ingest.distCUDA2(pts), 5, 2))
        try:
            import build_ref
            lib = build_ref.load_knn()
            ref = torch.zeros(P, device=dev)
            row["reference_simple_knn_ms"] = t(lambda: lib.ref_simple_knn(P, pts.data_ptr(), ref.data_ptr()), 3, 1)
            row["bit_identical"] = bool(torch.equal(ingest.distCUDA2(pts), ref))
        except Exception as ex:
            row["ref"] = str(ex)
        out[f"knn_dist2_P{P}"] = row
        del m, sc, st
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main()
