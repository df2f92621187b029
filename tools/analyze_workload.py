#!/usr/bin/env python
"""CPU analysis of the cfgB blend workload (test-infrastructure tool: it runs the C oracle's forward and then counts, per
tile, which (instance, pixel) pairs the blend kernels evaluate and which of them blend).  It answers the questions DESIGN.md
section 8 asks before the next kernel change -- how many of the instances the backward walks contribute to no pixel at all,
to no pixel of a 32-pixel half, and what share of the evaluated pairs those are -- without a GPU.

    python tools/analyze_workload.py [--P 500000] [--width 640] [--height 480]   ->  one JSON object"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=500_000)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--seed", type=int, default=2)
    args = ap.parse_args()
    import oracle as O
    from leg_slam_b200 import synthetic
    O.build()
    W, H = args.width, args.height
    sc = synthetic.make_scene(args.P, seed=args.seed)
    cam = synthetic.make_cameras(8, W, H, seed=args.seed)[0]
    a = synthetic.activate(sc)
    n = lambda t: t.numpy()  # noqa: E731
    f = O.forward(n(a["means3D"]), n(a["opacities"]), n(cam.viewmatrix), n(cam.projmatrix), n(cam.campos), W, H, cam.tanfovx,
                  cam.tanfovy, np.zeros(3, np.float32), shs=n(a["shs"]), degree=3, lang_feat=n(a["lang_feats"]),
                  scales=n(a["scales"]), rotations=n(a["rotations"]))
    tiles_x, tiles_y = (W + 7) // 8, (H + 7) // 8
    ncon = f["n_contrib"].reshape(H, W)
    m2d, co, pl, ranges = f["means2D"], f["conic_opacity"], f["point_list"], f["ranges"]
    tot = dict(R=0, R_cut=0, inst_unused=0, halves=0, halves_unused=0, halves_cull_pass=0, halves_cull_pass_unused=0,
               halves_geometric=0, halves_exact_rect=0,
               pairs_eval=0, pairs_eval_unused_inst=0, pairs_eval_unused_half=0, pairs_alpha_pass=0, pairs_blend=0)
    for ty in range(tiles_y):
        for tx in range(tiles_x):
            t = ty * tiles_x + tx
            beg, end = int(ranges[t, 0]), int(ranges[t, 1])
            cnt = end - beg
            tot["R"] += cnt
            if cnt == 0:
                continue
            ys, xs = np.meshgrid(np.arange(ty * 8, ty * 8 + 8), np.arange(tx * 8, tx * 8 + 8), indexing="ij")
            inside = (ys < H) & (xs < W)
            nc = np.where(inside, ncon[np.minimum(ys, H - 1), np.minimum(xs, W - 1)], 0).reshape(64).astype(np.int64)
            last = int(nc.max())              # tile_last: the backward cuts the list here
            if last == 0:
                continue
            ids = pl[beg:beg + last]
            dx = m2d[ids, 0][:, None] - xs.reshape(1, 64).astype(np.float32)
            dy = m2d[ids, 1][:, None] - ys.reshape(1, 64).astype(np.float32)
            c = co[ids]
            power = -0.5 * (c[:, 0:1] * dx * dx + c[:, 2:3] * dy * dy) - c[:, 1:2] * dx * dy
            alpha = np.minimum(0.99, c[:, 3:4] * np.exp(np.minimum(power, 0.0)))
            evald = np.arange(last)[:, None] < nc[None, :]      # pixel p replays instances 0 .. n_contrib[p]-1
            passed = evald & (power <= 0) & (alpha >= 1.0 / 255.0)  # all of them blend: n_contrib is the last BLENDED one
            used_inst = passed.any(axis=1)
            used_half = np.stack([passed[:, :32].any(axis=1), passed[:, 32:].any(axis=1)], axis=1)
            ev_half = np.stack([evald[:, :32].sum(axis=1), evald[:, 32:].sum(axis=1)], axis=1)
            # the blend kernels' conservative per-warp cull (csrc/common.cuh footprint_touches): bounding box of the
            # alpha >= 1/255 ellipse against the warp's 8x4 pixel rectangle
            with np.errstate(invalid="ignore", divide="ignore"):
                tau = np.log(255.0 * c[:, 3]) * 1.01 + 0.01
                det = c[:, 0] * c[:, 2] - c[:, 1] * c[:, 1]
                k = 2.0 * tau / det
                ex = np.sqrt(k * c[:, 2]) * 1.01 + 0.01
                ey = np.sqrt(k * c[:, 0]) * 1.01 + 0.01
            gx, gy = m2d[ids, 0], m2d[ids, 1]
            cull = []
            for h in range(2):
                x0, x1, y0, y1 = tx * 8, tx * 8 + 7.0, ty * 8 + 4 * h, ty * 8 + 4 * h + 3.0
                box = (gx + ex >= x0) & (gx - ex <= x1) & (gy + ey >= y0) & (gy - ey <= y1)
                ok = np.where(~(tau > 0), False, np.where(~(det > 0) | np.isnan(ex) | np.isnan(ey), True, box))
                cull.append(ok)
            cull = np.stack(cull, axis=1)
            # exact continuous test (csrc/common.cuh footprint_touches_exact, same float32 expressions): minimum of the quadratic
            # form over the warp's pixel rectangle -- 0 if the centre is inside, else the minimum over the four edges, each a
            # clamped 1-D quadratic -- against the same threshold tau plus a magnitude-proportional rounding slack
            f32 = np.float32
            qa, qb, qc = c[:, 0].astype(f32), c[:, 1].astype(f32), c[:, 2].astype(f32)
            with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
                tau32 = (np.log(f32(255.0) * c[:, 3].astype(f32)) * f32(1.01) + f32(0.01)).astype(f32)
                det32 = qa * qc - qb * qb
                nb_c, nb_a = -qb / qc, -qb / qa
            for h in range(2):
                X0, X1 = f32(tx * 8) - gx.astype(f32), f32(tx * 8 + 7.0) - gx.astype(f32)
                Y0, Y1 = f32(ty * 8 + 4 * h) - gy.astype(f32), f32(ty * 8 + 4 * h + 3.0) - gy.astype(f32)
                inside_c = (X0 <= 0) & (X1 >= 0) & (Y0 <= 0) & (Y1 >= 0)
                qmin = np.full(len(ids), 3.0e38, f32)
                with np.errstate(invalid="ignore", over="ignore"):
                    for X, Y in ((X0, Y0), (X1, Y1)):
                        dyv = np.minimum(np.maximum(nb_c * X, Y0), Y1)
                        qmin = np.minimum(qmin, qa * X * X + f32(2.0) * qb * X * dyv + qc * dyv * dyv)
                        dxv = np.minimum(np.maximum(nb_a * Y, X0), X1)
                        qmin = np.minimum(qmin, qa * dxv * dxv + f32(2.0) * qb * dxv * Y + qc * Y * Y)
                    mx, my = np.maximum(np.abs(X0), np.abs(X1)), np.maximum(np.abs(Y0), np.abs(Y1))
                    slack = f32(2.0e-6) * (qa * mx * mx + f32(2.0) * np.abs(qb) * mx * my + qc * my * my)
                    ok = inside_c | np.isnan(qmin) | (f32(0.5) * qmin <= tau32 + slack)
                ok = np.where(~(tau32 > 0), False, np.where(~(det32 > 0) | ~(qa > 0) | ~(qc > 0), True, ok))
                assert not (used_half[:, h] & ~ok).any()  # never rejects a half that blends
                tot["halves_exact_rect"] += int(ok.sum())
            # halves in which some pixel CENTRE passes the alpha tests, termination ignored: what an exact geometric cull
            # (usable by the forward, which cannot know the termination in advance) would let through
            geo = (power <= 0) & (alpha >= 1.0 / 255.0)
            tot["halves_geometric"] += int(geo[:, :32].any(axis=1).sum() + geo[:, 32:].any(axis=1).sum())
            tot["halves_cull_pass"] += int(cull.sum())
            tot["halves_cull_pass_unused"] += int((cull & ~used_half).sum())
            assert not (used_half & ~cull).any()  # the cull never rejects a half that blends
            tot["R_cut"] += last
            tot["inst_unused"] += int((~used_inst).sum())
            tot["halves"] += 2 * last
            tot["halves_unused"] += int((~used_half).sum())
            tot["pairs_eval"] += int(evald.sum())
            tot["pairs_eval_unused_inst"] += int(evald[~used_inst].sum())
            tot["pairs_eval_unused_half"] += int(ev_half[~used_half].sum())
            tot["pairs_alpha_pass"] += int(passed.sum())
    tot["pairs_blend"] = int(f["n_blended"])
    out = dict(config=dict(P=args.P, W=W, H=H, seed=args.seed), counts=tot,
               share=dict(list_cut_by_tile_last=round(1 - tot["R_cut"] / max(tot["R"], 1), 4),
                          instances_no_pixel=round(tot["inst_unused"] / max(tot["R_cut"], 1), 4),
                          halves_no_pixel=round(tot["halves_unused"] / max(tot["halves"], 1), 4),
                          halves_passing_the_kernels_cull=round(tot["halves_cull_pass"] / max(tot["halves"], 1), 4),
                          halves_with_a_pixel_centre_inside_the_alpha_ellipse=round(tot["halves_geometric"] / max(tot["halves"], 1), 4),
                          halves_passing_an_exact_ellipse_vs_rectangle_test=round(tot["halves_exact_rect"] / max(tot["halves"], 1), 4),
                          halves_passing_the_cull_but_blending_nowhere=round(tot["halves_cull_pass_unused"] / max(tot["halves_cull_pass"], 1), 4),
                          evaluated_pairs_in_unused_instances=round(tot["pairs_eval_unused_inst"] / max(tot["pairs_eval"], 1), 4),
                          evaluated_pairs_in_unused_halves=round(tot["pairs_eval_unused_half"] / max(tot["pairs_eval"], 1), 4),
                          evaluated_pairs_that_blend=round(tot["pairs_alpha_pass"] / max(tot["pairs_eval"], 1), 4)))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
