#!/usr/bin/env python
"""Generate tests/golden/geometry.npz from the UNMODIFIED reference geometry operators, on a GPU box.

    gpurun -- 'python tests/golden/make_geometry_golden.py gpurun_out/golden'    # then copy geometry.npz here

Runs src/stereo_vision.cu + src/operate_points.cu as compiled by oracle/build_ref.py (oracle/_ref/ref_geometry.so; nothing of
ours on the path) on small seeded inputs and stores inputs and outputs.  The file pins the numpy restatement
(oracle/ingest_ref.py, tests/test_ingest.py::test_restatement_matches_reference_golden_cpu) where no GPU is present.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import build_ref  # noqa: E402
import test_ingest as TI  # noqa: E402


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    ref = build_ref.load_geometry()
    dev = torch.device("cuda:0")
    n = lambda t: t.detach().cpu().numpy()  # noqa: E731
    out = {}
    g = torch.Generator().manual_seed(2024)
    # depth image -> camera points -> world points
    W, H = 40, 30
    depth = torch.rand(W * H, generator=g) * 5 + 0.1
    mask = torch.rand(W * H, generator=g) > 0.3
    intr = [36.5, 37.25, 19.5, 14.5]
    pts = ref.reproject_depth_pinhole(depth.to(dev), mask.to(dev), intr, W)
    T = TI._pose(g)
    world = ref.transform_points(pts.clone(), T.to(dev))
    out.update(rp_depth=n(depth), rp_mask=n(mask), rp_intr=np.array(intr, np.float32), rp_width=np.int32(W), rp_points=n(pts),
               tp_T=n(T), tp_out=n(world))
    # loop-closure correction
    P = 3000
    lp, lr, lnt, lun = TI._loop_closure_case(P, 4242)
    Tn = (torch.eye(4) + 0.01 * torch.randn(4, 4, generator=g)).contiguous()
    Tn[:, 3] = torch.tensor([0.0, 0.0, 0.0, 1.0])
    view = TI._pose(g, t=(0.1, 0.2, 0.5))
    a = [t.clone().to(dev) for t in (lp, lr, lnt, lun)]
    num = ref.scale_and_transform_then_mark_visible(a[0], a[1], a[2], a[3], Tn.to(dev), view.to(dev), torch.eye(4, device=dev), 0, 1.04)
    out.update(lc_points=n(lp), lc_rots=n(lr), lc_not_transformed=n(lnt), lc_unstable=n(lun), lc_T=n(Tn), lc_view=n(view),
               lc_scale=np.float32(1.04), lc_out_points=n(a[0]), lc_out_rots=n(a[1]), lc_out_not_transformed=n(a[2]),
               lc_num=np.int32(num))
    # inactive-geometry densification: fractional pixels, and integer pixels (distance ties)
    for tag, ints, maxd in (("ig", False, 150.0), ("igi", True, 40.0)):
        px, has, p3, colors = TI._keypoint_case(400, 96, 64, 77 if ints else 78, integer_pixels=ints)
        kin = [80.0, 81.0, 47.5, 31.5]
        rp_, rc_ = ref.inactive_geo_densify(px.to(dev), has.to(dev), p3.to(dev), colors.to(dev), maxd, kin, 96)
        out.update({f"{tag}_pixels": n(px), f"{tag}_has3D": n(has), f"{tag}_points": n(p3), f"{tag}_colors": n(colors),
                    f"{tag}_maxd": np.float32(maxd), f"{tag}_intr": np.array(kin, np.float32), f"{tag}_width": np.int32(96),
                    f"{tag}_out_points": n(rp_), f"{tag}_out_colors": n(rc_)})
    torch.cuda.synchronize()
    np.savez_compressed(os.path.join(out_dir, "geometry.npz"), **out)
    print("wrote", os.path.join(out_dir, "geometry.npz"), {k: getattr(v, "shape", ()) for k, v in out.items() if k.endswith("out_points")},
          "lc_num", int(num))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
