"""Parity-test cases: seeded synthetic inputs shared by the golden generator, the CPU oracle
tests and the GPU parity tests.  Inputs are regenerated from the seed everywhere, so only the
reference's OUTPUTS are stored under tests/golden/."""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from leg_slam_b200 import synthetic  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
LF_GOLDEN_CH = [0, 9, 18, 27, 36, 45, 54, 63]  # language-feature channels kept in the fixtures

# name -> settings.  Small enough for the CPU oracle to finish in well under a second.
CASES = {
    # SH degree 3, scale/rotation, language features: the mapping configuration
    "sh3_lf": dict(P=4000, W=96, H=64, seed=11, degree=3, mean_scale=0.06, lf=True, mode="sh"),
    # ragged image (not a multiple of the 8x8 tile), SH degree 1 of 16 stored coefficients
    "ragged_sh1": dict(P=3000, W=93, H=61, seed=12, degree=1, mean_scale=0.07, lf=True, mode="sh"),
    # viewer / heat-map path: precomputed colours + precomputed 3D covariance, no language features
    "precomp_nolf": dict(P=2500, W=72, H=56, seed=13, degree=0, mean_scale=0.06, lf=False, mode="precomp"),
    # dense overdraw: many large opaque Gaussians -> early termination (T < 1e-4) and long tile lists
    "dense_opaque": dict(P=3000, W=48, H=32, seed=14, degree=2, mean_scale=0.12, lf=True, mode="sh", opacity_shift=3.0),
}


# BASELINE.json configurations at their stated sizes (cfgA / cfgB / cfgD; "ragged" = cfgB with an image
# that is not a multiple of the 8x8 tile).  GPU-only: compared with the compiled reference side by side; cfgA also with the
# CPU oracle.  scene seed 2 / 640x480 is the bench workload.
BASELINE_CASES = {
    "cfgA": dict(P=10_000, W=320, H=240, seed=1, room=(6.0, 4.0, 2.8), fx=None),
    "cfgB": dict(P=500_000, W=640, H=480, seed=2, room=(6.0, 4.0, 2.8), fx=None),
    "cfgB_ragged": dict(P=500_000, W=637, H=475, seed=2, room=(6.0, 4.0, 2.8), fx=320.0),
    "cfgD": dict(P=2_000_000, W=1296, H=968, seed=4, room=(8.0, 6.0, 3.0), fx=1169.7),
}


def make_baseline_case(name, device="cpu"):
    """Like make_case, for the BASELINE.json configurations: SH degree 3, scale/rotation, 64-D language features."""
    c = BASELINE_CASES[name]
    sc = synthetic.make_scene(c["P"], seed=c["seed"], room=c["room"])
    a = synthetic.activate(sc)
    cam = synthetic.make_cameras(1, c["W"], c["H"], fx=c["fx"], room=c["room"], seed=c["seed"])[0]
    g = torch.Generator().manual_seed(c["seed"] + 100)
    HW = c["H"] * c["W"]
    empty = torch.empty(0)
    out = dict(name=name, P=c["P"], W=c["W"], H=c["H"], degree=3, include_lf=True, bg=torch.zeros(3), means3D=a["means3D"],
               opacities=a["opacities"], lang_feats=a["lang_feats"], viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix,
               campos=cam.campos, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, scale_modifier=1.0,
               dL_dcolor=torch.randn(3, c["H"], c["W"], generator=g) / HW, dL_dlf=torch.randn(64, c["H"], c["W"], generator=g) / HW,
               dL_ddepth=torch.randn(1, c["H"], c["W"], generator=g) / HW, shs=a["shs"], colors_precomp=empty, scales=a["scales"],
               rotations=a["rotations"], cov3D_precomp=empty)
    return {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in out.items()}


def covariance_from_scale_rot(scales, rots):
    """[P,6] upper triangle of (S R)^T (S R) with the reference's glm conventions (forward.cu:118-152),
    in float64 then rounded: used only to feed the cov3D_precomp path."""
    s = scales.double()
    q = rots.double()
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], -1),
        torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], -1),
        torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1)], 1)  # R[c][r]
    Mm = R * s[:, None, :]  # M[c][r] = s_r R[c][r]
    Sig = torch.einsum("pak,pbk->pab", Mm, Mm)  # Sigma[a][b] = sum_k M[a][k] M[b][k]
    return torch.stack([Sig[:, 0, 0], Sig[:, 0, 1], Sig[:, 0, 2], Sig[:, 1, 1], Sig[:, 1, 2], Sig[:, 2, 2]], -1).float()


def make_case(name, device="cpu"):
    """-> dict with every tensor argument of the rasterizer for this case (torch, on `device`)."""
    c = CASES[name]
    sc = synthetic.make_scene(c["P"], seed=c["seed"], mean_scale=c["mean_scale"])
    if "opacity_shift" in c:
        sc["opacity"] = sc["opacity"] + c["opacity_shift"]
    a = synthetic.activate(sc)
    cam = synthetic.make_cameras(1, c["W"], c["H"], seed=c["seed"])[0]
    g = torch.Generator().manual_seed(c["seed"] + 100)
    HW = c["H"] * c["W"]
    out = dict(name=name, P=c["P"], W=c["W"], H=c["H"], degree=c["degree"], include_lf=c["lf"],
               bg=torch.tensor([0.1, 0.25, 0.4]), means3D=a["means3D"], opacities=a["opacities"],
               lang_feats=a["lang_feats"], viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix,
               campos=cam.campos, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, scale_modifier=1.0,
               dL_dcolor=torch.randn(3, c["H"], c["W"], generator=g) / HW,
               dL_dlf=torch.randn(64, c["H"], c["W"], generator=g) / HW,
               dL_ddepth=torch.randn(1, c["H"], c["W"], generator=g) / HW)
    empty = torch.empty(0)
    if c["mode"] == "sh":
        out.update(shs=a["shs"], colors_precomp=empty, scales=a["scales"], rotations=a["rotations"],
                   cov3D_precomp=empty)
    else:
        out.update(shs=empty, colors_precomp=torch.rand(c["P"], 3, generator=g), scales=empty, rotations=empty,
                   cov3D_precomp=covariance_from_scale_rot(a["scales"], a["rotations"]))
    return {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in out.items()}


def fwd_args(cs):
    """Argument tuple of RasterizeGaussiansCUDA (include/rasterize_points.h:19-40)."""
    return (cs["bg"], cs["means3D"], cs["colors_precomp"], cs["lang_feats"], cs["opacities"], cs["scales"],
            cs["rotations"], cs["scale_modifier"], cs["cov3D_precomp"], cs["viewmatrix"], cs["projmatrix"],
            cs["tanfovx"], cs["tanfovy"], cs["H"], cs["W"], cs["shs"], cs["degree"], cs["campos"], False,
            cs["include_lf"])


def bwd_args(cs, radii, geom, R, binning, img):
    """Argument tuple of RasterizeGaussiansBackwardCUDA (include/rasterize_points.h:42-67)."""
    return (cs["bg"], cs["means3D"], radii, cs["colors_precomp"], cs["lang_feats"], cs["scales"], cs["rotations"],
            cs["scale_modifier"], cs["cov3D_precomp"], cs["viewmatrix"], cs["projmatrix"], cs["tanfovx"],
            cs["tanfovy"], cs["dL_dcolor"], cs["dL_dlf"], cs["dL_ddepth"], cs["shs"], cs["degree"], cs["campos"],
            geom, R, binning, img, cs["include_lf"])


GRAD_NAMES = ["dL_dmeans2D", "dL_dcolors", "dL_dlang_feats", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh",
              "dL_dscales", "dL_drotations"]


def oracle_forward(cs, O):
    """Run the CPU oracle on a case (O = the oracle module)."""
    n = lambda t: None if t.numel() == 0 else t.detach().cpu().numpy()  # noqa: E731
    return O.forward(n(cs["means3D"]), n(cs["opacities"]), n(cs["viewmatrix"]), n(cs["projmatrix"]), n(cs["campos"]),
                     cs["W"], cs["H"], cs["tanfovx"], cs["tanfovy"], n(cs["bg"]), shs=n(cs["shs"]),
                     degree=cs["degree"], colors_precomp=n(cs["colors_precomp"]), lang_feat=n(cs["lang_feats"]),
                     scales=n(cs["scales"]), rotations=n(cs["rotations"]), scale_modifier=cs["scale_modifier"],
                     cov3D_precomp=n(cs["cov3D_precomp"]), include_lf=cs["include_lf"])


def oracle_backward(cs, fwd, O):
    n = lambda t: None if t.numel() == 0 else t.detach().cpu().numpy()  # noqa: E731
    return O.backward(fwd, n(cs["means3D"]), n(cs["viewmatrix"]), n(cs["projmatrix"]), n(cs["campos"]),
                      cs["tanfovx"], cs["tanfovy"], n(cs["bg"]), n(cs["dL_dcolor"]), n(cs["dL_dlf"]),
                      n(cs["dL_ddepth"]), shs=n(cs["shs"]), degree=cs["degree"], lang_feat=n(cs["lang_feats"]),
                      scales=n(cs["scales"]), rotations=n(cs["rotations"]), scale_modifier=cs["scale_modifier"],
                      cov3D_precomp=n(cs["cov3D_precomp"]), include_lf=cs["include_lf"])


def rel_err(a, b):
    """max|a-b| / max|b| : the per-tensor relative error of BASELINE.md section 6."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if b.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


# ---- Adam / cosine cases -------------------------------------------------------------------
ADAM_SHAPES = {"xyz": (777, 3), "f_dc": (777, 1, 3), "f_rest": (777, 15, 3), "lang_feat": (777, 64),
               "opacity": (777, 1), "scaling": (777, 3), "rotation": (777, 4)}
# learning rates of the Replica config (cfg/gaussian_mapper/RGB-D/Replica/replica_rgbd.yaml:56-63)
ADAM_LRS = {"xyz": 3.2e-4, "f_dc": 2.5e-3, "f_rest": 2.5e-3 / 20.0, "lang_feat": 1.5e-3, "opacity": 0.05,
            "scaling": 5e-3, "rotation": 1e-3}
ADAM_STEPS = 3


def adam_case(seed=21):
    g = torch.Generator().manual_seed(seed)
    params = {k: torch.randn(*s, generator=g) for k, s in ADAM_SHAPES.items()}
    grads = [{k: torch.randn(*s, generator=g) * (10.0 ** float(torch.randint(-6, 1, (1,), generator=g)))
              for k, s in ADAM_SHAPES.items()} for _ in range(ADAM_STEPS)]
    for gr in grads:  # invisible Gaussians get exactly-zero gradients every iteration
        for k in gr:
            gr[k][::5] = 0.0
    return params, grads


def cosine_case(seed=31, P=4096 + 37, Q=5):
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(P, 64, generator=g) * (0.2 + torch.rand(P, 1, generator=g))
    feats[3] = 0.0  # a zero row exercises the eps clamp of F.normalize
    text = torch.randn(Q, 64, generator=g)
    return feats, text
