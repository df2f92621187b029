"""Adaptive density control (densify / split / prune) of the Gaussian set together with its Adam state, fused
(SURVEY.md section 8f row 1).  Host-side mirror of the reference's GaussianModel methods
(src/gaussian_model.cpp): addDensificationStats :834-847, densifyAndPrune :806-824 (clone :775-804, split
:729-773, prune :597-651, optimizer surgery :653-727), resetOpacity :567-595; cadence as in
src/gaussian_mapper.cpp:737-761.

The reference runs ~150 libtorch kernels per call (every index / cat / repeat is a pass over all 21 parameter and
moment tensors); here one classification pass, four scans and ONE gather produce the final tensors
(lgs_densify_plan / lgs_densify_apply, csrc/densify.cu).  No CPU path: CPU tensors raise LgsError.
"""
import ctypes

import torch

from . import _lib
from ._lib import check

PARAM_ORDER = ("xyz", "features_dc", "features_rest", "lang_feat", "opacity", "scaling", "rotation")
_COPY, _ZERO_NEW, _XYZ, _SCALING = 0, 1, 2, 3


def _s(t):
    return torch.cuda.current_stream(t.device).cuda_stream


class DensifyStats:
    """xyz_gradient_accum_, denom_, max_radii2D_, exist_since_iter_ of the reference's GaussianModel."""

    def __init__(self, P, device):
        self.xyz_gradient_accum = torch.zeros(P, 1, device=device)
        self.denom = torch.zeros(P, 1, device=device)
        self.max_radii2D = torch.zeros(P, device=device)
        self.exist_since_iter = torch.zeros(P, dtype=torch.int32, device=device)

    def add(self, radii, dL_dmeans2D):
        """One launch for `max_radii2D_[vis] = max(...)` + addDensificationStats(viewspace_points, vis)."""
        if not radii.is_cuda:
            raise _lib.LgsError("leg_slam_b200 has no CPU path: tensors must live on a CUDA device")
        P = radii.shape[0]
        with torch.cuda.device(radii.device):
            check(_lib.lib().lgs_densify_stats(P, radii.contiguous().data_ptr(), dL_dmeans2D.contiguous().data_ptr(),
                                               self.xyz_gradient_accum.data_ptr(), self.denom.data_ptr(),
                                               self.max_radii2D.data_ptr(), _s(radii)), "lgs_densify_stats")


def densify_and_prune(params, exp_avg, exp_avg_sq, stats, max_grad, min_opacity, extent, max_screen_size,
                      percent_dense=0.01, generator=None, normal01=None):
    """GaussianModel::densifyAndPrune on dicts of the 7 parameter tensors and their Adam moments.

    Returns (new_params, new_exp_avg, new_exp_avg_sq, new_stats, info); inputs are left untouched.  `normal01(n)`
    supplies the n x 3 standard-normal draws of densifyAndSplit (default: torch.randn with `generator`)."""
    L = _lib.lib()
    xyz = params["xyz"]
    if not xyz.is_cuda:
        raise _lib.LgsError("leg_slam_b200 has no CPU path: tensors must live on a CUDA device")
    P, dev = xyz.shape[0], xyz.device
    for k in PARAM_ORDER:
        for d in (params, exp_avg, exp_avg_sq):
            if d[k].dtype != torch.float32 or not d[k].is_contiguous() or d[k].shape[0] != P:
                raise TypeError(f"{k}: contiguous float32 [P, ...] tensors expected")
    totals = (ctypes.c_int * 4)()
    with torch.cuda.device(dev):
        plan = torch.empty(L.lgs_densify_plan_bytes(P), dtype=torch.uint8, device=dev)
        check(L.lgs_densify_plan(P, stats.xyz_gradient_accum.data_ptr(), stats.denom.data_ptr(), params["scaling"].data_ptr(),
                                 params["opacity"].data_ptr(), float(max_grad), float(min_opacity), float(extent),
                                 float(percent_dense), int(max_screen_size), plan.data_ptr(), totals, _s(xyz)),
              "lgs_densify_plan")
        nA, nB, nC, nS = (int(t) for t in totals)
        newP = nA + nB + 2 * nC
        if normal01 is None:
            normal01 = lambda n: torch.randn(n, 3, device=dev, generator=generator)  # noqa: E731
        samples = normal01(2 * nS).contiguous() if nS > 0 else torch.empty(0, 3, device=dev)
        src, dst, rows, modes = [], [], [], []
        out_p, out_m, out_v = {}, {}, {}
        for k in PARAM_ORDER:
            shape = (newP,) + tuple(params[k].shape[1:])
            out_p[k], out_m[k], out_v[k] = (torch.empty(shape, device=dev) for _ in range(3))
            row = params[k][0].numel() if P else 1
            mode = _XYZ if k == "xyz" else (_SCALING if k == "scaling" else _COPY)
            for s_, d_, m_ in ((params[k], out_p[k], mode), (exp_avg[k], out_m[k], _ZERO_NEW), (exp_avg_sq[k], out_v[k], _ZERO_NEW)):
                src.append(s_.data_ptr()), dst.append(d_.data_ptr()), rows.append(row), modes.append(m_)
        new_stats = DensifyStats(newP, dev)  # densificationPostfix zeroes accum / denom / max_radii2D (:723-725)
        src.append(stats.exist_since_iter.data_ptr()), dst.append(new_stats.exist_since_iter.data_ptr())
        rows.append(1), modes.append(_COPY)
        n = len(src)
        if newP > 0:
            scratch = torch.empty(2 * newP, dtype=torch.int32, device=dev)
            check(L.lgs_densify_apply(P, plan.data_ptr(), totals, n, (ctypes.c_void_p * n)(*src), (ctypes.c_void_p * n)(*dst),
                                      (ctypes.c_int * n)(*rows), (ctypes.c_int * n)(*modes), params["scaling"].data_ptr(),
                                      params["rotation"].data_ptr(), samples.data_ptr() if nS > 0 else None,
                                      scratch.data_ptr(), _s(xyz)), "lgs_densify_apply")
    info = dict(kept=nA, cloned=nB, split_kept=nC, split_selected=nS, new_P=newP)
    return out_p, out_m, out_v, new_stats, info


SH_C0 = 0.282094806432724  # include/sh_utils.h:32 `const float C0 = 0.28209479177387814f`: the float32 value (a CUDA tensor /
# scalar multiplies by the scalar's reciprocal, so the double literal would give a last-bit-different DC coefficient)


def increase_pcd(params, exp_avg, exp_avg_sq, stats, new_points, new_colors, iteration, sh_degree=3):
    """GaussianModel::increasePcd (tensor overload, src/gaussian_model.cpp:297-384): Gaussians for the new points of a
    keyframe -- colour as the DC coefficient (RGB2SH), zero higher-order SH and language features, isotropic scale from
    the mean squared distance to the 3 nearest new points (`ingest.distCUDA2`, clamped at 1e-7), identity rotation,
    opacity inverse_sigmoid(0.1) -- appended to the 7 parameter tensors with zero Adam moments (densificationPostfix,
    :653-727), the statistics reset at the new size.  Returns (new_params, new_exp_avg, new_exp_avg_sq, new_stats);
    inputs are left untouched.  One allocation + two copies per tensor (no `cat` temporaries); the scale needs the k-NN
    kernel, so there is no CPU path."""
    from . import ingest
    xyz = params["xyz"]
    if not xyz.is_cuda or not new_points.is_cuda:
        raise _lib.LgsError("leg_slam_b200 has no CPU path: tensors must live on a CUDA device")
    if new_points.dim() != 2 or new_points.shape[1] != 3 or new_colors.shape != new_points.shape:
        raise ValueError("new_points and new_colors must have dimensions (num_points, 3)")
    n, P, dev = int(new_points.shape[0]), int(xyz.shape[0]), xyz.device
    if n == 0:  # the reference returns before touching anything (:299-300)
        return params, exp_avg, exp_avg_sq, stats
    new_points = new_points.to(dev, torch.float32).contiguous()
    n_coef = (sh_degree + 1) ** 2
    if params["features_rest"].shape[1] != n_coef - 1:
        raise ValueError("features_rest does not match sh_degree")
    d2 = torch.clamp_min(ingest.distCUDA2(new_points.clone()), 0.0000001)
    x = torch.full((n, 1), 0.1, dtype=torch.float32, device=dev)
    new = dict(xyz=new_points,
               features_dc=((new_colors.to(dev, torch.float32) - 0.5) / SH_C0).unsqueeze(1),        # [n,1,3]
               features_rest=None, lang_feat=None,                                                   # zeros
               opacity=torch.log(x / (1 - x)),
               scaling=torch.log(torch.sqrt(d2)).unsqueeze(1).expand(n, 3),
               rotation=None)                                                                        # (1,0,0,0)
    out_p, out_m, out_v = {}, {}, {}
    for k in PARAM_ORDER:
        shape = (P + n,) + tuple(params[k].shape[1:])
        out_p[k] = torch.empty(shape, dtype=torch.float32, device=dev)
        out_p[k][:P].copy_(params[k])
        if new[k] is None:
            out_p[k][P:].zero_()
        else:
            out_p[k][P:].copy_(new[k])
        for src, dst in ((exp_avg, out_m), (exp_avg_sq, out_v)):
            dst[k] = torch.empty(shape, dtype=torch.float32, device=dev)
            dst[k][:P].copy_(src[k])
            dst[k][P:].zero_()
    out_p["rotation"][P:, 0] = 1.0
    new_stats = DensifyStats(P + n, dev)  # accum / denom / max_radii2D zeroed at the new size (:723-725)
    new_stats.exist_since_iter[:P].copy_(stats.exist_since_iter)
    new_stats.exist_since_iter[P:] = int(iteration)
    return out_p, out_m, out_v, new_stats


def create_from_pcd(points, colors, lang_feats=None, sh_degree=3):
    """GaussianModel::createFromPcd (reference src/gaussian_model.cpp:109-194) on tensors: the initial Gaussian set of a point
    cloud -- colour as the DC coefficient (RGB2SH), zero higher-order SH, the points' language features (zeros when the cloud
    carries none), isotropic log-scale from the mean squared distance to the 3 nearest points (`ingest.distCUDA2`, clamped
    at 1e-7), identity rotation, opacity inverse_sigmoid(0.1).  Returns (params, stats): the 7 parameter tensors in the
    mapper's layout and fresh statistics with exist_since_iter = 0.  The scale needs the k-NN kernel: no CPU path."""
    from . import ingest
    if not points.is_cuda:
        raise _lib.LgsError("leg_slam_b200 has no CPU path: tensors must live on a CUDA device")
    if points.dim() != 2 or points.shape[1] != 3 or colors.shape != points.shape:
        raise ValueError("points and colors must have dimensions (num_points, 3)")
    n, dev = int(points.shape[0]), points.device
    if lang_feats is not None and tuple(lang_feats.shape) != (n, 64):
        raise ValueError("lang_feats must have dimensions (num_points, 64)")
    f = dict(dtype=torch.float32, device=dev)
    xyz = points.to(**f).contiguous().clone()
    d2 = torch.clamp_min(ingest.distCUDA2(xyz.clone()), 0.0000001)
    x = torch.full((n, 1), 0.1, **f)
    rot = torch.zeros(n, 4, **f)
    rot[:, 0] = 1.0
    params = dict(xyz=xyz, features_dc=((colors.to(**f) - 0.5) / SH_C0).unsqueeze(1).contiguous(),
                  features_rest=torch.zeros(n, (sh_degree + 1) ** 2 - 1, 3, **f),
                  lang_feat=torch.zeros(n, 64, **f) if lang_feats is None else lang_feats.to(**f).contiguous().clone(),
                  opacity=torch.log(x / (1 - x)), scaling=torch.log(torch.sqrt(d2)).unsqueeze(1).repeat(1, 3), rotation=rot)
    return params, DensifyStats(n, dev)


def reset_opacity(params, exp_avg, exp_avg_sq):
    """GaussianModel::resetOpacity (:567-575): inverse_sigmoid(min(sigmoid(opacity), 1)) -- the reference clamps against
    ones, i.e. not at all -- and zeroes the opacity's Adam moments (:577-595)."""
    op = torch.sigmoid(params["opacity"])
    x = torch.min(op, torch.ones_like(op))
    params["opacity"] = torch.log(x / (1 - x))
    exp_avg["opacity"] = torch.zeros_like(params["opacity"])
    exp_avg_sq["opacity"] = torch.zeros_like(params["opacity"])
