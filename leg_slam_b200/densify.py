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


def reset_opacity(params, exp_avg, exp_avg_sq):
    """GaussianModel::resetOpacity (:567-575): inverse_sigmoid(min(sigmoid(opacity), 1)) -- the reference clamps against
    ones, i.e. not at all -- and zeroes the opacity's Adam moments (:577-595)."""
    op = torch.sigmoid(params["opacity"])
    x = torch.min(op, torch.ones_like(op))
    params["opacity"] = torch.log(x / (1 - x))
    exp_avg["opacity"] = torch.zeros_like(params["opacity"])
    exp_avg_sq["opacity"] = torch.zeros_like(params["opacity"])
