"""Fused boundary ops of the mapping iteration on liblgs.so: the
reference's activations (src/gaussian_model.cpp:46-68) forward/backward and the mapper's loss with its
gradient (src/gaussian_mapper.cpp:707-724, include/loss_utils.h), each one or a few launches instead
of a dozen ATen kernels.  `leg_slam_b200.loss` / `synthetic.activate` are the stock-torch statements of
the same maths; tests hold these to them."""
import torch

from . import _lib
from ._lib import check, ptr


def _s(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def activations_fwd(params, out=None, cat_sh=True):
    """raw parameter dict (xyz, features_dc, features_rest, lang_feat, opacity, scaling, rotation) ->
    dict(means3D, shs, lang_feats, opacities, scales, rotations); `out` reuses the four computed tensors.
    cat_sh=False skips the SH concatenation (shs is None): the split-SH rasterizer entry points read
    features_dc / features_rest in place."""
    L = _lib.lib()
    P = params["xyz"].shape[0]
    n_rest = params["features_rest"].shape[1]
    dev = params["xyz"].device
    if out is None:
        f = dict(dtype=torch.float32, device=dev)
        out = dict(scales=torch.empty(P, 3, **f), rotations=torch.empty(P, 4, **f), opacities=torch.empty(P, 1, **f),
                   shs=torch.empty(P, n_rest + 1, 3, **f) if cat_sh else None)
    with torch.cuda.device(dev):
        check(L.lgs_activations_fwd(P, n_rest, ptr(params["scaling"]), ptr(params["rotation"]), ptr(params["opacity"]),
                                    ptr(params["features_dc"]), ptr(params["features_rest"]), ptr(out["scales"]),
                                    ptr(out["rotations"]), ptr(out["opacities"]), ptr(out["shs"]) if cat_sh else None,
                                    _s(params["xyz"])),
              "lgs_activations_fwd")
    return dict(means3D=params["xyz"], shs=out["shs"] if cat_sh else None, lang_feats=params["lang_feat"], opacities=out["opacities"],
                scales=out["scales"], rotations=out["rotations"])


def activations_bwd(params, act, g_scales, g_rotations, g_opacities, g_shs, out, accumulate=False):
    """Gradients of the raw parameters into out[scaling|rotation|opacity|features_dc|features_rest];
    g_shs=None skips the SH part (already written by the split-SH backward)."""
    L = _lib.lib()
    P = params["xyz"].shape[0]
    n_rest = params["features_rest"].shape[1]
    with torch.cuda.device(params["xyz"].device):
        check(L.lgs_activations_bwd(P, n_rest, int(accumulate), ptr(params["rotation"]), ptr(act["scales"]),
                                    ptr(act["opacities"]), ptr(g_scales), ptr(g_rotations), ptr(g_opacities),
                                    None if g_shs is None else ptr(g_shs),
                                    ptr(out["scaling"]), ptr(out["rotation"]), ptr(out["opacity"]), ptr(out["features_dc"]),
                                    ptr(out["features_rest"]), _s(params["xyz"])), "lgs_activations_bwd")


class FusedMappingLoss:
    """loss and dL/d(image, lf, depth) in four launches; buffers are allocated once per image shape."""

    def __init__(self, lambda_dssim=0.2, faithful_sign=True):
        self.lam, self.sign = float(lambda_dssim), (1 if faithful_sign else -1)
        self._buf = {}

    def __call__(self, image, lf, depth, gt_image, gt_lf, gt_depth, mask=None):
        L = _lib.lib()
        H, W = image.shape[-2:]
        key = (H, W, image.device)
        b = self._buf.get(key)
        if b is None:
            f = dict(dtype=torch.float32, device=image.device)
            b = dict(gi=torch.empty(3, H, W, **f), gl=torch.empty(64, H, W, **f), gd=torch.empty(1, H, W, **f),
                     loss=torch.empty(8, **f),
                     scratch=torch.empty(L.lgs_mapping_loss_scratch_bytes(W, H), dtype=torch.uint8, device=image.device))
            self._buf[key] = b
        image, lf, depth, gt_image, gt_lf, gt_depth = (t.contiguous() for t in (image, lf, depth, gt_image, gt_lf, gt_depth))
        with torch.cuda.device(image.device):
            check(L.lgs_mapping_loss(W, H, gt_lf.shape[-1], gt_lf.shape[-2], ptr(image), ptr(lf), ptr(depth), ptr(gt_image),
                                     ptr(gt_lf), ptr(gt_depth), None if mask is None else ptr(mask.contiguous()), self.lam,
                                     self.sign, ptr(b["gi"]), ptr(b["gl"]), ptr(b["gd"]), ptr(b["loss"]), ptr(b["scratch"]),
                                     _s(image)), "lgs_mapping_loss")
        return b["loss"], b["gi"], b["gl"], b["gd"]
