"""L2 of the boundary: the autograd wrapper and the `GaussianRasterizer` module, mirroring

  GaussianRasterizationSettings / GaussianRasterizerFunction / rasterizeGaussians /
  GaussianRasterizer::{forward, markVisibleGaussians}
      reference include/gaussian_rasterizer.h:25-129, src/gaussian_rasterizer.cpp:18-236
  and the Python twin eval/submodules/diff-gaussian-rasterization-legs-slam/
      diff_gaussian_rasterization_legs_slam/__init__.py:21-210

Same names, argument meaning and error behaviour.  The backward follows the C++ wrapper
(three upstream gradients: colour, language feature, depth; 24-argument backward call,
src/gaussian_rasterizer.cpp:137-175) -- the shipped Python wrapper passes 22 arguments and
is forward-only in practice.
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from . import rasterize_points as _C


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    include_language_features: bool


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, lang_feats, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings):
        rs = raster_settings
        num_rendered, color, lf, depth, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(
            rs.bg, means3D, colors_precomp, lang_feats, opacities, scales, rotations, rs.scale_modifier,
            cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width,
            sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.include_language_features)
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, lang_feats, means3D, scales, rotations, cov3Ds_precomp, radii, sh,
                              geomBuffer, binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, lf, depth, radii

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_lf, grad_out_depth, _grad_radii=None):
        rs = ctx.raster_settings
        colors_precomp, lang_feats, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, \
            binningBuffer, imgBuffer = ctx.saved_tensors
        H, W = rs.image_height, rs.image_width
        # autograd hands None for outputs the loss did not touch
        if grad_out_color is None:
            grad_out_color = torch.zeros((3, H, W), dtype=torch.float32, device=means3D.device)
        if grad_out_lf is None:
            grad_out_lf = torch.zeros((_C.LF_NUM_CHANNELS, H, W), dtype=torch.float32, device=means3D.device)
        if grad_out_depth is None:
            grad_out_depth = torch.zeros((1, H, W), dtype=torch.float32, device=means3D.device)
        (grad_means2D, grad_colors_precomp, grad_lang_feats, grad_opacities, grad_means3D, grad_cov3Ds_precomp,
         grad_sh, grad_scales, grad_rotations) = _C.rasterize_gaussians_backward(
            rs.bg, means3D, radii, colors_precomp, lang_feats, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
            rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_out_color, grad_out_lf, grad_out_depth, sh,
            rs.sh_degree, rs.campos, geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer,
            rs.include_language_features)

        def fit(g, like):  # empty sentinel inputs get no gradient
            return g if like.numel() != 0 and g.numel() == like.numel() else None
        return (grad_means3D, grad_means2D, fit(grad_sh, sh), fit(grad_colors_precomp, colors_precomp),
                fit(grad_lang_feats, lang_feats), grad_opacities, fit(grad_scales, scales),
                fit(grad_rotations, rotations), fit(grad_cov3Ds_precomp, cov3Ds_precomp), None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, lang_feats, opacities, scales, rotations,
                        cov3Ds_precomp, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, lang_feats, opacities, scales,
                                     rotations, cov3Ds_precomp, raster_settings)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return _C.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, lang_feats=None, scales=None,
                rotations=None, cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        dev = means3D.device
        empty = lambda: torch.empty(0, dtype=torch.float32, device=dev)  # noqa: E731
        shs = empty() if shs is None else shs
        colors_precomp = empty() if colors_precomp is None else colors_precomp
        lang_feats = empty() if lang_feats is None else lang_feats
        scales = empty() if scales is None else scales
        rotations = empty() if rotations is None else rotations
        cov3D_precomp = empty() if cov3D_precomp is None else cov3D_precomp
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, lang_feats, opacities, scales, rotations,
                                   cov3D_precomp, rs)
