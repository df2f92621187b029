"""GaussianRenderer::render (reference src/gaussian_renderer.cpp:24-160, include/gaussian_renderer.h) on our rasterizer:
the host logic between a Gaussian model + a keyframe and the rasterizer call -- which of shs / colors_precomp and of
scales + rotations / cov3D_precomp is handed over (the others stay the reference's empty-tensor sentinels), the zero
`screenspace_points` leaf that receives the 2D-mean gradient, `radii > 0` as the visibility filter -- and the 6-tuple it
returns:  (rendered_image, rendered_lf, rendered_depth, screenspace_points, visibility_filter, radii).

`GaussianModelView` supplies the accessors of the reference's GaussianModel that render() calls (activations of
src/gaussian_model.cpp:46-88) over a dict of raw parameter tensors; `KeyframeView` the GaussianKeyframe fields
(src/gaussian_keyframe.cpp:111-193).  Autograd flows through (the mapper's training path does not use this file: it runs the
fused, autograd-free sequence of leg_slam_b200.mapper)."""
import math
from typing import NamedTuple

import torch

from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435)


class GaussianPipelineParams(NamedTuple):
    """include/gaussian_parameters.h: both false in every shipped configuration."""
    convert_SHs_: bool = False
    compute_cov3D_: bool = False


class KeyframeView:
    """The GaussianKeyframe fields render() reads, from a leg_slam_b200.synthetic.Camera."""

    def __init__(self, camera):
        self.FoVx_ = 2.0 * math.atan(camera.tanfovx)
        self.FoVy_ = 2.0 * math.atan(camera.tanfovy)
        self.world_view_transform_ = camera.viewmatrix
        self.full_proj_transform_ = camera.projmatrix
        self.camera_center_ = camera.campos
        self.image_height_, self.image_width_ = camera.height, camera.width


def build_rotation(r):
    """general_utils::build_rotation (include/general_utils.h:29-60): normalised quaternion (r,x,y,z) -> [P,3,3]."""
    q = r / r.norm(dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
        torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], 1)


def eval_sh(deg, sh, dirs):
    """sh_utils::eval_sh (include/sh_utils.h:63-130): sh [P,3,(deg+1)^2...], dirs [P,3] unit -> [P,3]."""
    result = SH_C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        result = result - SH_C1 * y * sh[..., 1] + SH_C1 * z * sh[..., 2] - SH_C1 * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            result = (result + SH_C2[0] * xy * sh[..., 4] + SH_C2[1] * yz * sh[..., 5] + SH_C2[2] * (2.0 * zz - xx - yy) * sh[..., 6]
                      + SH_C2[3] * xz * sh[..., 7] + SH_C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                result = (result + SH_C3[0] * y * (3 * xx - yy) * sh[..., 9] + SH_C3[1] * xy * z * sh[..., 10]
                          + SH_C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12]
                          + SH_C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + SH_C3[5] * z * (xx - yy) * sh[..., 14]
                          + SH_C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return result


class GaussianModelView:
    """The accessors of the reference's GaussianModel that GaussianRenderer::render uses (src/gaussian_model.cpp:46-88), over
    raw parameter tensors: xyz, features_dc, features_rest, lang_feat, opacity, scaling, rotation."""

    def __init__(self, params, sh_degree=3, active_sh_degree=None):
        self.p = params
        self.max_sh_degree_ = sh_degree
        self.active_sh_degree_ = sh_degree if active_sh_degree is None else active_sh_degree

    def getXYZ(self):
        return self.p["xyz"]

    def getScalingActivation(self):
        return torch.exp(self.p["scaling"])

    def getRotationActivation(self):
        return torch.nn.functional.normalize(self.p["rotation"])

    def getOpacityActivation(self):
        return torch.sigmoid(self.p["opacity"])

    def getFeatures(self):
        return torch.cat([self.p["features_dc"], self.p["features_rest"]], dim=1)

    def getLanguageFeatures(self):
        return self.p["lang_feat"]

    def getCovarianceActivation(self, scaling_modifier=1):
        R = build_rotation(self.p["rotation"])
        L = R @ torch.diag_embed(scaling_modifier * self.getScalingActivation())
        c = L @ L.transpose(1, 2)
        return torch.stack([c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2]], dim=1)


class GaussianRenderer:
    @staticmethod
    def render(viewpoint_camera, image_height, image_width, pc, pipe, bg_color, override_color=None, scaling_modifier=1.0,
               use_override_color=False, include_language_features=False):
        """-> (rendered_image, rendered_lf, rendered_depth, screenspace_points, visibility_filter, radii).
        Background tensor (bg_color) must be on the GPU, like the reference's."""
        xyz = pc.getXYZ()
        # zero leaf that makes autograd return the gradient of the 2D (screen-space) means (:41-48)
        screenspace_points = torch.zeros_like(xyz, requires_grad=True)
        try:
            screenspace_points.retain_grad()
        except Exception:
            pass
        tanfovx = math.tan(viewpoint_camera.FoVx_ * 0.5)
        tanfovy = math.tan(viewpoint_camera.FoVy_ * 0.5)
        raster_settings = GaussianRasterizationSettings(
            image_height, image_width, tanfovx, tanfovy, bg_color, scaling_modifier, viewpoint_camera.world_view_transform_,
            viewpoint_camera.full_proj_transform_, pc.active_sh_degree_, viewpoint_camera.camera_center_, False,
            include_language_features)
        rasterizer = GaussianRasterizer(raster_settings)
        means3D, means2D, opacity = xyz, screenspace_points, pc.getOpacityActivation()
        # precomputed 3D covariance, or scaling / rotation for the rasterizer to build it from (:76-90)
        scales = rotations = cov3D_precomp = None
        if pipe.compute_cov3D_:
            cov3D_precomp = pc.getCovarianceActivation()
        else:
            scales, rotations = pc.getScalingActivation(), pc.getRotationActivation()
        # override colours, SH -> RGB here, or SHs for the rasterizer to convert (:95-116)
        shs = colors_precomp = None
        if use_override_color:
            colors_precomp = override_color
        elif pipe.convert_SHs_:
            n = (pc.max_sh_degree_ + 1) ** 2
            shs_view = pc.getFeatures().transpose(1, 2).reshape(-1, 3, n)
            dir_pp = xyz - viewpoint_camera.camera_center_[None, :]
            dir_pp = dir_pp / dir_pp.norm(dim=1, keepdim=True)
            colors_precomp = torch.clamp_min(eval_sh(pc.active_sh_degree_, shs_view, dir_pp) + 0.5, 0.0)
        else:
            shs = pc.getFeatures()
        lang_feat = pc.getLanguageFeatures() if include_language_features else None
        rendered_image, rendered_lf, rendered_depth, radii = rasterizer(
            means3D, means2D, opacity, shs=shs, colors_precomp=colors_precomp, lang_feats=lang_feat, scales=scales,
            rotations=rotations, cov3D_precomp=cov3D_precomp)
        # frustum-culled Gaussians and those with radius 0 were not visible (:147-159)
        return rendered_image, rendered_lf, rendered_depth, screenspace_points, radii > 0, radii
