"""Deterministic synthetic Replica-/ScanNet-shaped scenes and cameras.

Datasets are not available offline, so every test and benchmark renders these.  Everything
is generated on the CPU from a seeded torch.Generator (bit-identical everywhere) and then
moved to the requested device.

Camera conventions follow the reference (src/gaussian_keyframe.cpp:111-193):
  viewmatrix  = world_view_transform_ = W2C^T                (row-major tensor of the transpose,
                                                               i.e. column-major W2C for the kernels)
  projmatrix  = full_proj_transform_  = viewmatrix @ P^T     with P from getProjectionMatrix
  campos      = inverse(viewmatrix)[3, :3]
  tanfovx/y   = tan(FoV/2); the projection ignores cx, cy (symmetric frustum)
"""
import math
from typing import NamedTuple

import torch

SH_C0 = 0.28209479177387814


class Camera(NamedTuple):
    width: int
    height: int
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor   # [4,4]
    projmatrix: torch.Tensor   # [4,4]
    campos: torch.Tensor       # [3]

    def to(self, device):
        return Camera(self.width, self.height, self.tanfovx, self.tanfovy, self.viewmatrix.to(device),
                      self.projmatrix.to(device), self.campos.to(device))


def projection_matrix(znear, zfar, fovx, fovy):
    """getProjectionMatrix (src/gaussian_keyframe.cpp:166-193)."""
    t = math.tan(fovy / 2) * znear
    r = math.tan(fovx / 2) * znear
    P = torch.zeros(4, 4, dtype=torch.float32)
    P[0, 0] = 2.0 * znear / (2 * r)
    P[1, 1] = 2.0 * znear / (2 * t)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def camera_from_pose(R_c2w, center, width, height, fx, fy, znear=0.01, zfar=100.0):
    """Pinhole camera looking down +z (OpenCV axes) at `center` with rotation R_c2w [3,3]."""
    R_c2w = R_c2w.to(torch.float64)
    center = center.to(torch.float64)
    w2c = torch.eye(4, dtype=torch.float64)
    w2c[:3, :3] = R_c2w.t()
    w2c[:3, 3] = -R_c2w.t() @ center
    fovx = 2 * math.atan(width / (2 * fx))
    fovy = 2 * math.atan(height / (2 * fy))
    view = w2c.t().to(torch.float32).contiguous()
    proj = projection_matrix(znear, zfar, fovx, fovy).t()
    full = (view.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    campos = view.inverse()[3, :3].contiguous()
    return Camera(width, height, math.tan(fovx / 2), math.tan(fovy / 2), view, full, campos)


def look_at(center, target, up=(0.0, 0.0, 1.0)):
    """Rotation camera->world with z = forward, x = right, y = down."""
    c = torch.as_tensor(center, dtype=torch.float64)
    f = torch.as_tensor(target, dtype=torch.float64) - c
    f = f / f.norm()
    upv = torch.as_tensor(up, dtype=torch.float64)
    x = torch.linalg.cross(f, upv)
    x = x / x.norm()
    y = torch.linalg.cross(f, x)
    return torch.stack([x, y, f], dim=1)


def make_cameras(n, width, height, fx=None, fy=None, room=(6.0, 4.0, 2.8), seed=0):
    """n poses on a seeded orbit inside the room, each looking across the room at the far walls."""
    fx = width / 2.0 if fx is None else fx   # FoVx 90 degrees (Replica: fx=600 @ 1200 px)
    fy = fx if fy is None else fy
    g = torch.Generator().manual_seed(1000 + seed)
    cams = []
    cx, cy = room[0] / 2, room[1] / 2
    for i in range(n):
        ang = 2 * math.pi * (i / max(n, 1)) + float(torch.rand(1, generator=g)) * 0.3
        rad = 0.6 + 0.5 * float(torch.rand(1, generator=g))
        h = 1.2 + 0.4 * float(torch.rand(1, generator=g))
        c = (cx + rad * math.cos(ang), cy + rad * math.sin(ang), h)
        # look outward across the centre towards the opposite wall, slightly downward
        tgt = (cx - 2.0 * math.cos(ang), cy - 2.0 * math.sin(ang), h - 0.3 + 0.4 * float(torch.rand(1, generator=g)))
        cams.append(camera_from_pose(look_at(c, tgt), torch.tensor(c), width, height, fx, fy))
    return cams


def make_scene(P, seed=0, room=(6.0, 4.0, 2.8), mean_scale=None, sh_degree=3, device="cpu"):
    """Raw (pre-activation) Gaussian parameters of a room-shaped scene, as the reference's
    GaussianModel holds them (src/gaussian_model.cpp:46-68): xyz [P,3], features_dc [P,1,3],
    features_rest [P,15,3], lang_feat [P,64], opacity [P,1] (logit), scaling [P,3] (log),
    rotation [P,4] (un-normalised quaternion r,x,y,z)."""
    g = torch.Generator().manual_seed(seed)
    lx, ly, lz = room
    if mean_scale is None:
        mean_scale = 0.015 * math.sqrt(500_000 / max(P, 1))  # ~4-6 px projected radius at any P
        mean_scale = min(mean_scale, 0.08)
    n_clutter = P // 20
    n_surf = P - n_clutter
    # surface points: pick a wall/floor/ceiling with probability proportional to its area
    areas = torch.tensor([lx * ly, lx * ly, lx * lz, lx * lz, ly * lz, ly * lz])
    face = torch.multinomial(areas / areas.sum(), n_surf, replacement=True, generator=g)
    u = torch.rand(n_surf, generator=g)
    v = torch.rand(n_surf, generator=g)
    xyz = torch.empty(n_surf, 3)
    for f, (ax_u, ax_v, ax_w, val) in enumerate([(0, 1, 2, 0.0), (0, 1, 2, lz), (0, 2, 1, 0.0), (0, 2, 1, ly),
                                                 (1, 2, 0, 0.0), (1, 2, 0, lx)]):
        m = face == f
        xyz[m, ax_u] = u[m] * room[ax_u]
        xyz[m, ax_v] = v[m] * room[ax_v]
        xyz[m, ax_w] = val
    xyz = xyz + 0.01 * torch.randn(n_surf, 3, generator=g)
    clutter = torch.rand(n_clutter, 3, generator=g) * torch.tensor(room)
    xyz = torch.cat([xyz, clutter], 0)
    xyz = xyz[torch.randperm(P, generator=g)]

    scaling = math.log(mean_scale) + 0.4 * torch.randn(P, 3, generator=g)
    scaling[:, 2] += math.log(0.2)  # surfel-like
    rotation = torch.randn(P, 4, generator=g)
    rotation = rotation / rotation.norm(dim=1, keepdim=True)
    opacity = 1.5 + 1.5 * torch.randn(P, 1, generator=g)
    f_dc = (torch.rand(P, 1, 3, generator=g) * 2 - 1)
    n_rest = (sh_degree + 1) ** 2 - 1
    f_rest = 0.05 * torch.randn(P, n_rest, 3, generator=g)
    lf = torch.randn(P, 64, generator=g)
    lf = lf / lf.norm(dim=1, keepdim=True) * (0.5 + torch.rand(P, 1, generator=g))
    out = dict(xyz=xyz, features_dc=f_dc, features_rest=f_rest, lang_feat=lf, opacity=opacity, scaling=scaling,
               rotation=rotation)
    return {k: t.to(torch.float32).contiguous().to(device) for k, t in out.items()}


def activate(params):
    """The reference's activations (src/gaussian_model.cpp:46-68): exp / normalize / sigmoid / cat."""
    return dict(
        means3D=params["xyz"],
        shs=torch.cat([params["features_dc"], params["features_rest"]], dim=1),
        lang_feats=params["lang_feat"],
        opacities=torch.sigmoid(params["opacity"]),
        scales=torch.exp(params["scaling"]),
        rotations=torch.nn.functional.normalize(params["rotation"]),
    )
