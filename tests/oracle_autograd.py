"""torch.autograd wrapper around the CPU oracle (test infrastructure): lets CPU tests drive the
host-side mapping logic (leg_slam_b200.mapper) with the restated reference rasterizer in place
of the CUDA one."""
import numpy as np
import torch

import oracle as O


class _OracleRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, shs, lang_feats, opacities, scales, rotations, cam, bg, degree):
        n = lambda t: t.detach().cpu().numpy()  # noqa: E731
        f = O.forward(n(means3D), n(opacities), n(cam.viewmatrix), n(cam.projmatrix), n(cam.campos), cam.width,
                      cam.height, cam.tanfovx, cam.tanfovy, n(bg), shs=n(shs), degree=degree, lang_feat=n(lang_feats),
                      scales=n(scales), rotations=n(rotations))
        ctx.f, ctx.cam, ctx.degree = f, cam, degree
        ctx.save_for_backward(means3D, shs, lang_feats, scales, rotations, bg)
        t = torch.from_numpy
        return t(f["out_color"]), t(f["out_lf"]), t(f["out_depth"]), t(f["radii"].copy())

    @staticmethod
    def backward(ctx, gc, gl, gd, _gr=None):
        means3D, shs, lang_feats, scales, rotations, bg = ctx.saved_tensors
        cam, f = ctx.cam, ctx.f
        H, W = cam.height, cam.width
        n = lambda t: t.detach().cpu().numpy()  # noqa: E731
        z = lambda c: np.zeros((c, H, W), np.float32)  # noqa: E731
        g = O.backward(f, n(means3D), n(cam.viewmatrix), n(cam.projmatrix), n(cam.campos), cam.tanfovx, cam.tanfovy,
                       n(bg), z(3) if gc is None else n(gc), z(64) if gl is None else n(gl),
                       z(1) if gd is None else n(gd), shs=n(shs), degree=ctx.degree, lang_feat=n(lang_feats),
                       scales=n(scales), rotations=n(rotations))
        t = torch.from_numpy
        return (t(g["dL_dmeans3D"]), t(g["dL_dsh"]), t(g["dL_dlang_feats"]), t(g["dL_dopacity"]), t(g["dL_dscales"]),
                t(g["dL_drotations"]), None, None, None)


def make_render_fn(bg, degree=3):
    def render(cam, a):
        return _OracleRasterize.apply(a["means3D"], a["shs"], a["lang_feats"], a["opacities"], a["scales"],
                                      a["rotations"], cam, bg, degree)
    return render
