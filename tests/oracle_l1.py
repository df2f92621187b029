"""TEST INFRASTRUCTURE -- L1 of the boundary (RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA / markVisible, reference
include/rasterize_points.h:18-71) on the CPU oracle, recording every call.

Lets CPU tests put the SAME kernels under two pieces of glue -- the reference's own GaussianRasterizerFunction /
GaussianRasterizer / GaussianRenderer (compiled unmodified into oracle/_ref/ref_model.so) and the package's twins
(leg_slam_b200/rasterizer.py, renderer.py) -- and compare, call for call, what each hands down and what each makes of the
results.  The three opaque buffers carry a handle to the oracle's forward state."""
import numpy as np
import torch

import oracle as O

LF_NUM_CHANNELS = 64


class RecordingL1:
    LF_NUM_CHANNELS = LF_NUM_CHANNELS

    def __init__(self):
        self.calls = []      # (name, [args]) with tensors cloned at call time
        self._fwd = {}

    @staticmethod
    def _snap(a):
        return a.detach().clone() if torch.is_tensor(a) else a

    @staticmethod
    def _n(t):
        return None if t.numel() == 0 else t.detach().numpy()

    def rasterize_gaussians(self, bg, means3D, colors, lang_feat, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                            viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                            prefiltered, include_lang_feat):
        args = [bg, means3D, colors, lang_feat, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix,
                tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos, prefiltered, include_lang_feat]
        self.calls.append(("rasterize_gaussians", [self._snap(a) for a in args]))
        n = self._n
        f = O.forward(n(means3D), n(opacity), n(viewmatrix), n(projmatrix), n(campos), image_width, image_height, tan_fovx,
                      tan_fovy, n(bg), shs=n(sh), degree=degree, colors_precomp=n(colors), lang_feat=n(lang_feat),
                      scales=n(scales), rotations=n(rotations), scale_modifier=scale_modifier, cov3D_precomp=n(cov3D_precomp),
                      include_lf=bool(include_lang_feat))
        h = len(self._fwd) + 1
        self._fwd[h] = f
        buf = lambda: torch.tensor([h], dtype=torch.int64).view(torch.uint8)  # noqa: E731
        t = torch.from_numpy
        return (int(f["num_rendered"]), t(f["out_color"]), t(f["out_lf"]), t(f["out_depth"]), t(f["radii"].copy()), buf(), buf(),
                buf())

    def rasterize_gaussians_backward(self, bg, means3D, radii, colors, lang_feat, scales, rotations, scale_modifier,
                                     cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, dL_dout_lf,
                                     dL_dout_depth, sh, degree, campos, geomBuffer, R, binningBuffer, imageBuffer,
                                     include_lang_feat):
        args = [bg, means3D, radii, colors, lang_feat, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix,
                tan_fovx, tan_fovy, dL_dout_color, dL_dout_lf, dL_dout_depth, sh, degree, campos, geomBuffer, R, binningBuffer,
                imageBuffer, include_lang_feat]
        self.calls.append(("rasterize_gaussians_backward", [self._snap(a) for a in args]))
        hs = {int(b.view(torch.int64)[0]) for b in (geomBuffer, binningBuffer, imageBuffer)}
        assert len(hs) == 1, "the three buffers of one forward call"
        f = self._fwd[hs.pop()]
        assert R == f["num_rendered"] and np.array_equal(radii.numpy(), f["radii"])
        n = self._n
        g = O.backward(f, n(means3D), n(viewmatrix), n(projmatrix), n(campos), tan_fovx, tan_fovy, n(bg), n(dL_dout_color),
                       n(dL_dout_lf), n(dL_dout_depth), shs=n(sh), degree=degree, lang_feat=n(lang_feat), scales=n(scales),
                       rotations=n(rotations), scale_modifier=scale_modifier, cov3D_precomp=n(cov3D_precomp),
                       include_lf=bool(include_lang_feat))
        t = torch.from_numpy
        return tuple(t(g[k]) for k in ("dL_dmeans2D", "dL_dcolors", "dL_dlang_feats", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D",
                                       "dL_dsh", "dL_dscales", "dL_drotations"))

    def mark_visible(self, means3D, viewmatrix, projmatrix):
        self.calls.append(("mark_visible", [self._snap(a) for a in (means3D, viewmatrix, projmatrix)]))
        return torch.from_numpy(O.mark_visible(means3D.detach().numpy(), viewmatrix.numpy()).astype(bool))


def same_calls(a, b, rtol=0.0):
    """Two call logs are the same calls: names, argument count, every scalar equal, every tensor of the same dtype / shape
    and bit-equal -- or, with rtol, float tensors within rtol of their largest magnitude -- (an empty sentinel equals an empty
    sentinel of any shape with 0 elements)."""
    assert [c[0] for c in a] == [c[0] for c in b], ([c[0] for c in a], [c[0] for c in b])
    for (name, x), (_, y) in zip(a, b):
        assert len(x) == len(y), name
        for i, (u, v) in enumerate(zip(x, y)):
            if torch.is_tensor(u) or torch.is_tensor(v):
                assert torch.is_tensor(u) and torch.is_tensor(v), (name, i)
                assert u.dtype == v.dtype, (name, i, u.dtype, v.dtype)
                if u.numel() == 0 or v.numel() == 0:
                    assert u.numel() == v.numel() == 0, (name, i, tuple(u.shape), tuple(v.shape))
                elif rtol and u.is_floating_point():
                    assert u.shape == v.shape, (name, i)
                    assert float((u - v).abs().max()) <= rtol * max(1e-12, float(v.abs().max())), (name, i)
                else:
                    assert u.shape == v.shape and torch.equal(u, v), (name, i)
            elif isinstance(u, float) or isinstance(v, float):
                assert np.float32(u) == np.float32(v), (name, i, u, v)     # the C++ signature takes these as float
            else:
                assert u == v and type(u) is type(v), (name, i, u, v)
    return True
