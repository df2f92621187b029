"""L1 of the drop-in boundary in Python: the three functions the reference exports from
`rasterize_points.cu` / pybind `_C`, same names, argument order, return tuples and error
behaviour, implemented on liblgs.so (include/lgs.h) -- never on torch ops.

  rasterize_gaussians(...)          <- RasterizeGaussiansCUDA          (src/rasterize_points.cu:37-120)
  rasterize_gaussians_backward(...) <- RasterizeGaussiansBackwardCUDA  (src/rasterize_points.cu:122-209)
  mark_visible(...)                 <- markVisible                     (src/rasterize_points.cu:211-228)

The C++ twin of this file (csrc/host/rasterize_points.cpp) exports the libtorch signatures
themselves; both sit on the same C ABI.
"""
import ctypes

import torch

from . import _lib
from ._lib import check, ptr

NUM_CHANNELS = 3
LF_NUM_CHANNELS = 64


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


_scratch = {}


def _workspace(dev, stream, nbytes):
    """Grow-only scratch per (device, stream) for buffers that only live inside one call (the render
    backward's hand-off records): avoids a fresh ~0.5 KB/instance allocation of varying size every
    iteration, which fragments the caching allocator."""
    key = (dev.index, stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _scratch.pop(key, None)
        buf = torch.empty(int(nbytes * 1.25) + (1 << 20), dtype=torch.uint8, device=dev)
        _scratch[key] = buf
    return buf


def _round_up(n, q):
    return ((int(n) + q - 1) // q) * q


def _f32c(t):
    """`.contiguous()` as the reference does on every input (float32 is assumed there too)."""
    if t is None:
        return None
    if t.numel() and t.dtype != torch.float32:
        raise TypeError(f"expected float32 tensor, got {t.dtype}")
    return t.contiguous()


def rasterize_gaussians(background, means3D, colors, lang_feat, opacity, scales, rotations, scale_modifier,
                        cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width,
                        sh, degree, campos, prefiltered, include_lang_feat, sh_rest=None, capacity=None, buffers=None):
    """-> (num_rendered, out_color[3,H,W], out_lang_feat[64,H,W], out_depth[1,H,W], radii[P] int32,
           geomBuffer, binningBuffer, imgBuffer)

    `sh_rest` (not in the reference signature): when given, `sh` is features_dc [P,1,3] and `sh_rest` is
    features_rest [P,M-1,3] -- the reference's two parameter tensors, read in place instead of their
    per-iteration torch::cat (lgs_forward_stage1_split_sh).
    `capacity` (not in the reference signature): when given, the forward runs WITHOUT the reference's blocking read-back of
    num_rendered (rasterizer_impl.cu:281-282): the binning buffer is sized for `capacity` instances, R stays on the device,
    and the returned num_rendered is `capacity` -- the value to hand to the backward.  `forward_status(geomBuffer, P)` fetches
    the true R and the overflow flag asynchronously.  `buffers` = dict of caller-owned work buffers to reuse (geom, img,
    binning; any missing or too small one is allocated and stored back)."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise ValueError("means3D must have dimensions (num_points, 3)")  # AT_ERROR, rasterize_points.cu:59-61
    L = _lib.lib()
    if not means3D.is_cuda:
        raise _lib.LgsError("leg_slam_b200 has no CPU path: tensors must live on a CUDA device")
    P, H, W = int(means3D.size(0)), int(image_height), int(image_width)
    dev = means3D.device
    fopt = dict(dtype=torch.float32, device=dev)
    byte = dict(dtype=torch.uint8, device=dev)
    include_lf = bool(include_lang_feat)

    if P == 0:  # the reference returns its zero-filled outputs untouched
        return (0, torch.zeros((NUM_CHANNELS, H, W), **fopt), torch.zeros((LF_NUM_CHANNELS, H, W), **fopt),
                torch.zeros((1, H, W), **fopt), torch.zeros((0,), dtype=torch.int32, device=dev),
                torch.empty(0, **byte), torch.empty(0, **byte), torch.empty(0, **byte))

    # every pixel of the written outputs is written by the kernel: no memset needed
    out_color = torch.empty((NUM_CHANNELS, H, W), **fopt)
    out_depth = torch.empty((1, H, W), **fopt)
    out_lf = (torch.empty if include_lf else torch.zeros)((LF_NUM_CHANNELS, H, W), **fopt)
    radii = torch.empty((P,), dtype=torch.int32, device=dev)

    background, means3D, colors, lang_feat, opacity, scales, rotations, cov3D_precomp, viewmatrix, projmatrix, \
        sh, campos = map(_f32c, (background, means3D, colors, lang_feat, opacity, scales, rotations, cov3D_precomp,
                                 viewmatrix, projmatrix, sh, campos))
    M = int(sh.size(1)) if sh is not None and sh.size(0) != 0 else 0
    s = _stream(means3D)

    def _buf(name, nbytes):
        if buffers is None:
            return torch.empty(nbytes, **byte)
        b = buffers.get(name)
        if b is None or b.numel() < nbytes or b.device != dev:
            b = torch.empty(nbytes, **byte)
            buffers[name] = b
        return b

    geom = _buf("geom", L.lgs_geom_bytes(P))
    img = _buf("img", L.lgs_image_bytes(W, H))
    R = ctypes.c_int(0)
    Rref = None if capacity is not None else ctypes.byref(R)
    with torch.cuda.device(dev):
        if sh_rest is not None:
            sh_rest = _f32c(sh_rest)
            M = 1 + int(sh_rest.size(1))
            check(L.lgs_forward_stage1_split_sh(P, int(degree), M, W, H, ptr(means3D), ptr(sh), ptr(sh_rest), ptr(opacity),
                                                ptr(scales), float(scale_modifier), ptr(rotations), ptr(cov3D_precomp),
                                                ptr(viewmatrix), ptr(projmatrix), ptr(campos), float(tan_fovx),
                                                float(tan_fovy), int(bool(prefiltered)), geom.data_ptr(), radii.data_ptr(),
                                                Rref, s), "lgs_forward_stage1_split_sh")
        else:
            check(L.lgs_forward_stage1(P, int(degree), M, W, H, ptr(means3D), ptr(sh), ptr(colors), ptr(opacity),
                                       ptr(scales), float(scale_modifier), ptr(rotations), ptr(cov3D_precomp),
                                       ptr(viewmatrix), ptr(projmatrix), ptr(campos), float(tan_fovx), float(tan_fovy),
                                       int(bool(prefiltered)), geom.data_ptr(), radii.data_ptr(), Rref, s),
                  "lgs_forward_stage1")
        if capacity is not None:
            R.value = int(capacity)
            binning = _buf("binning", L.lgs_binning_bytes(R.value))
        else:
            # capacity rounded up so that consecutive iterations (R drifts slowly) reuse the same cached block
            binning = _buf("binning", L.lgs_binning_bytes(_round_up(R.value, 1 << 18)))
        check(L.lgs_forward_stage2(P, W, H, R.value, ptr(background), ptr(lang_feat) if include_lf else None,
                                   geom.data_ptr(), binning.data_ptr(), img.data_ptr(), out_color.data_ptr(),
                                   out_lf.data_ptr(), out_depth.data_ptr(), int(include_lf), s),
              "lgs_forward_stage2")
    return R.value, out_color, out_lf, out_depth, radii, geom, binning, img


_status_host = {}


def forward_status(geomBuffer, P):
    """-> pinned int32[4] = (R, largest depth bits, overflow, sort error) of the frame in `geomBuffer`, copied asynchronously on
    the current stream (lgs_forward_status): valid after the caller's next synchronisation of that stream."""
    dev = geomBuffer.device
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    st = _status_host.get(key)
    if st is None:
        st = torch.zeros(4, dtype=torch.int32).pin_memory()
        _status_host[key] = st
    with torch.cuda.device(dev):
        check(_lib.lib().lgs_forward_status(geomBuffer.data_ptr(), int(P), st.data_ptr(), key[1]), "lgs_forward_status")
    return st


def backward_outputs(P, M, dev, include_lf, has_sh, has_scales):
    """Freshly allocated gradient tensors of one backward call (torch.empty where the kernels write
    every row, zeros where a path is absent)."""
    mk = (lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)) if P else \
         (lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev))
    z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)  # noqa: E731
    return dict(dL_dmeans2D=mk(P, 3), dL_dcolors=mk(P, NUM_CHANNELS),
                # without language features nothing accumulates into dL_dlang_feats: it must still be zeros
                dL_dlang_feats=mk(P, LF_NUM_CHANNELS) if include_lf else z(P, LF_NUM_CHANNELS),
                dL_dopacity=mk(P, 1), dL_dmeans3D=mk(P, 3), dL_dcov3D=mk(P, 6), dL_dconic=mk(P, 2, 2),
                # the kernel fills (or zero-fills) these only on the path that produces them
                dL_dsh=mk(P, M, 3) if has_sh else z(P, M, 3), dL_dscales=mk(P, 3) if has_scales else z(P, 3),
                dL_drotations=mk(P, 4) if has_scales else z(P, 4))


def rasterize_gaussians_backward_into(out, background, means3D, radii, colors, lang_feat, scales, rotations,
                                      scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy,
                                      dL_dout_color, dL_dout_lang_feat, dL_dout_depth, sh, degree, campos, geomBuffer, R,
                                      binningBuffer, imageBuffer, include_lang_feat, sh_rest=None, accumulate_sh=False):
    """RasterizeGaussiansBackwardCUDA writing into caller-provided tensors `out` (keys of
    backward_outputs; any contiguous float32 storage, e.g. slices of a flat gradient buffer).

    Split SH layout (`sh_rest` given, see rasterize_gaussians): dL/dSH is written to out["dL_dfeatures_dc"] [P,1,3]
    and out["dL_dfeatures_rest"] [P,M-1,3]; `accumulate_sh` adds to them instead of overwriting."""
    L = _lib.lib()
    P = int(means3D.size(0))
    if P == 0:
        return out
    H, W = int(dL_dout_color.size(1)), int(dL_dout_color.size(2))
    dev = means3D.device
    M = int(sh.size(1)) if sh is not None and sh.size(0) != 0 else 0
    include_lf = bool(include_lang_feat)
    background, means3D, colors, lang_feat, scales, rotations, cov3D_precomp, viewmatrix, projmatrix, sh, \
        campos, dL_dout_color, dL_dout_lang_feat, dL_dout_depth = map(
            _f32c, (background, means3D, colors, lang_feat, scales, rotations, cov3D_precomp, viewmatrix,
                    projmatrix, sh, campos, dL_dout_color, dL_dout_lang_feat, dL_dout_depth))
    scratch = _workspace(dev, _stream(means3D), L.lgs_backward_scratch_bytes(int(R), W, H))
    if sh_rest is not None:
        sh_rest = _f32c(sh_rest)
        M = 1 + int(sh_rest.size(1))
        with torch.cuda.device(dev):
            check(L.lgs_backward_split_sh(
                P, int(degree), M, int(R), W, H, ptr(background), ptr(means3D), ptr(sh), ptr(sh_rest),
                ptr(lang_feat) if include_lf else None, ptr(scales), float(scale_modifier), ptr(rotations),
                ptr(cov3D_precomp), ptr(viewmatrix), ptr(projmatrix), ptr(campos), float(tan_fovx), float(tan_fovy),
                ptr(radii.contiguous()), ptr(geomBuffer), ptr(binningBuffer), ptr(imageBuffer), ptr(dL_dout_color),
                ptr(dL_dout_lang_feat) if include_lf else None, ptr(dL_dout_depth), out["dL_dmeans2D"].data_ptr(),
                out["dL_dconic"].data_ptr(), out["dL_dopacity"].data_ptr(), out["dL_dcolors"].data_ptr(),
                out["dL_dlang_feats"].data_ptr(), None, out["dL_dmeans3D"].data_ptr(), out["dL_dcov3D"].data_ptr(),
                out["dL_dfeatures_dc"].data_ptr(), out["dL_dfeatures_rest"].data_ptr(), int(bool(accumulate_sh)),
                ptr(out["dL_dscales"]), ptr(out["dL_drotations"]), int(include_lf), 1, scratch.data_ptr(),
                _stream(means3D)), "lgs_backward_split_sh")
        return out
    with torch.cuda.device(dev):
        check(L.lgs_backward(
            P, int(degree), M, int(R), W, H, ptr(background), ptr(means3D), ptr(sh), ptr(colors),
            ptr(lang_feat) if include_lf else None, ptr(scales), float(scale_modifier), ptr(rotations),
            ptr(cov3D_precomp), ptr(viewmatrix), ptr(projmatrix), ptr(campos), float(tan_fovx), float(tan_fovy),
            ptr(radii.contiguous()), ptr(geomBuffer), ptr(binningBuffer), ptr(imageBuffer), ptr(dL_dout_color),
            ptr(dL_dout_lang_feat) if include_lf else None, ptr(dL_dout_depth), out["dL_dmeans2D"].data_ptr(),
            out["dL_dconic"].data_ptr(), out["dL_dopacity"].data_ptr(), out["dL_dcolors"].data_ptr(),
            out["dL_dlang_feats"].data_ptr(), None, out["dL_dmeans3D"].data_ptr(), out["dL_dcov3D"].data_ptr(),
            ptr(out["dL_dsh"]), ptr(out["dL_dscales"]), ptr(out["dL_drotations"]), int(include_lf), 1,
            scratch.data_ptr(), _stream(means3D)), "lgs_backward")
    return out


def rasterize_gaussians_backward(background, means3D, radii, colors, lang_feat, scales, rotations, scale_modifier,
                                 cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color,
                                 dL_dout_lang_feat, dL_dout_depth, sh, degree, campos, geomBuffer, R, binningBuffer,
                                 imageBuffer, include_lang_feat):
    """-> (dL_dmeans2D[P,3], dL_dcolors[P,3], dL_dlang_feats[P,64], dL_dopacity[P,1], dL_dmeans3D[P,3],
           dL_dcov3D[P,6], dL_dsh[P,M,3], dL_dscales[P,3], dL_drotations[P,4])"""
    P = int(means3D.size(0))
    M = int(sh.size(1)) if sh is not None and sh.size(0) != 0 else 0
    has_sh = M != 0 and sh is not None and sh.numel() != 0
    has_scales = scales is not None and scales.numel() != 0
    o = backward_outputs(P, M, means3D.device, bool(include_lang_feat), has_sh, has_scales)
    rasterize_gaussians_backward_into(o, background, means3D, radii, colors, lang_feat, scales, rotations, scale_modifier,
                                      cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color,
                                      dL_dout_lang_feat, dL_dout_depth, sh, degree, campos, geomBuffer, R, binningBuffer,
                                      imageBuffer, include_lang_feat)
    return (o["dL_dmeans2D"], o["dL_dcolors"], o["dL_dlang_feats"], o["dL_dopacity"], o["dL_dmeans3D"], o["dL_dcov3D"],
            o["dL_dsh"], o["dL_dscales"], o["dL_drotations"])


def mark_visible(means3D, viewmatrix, projmatrix):
    """-> bool[P], true where the Gaussian's view-space z exceeds 0.2."""
    L = _lib.lib()
    P = int(means3D.size(0))
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    if P != 0:
        means3D, viewmatrix, projmatrix = map(_f32c, (means3D, viewmatrix, projmatrix))
        with torch.cuda.device(means3D.device):
            check(L.lgs_mark_visible(P, ptr(means3D), ptr(viewmatrix), ptr(projmatrix), present.data_ptr(),
                                     _stream(means3D)), "lgs_mark_visible")
    return present
