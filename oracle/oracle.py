"""numpy front-end of the CPU oracle (oracle/lgs_oracle.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg -- never by
leg_slam_b200/.

`forward()` / `backward()` chain the restated reference stages exactly as
CudaRasterizer::Rasterizer::forward / backward do (cuda_rasterizer/rasterizer_impl.cu:198-343,
347-453) and return every intermediate the parity tests compare.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "lgs_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
SO = os.path.join(OUT_DIR, "liblgs_oracle.so")

_lib = None


def build(force=False):
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= os.path.getmtime(SRC):
        return SO
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", SRC, "-o", SO, "-lm"])
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_num_rendered.restype = ctypes.c_int64
        _lib.oracle_render_fwd.restype = ctypes.c_int64
        _lib.oracle_num_threads.restype = ctypes.c_int
        _lib.oracle_expon_lr.restype = ctypes.c_float
        _lib.oracle_expon_lr.argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int]
    return _lib


def num_threads():
    return lib().oracle_num_threads()


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


F, I, D64 = ctypes.c_float, ctypes.c_int, ctypes.c_double


def forward(means3D, opacities, viewmatrix, projmatrix, campos, W, H, tanfovx, tanfovy, bg, shs=None, degree=0,
            colors_precomp=None, lang_feat=None, scales=None, rotations=None, scale_modifier=1.0,
            cov3D_precomp=None, include_lf=True, render=True):
    """All inputs numpy float32 in the reference's layouts.  Returns a dict of every stage output."""
    L = lib()
    means3D, opacities, viewmatrix, projmatrix, campos, bg, shs, colors_precomp, lang_feat, scales, rotations, \
        cov3D_precomp = map(_f, (means3D, opacities, viewmatrix, projmatrix, campos, bg, shs, colors_precomp,
                                 lang_feat, scales, rotations, cov3D_precomp))
    P = means3D.shape[0]
    M = 0 if shs is None else shs.shape[1]
    o = dict(P=P, W=W, H=H)
    o["radii"] = np.zeros(P, np.int32)
    o["means2D"] = np.zeros((P, 2), np.float32)
    o["depths"] = np.zeros(P, np.float32)
    o["cov3D"] = np.zeros((P, 6), np.float32)
    o["conic_opacity"] = np.zeros((P, 4), np.float32)
    o["rgb"] = np.zeros((P, 3), np.float32)
    o["clamped"] = np.zeros((P, 3), np.uint8)
    o["tiles_touched"] = np.zeros(P, np.uint32)
    L.oracle_preprocess(I(P), I(degree), I(M), _p(means3D), _p(scales), F(scale_modifier), _p(rotations),
                        _p(opacities), _p(shs), _p(cov3D_precomp), _p(colors_precomp), _p(viewmatrix),
                        _p(projmatrix), _p(campos), I(W), I(H), F(tanfovx), F(tanfovy), _p(o["radii"]),
                        _p(o["means2D"]), _p(o["depths"]), _p(o["cov3D"]), _p(o["conic_opacity"]), _p(o["rgb"]),
                        _p(o["clamped"]), _p(o["tiles_touched"]))
    R = int(L.oracle_num_rendered(I(P), _p(o["tiles_touched"])))
    o["num_rendered"] = R
    tiles = ((W + 7) // 8) * ((H + 7) // 8)
    o["keys_unsorted"] = np.zeros(R, np.uint64)
    o["values_unsorted"] = np.zeros(R, np.uint32)
    o["keys_sorted"] = np.zeros(R, np.uint64)
    o["point_list"] = np.zeros(R, np.uint32)
    o["ranges"] = np.zeros((tiles, 2), np.uint32)
    L.oracle_binning(I(P), I(W), I(H), _p(o["means2D"]), _p(o["depths"]), _p(o["radii"]), _p(o["tiles_touched"]),
                     ctypes.c_int64(R), _p(o["keys_unsorted"]), _p(o["values_unsorted"]), _p(o["keys_sorted"]),
                     _p(o["point_list"]), _p(o["ranges"]))
    if not render:
        return o
    colors = colors_precomp if colors_precomp is not None else o["rgb"]
    o["colors"] = colors
    o["final_T"] = np.zeros(W * H, np.float32)
    o["n_contrib"] = np.zeros(W * H, np.uint32)
    o["out_color"] = np.zeros((3, H, W), np.float32)
    o["out_lf"] = np.zeros((64, H, W), np.float32)
    o["out_depth"] = np.zeros((1, H, W), np.float32)
    o["n_blended"] = int(L.oracle_render_fwd(
        I(W), I(H), _p(o["ranges"]), _p(o["point_list"]), _p(o["means2D"]), _p(colors), _p(lang_feat),
        _p(o["depths"]), _p(o["conic_opacity"]), _p(bg), I(int(include_lf)), _p(o["final_T"]), _p(o["n_contrib"]),
        _p(o["out_color"]), _p(o["out_lf"]), _p(o["out_depth"])))
    return o


def backward(fwd, means3D, viewmatrix, projmatrix, campos, tanfovx, tanfovy, bg, dL_dcolor, dL_dlf, dL_ddepth,
             shs=None, degree=0, lang_feat=None, scales=None, rotations=None, scale_modifier=1.0,
             cov3D_precomp=None, include_lf=True):
    """Gradients in the layout RasterizeGaussiansBackwardCUDA returns them (src/rasterize_points.cu:205-208)."""
    L = lib()
    means3D, viewmatrix, projmatrix, campos, bg, dL_dcolor, dL_dlf, dL_ddepth, shs, lang_feat, scales, rotations, \
        cov3D_precomp = map(_f, (means3D, viewmatrix, projmatrix, campos, bg, dL_dcolor, dL_dlf, dL_ddepth, shs,
                                 lang_feat, scales, rotations, cov3D_precomp))
    P, W, H = fwd["P"], fwd["W"], fwd["H"]
    M = 0 if shs is None else shs.shape[1]
    g = dict(dL_dmeans2D=np.zeros((P, 3), np.float32), dL_dconic=np.zeros((P, 4), np.float32),
             dL_dopacity=np.zeros((P, 1), np.float32), dL_dcolors=np.zeros((P, 3), np.float32),
             dL_dlang_feats=np.zeros((P, 64), np.float32), dL_ddepths=np.zeros((P, 1), np.float32),
             dL_dmeans3D=np.zeros((P, 3), np.float32), dL_dcov3D=np.zeros((P, 6), np.float32),
             dL_dsh=np.zeros((P, M, 3), np.float32), dL_dscales=np.zeros((P, 3), np.float32),
             dL_drotations=np.zeros((P, 4), np.float32))
    if fwd["num_rendered"] > 0:
        L.oracle_render_bwd(I(W), I(H), _p(fwd["ranges"]), _p(fwd["point_list"]), _p(bg), _p(fwd["means2D"]),
                            _p(fwd["conic_opacity"]), _p(fwd["colors"]), _p(lang_feat), _p(fwd["depths"]),
                            _p(fwd["final_T"]), _p(fwd["n_contrib"]), _p(dL_dcolor), _p(dL_dlf), _p(dL_ddepth),
                            I(int(include_lf)), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]),
                            _p(g["dL_dcolors"]), _p(g["dL_dlang_feats"]), _p(g["dL_ddepths"]))
    cov3D = cov3D_precomp if cov3D_precomp is not None else fwd["cov3D"]
    L.oracle_preprocess_bwd(I(P), I(degree), I(M), _p(means3D), _p(fwd["radii"]), _p(shs), _p(fwd["clamped"]),
                            _p(scales), _p(rotations), F(scale_modifier), _p(cov3D), _p(viewmatrix), _p(projmatrix),
                            _p(campos), I(W), I(H), F(tanfovx), F(tanfovy), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]),
                            _p(g["dL_dcolors"]), _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]), _p(g["dL_dsh"]),
                            _p(g["dL_dscales"]), _p(g["dL_drotations"]))
    return g


def adam(p, g, m, v, lr, beta1=0.9, beta2=0.999, eps=1e-15, step=1):
    """In place on float32 numpy arrays."""
    assert p.dtype == np.float32 and p.flags.c_contiguous
    lib().oracle_adam(ctypes.c_int64(p.size), _p(p), _p(_f(g)), _p(m), _p(v), D64(lr), D64(beta1), D64(beta2),
                      D64(eps), I(step))


def cosine(feats, text):
    feats, text = _f(feats), _f(text)
    out = np.zeros((feats.shape[0], text.shape[0]), np.float32)
    lib().oracle_cosine(I(feats.shape[0]), I(text.shape[0]), _p(feats), _p(text), _p(out))
    return out


def mark_visible(means3D, viewmatrix):
    means3D, viewmatrix = _f(means3D), _f(viewmatrix)
    out = np.zeros(means3D.shape[0], np.uint8)
    lib().oracle_mark_visible(I(means3D.shape[0]), _p(means3D), _p(viewmatrix), _p(out))
    return out.astype(bool)


def expon_lr(step, lr_init, lr_final, lr_delay_mult=1.0, lr_delay_steps=0, max_steps=1_000_000):
    return float(lib().oracle_expon_lr(int(step), float(lr_init), float(lr_final), float(lr_delay_mult), int(lr_delay_steps),
                                       int(max_steps)))
