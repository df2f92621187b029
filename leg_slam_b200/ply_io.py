""".ply checkpoints of the Gaussian set: the reference's GaussianModel::savePly /
loadPly (src/gaussian_model.cpp:854-1075; Python reader eval/gaussian_model.py:58-111) with the two things its
loader lacks -- the language features are read back, and the Adam state can ride along as extra float properties
(`adam_m_<i>`, `adam_v_<i>` in parameter order, step counts as header comments), so training can resume.

File format = what tinyply writes for the reference: `ply / format binary_little_endian 1.0 / element vertex P /
property float <name> ... / end_header` followed by P interleaved float32 records with the properties
x y z nx ny nz f_dc_0..2 f_rest_0..(3K-1) lf_0..63 opacity scale_0..2 rot_0..3, f_dc / f_rest stored channel-major
(features.transpose(1, 2).flatten(1)).  Files written here load in the reference and vice versa (the reference
ignores properties it does not ask for).  The interleaving runs on the GPU (lgs_ply_pack / lgs_ply_unpack)."""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check

PARAM_ORDER = ("xyz", "features_dc", "features_rest", "lang_feat", "opacity", "scaling", "rotation")


def _columns(n_rest, n_lf, with_adam):
    """[(property name, tensor name or None, element index within the tensor's row)] in file order."""
    cols = [("x", "xyz", 0), ("y", "xyz", 1), ("z", "xyz", 2), ("nx", None, 0), ("ny", None, 0), ("nz", None, 0)]
    cols += [(f"f_dc_{c}", "features_dc", c) for c in range(3)]              # [P,1,3] -> transpose -> [P,3,1]
    for i in range(3 * n_rest):                                                # [P,K,3] -> transpose -> [P,3,K]
        c, k = divmod(i, n_rest)
        cols.append((f"f_rest_{i}", "features_rest", 3 * k + c))
    cols += [(f"lf_{i}", "lang_feat", i) for i in range(n_lf)]
    cols += [("opacity", "opacity", 0)]
    cols += [(f"scale_{i}", "scaling", i) for i in range(3)]
    cols += [(f"rot_{i}", "rotation", i) for i in range(4)]
    if with_adam:
        rows = dict(xyz=3, features_dc=3, features_rest=3 * n_rest, lang_feat=n_lf, opacity=1, scaling=3, rotation=4)
        for tag in ("m", "v"):
            i = 0
            for name in PARAM_ORDER:
                for e in range(rows[name]):
                    cols.append((f"adam_{tag}_{i}", f"{tag}:{name}", e))
                    i += 1
    return cols


def _run(pack, P, cols, tensors, block):
    """tensors: dict name -> [P, ...] float32 CUDA tensor; cols as from _columns / the parsed header."""
    names = sorted({c[1] for c in cols if c[1] is not None and c[1] in tensors})
    tid = {n: i for i, n in enumerate(names)}
    dev = block.device
    col_t = torch.tensor([tid.get(c[1], -1) if c[1] is not None else -1 for c in cols], dtype=torch.int32, device=dev)
    col_e = torch.tensor([c[2] for c in cols], dtype=torch.int32, device=dev)
    n = len(names)
    ptrs = (ctypes.c_void_p * n)(*[tensors[k].data_ptr() for k in names])
    rows = (ctypes.c_int * n)(*[tensors[k][0].numel() if P else 1 for k in names])
    L = _lib.lib()
    fn = L.lgs_ply_pack if pack else L.lgs_ply_unpack
    with torch.cuda.device(dev):
        check(fn(P, len(cols), col_t.data_ptr(), col_e.data_ptr(), n, ptrs, rows, block.data_ptr(),
                 torch.cuda.current_stream(dev).cuda_stream), "lgs_ply_pack" if pack else "lgs_ply_unpack")


def save_ply(path, params, exp_avg=None, exp_avg_sq=None, steps=None):
    """GaussianModel::savePly (+ optional optimizer state).  params: dict of the 7 CUDA parameter tensors."""
    xyz = params["xyz"]
    if not xyz.is_cuda:
        raise _lib.LgsError("leg_slam_b200 has no CPU path: tensors must live on a CUDA device")
    P, n_rest, n_lf = xyz.shape[0], params["features_rest"].shape[1], params["lang_feat"].shape[1]
    with_adam = exp_avg is not None and exp_avg_sq is not None
    cols = _columns(n_rest, n_lf, with_adam)
    tensors = {k: params[k].detach().contiguous().float() for k in PARAM_ORDER}
    if with_adam:
        tensors.update({f"m:{k}": exp_avg[k].contiguous().float() for k in PARAM_ORDER})
        tensors.update({f"v:{k}": exp_avg_sq[k].contiguous().float() for k in PARAM_ORDER})
    block = torch.empty(P, len(cols), dtype=torch.float32, device=xyz.device)
    _run(True, P, cols, tensors, block)
    header = ["ply", "format binary_little_endian 1.0"]
    if with_adam and steps is not None:
        header += [f"comment lgs_adam_step {k} {int(steps[k])}" for k in PARAM_ORDER]
    header += [f"element vertex {P}"] + [f"property float {c[0]}" for c in cols] + ["end_header"]
    host = block.cpu().numpy()
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        host.astype("<f4", copy=False).tofile(f)


def read_header(f):
    """-> (P, [property names], {comment key: value}, data offset).  float32 properties only (what the reference writes)."""
    if f.readline().strip() != b"ply":
        raise ValueError("not a ply file")
    P, props, comments = None, [], {}
    in_vertex = False
    while True:
        line = f.readline()
        if not line:
            raise ValueError("ply header not terminated")
        tok = line.decode("ascii").split()
        if not tok:
            continue
        if tok[0] == "format" and tok[1] != "binary_little_endian":
            raise ValueError("only binary_little_endian ply files are supported")
        elif tok[0] == "comment":
            if len(tok) >= 4 and tok[1] == "lgs_adam_step":
                comments[tok[2]] = int(tok[3])
        elif tok[0] == "element":
            in_vertex = tok[1] == "vertex"
            if in_vertex:
                P = int(tok[2])
        elif tok[0] == "property" and in_vertex:
            if tok[1] not in ("float", "float32"):
                raise ValueError(f"property {tok[-1]}: only float32 vertex properties are supported")
            props.append(tok[2])
        elif tok[0] == "end_header":
            break
    if P is None:
        raise ValueError("ply file has no vertex element")
    return P, props, comments, f.tell()


def load_ply(path, device, max_sh_degree=3):
    """GaussianModel::loadPly, also restoring the language features and, when present, the Adam state.
    -> (params, exp_avg or None, exp_avg_sq or None, steps or None)"""
    with open(path, "rb") as f:
        P, props, comments, off = read_header(f)
        f.seek(off)
        host = np.fromfile(f, dtype="<f4", count=P * len(props)).reshape(P, len(props))
    n_rest = (max_sh_degree + 1) ** 2 - 1
    n_have = sum(1 for p in props if p.startswith("f_rest_"))
    if n_have != 3 * n_rest:
        raise ValueError(f"file holds {n_have} f_rest properties, max_sh_degree={max_sh_degree} needs {3 * n_rest}")
    n_lf = sum(1 for p in props if p.startswith("lf_"))
    with_adam = any(p.startswith("adam_m_") for p in props)
    want = {c[0]: (c[1], c[2]) for c in _columns(n_rest, n_lf, with_adam)}
    cols = [(p,) + want.get(p, (None, 0)) for p in props]  # unknown properties are skipped, like the reference does
    dev = torch.device(device)
    f32 = dict(dtype=torch.float32, device=dev)
    shapes = dict(xyz=(P, 3), features_dc=(P, 1, 3), features_rest=(P, n_rest, 3), lang_feat=(P, n_lf), opacity=(P, 1),
                  scaling=(P, 3), rotation=(P, 4))
    tensors = {k: torch.zeros(shapes[k], **f32) for k in PARAM_ORDER}
    if with_adam:
        tensors.update({f"m:{k}": torch.zeros(shapes[k], **f32) for k in PARAM_ORDER})
        tensors.update({f"v:{k}": torch.zeros(shapes[k], **f32) for k in PARAM_ORDER})
    block = torch.from_numpy(np.ascontiguousarray(host)).to(dev)
    _run(False, P, cols, tensors, block)
    params = {k: tensors[k] for k in PARAM_ORDER}
    if not with_adam:
        return params, None, None, None
    return (params, {k: tensors[f"m:{k}"] for k in PARAM_ORDER}, {k: tensors[f"v:{k}"] for k in PARAM_ORDER},
            {k: comments.get(k, 0) for k in PARAM_ORDER})
