"""TEST INFRASTRUCTURE -- numpy restatement of the reference's .ply checkpoint format.

write_ply follows GaussianModel::savePly (/root/reference/src/gaussian_model.cpp:972-1075: property order, the
transpose(1,2).flatten(1) of f_dc / f_rest, zero normals) with tinyply's binary writer (header lines `property float
<name>`, little-endian float32 records).  read_ply follows the reference's Python reader
(/root/reference/eval/gaussian_model.py:58-111: properties looked up BY NAME, f_rest reshaped (P, 3, K) then
transposed).  PINNED (tests/test_ply_io.py::test_restatement_pinned_by_the_reference_tinyply_cpu): the file written here is
byte-identical to the one the reference's own tinyply (third_party/tinyply, compiled unmodified into oracle/_ref/ref_ply.so)
writes for savePly's call sequence, and the reference's reader -- tinyply with loadPly's property requests and reshapes
(gaussian_model.cpp:882-956) -- returns the same tensors.  Also held to GaussianModel::savePly ITSELF: the unmodified class
(oracle/_ref/ref_model.so, CPU tensors) writes a byte-identical file and its loadPly parses the one written here
(tests/test_reference_model.py::test_ply_restatement_equals_reference_model_save_and_load); read_ply is held to the reference's
Python reader itself, imported unmodified (tests/test_reference_eval.py)."""
import numpy as np


def write_ply(path, xyz, f_dc, f_rest, lf, opacity, scale, rot):
    P = xyz.shape[0]
    cols = [xyz, np.zeros_like(xyz), np.transpose(f_dc, (0, 2, 1)).reshape(P, -1), np.transpose(f_rest, (0, 2, 1)).reshape(P, -1),
            lf, opacity.reshape(P, 1), scale, rot]
    names = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(cols[2].shape[1])] + \
            [f"f_rest_{i}" for i in range(cols[3].shape[1])] + [f"lf_{i}" for i in range(lf.shape[1])] + ["opacity"] + \
            [f"scale_{i}" for i in range(scale.shape[1])] + [f"rot_{i}" for i in range(rot.shape[1])]
    block = np.concatenate([np.asarray(c, np.float32) for c in cols], axis=1)
    assert block.shape[1] == len(names)
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % P).encode())
        f.write("".join(f"property float {n}\n" for n in names).encode())
        f.write(b"end_header\n")
        block.astype("<f4").tofile(f)


def read_ply(path, max_sh_degree=3):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"ply"
        names, P = [], 0
        while True:
            tok = f.readline().decode().split()
            if tok[0] == "element" and tok[1] == "vertex":
                P = int(tok[2])
            elif tok[0] == "property":
                names.append(tok[2])
            elif tok[0] == "end_header":
                break
        data = np.fromfile(f, dtype="<f4", count=P * len(names)).reshape(P, len(names))
    col = {n: data[:, i] for i, n in enumerate(names)}
    by = lambda prefix: sorted([n for n in names if n.startswith(prefix)], key=lambda x: int(x.split("_")[-1]))  # noqa: E731
    xyz = np.stack([col["x"], col["y"], col["z"]], axis=1)
    f_dc = np.stack([col["f_dc_0"], col["f_dc_1"], col["f_dc_2"]], axis=1)[:, :, None]                  # (P, 3, 1)
    f_rest = np.stack([col[n] for n in by("f_rest_")], axis=1).reshape(P, 3, (max_sh_degree + 1) ** 2 - 1)
    return dict(xyz=xyz, features_dc=np.transpose(f_dc, (0, 2, 1)).copy(), features_rest=np.transpose(f_rest, (0, 2, 1)).copy(),
                lang_feat=np.stack([col[n] for n in by("lf_")], axis=1), opacity=col["opacity"][:, None].copy(),
                scaling=np.stack([col[n] for n in by("scale_")], axis=1), rotation=np.stack([col[n] for n in by("rot")], axis=1))
