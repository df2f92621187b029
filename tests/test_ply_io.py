""".ply checkpoints: leg_slam_b200.ply_io against the numpy restatement of the
reference's writer / reader (oracle/ply_ref.py)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import ply_ref as PR  # noqa: E402


def _model(P, seed=0):
    g = np.random.default_rng(seed)
    f = lambda *s: g.standard_normal(s).astype(np.float32)  # noqa: E731
    return dict(xyz=f(P, 3), features_dc=f(P, 1, 3), features_rest=f(P, 15, 3), lang_feat=f(P, 64), opacity=f(P, 1),
                scaling=f(P, 3), rotation=f(P, 4))


def test_restatement_roundtrip_and_header_cpu(tmp_path):
    m = _model(37)
    path = tmp_path / "a.ply"
    PR.write_ply(path, m["xyz"], m["features_dc"], m["features_rest"], m["lang_feat"], m["opacity"], m["scaling"], m["rotation"])
    raw = open(path, "rb").read()
    head = raw[:raw.index(b"end_header\n")].decode().split("\n")
    assert head[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 37"]
    props = [h.split()[-1] for h in head if h.startswith("property float ")]
    assert props[:9] == ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] and len(props) == 126
    assert props[9 + 45] == "lf_0" and props[-8:] == ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    assert len(raw) == raw.index(b"end_header\n") + 11 + 37 * 126 * 4
    back = PR.read_ply(path)
    for k in m:
        np.testing.assert_array_equal(back[k], m[k], err_msg=k)
    # f_rest is stored channel-major: property f_rest_1 is coefficient 2 of the RED channel
    rec = np.frombuffer(raw[raw.index(b"end_header\n") + 11:], "<f4").reshape(37, 126)
    np.testing.assert_array_equal(rec[:, 9 + 1], m["features_rest"][:, 1, 0])
    np.testing.assert_array_equal(rec[:, 9 + 15], m["features_rest"][:, 0, 1])


@pytest.fixture(scope="module")
def ref_ply():
    """The reference's own tinyply behind savePly's / loadPly's call sequence (oracle/ref_ply_wrap.cpp -> oracle/_ref/ref_ply.so)."""
    import build_ref
    try:
        if os.path.isdir(build_ref.PLY_REF):
            build_ref.build_ply(verbose=False)
        return build_ref.load_ply()
    except FileNotFoundError as e:
        pytest.skip(str(e))


def _ptr(a):
    import ctypes
    return a.ctypes.data_as(ctypes.c_void_p)


def _ref_write(lib, path, m):
    P = m["xyz"].shape[0]
    dc = np.ascontiguousarray(np.transpose(m["features_dc"], (0, 2, 1)).reshape(P, -1))      # savePly: transpose(1,2).flatten(1)
    rest = np.ascontiguousarray(np.transpose(m["features_rest"], (0, 2, 1)).reshape(P, -1))
    normals = np.zeros_like(m["xyz"])
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (m["xyz"], normals, dc, rest, m["lang_feat"], m["opacity"], m["scaling"],
                                                                m["rotation"])]
    assert lib.ref_ply_write(str(path).encode(), P, dc.shape[1], rest.shape[1], m["lang_feat"].shape[1], *[_ptr(a) for a in arrs]) == 0


def _ref_read(lib, path, max_sh_degree=3, n_lf=64):
    """loadPly's requests through tinyply, then its from_blob(...).transpose(1, 2) (gaussian_model.cpp:946-956)."""
    P = int(lib.ref_ply_count(str(path).encode()))
    assert P >= 0
    n_rest = ((max_sh_degree + 1) ** 2 - 1) * 3
    z = lambda *s: np.zeros(s, np.float32)  # noqa: E731
    xyz, dc, rest, lf, op, sc, rot = z(P, 3), z(P, 3), z(P, n_rest), z(P, n_lf), z(P, 1), z(P, 3), z(P, 4)
    assert lib.ref_ply_read(str(path).encode(), max_sh_degree, n_lf, *[_ptr(a) for a in (xyz, dc, rest, lf, op, sc, rot)]) == 0
    return dict(xyz=xyz, features_dc=np.ascontiguousarray(dc.reshape(P, 3, 1).transpose(0, 2, 1)),
                features_rest=np.ascontiguousarray(rest.reshape(P, 3, n_rest // 3).transpose(0, 2, 1)), lang_feat=lf, opacity=op,
                scaling=sc, rotation=rot)


@pytest.mark.parametrize("P", [1, 37, 5003])
def test_restatement_pinned_by_the_reference_tinyply_cpu(tmp_path, ref_ply, P):
    """SURVEY.md 8f row 3 oracle pin: the numpy restatement (oracle/ply_ref.py) writes, byte for byte, the file the reference's
    own tinyply writes for savePly's call sequence, and the reference's reader (tinyply + loadPly's property requests and
    reshapes) returns exactly the tensors that went in -- from either file."""
    m = _model(P, seed=P)
    a, b = tmp_path / "restated.ply", tmp_path / "tinyply.ply"
    PR.write_ply(a, m["xyz"], m["features_dc"], m["features_rest"], m["lang_feat"], m["opacity"], m["scaling"], m["rotation"])
    _ref_write(ref_ply, b, m)
    assert open(a, "rb").read() == open(b, "rb").read()
    for path in (a, b):
        back = _ref_read(ref_ply, path)
        for k in m:
            np.testing.assert_array_equal(back[k], m[k], err_msg=k)
    restated = PR.read_ply(b)   # and the restated reader agrees with the reference's on the reference-written file
    for k in m:
        np.testing.assert_array_equal(restated[k], m[k], err_msg=k)


def test_column_table_and_header_parser_cpu(tmp_path):
    """Host logic of leg_slam_b200.ply_io without a GPU: the column table is the reference's property order and the
    header parser reads what the restated writer writes (plus the optimizer comments)."""
    from leg_slam_b200 import ply_io
    m = _model(5)
    path = tmp_path / "h.ply"
    PR.write_ply(path, m["xyz"], m["features_dc"], m["features_rest"], m["lang_feat"], m["opacity"], m["scaling"], m["rotation"])
    with open(path, "rb") as f:
        P, props, comments, off = ply_io.read_header(f)
    cols = ply_io._columns(15, 64, False)
    assert P == 5 and props == [c[0] for c in cols] and comments == {}
    assert off == open(path, "rb").read().index(b"end_header\n") + 11
    # f_rest_i is channel-major: property i holds coefficient i % 15 of channel i // 15 = element 3*(i%15) + i//15 of the row
    assert cols[9 + 16] == ("f_rest_16", "features_rest", 3 * 1 + 1)
    with_adam = ply_io._columns(15, 64, True)
    assert len(with_adam) == 126 + 2 * 123 and with_adam[126][0] == "adam_m_0" and with_adam[-1] == ("adam_v_122", "v:rotation", 3)
    bad = tmp_path / "bad.ply"
    bad.write_bytes(b"ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nend_header\n0\n")
    with open(bad, "rb") as f, pytest.raises(ValueError):
        ply_io.read_header(f)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.gpu
def test_save_matches_reference_format_and_loads_back(tmp_path, dev, ref_ply):
    from leg_slam_b200 import ply_io
    m = _model(5003, seed=2)
    t = {k: torch.from_numpy(v).to(dev) for k, v in m.items()}
    ours, ref = tmp_path / "ours.ply", tmp_path / "ref.ply"
    ply_io.save_ply(ours, t)
    PR.write_ply(ref, m["xyz"], m["features_dc"], m["features_rest"], m["lang_feat"], m["opacity"], m["scaling"], m["rotation"])
    assert open(ours, "rb").read() == open(ref, "rb").read()      # byte-identical files
    _ref_write(ref_ply, tmp_path / "tinyply.ply", m)              # ... and identical to what the reference's tinyply writes
    assert open(ours, "rb").read() == open(tmp_path / "tinyply.ply", "rb").read()
    for k, v in _ref_read(ref_ply, ours).items():                 # the reference's reader (tinyply + loadPly's requests) on our file
        np.testing.assert_array_equal(v, m[k], err_msg=k)
    back = PR.read_ply(ours)                                      # the reference's Python reader logic on our file
    for k in m:
        np.testing.assert_array_equal(back[k], m[k], err_msg=k)
    params, ea, eas, steps = ply_io.load_ply(ref, dev)            # our loader on a reference-format file
    assert ea is None and steps is None
    for k in m:
        assert params[k].shape == t[k].shape and torch.equal(params[k], t[k]), k


@pytest.mark.gpu
def test_checkpoint_with_optimizer_state_resumes(tmp_path, dev, ref_ply):
    from leg_slam_b200 import ply_io
    m = _model(1201, seed=4)
    t = {k: torch.from_numpy(v).to(dev) for k, v in m.items()}
    g = torch.Generator().manual_seed(1)
    ea = {k: torch.randn(v.shape, generator=g).to(dev) for k, v in t.items()}
    eas = {k: torch.rand(v.shape, generator=g).to(dev) for k, v in t.items()}
    steps = {k: 7 + i for i, k in enumerate(ply_io.PARAM_ORDER)}
    path = tmp_path / "ckpt.ply"
    ply_io.save_ply(path, t, ea, eas, steps)
    back = PR.read_ply(path)   # still a valid reference-format file: extra properties are ignored by name lookup
    for k in m:
        np.testing.assert_array_equal(back[k], m[k], err_msg=k)
    for k, v in _ref_read(ref_ply, path).items():   # the reference's tinyply skips the comments and the extra properties too
        np.testing.assert_array_equal(v, m[k], err_msg=k)
    p2, ea2, eas2, steps2 = ply_io.load_ply(path, dev)
    assert steps2 == steps
    for k in t:
        assert torch.equal(p2[k], t[k]) and torch.equal(ea2[k], ea[k]) and torch.equal(eas2[k], eas[k]), k
    with pytest.raises(ValueError):
        ply_io.load_ply(path, dev, max_sh_degree=2)


@pytest.mark.gpu
def test_mapper_resumes_from_checkpoint(tmp_path, dev):
    """Train, checkpoint, keep training; a second mapper loaded from the checkpoint takes the same next step."""
    from leg_slam_b200 import mapper as M, synthetic
    W, H = 96, 64
    sc = synthetic.make_scene(3000, seed=71, mean_scale=0.06, device=dev)
    cams = synthetic.make_cameras(1, W, H, seed=71)
    g = torch.Generator().manual_seed(72)
    win = [M.Keyframe(c.to(dev), torch.rand(3, H, W, generator=g).to(dev), torch.randn(64, 37, 37, generator=g).to(dev),
                      (torch.rand(1, H, W, generator=g) * 3).to(dev)) for c in cams]
    a = M.Mapper(sc, sh_degree=3)
    for _ in range(3):
        a.train_step(win)
    path = tmp_path / "m.ply"
    a.save_checkpoint(path)
    b = M.Mapper(synthetic.make_scene(10, seed=1, device=dev), sh_degree=3)   # different set, replaced by the checkpoint
    b.load_checkpoint(path)
    assert b.optimizer.state[b.params["xyz"]]["step"] == 3
    la, lb = a.train_step(win), b.train_step(win)
    # same parameters, same moments, same step count -> the same update up to the backward's atomic order
    assert abs(float(la) - float(lb)) <= 1e-6 * abs(float(la))
    for k in M.PARAM_ORDER:
        d = (a.params[k].detach() - b.params[k].detach()).abs().max()
        assert float(d) <= 3e-3 * 4 * M.DEFAULT_LRS[k] + 1e-12, (k, float(d))
