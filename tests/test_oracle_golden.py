"""CPU: pin the oracle (oracle/lgs_oracle.c) against the golden outputs of the UNMODIFIED
reference kernels (tests/golden/*.npz, produced on a B200 by tests/golden/make_golden.py).
Integer work (radii, keys, sorted order, ranges) and the float chain that feeds the keys
(depths, pixel centres, conics) must be bit-exact; images 1e-4; gradients 1e-3."""
import numpy as np
import pytest

import cases
from conftest import golden


@pytest.mark.parametrize("name", list(cases.CASES))
def test_oracle_forward_matches_reference(name, oracle_mod):
    gd = golden(name)
    cs = cases.make_case(name)
    f = cases.oracle_forward(cs, oracle_mod)
    vis = gd["visible"]
    assert f["num_rendered"] == int(gd["num_rendered"])
    np.testing.assert_array_equal(f["radii"], gd["radii"])
    np.testing.assert_array_equal(f["tiles_touched"], gd["tiles_touched"].view(np.uint32))
    np.testing.assert_array_equal(f["depths"][vis].view(np.uint32), gd["depths"][vis].view(np.uint32))
    np.testing.assert_array_equal(f["means2D"][vis].view(np.uint32), gd["means2D"][vis].view(np.uint32))
    np.testing.assert_array_equal(f["conic_opacity"][vis].view(np.uint32), gd["conic_opacity"][vis].view(np.uint32))
    if "cov3D" in gd.files:
        np.testing.assert_array_equal(f["cov3D"][vis].view(np.uint32), gd["cov3D"][vis].view(np.uint32))
        assert cases.rel_err(f["rgb"][vis], gd["rgb"][vis]) <= 1e-6
    np.testing.assert_array_equal(f["keys_unsorted"], gd["keys_unsorted"].view(np.uint64))
    np.testing.assert_array_equal(f["values_unsorted"], gd["values_unsorted"].view(np.uint32))
    np.testing.assert_array_equal(f["keys_sorted"], gd["keys_sorted"].view(np.uint64))
    np.testing.assert_array_equal(f["point_list"], gd["point_list"].view(np.uint32))
    np.testing.assert_array_equal(f["ranges"], gd["ranges"].view(np.uint32))
    # libm expf vs CUDA expf differ by ulps: a vanishing fraction of threshold-straddling
    # fragments may flip, everything else agrees to rounding
    assert (f["n_contrib"] != gd["n_contrib"].view(np.uint32)).mean() <= 1e-3
    assert cases.rel_err(f["final_T"], gd["final_T"]) <= 1e-4
    assert cases.rel_err(f["out_color"], gd["out_color"]) <= 1e-4
    assert cases.rel_err(f["out_depth"], gd["out_depth"]) <= 1e-4
    assert cases.rel_err(f["out_lf"][cases.LF_GOLDEN_CH], gd["out_lf_sub"]) <= 1e-4
    if not cs["include_lf"]:
        assert not gd["out_lf_sub"].any()


@pytest.mark.parametrize("name", list(cases.CASES))
def test_oracle_backward_matches_reference(name, oracle_mod):
    gd = golden(name)
    cs = cases.make_case(name)
    f = cases.oracle_forward(cs, oracle_mod)
    g = cases.oracle_backward(cs, f, oracle_mod)
    for gname in cases.GRAD_NAMES:
        o = g[gname][:, cases.LF_GOLDEN_CH] if gname == "dL_dlang_feats" else g[gname]
        assert o.shape == gd[gname].shape, gname
        assert cases.rel_err(o, gd[gname]) <= 1e-3, gname


def test_oracle_adam_matches_torch(oracle_mod):
    gd = golden("adam")
    params, grads = cases.adam_case()
    p = {k: v.clone().numpy() for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(x) for k, x in p.items()}
    for step, gr in enumerate(grads):
        for k in p:
            oracle_mod.adam(p[k].reshape(-1), gr[k].numpy().reshape(-1), m[k].reshape(-1), v[k].reshape(-1),
                            cases.ADAM_LRS[k], step=step + 1)
            assert cases.rel_err(p[k], gd[f"p{step}_{k}"]) <= 1e-6, (k, step)
    for k in p:
        assert cases.rel_err(m[k], gd[f"m_{k}"]) <= 1e-6 and cases.rel_err(v[k], gd[f"v_{k}"]) <= 1e-6


def test_oracle_cosine_matches_torch(oracle_mod):
    gd = golden("cosine")
    feats, text = cases.cosine_case()
    sim = oracle_mod.cosine(feats.numpy(), text.numpy())
    assert np.abs(sim - gd["sim"]).max() <= 1e-6
    assert np.abs(sim - gd["sim_fp32_torch"]).max() <= 5e-6


def test_oracle_self_consistency(oracle_mod):
    """Properties that need no golden: permutation, stability, ranges, idempotence."""
    cs = cases.make_case("dense_opaque")
    f = cases.oracle_forward(cs, oracle_mod)
    f2 = cases.oracle_forward(cs, oracle_mod)
    for k in ("keys_sorted", "point_list", "ranges", "n_contrib", "out_color", "out_lf"):
        np.testing.assert_array_equal(f[k], f2[k])
    assert np.all(np.diff(f["keys_sorted"].astype(np.int64)) >= 0)
    assert sorted(f["keys_unsorted"].tolist()) == f["keys_sorted"].tolist()
    same = f["keys_sorted"][1:] == f["keys_sorted"][:-1]
    assert np.all(f["point_list"][1:][same].astype(np.int64) > f["point_list"][:-1][same].astype(np.int64))
    lens = f["ranges"][:, 1].astype(np.int64) - f["ranges"][:, 0]
    assert lens.sum() == f["num_rendered"] == f["tiles_touched"].sum()
    # early termination happened somewhere (T < 1e-4 stops blending) and T never goes negative
    assert (f["final_T"] < 1e-3).any() and (f["final_T"] >= 0).all()
    # no-LF call leaves the feature image untouched and the other images identical
    cs2 = dict(cs, include_lf=False)
    f3 = cases.oracle_forward(cs2, oracle_mod)
    np.testing.assert_array_equal(f3["out_color"], f["out_color"])
    assert not f3["out_lf"].any()
