"""Adaptive density control: the fused CUDA path (lgs_densify_*) against the torch
restatement of the reference's densifyAndPrune sequence (oracle/densify_ref.py), plus the restatement's own
invariants on CPU.  The restatement itself is held, bit for bit, to the unmodified reference GaussianModel class running on CPU
tensors (tests/test_reference_model.py; fixture tests/golden/model.npz)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import densify_ref as DR  # noqa: E402


def make_model(P, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    params = dict(xyz=r(P, 3), features_dc=r(P, 1, 3), features_rest=r(P, 15, 3) * 0.1, lang_feat=r(P, 64),
                  opacity=r(P, 1) * 3.0,                      # sigmoid spans (0, 1): some fall under min_opacity
                  scaling=r(P, 3) * 0.8 - 3.5,                # exp -> around 0.03, both sides of percent_dense * extent
                  rotation=r(P, 4))
    params = {k: v.to(dev).contiguous() for k, v in params.items()}
    m = DR.Model(params)
    for k in DR.PARAMS:  # non-trivial Adam moments
        m.m[k] = (r(*m.p[k].shape) * 0.01).to(dev)
        m.v[k] = (r(*m.p[k].shape) ** 2 * 1e-4).to(dev)
    m.exist_since_iter = torch.randint(0, 50, (P,), generator=g, dtype=torch.int32).to(dev)
    # statistics: a few views' worth, some never seen (0/0 -> nan -> 0)
    seen = torch.rand(P, generator=g) < 0.8
    m.denom = (seen.float() * torch.randint(1, 6, (P,), generator=g).float()).unsqueeze(1).to(dev)
    m.xyz_gradient_accum = (m.denom.cpu() * torch.rand(P, 1, generator=g) * 4e-4).to(dev)
    m.max_radii2D = (torch.rand(P, generator=g) * 40).to(dev)
    return m


def clone_model(m):
    c = DR.Model(m.p)
    c.m = {k: v.clone() for k, v in m.m.items()}
    c.v = {k: v.clone() for k, v in m.v.items()}
    c.exist_since_iter = m.exist_since_iter.clone()
    c.xyz_gradient_accum, c.denom, c.max_radii2D = m.xyz_gradient_accum.clone(), m.denom.clone(), m.max_radii2D.clone()
    return c


ARGS = dict(max_grad=2e-4, min_opacity=0.05, extent=4.0)


def test_restatement_invariants_cpu():
    m = make_model(3000, torch.device("cpu"), seed=3)
    ref = clone_model(m)
    g = torch.Generator().manual_seed(7)
    z = {}
    def normal01(n):
        z["n"] = n
        return torch.randn(n, 3, generator=g)
    grads = torch.nan_to_num(m.xyz_gradient_accum / m.denom, nan=0.0).squeeze()
    smax = torch.exp(m.p["scaling"]).max(dim=1).values
    n_clone = int(((grads >= ARGS["max_grad"]) & (smax <= 0.01 * ARGS["extent"])).sum())
    n_split = int(((grads >= ARGS["max_grad"]) & (smax > 0.01 * ARGS["extent"])).sum())
    assert n_clone > 0 and n_split > 0  # the fixture exercises both branches
    ref.densify_and_prune(ARGS["max_grad"], ARGS["min_opacity"], ARGS["extent"], 0, normal01)
    assert z["n"] == 2 * n_split
    P2 = ref.p["xyz"].shape[0]
    for k in DR.PARAMS:
        assert ref.p[k].shape[0] == P2 and ref.m[k].shape == ref.p[k].shape and ref.v[k].shape == ref.p[k].shape
    assert ref.xyz_gradient_accum.shape == (P2, 1) and not ref.xyz_gradient_accum.any() and not ref.max_radii2D.any()
    assert bool((torch.sigmoid(ref.p["opacity"]) >= ARGS["min_opacity"]).all())  # everything below was pruned
    # the moments of appended points are zero, those of survivors are carried over
    n_keep = int((~(((grads >= ARGS["max_grad"]) & (smax > 0.01 * ARGS["extent"]))) &
                  (torch.sigmoid(m.p["opacity"]).squeeze(-1) >= ARGS["min_opacity"])).sum())
    assert not ref.m["lang_feat"][n_keep:].any() and ref.m["lang_feat"][:n_keep].any()
    ref.reset_opacity()
    assert not ref.m["opacity"].any() and not ref.v["opacity"].any()


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("max_screen_size", [0, 20])
@pytest.mark.parametrize("P", [1, 4097, 60000])
def test_fused_densify_matches_restatement(P, max_screen_size, dev):
    from leg_slam_b200 import densify as D
    m = make_model(P, dev, seed=P)
    ref = clone_model(m)
    zs = {}
    def normal01(n):
        if "z" not in zs:
            zs["z"] = torch.randn(n, 3, generator=torch.Generator().manual_seed(11)).to(dev)
        assert zs["z"].shape[0] == n
        return zs["z"]
    ref.densify_and_prune(ARGS["max_grad"], ARGS["min_opacity"], ARGS["extent"], max_screen_size, normal01)
    st = D.DensifyStats(P, dev)
    st.xyz_gradient_accum, st.denom, st.max_radii2D = m.xyz_gradient_accum.clone(), m.denom.clone(), m.max_radii2D.clone()
    st.exist_since_iter = m.exist_since_iter.clone()
    p2, m2, v2, st2, info = D.densify_and_prune(m.p, m.m, m.v, st, ARGS["max_grad"], ARGS["min_opacity"], ARGS["extent"],
                                                max_screen_size, normal01=normal01)
    torch.cuda.synchronize()
    assert info["new_P"] == ref.p["xyz"].shape[0]
    n_old = info["kept"] + info["cloned"]
    for k in DR.PARAMS:
        a, b = p2[k], ref.p[k]
        assert a.shape == b.shape, k
        if k in ("xyz", "scaling"):  # children are computed: same formula, different instruction order
            assert torch.equal(a[:n_old], b[:n_old]), k
            if a.shape[0] > n_old:
                err = (a[n_old:] - b[n_old:]).abs().max() / b[n_old:].abs().max().clamp_min(1e-12)
                assert float(err) <= 2e-6, (k, float(err))
        else:
            assert torch.equal(a, b), k
        assert torch.equal(m2[k], ref.m[k]) and torch.equal(v2[k], ref.v[k]), k
    assert torch.equal(st2.exist_since_iter, ref.exist_since_iter)
    assert not st2.xyz_gradient_accum.any() and not st2.denom.any() and not st2.max_radii2D.any()
    # inputs untouched
    assert m.p["xyz"].shape[0] == P


@pytest.mark.gpu
def test_fused_stats_match_restatement(dev):
    from leg_slam_b200 import densify as D
    P = 50000
    m = make_model(P, dev, seed=5)
    g = torch.Generator().manual_seed(9)
    radii = (torch.randint(-2, 30, (P,), generator=g, dtype=torch.int32)).clamp_min(0).to(dev)
    grad = (torch.randn(P, 3, generator=g) * 1e-3).to(dev)
    st = D.DensifyStats(P, dev)
    st.xyz_gradient_accum, st.denom, st.max_radii2D = m.xyz_gradient_accum.clone(), m.denom.clone(), m.max_radii2D.clone()
    st.add(radii, grad)
    m.add_stats(radii, grad)
    torch.cuda.synchronize()
    assert torch.equal(st.denom, m.denom) and torch.equal(st.max_radii2D, m.max_radii2D)
    np.testing.assert_allclose(st.xyz_gradient_accum.cpu().numpy(), m.xyz_gradient_accum.cpu().numpy(), rtol=1e-6, atol=0)  # norm: fused multiply-add vs torch.norm


@pytest.mark.gpu
def test_mapper_densify_keeps_training(dev):
    """Statistics accumulate inside train_step, densify_and_prune rebuilds parameters + Adam state + the flat
    gradient buffer, and the mapping iteration keeps running on the new point set."""
    from leg_slam_b200 import mapper as M, synthetic
    W, H = 96, 64
    sc = synthetic.make_scene(6000, seed=61, mean_scale=0.06, device=dev)
    cams = synthetic.make_cameras(2, W, H, seed=61)
    g = torch.Generator().manual_seed(62)
    win = [M.Keyframe(c.to(dev), torch.rand(3, H, W, generator=g).to(dev), torch.randn(64, 37, 37, generator=g).to(dev),
                      (torch.rand(1, H, W, generator=g) * 3).to(dev)) for c in cams]
    mp = M.Mapper(sc, sh_degree=3, track_densify_stats=True)
    for _ in range(4):
        l0 = mp.train_step(win)
    assert float(mp.stats.denom.max()) == 8.0  # 4 iterations x 2 views for the always-visible Gaussians
    P0 = mp.params["xyz"].shape[0]
    m_before = mp.optimizer.state[mp.params["lang_feat"]]["exp_avg"].clone()
    info = mp.densify_and_prune(max_grad=1e-7, min_opacity=0.005, extent=6.0, max_screen_size=0,
                                generator=torch.Generator(device=dev).manual_seed(1))
    P1 = mp.params["xyz"].shape[0]
    assert P1 == info["new_P"] and P1 != P0 and info["cloned"] + info["split_selected"] > 0
    st = mp.optimizer.state[mp.params["lang_feat"]]
    assert st["step"] == 4 and st["exp_avg"].shape == (P1, 64)
    assert not st["exp_avg"][info["kept"]:].any() and m_before.any() and st["exp_avg"][:info["kept"]].any()
    assert 123 * P1 <= mp.grads.flat.numel() <= 123 * P1 + 7 * 3  # 16-byte aligned views
    for _ in range(2):
        l1 = mp.train_step(win)
    assert torch.isfinite(l1) and mp.stats.denom.shape == (P1, 1)


def _new_points(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 3, generator=g) * 2.0 - 1.0, torch.rand(n, 3, generator=g)


def test_increase_pcd_restatement_invariants_cpu():
    """GaussianModel::increasePcd as restated (oracle/densify_ref.py, reference src/gaussian_model.cpp:297-384): the new rows
    carry the colour as DC coefficient, zero higher SH / language features, log-sqrt 3-NN scale, identity rotation, opacity
    inverse_sigmoid(0.1), zero Adam moments; the old rows and moments are untouched; the statistics restart at the new size."""
    import ingest_ref as IR
    m = make_model(500, torch.device("cpu"), seed=5)
    ref = clone_model(m)
    pts, cols = _new_points(37, 6)
    dist2 = lambda p: torch.from_numpy(IR.knn_mean_dist2(p.numpy()))  # noqa: E731
    m.increase_pcd(pts, cols, 17, dist2)
    P, n = 500, 37
    for k in DR.PARAMS:
        assert m.p[k].shape[0] == P + n and torch.equal(m.p[k][:P], ref.p[k])
        assert torch.equal(m.m[k][:P], ref.m[k]) and torch.equal(m.v[k][:P], ref.v[k])
        assert not m.m[k][P:].any() and not m.v[k][P:].any()
    assert torch.equal(m.p["xyz"][P:], pts)
    np.testing.assert_allclose(m.p["features_dc"][P:, 0].numpy() * 0.28209479177387814 + 0.5, cols.numpy(), atol=1e-6)
    assert not m.p["features_rest"][P:].any() and not m.p["lang_feat"][P:].any()
    assert torch.equal(m.p["rotation"][P:], torch.tensor([1.0, 0, 0, 0]).expand(n, 4))
    np.testing.assert_allclose(torch.sigmoid(m.p["opacity"][P:]).numpy(), 0.1, atol=1e-6)
    d2 = np.maximum(IR.knn_mean_dist2(pts.numpy()), 1e-7)
    np.testing.assert_allclose(torch.exp(m.p["scaling"][P:]).numpy(), np.sqrt(d2)[:, None].repeat(3, 1), rtol=1e-5)
    assert torch.equal(m.exist_since_iter[:P], ref.exist_since_iter) and (m.exist_since_iter[P:] == 17).all()
    assert m.denom.shape == (P + n, 1) and not m.denom.any() and not m.xyz_gradient_accum.any() and not m.max_radii2D.any()
    before = m.p["xyz"].shape[0]
    m.increase_pcd(pts[:0], cols[:0], 18, dist2)  # no new points: nothing changes (:299-300)
    assert m.p["xyz"].shape[0] == before


@pytest.mark.gpu
def test_increase_pcd_matches_restatement(dev):
    """leg_slam_b200.densify.increase_pcd against the restatement (same distCUDA2 kernel for the scale): bit-exact."""
    from leg_slam_b200 import densify as D, ingest
    m = make_model(4000, dev, seed=8)
    ref, orig = clone_model(m), clone_model(m)
    pts, cols = _new_points(1234, 9)
    pts, cols = pts.to(dev), cols.to(dev)
    st = D.DensifyStats(4000, dev)
    st.exist_since_iter = m.exist_since_iter.clone()
    p2, m2, v2, st2 = D.increase_pcd(m.p, m.m, m.v, st, pts, cols, 23)
    ref.increase_pcd(pts, cols, 23, ingest.distCUDA2)
    for k in DR.PARAMS:
        assert p2[k].is_contiguous() and torch.equal(p2[k], ref.p[k]), k
        assert torch.equal(m2[k], ref.m[k]) and torch.equal(v2[k], ref.v[k]), k
        assert torch.equal(m.p[k], orig.p[k]) and torch.equal(m.m[k], orig.m[k])  # inputs untouched
    assert torch.equal(st2.exist_since_iter, ref.exist_since_iter)
    assert st2.denom.shape == ref.denom.shape and not st2.denom.any() and not st2.max_radii2D.any()
    try:  # the DC coefficient against the reference's own RGB2SH (unmodified include/sh_utils.h) on the same device
        import build_ref
        assert torch.equal(p2["features_dc"][4000:, 0], build_ref.load_utils().RGB2SH(cols))
    except FileNotFoundError:
        pass
    same = D.increase_pcd(m.p, m.m, m.v, st, pts[:0], cols[:0], 24)
    assert same[0] is m.p and same[3] is st
    with pytest.raises(Exception, match="no CPU path"):
        D.increase_pcd(m.p, m.m, m.v, st, pts.cpu(), cols.cpu(), 24)


@pytest.mark.gpu
def test_mapper_increase_pcd_and_reset_opacity(dev):
    """A keyframe's new points join the Gaussian set between two mapping iterations: Adam state of the old rows and the step
    counts carry over, the new rows start from zero moments, and training continues; resetOpacity zeroes the opacity moments."""
    from leg_slam_b200 import mapper as M, synthetic
    W, H = 96, 64
    sc = synthetic.make_scene(5000, seed=63, mean_scale=0.06, device=dev)
    cams = synthetic.make_cameras(2, W, H, seed=63)
    g = torch.Generator().manual_seed(64)
    win = [M.Keyframe(c.to(dev), torch.rand(3, H, W, generator=g).to(dev), torch.randn(64, 37, 37, generator=g).to(dev),
                      (torch.rand(1, H, W, generator=g) * 3).to(dev)) for c in cams]
    mp = M.Mapper(sc, sh_degree=3, track_densify_stats=True)
    for _ in range(3):
        mp.train_step(win)
    P0 = mp.params["xyz"].shape[0]
    xyz_before = mp.params["xyz"].detach().clone()
    m_before = mp.optimizer.state[mp.params["xyz"]]["exp_avg"].clone()
    pts, cols = _new_points(777, 65)
    assert mp.increase_pcd(pts.to(dev) + torch.tensor([3.0, 2.0, 1.4], device=dev), cols.to(dev), iteration=3) == 777
    P1 = mp.params["xyz"].shape[0]
    assert P1 == P0 + 777 and torch.equal(mp.params["xyz"].detach()[:P0], xyz_before)
    st = mp.optimizer.state[mp.params["xyz"]]
    assert st["step"] == 3 and torch.equal(st["exp_avg"][:P0], m_before) and not st["exp_avg"][P0:].any()
    assert (mp.stats.exist_since_iter[P0:] == 3).all() and mp.stats.denom.shape == (P1, 1) and not mp.stats.denom.any()
    l1 = mp.train_step(win)
    assert torch.isfinite(l1) and mp.optimizer.state[mp.params["xyz"]]["step"] == 4
    op_before = torch.sigmoid(mp.params["opacity"].detach()).clone()
    mp.reset_opacity()
    so = mp.optimizer.state[mp.params["opacity"]]
    assert not so["exp_avg"].any() and not so["exp_avg_sq"].any() and so["step"] == 4
    np.testing.assert_allclose(torch.sigmoid(mp.params["opacity"].detach()).cpu().numpy(), op_before.cpu().numpy(), atol=1e-6)
    assert torch.isfinite(mp.train_step(win))


@pytest.mark.gpu
def test_mapper_loop_closure_correction(dev):
    """GaussianModel::scaledTransformVisiblePointsOfKeyframe on the mapper (reference src/gaussian_model.cpp:422-481): the rows
    the unmodified reference operator selects and moves (oracle/_ref/ref_geometry.so on clones of the same tensors) are the
    rows the mapper's parameters end with; every rotation row is the activated one; xyz / rotation moments are zeroed, their
    step counts and every other tensor's Adam state are untouched; training continues."""
    import build_ref
    from leg_slam_b200 import mapper as M, synthetic
    try:
        ref = build_ref.load_geometry()
    except FileNotFoundError as ex:
        pytest.skip(str(ex))
    W, H = 96, 64
    sc = synthetic.make_scene(6000, seed=71, mean_scale=0.06, device=dev)
    cams = [c.to(dev) for c in synthetic.make_cameras(2, W, H, seed=71)]
    g = torch.Generator().manual_seed(72)
    win = [M.Keyframe(c, torch.rand(3, H, W, generator=g).to(dev), torch.randn(64, 37, 37, generator=g).to(dev),
                      (torch.rand(1, H, W, generator=g) * 3).to(dev)) for c in cams]
    mp = M.Mapper(sc, sh_degree=3, track_densify_stats=True)
    for _ in range(3):
        mp.train_step(win)
    P = mp.params["xyz"].shape[0]
    mp.stats.exist_since_iter.copy_(torch.randint(0, 40, (P,), generator=g).to(dev))  # ages relative to the keyframe
    flags = (torch.rand(P, generator=g) > 0.2).to(dev)
    diff = torch.eye(4)
    diff[:3, :3] = torch.tensor([[0.9998, -0.02, 0.0], [0.02, 0.9998, 0.0], [0.0, 0.0, 1.0]])
    diff[:3, 3] = torch.tensor([0.05, -0.02, 0.01])
    diff_t = diff.t().contiguous().to(dev)  # pose tensors are stored transposed (src/gaussian_mapper.cpp:931-933)
    cam = cams[0]
    # the reference's sequence on clones: activation, unstable flags, the unmodified operator
    r_pts = mp.params["xyz"].detach().clone()
    r_rots = torch.nn.functional.normalize(mp.params["rotation"].detach())
    r_flags = flags.clone()
    r_unstable = torch.abs(mp.stats.exist_since_iter - 17) < 15
    n_ref = ref.scale_and_transform_then_mark_visible(r_pts, r_rots, r_flags, r_unstable, diff_t, cam.viewmatrix, cam.projmatrix, 2, 1.03)
    lf_m = mp.optimizer.state[mp.params["lang_feat"]]["exp_avg"].clone()
    assert mp.optimizer.state[mp.params["xyz"]]["exp_avg"].any()
    n = mp.scaled_transform_visible_points_of_keyframe(flags, diff_t, cam.viewmatrix, cam.projmatrix, 17, 15, 2, 1.03)
    assert n == n_ref and 2 < n < P + 2
    assert torch.equal(flags, r_flags)
    assert torch.equal(mp.params["xyz"].detach(), r_pts) and torch.equal(mp.params["rotation"].detach(), r_rots)
    for k in ("xyz", "rotation"):
        st = mp.optimizer.state[mp.params[k]]
        assert st["step"] == 3 and not st["exp_avg"].any() and not st["exp_avg_sq"].any()
    st = mp.optimizer.state[mp.params["lang_feat"]]
    assert st["step"] == 3 and torch.equal(st["exp_avg"], lf_m)
    assert torch.isfinite(mp.train_step(win)) and mp.optimizer.state[mp.params["xyz"]]["step"] == 4
    plain = M.Mapper(sc, sh_degree=3)
    with pytest.raises(ValueError, match="track_densify_stats"):
        plain.scaled_transform_visible_points_of_keyframe(flags, diff_t, cam.viewmatrix, cam.projmatrix, 17, 15)


@pytest.mark.gpu
def test_create_from_pcd_and_scaled_transformation(dev):
    """GaussianModel::createFromPcd (reference src/gaussian_model.cpp:109-194) against the same composition of the
    reference's own pieces -- RGB2SH and inverse_sigmoid from its unmodified headers (oracle/_ref/ref_utils.so), simple-knn
    (ref_simple_knn.so) -- and applyScaledTransformation (:387-420) against transformPoints of the unmodified
    src/operate_points.cu (ref_geometry.so); the created model trains."""
    import build_ref
    from leg_slam_b200 import densify as D, mapper as M, synthetic
    try:
        ru, knn, geo = build_ref.load_utils(), build_ref.load_knn(), build_ref.load_geometry()
    except FileNotFoundError as ex:
        pytest.skip(str(ex))
    g = torch.Generator().manual_seed(91)
    n = 4000
    pts = (torch.rand(n, 3, generator=g) * torch.tensor([6.0, 4.0, 2.8])).to(dev)
    cols = torch.rand(n, 3, generator=g).to(dev)
    lfs = torch.randn(n, 64, generator=g).to(dev)
    params, stats = D.create_from_pcd(pts, cols, lfs, sh_degree=3)
    assert [tuple(params[k].shape) for k in M.PARAM_ORDER] == [(n, 3), (n, 1, 3), (n, 15, 3), (n, 64), (n, 1), (n, 3), (n, 4)]
    assert torch.equal(params["xyz"], pts) and params["xyz"].data_ptr() != pts.data_ptr()
    assert torch.equal(params["features_dc"][:, 0], ru.RGB2SH(cols)) and not params["features_rest"].any()
    assert torch.equal(params["lang_feat"], lfs)
    assert torch.equal(params["opacity"], ru.inverse_sigmoid(torch.full((n, 1), 0.1, device=dev)))
    d2 = torch.zeros(n, device=dev)
    torch.cuda.synchronize()
    assert knn.ref_simple_knn(n, pts.contiguous().data_ptr(), d2.data_ptr()) == 0
    want = torch.log(torch.sqrt(torch.clamp_min(d2, 0.0000001)))
    assert torch.equal(params["scaling"], want[:, None].repeat(1, 3))
    assert torch.equal(params["rotation"], torch.tensor([1.0, 0, 0, 0], device=dev).repeat(n, 1))
    assert not stats.exist_since_iter.any() and stats.denom.shape == (n, 1)
    nolf, _ = D.create_from_pcd(pts, cols, None, sh_degree=3)
    assert not nolf["lang_feat"].any()
    with pytest.raises(ValueError, match="num_points, 3"):
        D.create_from_pcd(pts, cols[:, :2])
    with pytest.raises(Exception, match="no CPU path"):
        D.create_from_pcd(pts.cpu(), cols.cpu())
    # the created model maps; then the map-wide scaled transformation
    W, H = 96, 64
    cams = [c.to(dev) for c in synthetic.make_cameras(2, W, H, seed=92)]
    win = [M.Keyframe(c, torch.rand(3, H, W, generator=g).to(dev), torch.randn(64, 37, 37, generator=g).to(dev),
                      (torch.rand(1, H, W, generator=g) * 3).to(dev)) for c in cams]
    params["xyz"] = params["xyz"] - torch.tensor([3.0, 2.0, 1.4], device=dev)  # the synthetic cameras orbit the origin
    mp = M.Mapper(params, sh_degree=3, track_densify_stats=True)
    for _ in range(2):
        assert torch.isfinite(mp.train_step(win))
    import test_ingest as TI
    T = TI._pose(g).to(dev)
    xyz0, sc0 = mp.params["xyz"].detach().clone(), mp.params["scaling"].detach().clone()
    op_m = mp.optimizer.state[mp.params["opacity"]]["exp_avg"].clone()
    mp.apply_scaled_transformation(1.25, T)
    assert torch.equal(mp.params["xyz"].detach(), geo.transform_points(xyz0 * 1.25, T))
    assert torch.equal(mp.params["scaling"].detach(), sc0 * 1.25)
    for k in ("xyz", "scaling"):
        st = mp.optimizer.state[mp.params[k]]
        assert st["step"] == 2 and not st["exp_avg"].any() and not st["exp_avg_sq"].any()
    assert torch.equal(mp.optimizer.state[mp.params["opacity"]]["exp_avg"], op_m)
    assert torch.isfinite(mp.train_step(win))
