"""GPU tests of the C++ host layers above the C ABI: the L2 autograd layer (include/gaussian_rasterizer.h, _L2.so) against
its Python twin, and the C++ LgsFusedAdam (include/lgs_adam.h) against torch.optim.Adam.  (Their build, argument validation
and CPU refusal are covered by tests/test_host_cpp.py on CPU.)  First run on a B200 in round 2:
profiles/r02_unverified_tests_first_run.log."""

import pytest
import torch

import cases

pytestmark = pytest.mark.gpu


def test_l2_cpp_autograd_equals_python_wrapper():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from leg_slam_b200 import GaussianRasterizationSettings, GaussianRasterizer, build_host
    build_host.build()
    from leg_slam_b200 import _L2
    dev = torch.device("cuda:0")
    cs = cases.make_case("sh3_lf", dev)
    e = torch.empty(0, device=dev)
    g = torch.Generator().manual_seed(5)
    gc, gl, gd = (torch.randn(c, cs["H"], cs["W"], generator=g).to(dev) for c in (3, 64, 1))
    names = ("means3D", "opacities", "shs", "lang_feats", "scales", "rotations")

    def leaves():
        d = {k: cs[k].detach().clone().requires_grad_(True) for k in names}
        d["means2D"] = torch.zeros_like(cs["means3D"]).requires_grad_(True)
        return d

    a = leaves()
    rs = _L2.GaussianRasterizationSettings(cs["H"], cs["W"], cs["tanfovx"], cs["tanfovy"], cs["bg"], 1.0, cs["viewmatrix"],
                                           cs["projmatrix"], cs["degree"], cs["campos"], False, True)
    out_a = _L2.GaussianRasterizer(rs).forward(a["means3D"], a["means2D"], a["opacities"], True, False, True, True, True, False,
                                               a["shs"], e, a["lang_feats"], a["scales"], a["rotations"], e)
    ((out_a[0] * gc).sum() + (out_a[1] * gl).sum() + (out_a[2] * gd).sum()).backward()

    b = leaves()
    ps = GaussianRasterizationSettings(cs["H"], cs["W"], cs["tanfovx"], cs["tanfovy"], cs["bg"], 1.0, cs["viewmatrix"],
                                       cs["projmatrix"], cs["degree"], cs["campos"], False, True)
    out_b = GaussianRasterizer(ps)(b["means3D"], b["means2D"], b["opacities"], shs=b["shs"], lang_feats=b["lang_feats"],
                                   scales=b["scales"], rotations=b["rotations"])
    ((out_b[0] * gc).sum() + (out_b[1] * gl).sum() + (out_b[2] * gd).sum()).backward()

    for x, y in zip(out_a, out_b):
        assert torch.equal(x, y)
    assert not out_a[3].requires_grad
    for k in names + ("means2D",):
        assert a[k].grad is not None and a[k].grad.shape == b[k].grad.shape, k
        assert cases.rel_err(a[k].grad.cpu().numpy(), b[k].grad.cpu().numpy()) <= 1e-3, k


def test_cpp_fused_adam_equals_torch_adam():
    """LgsFusedAdam over the reference's group layout (one tensor per group, its learning rates, eps 1e-15) against
    torch.optim.Adam: parameters <= 1e-6 relative after 3 steps, moments too, libtorch's step counts kept."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from leg_slam_b200 import build_host
    build_host.build()
    from leg_slam_b200 import _L2
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(12)
    shapes = [(1000, 3), (1000, 1, 3), (1000, 15, 3), (1000, 64), (1000, 1), (1000, 3), (1000, 4)]
    lrs = [3.2e-4, 2.5e-3, 1.25e-4, 1.5e-3, 0.05, 5e-3, 1e-3]
    p0 = [torch.randn(s, generator=g) for s in shapes]
    grads = [[torch.randn(s, generator=g) * 1e-2 for s in shapes] for _ in range(3)]
    ours = [t.clone().to(dev) for t in p0]
    out = _L2.fused_adam_run(ours, [[t.to(dev) for t in gs] for gs in grads], lrs, 1e-15)
    ref = [torch.nn.Parameter(t.clone().to(dev)) for t in p0]
    opt = torch.optim.Adam([dict(params=[t], lr=lr) for t, lr in zip(ref, lrs)], lr=0.0, eps=1e-15)
    for gs in grads:
        for t, gr in zip(ref, gs):
            t.grad = gr.to(dev)
        opt.step()
    n = len(shapes)
    assert out[-1].tolist() == [3] * n
    for i in range(n):
        assert cases.rel_err(ours[i].detach().cpu().numpy(), ref[i].detach().cpu().numpy()) <= 1e-6, i
        assert cases.rel_err(out[i].cpu().numpy(), opt.state[ref[i]]["exp_avg"].cpu().numpy()) <= 1e-6, i
        assert cases.rel_err(out[n + i].cpu().numpy(), opt.state[ref[i]]["exp_avg_sq"].cpu().numpy()) <= 1e-6, i


def test_second_device_in_one_process():
    """Kernel attributes (dynamic shared memory opt-in, SM count) belong to a device: a forward + backward + query on cuda:1
    after the same on cuda:0, in ONE process, must work and agree (the one-time configuration used to be cached in
    process-wide statics and applied to the first device only)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from leg_slam_b200 import cosine_query, rasterize_points as rp
    outs = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        cs = cases.make_case("sh3_lf", dev)
        R, color, lf, depth, radii, geom, binning, img = rp.rasterize_gaussians(*cases.fwd_args(cs))
        grads = rp.rasterize_gaussians_backward(*cases.bwd_args(cs, radii, geom, R, binning, img))
        feats, text = cases.cosine_case()
        sim = cosine_query(feats.to(dev), text.to(dev))
        torch.cuda.synchronize(dev)
        outs.append((R, color.cpu(), lf.cpu(), radii.cpu(), [g.cpu() for g in grads], sim.cpu()))
    a, b = outs
    assert a[0] == b[0] and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
    assert torch.equal(a[5], b[5])
    for x, y in zip(a[4], b[4]):
        assert cases.rel_err(x.numpy(), y.numpy()) <= 1e-4  # atomics reorder between runs


def test_cpp_geometry_operators_equal_the_compiled_reference():
    """The libtorch geometry operators with the reference's signatures (include/operate_points.h, stereo_vision.h, spatial.h;
    csrc/host/geometry_ops.cpp through `_C`) against the unmodified reference operators (oracle/_ref/ref_geometry.so,
    ref_simple_knn.so), call for call: rebinding, in-place updates, counts and outputs bit-identical."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import build_ref
    import test_ingest as TI
    from leg_slam_b200 import build_host
    build_host.build()
    from leg_slam_b200 import _C
    try:
        ref = build_ref.load_geometry()
        knn = build_ref.load_knn()
    except FileNotFoundError as ex:
        pytest.skip(str(ex))
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(21)
    # depth image -> camera points -> world points
    W, H = 320, 240
    depth = (torch.rand(W * H, generator=g) * 4 + 0.2).to(dev)
    mask = (torch.rand(W * H, generator=g) > 0.25).to(dev)
    intr = [300.0, 301.0, 159.5, 119.5]
    pa, pb = _C.reproject_depth_pinhole(depth, mask, intr, W), ref.reproject_depth_pinhole(depth, mask, intr, W)
    assert torch.equal(pa, pb)
    T = TI._pose(g).to(dev)
    wa, wb = _C.transform_points(pa, T), ref.transform_points(pb.clone(), T)
    assert torch.equal(wa, wb) and wa.data_ptr() != pa.data_ptr()
    # k-NN scale initialisation
    sub = wa[mask][:20000].contiguous()
    da = _C.dist_cuda2(sub)
    db = torch.zeros(sub.shape[0], device=dev)
    torch.cuda.synchronize()
    assert knn.ref_simple_knn(sub.shape[0], sub.data_ptr(), db.data_ptr()) == 0
    assert torch.equal(da, db)
    # loop-closure correction, in place, on contiguous tensors and on a strided view (staged and copied back)
    pts, rots, nt, un = TI._loop_closure_case(30000, 77)
    view, proj = TI._pose(g, t=(0.1, 0.2, 0.5)).to(dev), torch.eye(4, device=dev)
    a = [t.clone().to(dev) for t in (pts, rots, nt, un)]
    b = [t.clone().to(dev) for t in (pts, rots, nt, un)]
    na = _C.scale_and_transform_then_mark_visible(a[0], a[1], a[2], a[3], T, view, proj, 3, 1.05)
    nb = ref.scale_and_transform_then_mark_visible(b[0], b[1], b[2], b[3], T, view, proj, 3, 1.05)
    assert na == nb and 3 < na < 30003
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    wide = torch.zeros(30000, 6, device=dev)
    wide[:, :3] = pts.to(dev)
    c = [wide[:, :3], rots.clone().to(dev), nt.clone().to(dev), un.clone().to(dev)]
    assert not c[0].is_contiguous()
    nc = _C.scale_and_transform_then_mark_visible(c[0], c[1], c[2], c[3], T, view, proj, 3, 1.05)
    assert nc == nb and torch.equal(wide[:, :3], b[0]) and torch.equal(c[1], b[1]) and torch.equal(c[2], b[2])
    assert float(wide[:, 3:].abs().max()) == 0.0
    # inactive-geometry densification
    px, has, p3, colors = [t.to(dev) for t in TI._keypoint_case(3000, 640, 480, 5)]
    kin = [600.0, 600.0, 319.5, 239.5]
    ra = _C.inactive_geo_densify(px, has, p3, colors, 400.0, kin, 640)
    rb = ref.inactive_geo_densify(px, has, p3, colors, 400.0, kin, 640)
    assert 0 < ra[0].shape[0] < 3000
    assert torch.equal(ra[0], rb[0]) and torch.equal(ra[1], rb[1])
    e = _C.inactive_geo_densify(px[:0], has[:0], p3[:0], colors, 400.0, kin, 640)
    assert e[0] is None and e[1] is None  # undefined tensors, like the reference


def test_cpp_gaussian_model_equals_the_python_mapper_side():
    """GaussianModel in C++ (include/gaussian_model.h, _L2.so) against the Python statements of the same reference methods
    (leg_slam_b200.densify / ingest / optim, each already held to the reference), on the same tensors: createFromPcd, Adam
    steps through optimizer_ (LgsFusedAdam), addDensificationStats, densifyAndPrune (same generator state), increasePcd,
    resetOpacity, the loop-closure correction (against the unmodified reference operator) and applyScaledTransformation."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import test_ingest as TI
    from leg_slam_b200 import FusedAdam, build_host, densify as D, ingest
    build_host.build()
    from leg_slam_b200 import _L2
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(31)
    n = 6000
    pts = ((torch.rand(n, 3, generator=gen) - 0.5) * torch.tensor([6.0, 4.0, 2.8])).to(dev)
    cols, lfs = torch.rand(n, 3, generator=gen).to(dev), torch.randn(n, 64, generator=gen).to(dev)
    names = ("xyz_", "features_dc_", "features_rest_", "language_features_", "opacity_", "scaling_", "rotation_")
    cur = lambda g: [getattr(g, k).detach() for k in names]  # noqa: E731

    g = _L2.GaussianModel(3)
    g.createFromPcd(pts, cols, lfs, 2.0)
    py, stats = D.create_from_pcd(pts, cols, lfs, sh_degree=3)
    for a, k in zip(cur(g), D.PARAM_ORDER):
        assert torch.equal(a, py[k]), k
    assert not g.exist_since_iter_.any() and g.max_radii2D_.shape == (n,)
    args = _L2.GaussianOptimizationParams()
    g.trainingSetup(args)
    assert g.params_are_the_optimizers() and abs(g.learning_rate(0) - 3.2e-4) < 1e-9
    # three optimizer steps on seeded gradients: LgsFusedAdam behind optimizer_ vs the Python FusedAdam, same kernel
    ps = [torch.nn.Parameter(py[k].clone()) for k in D.PARAM_ORDER]
    opt = FusedAdam([dict(params=[p], lr=g.learning_rate(i)) for i, p in enumerate(ps)], lr=0.0, eps=1e-15)
    for s in range(3):
        grads = [(torch.randn(p.shape, generator=gen) * 0.01).to(dev) for p in ps]
        g.step([x.clone() for x in grads])
        for p, x in zip(ps, grads):
            p.grad = x
        opt.step()
    for i, (a, p) in enumerate(zip(cur(g), ps)):
        st = g.adam_state(i)
        assert torch.equal(a, p.detach()) and st[0] == 3, i
        assert torch.equal(st[1], opt.state[p]["exp_avg"]) and torch.equal(st[2], opt.state[p]["exp_avg_sq"]), i
    # statistics of two "views"
    for s in range(2):
        vs = torch.zeros(n, 3, device=dev, requires_grad=True)
        vs.grad = (torch.randn(n, 3, generator=gen) * 4e-4).to(dev)
        filt = (torch.rand(n, generator=gen) > 0.4).to(dev)
        g.addDensificationStats(vs, filt)
        stats.xyz_gradient_accum[filt] += vs.grad[filt, :2].norm(dim=-1, keepdim=True)
        stats.denom[filt] += 1
    assert torch.equal(g.xyz_gradient_accum_, stats.xyz_gradient_accum) and torch.equal(g.denom_, stats.denom)
    g.exist_since_iter_ = torch.randint(0, 30, (n,), generator=gen).to(torch.int32).to(dev)
    stats.exist_since_iter = g.exist_since_iter_.clone()
    # densifyAndPrune: same classification, same gather, same normal draws
    p_now = {k: a.clone() for k, a in zip(D.PARAM_ORDER, cur(g))}
    m_now = {k: g.adam_state(i)[1].clone() for i, k in enumerate(D.PARAM_ORDER)}
    v_now = {k: g.adam_state(i)[2].clone() for i, k in enumerate(D.PARAM_ORDER)}
    torch.manual_seed(77)
    g.densifyAndPrune(6e-4, 0.1, 20.0, 20)
    torch.manual_seed(77)
    p2, m2, v2, st2, info = D.densify_and_prune(p_now, m_now, v_now, stats, 6e-4, 0.1, 20.0, 20, percent_dense=g.percentDense())
    assert info["cloned"] > 0 and info["split_selected"] > 0 and 0 < info["kept"] < n and info["new_P"] != n, info
    for i, (a, k) in enumerate(zip(cur(g), D.PARAM_ORDER)):
        st = g.adam_state(i)
        assert torch.equal(a, p2[k]) and torch.equal(st[1], m2[k]) and torch.equal(st[2], v2[k]) and st[0] == 3, k
    assert torch.equal(g.exist_since_iter_, st2.exist_since_iter) and g.params_are_the_optimizers()
    P1 = info["new_P"]
    assert g.denom_.shape == (P1, 1) and not g.denom_.any() and not g.xyz_gradient_accum_.any() and g.max_radii2D_.shape == (P1,)
    # a keyframe's new points
    newp = ((torch.rand(500, 3, generator=gen) - 0.5) * 3).to(dev)
    newc = torch.rand(500, 3, generator=gen).to(dev)
    g.increasePcd(newp, newc, 42)
    p3, m3, v3, st3 = D.increase_pcd(p2, m2, v2, st2, newp, newc, 42)
    for i, (a, k) in enumerate(zip(cur(g), D.PARAM_ORDER)):
        st = g.adam_state(i)
        assert torch.equal(a, p3[k]) and torch.equal(st[1], m3[k]) and torch.equal(st[2], v3[k]) and st[0] == 3, k
    assert torch.equal(g.exist_since_iter_, st3.exist_since_iter) and (g.exist_since_iter_[P1:] == 42).all()
    # resetOpacity
    D.reset_opacity(p3, m3, v3)
    g.resetOpacity()
    st = g.adam_state(4)
    assert torch.equal(g.opacity_.detach(), p3["opacity"]) and st[0] == 3 and not st[1].any() and not st[2].any()
    # loop closure: the reference's sequence with its unmodified operator on clones
    import build_ref
    try:
        ref = build_ref.load_geometry()
    except FileNotFoundError as ex:
        pytest.skip(str(ex))
    P2 = g.xyz_.shape[0]
    flags = (torch.rand(P2, generator=gen) > 0.2).to(dev)
    T = TI._pose(gen).to(dev)
    view, proj = TI._pose(gen, t=(0.1, 0.2, 0.5)).to(dev), torch.eye(4, device=dev)
    r_pts, r_rots, r_flags = g.xyz_.detach().clone(), torch.nn.functional.normalize(g.rotation_.detach()), flags.clone()
    r_un = torch.abs(g.exist_since_iter_ - 17) < 15
    n_ref = ref.scale_and_transform_then_mark_visible(r_pts, r_rots, r_flags, r_un, T, view, proj, 4, 1.02)
    sc_before = g.adam_state(5)[1].clone()
    assert g.scaledTransformVisiblePointsOfKeyframe(flags, T, view, proj, 17, 15, 4, 1.02) == n_ref and 4 < n_ref < P2 + 4
    assert torch.equal(flags, r_flags) and torch.equal(g.xyz_.detach(), r_pts) and torch.equal(g.rotation_.detach(), r_rots)
    for i in (0, 6):
        st = g.adam_state(i)
        assert st[0] == 3 and not st[1].any() and not st[2].any()
    assert torch.equal(g.adam_state(5)[1], sc_before) and g.params_are_the_optimizers()
    # map-wide scaled transformation
    xyz0, sc0 = g.xyz_.detach().clone(), g.scaling_.detach().clone()
    g.applyScaledTransformation(1.25, T)
    assert torch.equal(g.xyz_.detach(), ingest.transformPoints(xyz0 * 1.25, T)) and torch.equal(g.scaling_.detach(), sc0 * 1.25)
    assert not g.adam_state(5)[1].any() and g.adam_state(5)[0] == 3 and g.params_are_the_optimizers()
    g.step([torch.zeros_like(a) for a in cur(g)])
    assert g.adam_state(0)[0] == 4


def _cpp_model_from_scene(_L2, sc, lrs=None):
    g = _L2.GaussianModel(3)
    for name, k in (("xyz_", "xyz"), ("features_dc_", "features_dc"), ("features_rest_", "features_rest"),
                    ("language_features_", "lang_feat"), ("opacity_", "opacity"), ("scaling_", "scaling"), ("rotation_", "rotation")):
        setattr(g, name, sc[k].detach().clone().contiguous().requires_grad_())
    P = sc["xyz"].shape[0]
    g.exist_since_iter_ = torch.zeros(P, dtype=torch.int32, device=sc["xyz"].device)
    g.max_radii2D_ = torch.zeros(P, device=sc["xyz"].device)
    g.spatial_lr_scale_ = 1.0
    a = _L2.GaussianOptimizationParams()
    a.position_lr_init_ = 3.2e-4  # leg_slam_b200.mapper.DEFAULT_LRS (the reference's Replica configuration)
    g.trainingSetup(a)
    g.setShDegree(3)
    return g


def _cpp_keyframe(_L2, kf):
    import math
    k = _L2.GaussianKeyframe()
    c = kf.camera
    k.FoVx_, k.FoVy_ = 2.0 * math.atan(c.tanfovx), 2.0 * math.atan(c.tanfovy)
    k.image_height_, k.image_width_ = c.height, c.width
    k.world_view_transform_, k.full_proj_transform_, k.camera_center_ = c.viewmatrix, c.projmatrix, c.campos
    k.language_features_ = kf.gt_lf
    return k


def test_cpp_renderer_and_mapping_iteration():
    """GaussianRenderer::render in C++ (include/gaussian_renderer.h) against the Python statement of the same function on the
    same model -- all input selections (SHs / SH -> RGB outside / override colour; scaling + rotation / precomputed covariance)
    -- and three mapping iterations through mappingIterationBackward + mappingIterationStep (render, fused loss, autograd
    through the C++ rasterizer node, densification statistics, LgsFusedAdam) against the Python mapper's fused path."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import numpy as np
    from leg_slam_b200 import build_host, mapper as M, renderer as R, synthetic
    build_host.build()
    from leg_slam_b200 import _L2
    dev = torch.device("cuda:0")
    W, H = 96, 64
    sc = synthetic.make_scene(4000, seed=51, mean_scale=0.06, device=dev)
    cams = synthetic.make_cameras(2, W, H, seed=51)
    gen = torch.Generator().manual_seed(52)
    win = [M.Keyframe(c.to(dev), torch.rand(3, H, W, generator=gen).to(dev), torch.randn(64, 37, 37, generator=gen).to(dev),
                      (torch.rand(1, H, W, generator=gen) * 3).to(dev)) for c in cams]
    g = _cpp_model_from_scene(_L2, sc)
    kfs = [_cpp_keyframe(_L2, kf) for kf in win]
    bg = torch.zeros(3, device=dev)
    none = torch.empty(0, device=dev)
    pyview = R.GaussianModelView({k: sc[k] for k in M.PARAM_ORDER}, sh_degree=3)
    override = torch.rand(4000, 3, generator=gen).to(dev)
    with torch.no_grad():
        for conv, cov, use_override in ((False, False, False), (True, False, False), (False, True, False), (False, False, True)):
            a = _L2.render(kfs[0], H, W, g, _L2.GaussianPipelineParams(conv, cov), bg, override if use_override else none, 1.0,
                           use_override, True)
            b = R.GaussianRenderer.render(R.KeyframeView(win[0].camera), H, W, pyview, R.GaussianPipelineParams(conv, cov), bg,
                                          override if use_override else None, 1.0, use_override, True)
            assert torch.equal(a[5], b[5]) and torch.equal(a[4], b[4]), (conv, cov, use_override)
            for x, y in zip(a[:3], b[:3]):
                assert float((x - y).abs().max()) <= 1e-5 * max(1.0, float(y.abs().max())), (conv, cov, use_override)
            assert a[3].shape == (4000, 3) and not a[3].any()
    # three iterations, one keyframe each
    ours = M.Mapper(sc, sh_degree=3, track_densify_stats=True)
    pipe = _L2.GaussianPipelineParams()
    mask = torch.empty(0, device=dev)
    for it in range(3):
        kf = win[it % 2]
        l_cpp = _L2.mapping_iteration_backward(g, kfs[it % 2], pipe, bg, kf.gt_image, kf.gt_depth, mask, 0.2, True)
        _L2.mapping_iteration_step(g)
        l_py = ours.train_step([kf])
        assert abs(float(l_cpp) - float(l_py)) <= 1e-4 * abs(float(l_py)), (it, float(l_cpp), float(l_py))
    assert g.xyz_.grad is None  # zero_grad(true)
    cur = [getattr(g, n).detach() for n in ("xyz_", "features_dc_", "features_rest_", "language_features_", "opacity_", "scaling_",
                                            "rotation_")]
    for a, k in zip(cur, M.PARAM_ORDER):
        d = np.abs(a.cpu().numpy() - ours.params[k].detach().cpu().numpy())
        step = 3 * M.DEFAULT_LRS[k]
        assert (d > 0.05 * step).mean() <= 2e-3, (k, float(d.max()), step)
        assert g.adam_state(M.PARAM_ORDER.index(k))[0] == 3
    assert torch.equal(g.denom_, ours.stats.denom) and torch.equal(g.max_radii2D_, ours.stats.max_radii2D)
    ga, pa = g.xyz_gradient_accum_.cpu().numpy(), ours.stats.xyz_gradient_accum.cpu().numpy()
    assert np.abs(ga - pa).max() <= 2e-3 * np.abs(pa).max()


def test_cpp_gaussian_model_ply_checkpoints(tmp_path):
    """GaussianModel::savePly / loadPly in C++ against leg_slam_b200.ply_io (whose files are byte-identical to what the
    reference's own tinyply writes, tests/test_ply_io.py): same bytes with and without the optimizer state, and a model
    loaded from either file holds the tensors, the Adam moments and the step counts that went in."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from leg_slam_b200 import build_host, mapper as M, ply_io, synthetic
    build_host.build()
    from leg_slam_b200 import _L2
    dev = torch.device("cuda:0")
    sc = synthetic.make_scene(3001, seed=5, mean_scale=0.06, device=dev)
    g = _cpp_model_from_scene(_L2, sc)
    names = ("xyz_", "features_dc_", "features_rest_", "language_features_", "opacity_", "scaling_", "rotation_")
    gen = torch.Generator().manual_seed(6)
    for _ in range(2):
        g.step([(torch.randn(getattr(g, n).shape, generator=gen) * 0.01).to(dev) for n in names])
    # copies: the optimizer updates the model's tensors and moments in place further down
    params = {k: getattr(g, n).detach().clone() for k, n in zip(M.PARAM_ORDER, names)}
    m = {k: g.adam_state(i)[1].clone() for i, k in enumerate(M.PARAM_ORDER)}
    v = {k: g.adam_state(i)[2].clone() for i, k in enumerate(M.PARAM_ORDER)}
    a, b, c, d = (str(tmp_path / n) for n in ("cpp_state.ply", "py_state.ply", "cpp_plain.ply", "py_plain.ply"))
    g.savePly(a, True)
    ply_io.save_ply(b, params, m, v, {k: 2 for k in M.PARAM_ORDER})
    g.savePly(c, False)
    ply_io.save_ply(d, params)
    assert open(a, "rb").read() == open(b, "rb").read()
    assert open(c, "rb").read() == open(d, "rb").read()
    fresh = _L2.GaussianModel(3)
    fresh.loadPly(c)  # no optimizer yet: plain leaves
    for k, n in zip(M.PARAM_ORDER, names):
        t = getattr(fresh, n)
        assert torch.equal(t.detach(), params[k]) and t.requires_grad, k
    assert fresh.active_sh_degree_ == 3 and fresh.denom_.shape == (3001, 1) and fresh.exist_since_iter_.shape == (3001,)
    resumed = _cpp_model_from_scene(_L2, {k: torch.zeros_like(t) for k, t in sc.items() if k in M.PARAM_ORDER})
    resumed.loadPly(b)  # written by the Python side, with the optimizer state
    assert resumed.params_are_the_optimizers()
    for i, (k, n) in enumerate(zip(M.PARAM_ORDER, names)):
        st = resumed.adam_state(i)
        assert torch.equal(getattr(resumed, n).detach(), params[k]) and st[0] == 2, k
        assert torch.equal(st[1], m[k]) and torch.equal(st[2], v[k]), k
    # both continue identically
    grads = [(torch.randn(getattr(g, n).shape, generator=gen) * 0.01).to(dev) for n in names]
    g.step([x.clone() for x in grads])
    resumed.step([x.clone() for x in grads])
    for n in names:
        assert torch.equal(getattr(g, n).detach(), getattr(resumed, n).detach()), n
    p2, m2, v2, steps2 = ply_io.load_ply(a, dev, 3)  # and the Python side reads the C++ file
    assert all(torch.equal(p2[k], params[k]) and torch.equal(m2[k], m[k]) for k in M.PARAM_ORDER) and steps2["xyz"] == 2
    with pytest.raises(RuntimeError, match="cannot open"):
        fresh.loadPly(str(tmp_path / "missing.ply"))
