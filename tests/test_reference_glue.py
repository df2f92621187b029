"""The reference's own glue above L1 -- GaussianRasterizerFunction, GaussianRasterizer::forward / markVisibleGaussians
(src/gaussian_rasterizer.cpp:18-236) and GaussianRenderer::render (src/gaussian_renderer.cpp:24-160), compiled UNMODIFIED into
oracle/_ref/ref_model.so -- against the package's twins (leg_slam_b200/rasterizer.py, renderer.py), on CPU.

Both sides sit on the same L1: tests/oracle_l1.py runs the CPU oracle behind RasterizeGaussiansCUDA /
RasterizeGaussiansBackwardCUDA / markVisible and records every call.  Held: the two glues make the SAME calls (every argument of
the 20- and the 24-argument signatures, empty-tensor sentinels included), return the same tensors, and route the nine gradients
to the same leaves."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import build_ref  # noqa: E402
import cases  # noqa: E402,F401  (sys.path)
import oracle_l1  # noqa: E402
from leg_slam_b200 import rasterizer as RZ, renderer as RD, synthetic  # noqa: E402

P, W, H = 300, 48, 32
NAMES = ("xyz", "features_dc", "features_rest", "lang_feat", "opacity", "scaling", "rotation")


@pytest.fixture(scope="module")
def RM():
    try:
        if os.path.isdir(os.path.join(build_ref.MODEL_REF, "src")):
            build_ref.build_model(verbose=False)     # no-op when oracle/_ref/ref_model.so exists
        return build_ref.load_model()
    except FileNotFoundError as ex:
        pytest.skip(str(ex))


@pytest.fixture()
def both(RM, monkeypatch):
    """(reference-side L1 log, package-side L1 log): fresh recorders under both glues."""
    import oracle as O
    a, b = oracle_l1.RecordingL1(), oracle_l1.RecordingL1()
    RM.set_rasterizer(a.rasterize_gaussians, a.rasterize_gaussians_backward, a.mark_visible)
    monkeypatch.setattr(RZ, "_C", b)
    n = O.num_threads()
    O.lib().omp_set_num_threads(1)   # the oracle's backward accumulates with `omp atomic`: one thread = one summation order
    yield a, b
    O.lib().omp_set_num_threads(n)
    RM.set_rasterizer(None, None, None)


def scene():
    sc = synthetic.make_scene(P, seed=5, mean_scale=0.08)
    cam = synthetic.make_cameras(1, W, H, seed=5)[0]
    return sc, cam


def weights(seed=3):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(3, H, W, generator=g), torch.randn(64, H, W, generator=g), torch.randn(1, H, W, generator=g)


@pytest.mark.parametrize("convert_SHs,compute_cov3D,use_override,include_lf,active_deg,modifier,touch", [
    (False, False, False, True, 3, 1.0, "all"),      # the mapper's configuration (every shipped cfg: both pipe flags false)
    (False, False, False, False, 1, 1.0, "all"),     # no language features, a growing SH degree
    (False, True, False, True, 3, 1.0, "all"),       # precomputed 3D covariance
    (False, False, True, True, 3, 1.0, "all"),       # override colours (the query's heat-map render)
    (True, False, False, True, 2, 1.0, "all"),       # SH -> RGB on the host side
    (False, False, False, True, 3, 0.7, "all"),      # the viewer's scaling modifier
    (False, False, False, True, 3, 1.0, "colour"),   # a loss that touches one output: the other upstream gradients are zeros
])
def test_render_makes_the_same_l1_calls_and_routes_the_same_gradients(RM, both, convert_SHs, compute_cov3D, use_override,
                                                                      include_lf, active_deg, modifier, touch):
    la, lb = both
    sc, cam = scene()
    bg = torch.tensor([0.1, 0.2, 0.3])
    g = torch.Generator().manual_seed(11)
    override = torch.rand(P, 3, generator=g)
    wc, wl, wd = weights()
    kv = RD.KeyframeView(cam)

    def loss_of(img, lf, depth):
        if touch == "colour":
            return (img * wc).sum()
        return (img * wc).sum() + (lf * wl).sum() + (depth * wd).sum()

    # reference
    ref = RM.GaussianModel(3)
    ref.set_state([sc[k] for k in NAMES], torch.zeros(P, dtype=torch.int32), 1.0)
    ref.set_sh_degree(active_deg)
    ov_r = override.clone().requires_grad_() if use_override else torch.empty(0)
    out_r = RM.render(ref, kv.FoVx_, kv.FoVy_, cam.viewmatrix, cam.projmatrix, cam.campos, H, W, convert_SHs, compute_cov3D, bg,
                      ov_r, modifier, use_override, include_lf)
    loss_of(*out_r[:3]).backward()
    grads_r = [p.grad for p in ref.params()]

    # the package's twin
    params = {k: sc[k].detach().clone().requires_grad_() for k in NAMES}
    pc = RD.GaussianModelView(params, sh_degree=3, active_sh_degree=active_deg)
    ov_o = override.clone().requires_grad_() if use_override else None
    out_o = RD.GaussianRenderer.render(kv, H, W, pc, RD.GaussianPipelineParams(convert_SHs, compute_cov3D), bg, ov_o, modifier,
                                       use_override, include_lf)
    loss_of(*out_o[:3]).backward()

    exact = not convert_SHs and not compute_cov3D   # those two run a host-side formula first (eval_sh / L L^T): float rounding
    assert oracle_l1.same_calls(la.calls, lb.calls, rtol=0.0 if exact else 2e-5)
    assert [c[0] for c in la.calls] == ["rasterize_gaussians", "rasterize_gaussians_backward"]
    fwd = la.calls[0][1]
    # the sentinels: exactly one of sh / colours and of scales + rotations / cov3D reaches L1, the others are empty tensors
    assert (fwd[15].numel() == 0) == (use_override or convert_SHs) and (fwd[2].numel() == 0) != (fwd[15].numel() == 0)
    assert (fwd[8].numel() == 0) != compute_cov3D and (fwd[5].numel() == 0) == compute_cov3D == (fwd[6].numel() == 0)
    assert (fwd[3].numel() == 0) != include_lf and fwd[19] == include_lf and fwd[16] == active_deg and fwd[18] is False

    tol = 0.0 if exact else 2e-5
    def close(x, y, what):
        assert x.shape == y.shape and x.dtype == y.dtype, what
        if tol == 0.0:
            assert torch.equal(x, y), what
        else:
            assert float((x.detach() - y.detach()).abs().max()) <= tol * max(1e-12, float(y.detach().abs().max())), what

    for i, n in enumerate(("image", "lf", "depth")):
        close(out_r[i], out_o[i], n)
    assert torch.equal(out_r[4], out_o[4]) and out_r[4].dtype == torch.bool and torch.equal(out_r[4], out_r[5] > 0)
    assert torch.equal(out_r[5], out_o[5]) and out_r[5].dtype == torch.int32
    assert 0 < int(out_r[4].sum()) < P
    # the screen-space leaf receives dL_dmeans2D on both sides
    assert out_r[3].shape == (P, 3) and not out_r[3].any() and not out_o[3].any()
    close(out_r[3].grad, out_o[3].grad, "means2D grad")
    assert out_r[3].grad.abs().sum() > 0
    for k, gr in zip(NAMES, grads_r):
        go = params[k].grad
        used = not ((k in ("features_dc", "features_rest") and use_override) or (k == "lang_feat" and not include_lf))
        if not used:
            assert gr is None or not gr.any(), k
            assert go is None or not go.any(), k
            continue
        assert gr is not None and go is not None, k
        close(gr, go, k)
    if use_override:
        close(ov_r.grad, ov_o.grad, "override colour grad")
        assert ov_r.grad.abs().sum() > 0
    if active_deg < 3 and not use_override:   # coefficients above the active degree get exactly zero on both sides
        n_act = (active_deg + 1) ** 2 - 1
        assert not grads_r[2][:, n_act:].any() and not params["features_rest"].grad[:, n_act:].any()
        assert grads_r[2][:, :n_act].any()


def test_rasterizer_forward_validation_and_sentinels(RM, both):
    """GaussianRasterizer::forward's two refusals (reference :198-207, same messages in the package) and, for a valid call, the
    sentinel tensors it substitutes (:209-221)."""
    la, lb = both
    sc, cam = scene()
    bg = torch.zeros(3)
    e = torch.empty(0)
    xyz, op = sc["xyz"], torch.sigmoid(sc["opacity"])
    shs = torch.cat([sc["features_dc"], sc["features_rest"]], dim=1)
    scales, rots = torch.exp(sc["scaling"]), torch.nn.functional.normalize(sc["rotation"])
    cols = torch.rand(P, 3)
    cov = RD.GaussianModelView({k: sc[k] for k in NAMES}).getCovarianceActivation()
    rs_args = (H, W, cam.tanfovx, cam.tanfovy, bg, 1.0, cam.viewmatrix, cam.projmatrix, 3, cam.campos, False, False)
    rs = RZ.GaussianRasterizationSettings(*rs_args)
    ours = RZ.GaussianRasterizer(rs)
    means2D = torch.zeros_like(xyz)
    bad = [  # (has_shs, has_cols, has_scales, has_rots, has_cov) -> which message
        ((False, False, True, True, False), "excatly one of either SHs or precomputed colors"),
        ((True, True, True, True, False), "excatly one of either SHs or precomputed colors"),
        ((True, False, False, False, False), "exactly one of either scale/rotation pair or precomputed 3D covariance"),
        ((True, False, True, False, False), "exactly one of either scale/rotation pair or precomputed 3D covariance"),
        ((True, False, True, True, True), "exactly one of either scale/rotation pair or precomputed 3D covariance"),
        ((True, False, False, True, True), "exactly one of either scale/rotation pair or precomputed 3D covariance"),
    ]
    for (hs, hc, hsc, hr, hcov), msg in bad:
        with pytest.raises(RuntimeError, match=msg):
            RM.rasterizer_forward(*rs_args, xyz, means2D, op, hs, hc, False, hsc, hr, hcov, shs if hs else e, cols if hc else e, e,
                                  scales if hsc else e, rots if hr else e, cov if hcov else e)
        with pytest.raises(Exception, match=msg):
            ours(xyz, means2D, op, shs=shs if hs else None, colors_precomp=cols if hc else None,
                 scales=scales if hsc else None, rotations=rots if hr else None, cov3D_precomp=cov if hcov else None)
    assert not la.calls and not lb.calls          # refused before L1 on both sides
    for hs, hcov in ((True, False), (False, True)):
        r = RM.rasterizer_forward(*rs_args, xyz, means2D, op, hs, not hs, False, not hcov, not hcov, hcov, shs if hs else e,
                                  e if hs else cols, e, e if hcov else scales, e if hcov else rots, cov if hcov else e)
        o = ours(xyz, means2D, op, shs=shs if hs else None, colors_precomp=None if hs else cols,
                 scales=None if hcov else scales, rotations=None if hcov else rots, cov3D_precomp=cov if hcov else None)
        assert len(r) == len(o) == 4 and all(torch.equal(x, y) for x, y in zip(r, o))
    assert oracle_l1.same_calls(la.calls, lb.calls) and len(la.calls) == 2


def test_mark_visible_gaussians(RM, both):
    """GaussianRasterizer::markVisibleGaussians (reference :18-25): positions + the settings' two matrices go to L1's markVisible."""
    la, lb = both
    sc, cam = scene()
    bg = torch.zeros(3)
    r = RM.mark_visible_gaussians(cam.viewmatrix, cam.projmatrix, cam.campos, bg, sc["xyz"])
    rs = RZ.GaussianRasterizationSettings(1, 1, 1.0, 1.0, bg, 1.0, cam.viewmatrix, cam.projmatrix, 0, cam.campos, False, False)
    o = RZ.GaussianRasterizer(rs).markVisible(sc["xyz"])
    assert r.dtype == torch.bool and torch.equal(r, o) and 0 < int(r.sum()) < P
    assert oracle_l1.same_calls(la.calls, lb.calls) and [c[0] for c in la.calls] == ["mark_visible"]


def test_mapping_iterations_of_reference_classes_equal_mapper_train_step(RM, both):
    """Whole mapping iterations assembled ONLY from unmodified reference code, in trainForOneIteration's order (reference
    src/gaussian_mapper.cpp:662-796): GaussianModel::updateLearningRate -> GaussianRenderer::render (language features on) ->
    loss_utils chained as :707-721 (oracle/_ref/ref_loss.so) -> backward through GaussianRasterizerFunction -> max_radii2D /
    addDensificationStats -> the model's own libtorch Adam::step + zero_grad -- against Mapper.train_step's autograd path.
    Both arms render through the CPU oracle (the reference arm via the recording L1, the mapper via tests/oracle_autograd.py),
    so what is compared is everything AROUND the kernels: activations, argument routing, loss, gradient routing, optimizer."""
    import oracle_autograd
    from leg_slam_b200 import mapper as M
    try:
        ref_loss = build_ref.load_loss()
    except FileNotFoundError as ex:
        pytest.skip(str(ex))
    la, _ = both
    n_it = 3
    sc, cam = scene()
    g = torch.Generator().manual_seed(42)
    bg = torch.zeros(3)
    gt = dict(image=torch.rand(3, H, W, generator=g), lf=torch.randn(64, 37, 37, generator=g),
              depth=torch.rand(1, H, W, generator=g) * 3)
    mask = (torch.rand(1, H, W, generator=g) > 0.1).float().expand(3, H, W).contiguous()
    kv = RD.KeyframeView(cam)

    ref = RM.GaussianModel(3)
    ref.set_state([sc[k] for k in NAMES], torch.zeros(P, dtype=torch.int32), 2.0)
    ref.set_sh_degree(3)
    ref.training_setup(position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01, position_lr_max_steps=n_it,
                       feature_lr=0.0025, language_feature_lr=0.0015, opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001,
                       percent_dense=0.01)
    mp = M.Mapper({k: sc[k] for k in NAMES}, lrs=dict(zip(M.PARAM_ORDER, ref.lrs())), sh_degree=3, fused=False, use_cuda_graph=False,
                  optimizer_factory=lambda gr: torch.optim.Adam(gr, lr=0.0, eps=1e-15),
                  render_fn=oracle_autograd.make_render_fn(bg))
    f32 = lambda x: float(np.float32(x))  # noqa: E731
    mp.set_position_lr_schedule(f32(0.00016), f32(0.0000016), f32(0.01), n_it, spatial_lr_scale=2.0)
    kf = M.Keyframe(cam, gt["image"], gt["lf"], gt["depth"], mask)
    xyz_lr_sum = 0.0
    for it in range(n_it):
        lr = ref.update_learning_rate(it)
        assert mp.update_learning_rate(it) == lr
        xyz_lr_sum += lr
        image, lf, depth, viewspace, visible, radii = RM.render(ref, kv.FoVx_, kv.FoVy_, cam.viewmatrix, cam.projmatrix, cam.campos,
                                                                H, W, False, False, bg, torch.empty(0), 1.0, False, True)
        loss = ref_loss.mapping_loss(image, lf, depth, gt["image"], gt["lf"], gt["depth"], mask, 0.2)
        loss.backward()
        with torch.no_grad():
            mr = ref.max_radii2D
            mr[visible] = torch.max(mr[visible], radii[visible].to(mr.dtype))
            ref.max_radii2D = mr
            ref.add_densification_stats(viewspace.grad, visible)
            ref.step()
            ref.zero_grad()
        l_ours = mp.train_step([kf])
        assert abs(float(l_ours) - float(loss.detach())) <= 1e-5 * abs(float(loss.detach())), (it, float(l_ours), float(loss.detach()))
    assert [c[0] for c in la.calls] == ["rasterize_gaussians", "rasterize_gaussians_backward"] * n_it
    assert float(ref.denom.max()) == n_it and ref.xyz_gradient_accum.any()
    for k, r, lr in zip(M.PARAM_ORDER, ref.params(), ref.lrs()):
        # differences are measured against the size of the steps taken (Adam's first steps are sign-like: ~lr per element)
        step = xyz_lr_sum if k == "xyz" else n_it * lr
        d = (mp.params[k].detach() - r.detach()).abs()
        moved = (r.detach() - sc[k]).abs()
        assert float(moved.max()) > 0.5 * step / n_it, k               # the reference arm did train this tensor
        assert float(d.max()) <= 1e-3 * step, (k, float(d.max()), step)   # measured: <= 1.6e-5 of the steps taken
        st = ref.moments(M.PARAM_ORDER.index(k))
        assert st[0] == n_it and mp.optimizer.state[mp.params[k]]["step"] == n_it


REF_PY_PKG = "/root/reference/eval/submodules/diff-gaussian-rasterization-legs-slam/diff_gaussian_rasterization_legs_slam"


def _import_reference_python_wrapper(l1):
    """The reference's own Python wrapper (eval/submodules/.../diff_gaussian_rasterization_legs_slam/__init__.py), imported
    from where it lies with `l1` standing in for its compiled `_C` submodule."""
    import importlib.util
    import types
    name = "ref_dgr_legs_slam"
    for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
        del sys.modules[k]
    c = types.ModuleType(name + "._C")
    c.rasterize_gaussians, c.rasterize_gaussians_backward, c.mark_visible = (
        l1.rasterize_gaussians, l1.rasterize_gaussians_backward, l1.mark_visible)
    sys.modules[name + "._C"] = c
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_PY_PKG, "__init__.py"),
                                                  submodule_search_locations=[REF_PY_PKG])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not os.path.isdir(REF_PY_PKG), reason="reference tree not present (it does not travel to the GPU box)")
def test_python_wrapper_of_the_reference_makes_the_same_forward_calls(monkeypatch):
    """The eval package's GaussianRasterizer (reference eval/submodules/.../__init__.py:159-210), imported unmodified, next to
    leg_slam_b200.rasterizer.GaussianRasterizer on the same recording L1, the way the eval scripts use it (forward only, under
    no_grad -- SURVEY.md 8b: its backward passes 22 of the 24 arguments): same `_C.rasterize_gaussians` arguments for every
    input selection, same results, same refusals, same markVisible call."""
    import oracle as O
    la, lb = oracle_l1.RecordingL1(), oracle_l1.RecordingL1()
    ref = _import_reference_python_wrapper(la)
    monkeypatch.setattr(RZ, "_C", lb)
    n = O.num_threads()
    O.lib().omp_set_num_threads(1)
    try:
        sc, cam = scene()
        bg = torch.tensor([0.2, 0.1, 0.0])
        xyz, op = sc["xyz"], torch.sigmoid(sc["opacity"])
        shs = torch.cat([sc["features_dc"], sc["features_rest"]], dim=1)
        scales, rots = torch.exp(sc["scaling"]), torch.nn.functional.normalize(sc["rotation"])
        cols = torch.rand(P, 3, generator=torch.Generator().manual_seed(1))
        cov = RD.GaussianModelView({k: sc[k] for k in NAMES}).getCovarianceActivation()
        means2D = torch.zeros_like(xyz)
        assert ref.GaussianRasterizationSettings._fields == RZ.GaussianRasterizationSettings._fields
        selections = [dict(shs=shs, lang_feats=sc["lang_feat"], scales=scales, rotations=rots),
                      dict(shs=shs, scales=scales, rotations=rots),
                      dict(colors_precomp=cols, lang_feats=sc["lang_feat"], scales=scales, rotations=rots),   # heat-map render
                      dict(shs=shs, lang_feats=sc["lang_feat"], cov3D_precomp=cov)]
        with torch.no_grad():
            for sel in selections:
                args = (H, W, cam.tanfovx, cam.tanfovy, bg, 1.0, cam.viewmatrix, cam.projmatrix, 3, cam.campos, False, "lang_feats" in sel)
                r = ref.GaussianRasterizer(ref.GaussianRasterizationSettings(*args))(xyz, means2D, op, **sel)
                o = RZ.GaussianRasterizer(RZ.GaussianRasterizationSettings(*args))(xyz, means2D, op, **sel)
                assert len(r) == len(o) == 4 and all(torch.equal(x, y) for x, y in zip(r, o))
            vr = ref.GaussianRasterizer(ref.GaussianRasterizationSettings(*args)).markVisible(xyz)
            vo = RZ.GaussianRasterizer(RZ.GaussianRasterizationSettings(*args)).markVisible(xyz)
            assert torch.equal(vr, vo)
        assert oracle_l1.same_calls(la.calls, lb.calls)
        assert [c[0] for c in la.calls] == ["rasterize_gaussians"] * len(selections) + ["mark_visible"]
        for bad, msg in ((dict(scales=scales, rotations=rots), "excatly one of either SHs or precomputed colors"),
                         (dict(shs=shs, colors_precomp=cols, scales=scales, rotations=rots), "excatly one of either SHs"),
                         (dict(shs=shs, scales=scales), "exactly one of either scale/rotation pair or precomputed 3D covariance"),
                         (dict(shs=shs, scales=scales, rotations=rots, cov3D_precomp=cov), "exactly one of either scale/rotation pair")):
            for mod in (ref, RZ):
                with pytest.raises(Exception, match=msg):
                    mod.GaussianRasterizer(mod.GaussianRasterizationSettings(*args))(xyz, means2D, op, **bad)
        assert len(la.calls) == len(lb.calls) == len(selections) + 1
    finally:
        O.lib().omp_set_num_threads(n)


def test_train_for_one_iteration_lines_equal_mapper_train_step(RM, both):
    """The same comparison with the reference arm made of trainForOneIteration's OWN LINES: render -> loss -> backward
    (src/gaussian_mapper.cpp:686-724), density control (:737-761; statistics only here, the Gaussian set stays) and the
    optimizer step (:793-797), cut out of the file at build time and compiled as they stand (oracle/ref_model_wrap.cpp
    RefDensityControl::train_iteration) around the reference's model, renderer, glue and loss_utils, over the recording L1."""
    import oracle_autograd
    from leg_slam_b200 import mapper as M
    la, _ = both
    n_it = 3
    sc, cam = scene()
    g = torch.Generator().manual_seed(43)
    bg = torch.zeros(3)
    gt = dict(image=torch.rand(3, H, W, generator=g), lf=torch.randn(64, 37, 37, generator=g),
              depth=torch.rand(1, H, W, generator=g) * 3)
    mask = (torch.rand(1, H, W, generator=g) > 0.1).float().expand(3, H, W).contiguous()
    kv = RD.KeyframeView(cam)
    ref = RM.GaussianModel(3)
    ref.set_state([sc[k] for k in NAMES], torch.zeros(P, dtype=torch.int32), 2.0)
    ref.set_sh_degree(3)
    ref.training_setup(position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01, position_lr_max_steps=n_it,
                       feature_lr=0.0025, language_feature_lr=0.0015, opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001,
                       percent_dense=0.01)
    it_ref = RM.DensityControl(ref, iterations=1000, densification_interval=100, opacity_reset_interval=0, densify_from_iter=500,
                               densify_until_iter=1000, densify_grad_threshold=2e-4, densify_min_opacity=0.005,
                               prune_big_point_after_iter=0, white_background=False, cameras_extent=4.0)
    it_ref.set_lambda_dssim(0.2)
    it_ref.set_background(bg)
    mp = M.Mapper({k: sc[k] for k in NAMES}, lrs=dict(zip(M.PARAM_ORDER, ref.lrs())), sh_degree=3, fused=False, use_cuda_graph=False,
                  optimizer_factory=lambda gr: torch.optim.Adam(gr, lr=0.0, eps=1e-15),
                  render_fn=oracle_autograd.make_render_fn(bg))
    f32 = lambda x: float(np.float32(x))  # noqa: E731
    mp.set_position_lr_schedule(f32(0.00016), f32(0.0000016), f32(0.01), n_it, spatial_lr_scale=2.0)
    kf = M.Keyframe(cam, gt["image"], gt["lf"], gt["depth"], mask)
    xyz_lr_sum = 0.0
    for it in range(1, n_it + 1):
        lr = ref.update_learning_rate(it)
        assert mp.update_learning_rate(it) == lr
        xyz_lr_sum += lr
        loss, radii, visible = it_ref.train_iteration(it, kv.FoVx_, kv.FoVy_, cam.viewmatrix, cam.projmatrix, cam.campos, gt["lf"], H, W,
                                                      gt["image"], gt["depth"], mask)
        l_ours = mp.train_step([kf])
        assert abs(float(l_ours) - float(loss)) <= 1e-5 * abs(float(loss)), (it, float(l_ours), float(loss))
        assert torch.equal(visible, radii > 0) and 0 < int(visible.sum()) < P
    assert [c[0] for c in la.calls] == ["rasterize_gaussians", "rasterize_gaussians_backward"] * n_it
    assert float(ref.denom.max()) == n_it and ref.xyz_gradient_accum.any() and ref.max_radii2D.any()
    for k, r, lr in zip(M.PARAM_ORDER, ref.params(), ref.lrs()):
        step = xyz_lr_sum if k == "xyz" else n_it * lr
        d = (mp.params[k].detach() - r.detach()).abs()
        assert float((r.detach() - sc[k]).abs().max()) > 0.5 * step / n_it, k
        assert float(d.max()) <= 1e-3 * step, (k, float(d.max()), step)
        assert ref.moments(M.PARAM_ORDER.index(k))[0] == n_it and r.grad is None      # zero_grad(true) of :796
