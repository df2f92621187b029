"""The C++ host layers above the C ABI: the L0 class with the reference's signatures
(include/cuda_rasterizer/rasterizer.h, liblgs_host.so) and the L1 libtorch functions + pybind
module `_C` (include/rasterize_points.h, _C.so)."""
import os
import subprocess

import numpy as np
import pytest
import torch

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_libs():
    from leg_slam_b200 import build_host
    return build_host.build()


def test_l0_library_exports_reference_class(host_libs):
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", host_libs[0]], capture_output=True, text=True).stdout
    for sym in ("CudaRasterizer::Rasterizer::forward(", "CudaRasterizer::Rasterizer::backward(",
                "CudaRasterizer::Rasterizer::markVisible(", "lgs_host_set_stream"):
        assert sym in out, sym


def test_l1_module_exports_reference_names(host_libs):
    from leg_slam_b200 import _C
    for name in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible"):
        assert callable(getattr(_C, name))
    cs = cases.make_case("sh3_lf")
    with pytest.raises(RuntimeError, match="num_points, 3"):
        bad = dict(cs, means3D=cs["means3D"][:, :2])
        _C.rasterize_gaussians(*cases.fwd_args(bad))
    with pytest.raises(RuntimeError, match="no CPU path"):
        _C.rasterize_gaussians(*cases.fwd_args(cs))


@pytest.mark.gpu
def test_l1_cpp_equals_python_binding(host_libs):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from leg_slam_b200 import _C, rasterize_points as rp
    dev = torch.device("cuda:0")
    for name in ("sh3_lf", "precomp_nolf"):
        cs = cases.make_case(name, dev)
        a = _C.rasterize_gaussians(*cases.fwd_args(cs))
        b = rp.rasterize_gaussians(*cases.fwd_args(cs))
        assert a[0] == b[0]
        for x, y in zip(a[1:5], b[1:5]):
            assert torch.equal(x, y)
        ga = _C.rasterize_gaussians_backward(*cases.bwd_args(cs, a[4], a[5], a[0], a[6], a[7]))
        gb = rp.rasterize_gaussians_backward(*cases.bwd_args(cs, b[4], b[5], b[0], b[6], b[7]))
        for n, x, y in zip(cases.GRAD_NAMES, ga, gb):
            assert x.shape == y.shape, n
            assert cases.rel_err(x.cpu().numpy(), y.cpu().numpy()) <= 1e-3, n
        assert torch.equal(_C.mark_visible(cs["means3D"], cs["viewmatrix"], cs["projmatrix"]),
                           rp.mark_visible(cs["means3D"], cs["viewmatrix"], cs["projmatrix"]))


@pytest.mark.gpu
def test_l0_cpp_class_vs_oracle(host_libs, oracle_mod, tmp_path):
    """A C++ program that uses the class the way the reference's rasterize_points.cu does."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    exe = str(tmp_path / "l0_driver")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    pkg = os.path.join(ROOT, "leg_slam_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "l0_driver.cpp"),
                           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda, "include"), "-o", exe,
                           "-L" + pkg, "-llgs_host", "-llgs", "-L" + os.path.join(cuda, "lib64"), "-lcudart",
                           "-Wl,-rpath," + pkg])
    cs = cases.make_case("sh3_lf")
    P, W, H = cs["P"], cs["W"], cs["H"]
    with open(tmp_path / "in.bin", "wb") as f:
        np.array([P, W, H, cs["degree"], 16, 1], np.int32).tofile(f)
        np.array([cs["tanfovx"], cs["tanfovy"], 1.0], np.float32).tofile(f)
        for k in ("bg", "means3D", "shs", "lang_feats", "opacities", "scales", "rotations", "viewmatrix", "projmatrix",
                  "campos", "dL_dcolor", "dL_dlf", "dL_ddepth"):
            cs[k].contiguous().numpy().astype(np.float32).tofile(f)
    r = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    f = cases.oracle_forward(cs, oracle_mod)
    g = cases.oracle_backward(cs, f, oracle_mod)
    with open(tmp_path / "out.bin", "rb") as fh:
        R = int(np.fromfile(fh, np.int32, 1)[0])
        radii = np.fromfile(fh, np.int32, P)
        present = np.fromfile(fh, np.uint8, P).astype(bool)
        rdf = lambda *s: np.fromfile(fh, np.float32, int(np.prod(s))).reshape(s)  # noqa: E731
        color, lf, depth = rdf(3, H, W), rdf(64, H, W), rdf(1, H, W)
        grads = dict(dL_dmeans2D=rdf(P, 3), dL_dcolors=rdf(P, 3), dL_dlang_feats=rdf(P, 64), dL_dopacity=rdf(P, 1),
                     dL_dmeans3D=rdf(P, 3), dL_dcov3D=rdf(P, 6), dL_dsh=rdf(P, 16, 3), dL_dscales=rdf(P, 3),
                     dL_drotations=rdf(P, 4))
    assert R == f["num_rendered"] and np.array_equal(radii, f["radii"])
    assert np.array_equal(present, oracle_mod.mark_visible(cs["means3D"].numpy(), cs["viewmatrix"].numpy()))
    assert cases.rel_err(color, f["out_color"]) <= 1e-4 and cases.rel_err(lf, f["out_lf"]) <= 1e-4
    assert cases.rel_err(depth, f["out_depth"]) <= 1e-4
    for n in cases.GRAD_NAMES:
        assert cases.rel_err(grads[n], g[n]) <= 1e-3, n


def test_l2_cpp_module_validates_like_the_reference(host_libs):
    """The C++ L2 layer (include/gaussian_rasterizer.h -> _L2.so): same two argument-validation messages as
    GaussianRasterizer::forward (reference src/gaussian_rasterizer.cpp:196-206), raised before anything touches a device,
    and CPU tensors are refused by L1 underneath."""
    from leg_slam_b200 import _L2
    cs = cases.make_case("sh3_lf")
    rs = _L2.GaussianRasterizationSettings(cs["H"], cs["W"], cs["tanfovx"], cs["tanfovy"], cs["bg"], 1.0, cs["viewmatrix"],
                                           cs["projmatrix"], cs["degree"], cs["campos"], False, True)
    r = _L2.GaussianRasterizer(rs)
    e = torch.empty(0)
    m3, m2, op = cs["means3D"], torch.zeros_like(cs["means3D"]), cs["opacities"]

    def call(has_shs, has_col, has_sc, has_rot, has_cov):
        return r.forward(m3, m2, op, has_shs, has_col, True, has_sc, has_rot, has_cov, cs["shs"] if has_shs else e,
                         m3 if has_col else e, cs["lang_feats"], cs["scales"] if has_sc else e,
                         cs["rotations"] if has_rot else e, torch.zeros(cs["P"], 6) if has_cov else e)
    for flags in ((False, False, True, True, False), (True, True, True, True, False)):
        with pytest.raises(RuntimeError, match="SHs or precomputed colors"):
            call(*flags)
    for flags in ((True, False, True, False, False), (True, False, False, False, False), (True, False, True, True, True),
                  (True, False, False, True, True)):
        with pytest.raises(RuntimeError, match="scale/rotation pair or precomputed 3D covariance"):
            call(*flags)
    with pytest.raises(RuntimeError, match="no CPU path"):
        call(True, False, True, True, False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        r.markVisibleGaussians(m3)


def test_cpp_fused_adam_refuses_cpu_and_checks_arguments(host_libs):
    """LgsFusedAdam (include/lgs_adam.h, a torch::optim::Adam with a fused step): no CPU path, argument checks."""
    from leg_slam_b200 import _L2
    p = [torch.randn(10, 3), torch.randn(10, 1)]
    g = [[torch.randn(10, 3), torch.randn(10, 1)]]
    with pytest.raises(RuntimeError, match="no CPU path"):
        _L2.fused_adam_run(p, g, [1e-3, 5e-2], 1e-15)
    with pytest.raises(RuntimeError, match="one learning rate per parameter"):
        _L2.fused_adam_run(p, g, [1e-3], 1e-15)
    with pytest.raises(RuntimeError, match="one gradient per parameter"):
        _L2.fused_adam_run(p, [[g[0][0]]], [1e-3, 5e-2], 1e-15)


def test_l1_geometry_operators_validate_like_the_reference(host_libs):
    """include/operate_points.h / stereo_vision.h / spatial.h on liblgs (csrc/host/geometry_ops.cpp, bound in `_C` for the
    tests): the reference's AT_ERROR messages (src/operate_points.cu:62-64,106-118, src/stereo_vision.cu:140-142,175-180),
    its empty-input behaviour, and no CPU path."""
    from leg_slam_b200 import _C
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", os.path.join(ROOT, "leg_slam_b200", "_C.so")],
                         capture_output=True, text=True).stdout
    for sym in ("transformPoints(at::Tensor&, at::Tensor&)", "scaleAndTransformThenMarkVisiblePoints(", "reprojectDepthPinhole(",
                "monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints(", "distCUDA2(at::Tensor const&)"):
        assert sym in out, sym
    I = torch.eye(4)
    ones = torch.ones(5, dtype=torch.bool)
    with pytest.raises(RuntimeError, match=r"points must have dimensions \(num_points, 3\)"):
        _C.transform_points(torch.zeros(5, 2), I)
    with pytest.raises(RuntimeError, match=r"points must have dimensions \(num_points, 3\)"):
        _C.scale_and_transform_then_mark_visible(torch.zeros(5, 2), torch.zeros(5, 4), ones, ones, I, I, I, 0, 1.0)
    with pytest.raises(RuntimeError, match=r"points_mask must have dimensions \(num_points\)"):
        _C.scale_and_transform_then_mark_visible(torch.zeros(5, 3), torch.zeros(5, 4), ones[:4], ones, I, I, I, 0, 1.0)
    with pytest.raises(RuntimeError, match=r"points must have dimensions \(num_points\)"):
        _C.reproject_depth_pinhole(torch.zeros(5, 2), ones, [1.0, 1.0, 0.0, 0.0], 5)
    with pytest.raises(RuntimeError, match=r"kps_pixel must have dimensions \(num_points, 2\)"):
        _C.inactive_geo_densify(torch.zeros(5, 3), ones, torch.zeros(5, 3), torch.zeros(100), 1.0, [1.0, 1.0, 0.0, 0.0], 8)
    with pytest.raises(RuntimeError, match=r"kps_has3D must have dimensions \(num_points\)"):
        _C.inactive_geo_densify(torch.zeros(5, 2), ones[:, None], torch.zeros(5, 3), torch.zeros(100), 1.0, [1.0, 1.0, 0.0, 0.0], 8)
    with pytest.raises(RuntimeError, match=r"kps_point_local must have dimensions \(num_points, 3\)"):
        _C.inactive_geo_densify(torch.zeros(5, 2), ones, torch.zeros(5, 2), torch.zeros(100), 1.0, [1.0, 1.0, 0.0, 0.0], 8)
    for call in (lambda: _C.transform_points(torch.zeros(5, 3), I),
                 lambda: _C.scale_and_transform_then_mark_visible(torch.zeros(5, 3), torch.zeros(5, 4), ones, ones, I, I, I, 0, 1.0),
                 lambda: _C.reproject_depth_pinhole(torch.zeros(5), ones, [1.0, 1.0, 0.0, 0.0], 5),
                 lambda: _C.inactive_geo_densify(torch.zeros(5, 2), ones, torch.zeros(5, 3), torch.zeros(100), 1.0, [1.0, 1.0, 0.0, 0.0], 8),
                 lambda: _C.dist_cuda2(torch.zeros(5, 3))):
        with pytest.raises(RuntimeError, match="no CPU path"):
            call()
    # empty inputs never reach a device (the reference skips its launches for P == 0 / N == 0)
    e3 = torch.zeros(0, 3)
    assert _C.transform_points(e3, I).shape == (0, 3)
    assert _C.scale_and_transform_then_mark_visible(e3, torch.zeros(0, 4), ones[:0], ones[:0], I, I, I, 7, 1.0) == 7
    assert _C.dist_cuda2(e3).shape == (0,)


def test_cpp_gaussian_model_settings_and_surgery_cpu(host_libs, oracle_mod):
    """GaussianModel in C++ (include/gaussian_model.h, _L2.so): SH-degree bookkeeping, trainingSetup's seven groups and rates,
    the xyz schedule bit-identical to the libm statement of exponLrFunc in the C oracle, the rate setters, and the
    optimizer-state surgery that is plain libtorch (prunePoints, resetOpacity) on CPU tensors; everything that needs a
    kernel refuses CPU tensors."""
    from leg_slam_b200 import _L2
    g = _L2.GaussianModel(3)
    assert (g.active_sh_degree_, g.max_sh_degree_) == (0, 3)
    for want in (1, 2, 3, 3):
        g.oneUpShDegree()
        assert g.active_sh_degree_ == want
    g.setShDegree(7)
    assert g.active_sh_degree_ == 3
    g.setShDegree(1)
    assert g.active_sh_degree_ == 1
    gen = torch.Generator().manual_seed(1)
    n = 12
    g.xyz_, g.features_dc_, g.features_rest_ = torch.randn(n, 3, generator=gen), torch.randn(n, 1, 3, generator=gen), torch.randn(n, 15, 3, generator=gen)
    g.language_features_, g.opacity_ = torch.randn(n, 64, generator=gen), torch.randn(n, 1, generator=gen)
    g.scaling_, g.rotation_ = torch.randn(n, 3, generator=gen), torch.randn(n, 4, generator=gen)
    for name in ("xyz_", "features_dc_", "features_rest_", "language_features_", "opacity_", "scaling_", "rotation_"):
        setattr(g, name, getattr(g, name).requires_grad_())  # leaves, as createFromPcd leaves them
    g.exist_since_iter_ = torch.arange(n, dtype=torch.int32)
    g.max_radii2D_ = torch.zeros(n)
    g.spatial_lr_scale_ = 5.3
    a = _L2.GaussianOptimizationParams()
    assert abs(a.position_lr_init_ - 1.6e-4) < 1e-10 and a.position_lr_max_steps_ == 30_000
    g.trainingSetup(a)
    assert g.params_are_the_optimizers()
    f32 = np.float32
    init, final = float(f32(1.6e-4) * f32(5.3)), float(f32(1.6e-6) * f32(5.3))
    want = [init, float(f32(2.5e-3)), float(f32(2.5e-3)) / 20.0, float(f32(1.5e-3)), float(f32(0.05)), float(f32(5e-3)), float(f32(1e-3))]
    assert [g.learning_rate(i) for i in range(7)] == want
    assert g.xyz_gradient_accum_.shape == (n, 1) and g.denom_.shape == (n, 1)
    for step in (0, 1, 17, 999, 15_000, 29_999, 30_000, 31_000, -1):
        ref = oracle_mod.expon_lr(step, init, final, 0.01, 0, 30_000)
        assert g.updateLearningRate(step) == ref and g.learning_rate(0) == ref, step
    g.setPositionLearningRate(2e-4)
    assert g.learning_rate(0) == float(f32(2e-4) * f32(5.3))
    g.setFeatureLearningRate(1e-3)
    assert g.learning_rate(1) == float(f32(1e-3)) and g.learning_rate(2) == float(f32(1e-3)) / 20.0
    g.setLanguageFeatureLearningRate(2e-3), g.setOpacityLearningRate(0.04), g.setScalingLearningRate(4e-3), g.setRotationLearningRate(2e-3)
    assert [g.learning_rate(i) for i in (3, 4, 5, 6)] == [float(f32(v)) for v in (2e-3, 0.04, 4e-3, 2e-3)]
    assert g.adam_state(0)[0] == -1  # no step yet: no state
    with pytest.raises(RuntimeError, match="no CPU path"):
        g.step([torch.zeros_like(t) for t in (g.xyz_, g.features_dc_, g.features_rest_, g.language_features_, g.opacity_, g.scaling_, g.rotation_)])
    with pytest.raises(RuntimeError, match="no CPU path"):
        g.densifyAndPrune(1e-3, 0.005, 5.0, 20)
    with pytest.raises(RuntimeError, match="no CPU path"):
        g.createFromPcd(torch.zeros(4, 3), torch.zeros(4, 3), torch.empty(0), 1.0)
    # libtorch-only surgery
    op = g.opacity_.detach().clone()
    g.resetOpacity()
    assert torch.allclose(torch.sigmoid(g.opacity_.detach()), torch.sigmoid(op), atol=1e-6) and g.params_are_the_optimizers()
    mask = torch.zeros(n, dtype=torch.bool)
    mask[[1, 5, 6]] = True
    xyz = g.xyz_.detach().clone()
    g.prunePoints(mask)
    assert torch.equal(g.xyz_.detach(), xyz[~mask]) and g.rotation_.shape == (n - 3, 4) and g.params_are_the_optimizers()
    assert g.exist_since_iter_.tolist() == [i for i in range(n) if i not in (1, 5, 6)] and g.denom_.shape == (n - 3, 1)
    cov = g.getCovarianceActivation(1)
    q = torch.nn.functional.normalize(g.rotation_.detach())
    r, x, y, z = q.unbind(1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y), 2 * (x * y + r * z), 1 - 2 * (x * x + z * z),
                     2 * (y * z - r * x), 2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).view(-1, 3, 3)
    L = R @ torch.diag_embed(torch.exp(g.scaling_.detach()))
    S = L @ L.transpose(1, 2)
    want_cov = torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], 1)
    assert torch.allclose(cov, want_cov, rtol=1e-5, atol=1e-6)
    assert torch.equal(g.getFeatures(), torch.cat([g.features_dc_, g.features_rest_], 1))


def test_cpp_library_exports_the_reference_classes(host_libs):
    """liblgs_torch.so -- the libtorch layers as a plain C++ library -- exports the reference's functions and classes by
    their own (mangled) names, and the C++ mapper driver (tests/cpp/mapper_driver.cpp, no Python in the process) links
    against it."""
    from leg_slam_b200 import build_host
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", build_host.LIB_TORCH], capture_output=True, text=True).stdout
    for sym in ("RasterizeGaussiansCUDA(", "RasterizeGaussiansBackwardCUDA(", "markVisible(", "transformPoints(", "distCUDA2(",
                "scaleAndTransformThenMarkVisiblePoints(", "reprojectDepthPinhole(",
                "monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints(", "GaussianRasterizerFunction::forward(",
                "GaussianRasterizer::forward(", "LgsFusedAdam::step(", "GaussianModel::trainingSetup(",
                "GaussianModel::densifyAndPrune(", "GaussianModel::increasePcd(", "GaussianModel::updateLearningRate(",
                "GaussianModel::scaledTransformVisiblePointsOfKeyframe(", "GaussianModel::savePly(", "GaussianModel::loadPly(",
                "GaussianRenderer::render(", "mappingIterationBackward(", "mappingIterationStep("):
        assert sym in out, sym
    assert "PyInit" not in out
    exe = build_host.build_mapper_driver()
    assert os.access(exe, os.X_OK)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1  # usage error: the program loads (every shared-library dependency resolves) and needs two paths


@pytest.mark.gpu
def test_cpp_mapper_driver_follows_the_python_mapper(host_libs, tmp_path):
    """A C++ program on liblgs_torch.so that does what GaussianMapper::trainForOneIteration does with the reference's classes
    (settings, render, loss, backward, statistics, optimizer step; reference src/gaussian_mapper.cpp:662-796), against the
    Python mapper's fused path on the same scene, keyframe and learning-rate schedule."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import math
    from leg_slam_b200 import build_host, mapper as M, synthetic
    exe = build_host.build_mapper_driver()
    dev = torch.device("cuda:0")
    W, H, P, N_IT = 96, 64, 4000, 4
    sc = synthetic.make_scene(P, seed=51, mean_scale=0.06, device=dev)
    cam = synthetic.make_cameras(1, W, H, seed=51)[0].to(dev)
    g = torch.Generator().manual_seed(52)
    kf = M.Keyframe(cam, torch.rand(3, H, W, generator=g).to(dev), torch.randn(64, 37, 37, generator=g).to(dev),
                    (torch.rand(1, H, W, generator=g) * 3).to(dev))
    with open(tmp_path / "in.bin", "wb") as f:
        np.array([P, W, H, N_IT, 37, 37], np.int32).tofile(f)
        np.array([2.0 * math.atan(cam.tanfovx), 2.0 * math.atan(cam.tanfovy)], np.float32).tofile(f)
        for t in ([sc[k] for k in M.PARAM_ORDER] + [cam.viewmatrix, cam.projmatrix, cam.campos, kf.gt_image, kf.gt_lf, kf.gt_depth]):
            t.detach().contiguous().cpu().numpy().astype(np.float32).tofile(f)
    r = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    with open(tmp_path / "out.bin", "rb") as fh:
        losses = np.fromfile(fh, np.float32, N_IT)
        xyz = np.fromfile(fh, np.float32, 3 * P).reshape(P, 3)
        opacity = np.fromfile(fh, np.float32, P).reshape(P, 1)
        denom = np.fromfile(fh, np.float32, P)
        max_radii = np.fromfile(fh, np.float32, P)
        lr_last = float(np.fromfile(fh, np.float32, 1)[0])
    mp = M.Mapper(sc, sh_degree=3, track_densify_stats=True)
    mp.set_position_lr_schedule(3.2e-4, 3.2e-6, 0.01, N_IT)
    lrs = []
    for it in range(N_IT):
        lrs.append(mp.update_learning_rate(it))
        l_py = float(mp.train_step([kf]))
        assert abs(float(losses[it]) - l_py) <= 1e-4 * abs(l_py), (it, float(losses[it]), l_py)
    assert abs(lr_last - lrs[-1]) <= 2e-7 * lrs[-1]
    assert np.array_equal(denom, mp.stats.denom.cpu().numpy().reshape(-1))
    assert np.array_equal(max_radii, mp.stats.max_radii2D.cpu().numpy())
    for got, k, step in ((xyz, "xyz", sum(lrs)), (opacity, "opacity", N_IT * M.DEFAULT_LRS["opacity"])):
        d = np.abs(got - mp.params[k].detach().cpu().numpy())
        assert (d > 0.05 * step).mean() <= 3e-3, (k, float(d.max()), step)


def test_public_headers_are_self_contained():
    """Every header under include/ compiles on its own (what a maintainer who includes just one of them needs): the C ABI as
    C and as C++, the libtorch-level headers against the pip libtorch."""
    import sysconfig
    from concurrent.futures import ThreadPoolExecutor
    from torch.utils import cpp_extension as ce
    inc = os.path.join(ROOT, "include")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    torch_inc = ["-I" + inc, "-I" + os.path.join(cuda, "include"), "-I" + sysconfig.get_paths()["include"]] + ["-I" + p for p in ce.include_paths()]
    jobs = [["gcc", "-std=c11", "-fsyntax-only", "-x", "c", os.path.join(inc, "lgs.h")],
            ["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", os.path.join(inc, "lgs.h")],
            ["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", "-I" + inc, "-I" + os.path.join(cuda, "include"),
             os.path.join(inc, "cuda_rasterizer", "rasterizer.h")]]
    for h in ("rasterize_points.h", "operate_points.h", "stereo_vision.h", "spatial.h", "lgs_adam.h", "gaussian_rasterizer.h",
              "gaussian_keyframe.h", "gaussian_model.h", "gaussian_renderer.h"):
        jobs.append(["g++", "-std=c++17", "-fsyntax-only", "-D_GLIBCXX_USE_CXX11_ABI=1", "-x", "c++"] + torch_inc + [os.path.join(inc, h)])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd[-1], r.returncode, r.stderr[-1500:]
    with ThreadPoolExecutor(max_workers=8) as ex:
        for name, rc, err in ex.map(run, jobs):
            assert rc == 0, (name, err)
