"""The UNMODIFIED reference GaussianModel (src/gaussian_model.cpp compiled into oracle/_ref/ref_model.so, running on CPU
tensors) as the oracle of the rows beside the path: the torch restatements the GPU tests hold the
CUDA code to -- oracle/densify_ref.py, oracle/ply_ref.py -- the C oracle's Adam, and the package's host logic (learning-rate
schedule, per-group rates) are held to the reference's own methods here.  No GPU: the reference class is libtorch code and
takes data_device = "cpu" (src/gaussian_model.cpp:37-41).

With this file the chain for density control is: CUDA (leg_slam_b200/csrc/densify.cu) == oracle/densify_ref.py on a B200
(tests/test_densify.py) and oracle/densify_ref.py == the reference's GaussianModel, bit for bit, here."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import build_ref  # noqa: E402
import densify_ref as DR  # noqa: E402
import ingest_ref as IR  # noqa: E402
import ply_ref  # noqa: E402

OPT = dict(position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01, position_lr_max_steps=30000,
           feature_lr=0.0025, language_feature_lr=0.0015, opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001,
           percent_dense=0.01)


@pytest.fixture(scope="module")
def RM():
    try:
        if os.path.isdir(os.path.join(build_ref.MODEL_REF, "src")):
            build_ref.build_model(verbose=False)     # no-op when oracle/_ref/ref_model.so exists
        return build_ref.load_model()
    except FileNotFoundError as ex:
        pytest.skip(str(ex))


def make_params(P, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    t = [r(P, 3), r(P, 1, 3), r(P, 15, 3) * 0.1, r(P, 64), r(P, 1) * 3.0, r(P, 3) * 0.8 - 3.5, r(P, 4)]
    exist = torch.randint(0, 50, (P,), generator=g, dtype=torch.int32)
    return t, exist, g


def make_pair(RM, P, seed, steps=2, spatial_lr_scale=2.5):
    """A reference model with `steps` Adam steps behind it (so every group has non-trivial moments) and the restatement's Model
    carrying the same tensors."""
    t, exist, g = make_params(P, seed)
    ref = RM.GaussianModel(3)
    ref.set_state(t, exist, spatial_lr_scale)
    ref.training_setup(**OPT)
    for _ in range(steps):
        ref.set_grads([torch.randn(*x.shape, generator=g) * 0.01 for x in t])
        ref.step()
    ours = DR.Model(dict(zip(DR.PARAMS, [p.detach() for p in ref.params()])), percent_dense=OPT["percent_dense"])
    for i, k in enumerate(DR.PARAMS):
        _, ours.m[k], ours.v[k] = [x.clone() if torch.is_tensor(x) else x for x in ref.moments(i)]
    ours.exist_since_iter = exist.clone()
    return ref, ours, g


def assert_same(ref, ours, steps):
    for i, k in enumerate(DR.PARAMS):
        assert torch.equal(ref.params()[i].detach(), ours.p[k]), k
        st = ref.moments(i)
        assert st is not None and st[0] == steps, k
        assert torch.equal(st[1], ours.m[k]) and torch.equal(st[2], ours.v[k]), k
        assert ref.optimizer_params()[i].data_ptr() == ref.params()[i].data_ptr()
    assert ref.state_size() == 7
    assert torch.equal(ref.exist_since_iter, ours.exist_since_iter)
    for a, b in ((ref.xyz_gradient_accum, ours.xyz_gradient_accum), (ref.denom, ours.denom), (ref.max_radii2D, ours.max_radii2D)):
        assert a.shape == b.shape and torch.equal(a, b)


def seeded_normal01(seed):
    """at::normal(means, stds) on CPU tensors draws standard normals of the output's shape from the default generator, then
    scales (src/gaussian_model.cpp:746): the restatement's `normal01` draws the same numbers after the same seed."""
    def f(n):
        torch.manual_seed(seed)
        return torch.empty(n, 3).normal_()
    return f


@pytest.mark.parametrize("max_screen_size", [0, 20])
@pytest.mark.parametrize("P", [1, 300, 5000])
def test_densify_and_prune_restatement_equals_reference_model(RM, P, max_screen_size):
    """addDensificationStats over three views, then densifyAndPrune (reference src/gaussian_model.cpp:806-847 with everything it
    calls: :597-804) -- parameters, both Adam moments, step counts, exist_since_iter and the statistics vectors bit-identical
    between oracle/densify_ref.py and the reference class."""
    ref, ours, g = make_pair(RM, P, seed=100 + P)
    ref.xyz_gradient_accum, ref.denom = torch.zeros(P, 1), torch.zeros(P, 1)   # trainingSetup has set them; explicit here
    for _ in range(3):
        radii = torch.randint(-5, 40, (P,), generator=g, dtype=torch.int32).clamp_min(0)
        grad = torch.randn(P, 3, generator=g) * 3e-4
        f = radii > 0
        # src/gaussian_mapper.cpp:739-742 (the mapper's part of the statistics), then the model's
        mr = ref.max_radii2D
        mr[f] = torch.max(mr[f], radii[f].to(mr.dtype))
        ref.max_radii2D = mr
        ref.add_densification_stats(grad, f)
        ours.add_stats(radii, grad)
    assert_same(ref, ours, 2)
    n0 = ours.p["xyz"].shape[0]
    torch.manual_seed(7)
    ref.densify_and_prune(2e-4, 0.05, 4.0, max_screen_size)
    ours.densify_and_prune(2e-4, 0.05, 4.0, max_screen_size, seeded_normal01(7))
    assert_same(ref, ours, 2)
    if P >= 300:
        assert ours.p["xyz"].shape[0] != n0   # something was cloned / split / pruned
    assert not ref.xyz_gradient_accum.any() and not ref.denom.any() and not ref.max_radii2D.any()


def test_clone_split_prune_pieces_equal_reference_model(RM):
    """densifyAndClone, densifyAndSplit (N = 2 and 3) and prunePoints one by one (reference :775-804, :729-773, :597-651)."""
    P = 800
    for N in (2, 3):
        ref, ours, g = make_pair(RM, P, seed=31)
        grads = torch.rand(P, 1, generator=g) * 5e-4
        ref.densify_and_clone(grads, 2e-4, 4.0)
        ours.densify_and_clone(grads, 2e-4, 4.0)
        assert_same(ref, ours, 2)
        assert ours.p["xyz"].shape[0] > P
        torch.manual_seed(3)
        ref.densify_and_split(grads, 2e-4, 4.0, N)
        n_sel = ours.densify_and_split(grads, 2e-4, 4.0, seeded_normal01(3), N=N)
        assert n_sel > 0
        assert_same(ref, ours, 2)
        mask = torch.rand(ours.p["xyz"].shape[0], generator=g) < 0.3
        ref.prune_points(mask)
        ours.prune_points(mask)
        assert_same(ref, ours, 2)


def test_reset_opacity_equals_reference_model(RM):
    """resetOpacity (reference :567-595): the clamp against ones leaves the values but passes them
    through sigmoid and inverse_sigmoid; the opacity moments are zeroed, their step count and every other group are kept."""
    ref, ours, _ = make_pair(RM, 500, seed=5, steps=3)
    before = [ref.moments(i)[1].clone() for i in range(7)]
    ref.reset_opacity()
    ours.reset_opacity()
    assert_same(ref, ours, 3)
    assert not ref.moments(4)[1].any() and not ref.moments(4)[2].any()
    for i in (0, 1, 2, 3, 5, 6):
        assert torch.equal(ref.moments(i)[1], before[i])


def test_increase_pcd_restatement_equals_reference_model(RM):
    """increasePcd, tensor and std::vector overloads (reference :297-384, :196-295), with distCUDA2 supplied by the numpy
    restatement of simple-knn (oracle/ingest_ref.py; pinned by the compiled simple-knn on a B200, tests/test_ingest.py)."""
    dist2 = lambda pts: torch.from_numpy(IR.knn_mean_dist2(pts.numpy()))  # noqa: E731
    RM.set_dist2(dist2)
    try:
        ref, ours, g = make_pair(RM, 400, seed=9)
        pts = torch.rand(150, 3, generator=g) * torch.tensor([6.0, 4.0, 2.8])
        cols = torch.rand(150, 3, generator=g)
        ref.increase_pcd(pts, cols, 42)
        ours.increase_pcd(pts, cols, 42, dist2)
        assert_same(ref, ours, 2)
        assert ours.p["xyz"].shape[0] == 550 and int(ours.exist_since_iter[-1]) == 42
        assert torch.equal(ref.sparse_points_xyz, pts) and torch.equal(ref.sparse_points_color, cols)
        pts2, cols2 = torch.rand(7, 3, generator=g), torch.rand(7, 3, generator=g)
        ref.increase_pcd_vec(pts2.flatten().tolist(), cols2.flatten().tolist(), 43)
        ours.increase_pcd(pts2, cols2, 43, dist2)
        assert_same(ref, ours, 2)
        assert ref.sparse_points_xyz.shape == (157, 3)
        ref.increase_pcd_vec([], [], 44)   # no points: nothing happens (:200-201)
        assert_same(ref, ours, 2)
    finally:
        RM.set_dist2(None)
    with pytest.raises(RuntimeError, match="no callable set"):
        ref.increase_pcd(pts, cols, 45)


def test_create_from_pcd_composition_equals_reference_model(RM):
    """createFromPcd (reference :109-194): the composition tests/test_densify.py::test_create_from_pcd_and_scaled_transformation
    holds the CUDA path to on a B200 (RGB2SH of the colours in the DC coefficient, zero rest, log sqrt clamp knn in all three
    scales, identity quaternions, inverse_sigmoid(0.1), zero ages) is what the reference class produces.  The class reads the
    points from std::map<id, Point3D> (doubles, ordered by id)."""
    dist2 = lambda pts: torch.from_numpy(IR.knn_mean_dist2(pts.numpy()))  # noqa: E731
    RM.set_dist2(dist2)
    try:
        g = torch.Generator().manual_seed(91)
        n = 300
        pts = torch.rand(n, 3, generator=g) * torch.tensor([6.0, 4.0, 2.8])
        cols, lfs = torch.rand(n, 3, generator=g), torch.randn(n, 64, generator=g)
        ref = RM.GaussianModel(3)
        ref.create_from_pcd(pts, cols, lfs, 3.5)
    finally:
        RM.set_dist2(None)
    xyz, f_dc, f_rest, lf, op, sc, rot = [p.detach() for p in ref.params()]
    assert ref.spatial_lr_scale == 3.5
    assert torch.equal(xyz, pts) and torch.equal(lf, lfs)
    C0 = 0.282094806432724   # the float the reference's `0.28209479177387814f` is (include/sh_utils.h:32)
    assert f_dc.shape == (n, 1, 3) and torch.equal(f_dc[:, 0], (cols - 0.5) / C0)
    assert f_rest.shape == (n, 15, 3) and not f_rest.any()
    want = torch.log(torch.sqrt(torch.clamp_min(dist2(pts), 0.0000001)))
    assert torch.equal(sc, want[:, None].repeat(1, 3))
    assert torch.equal(rot, torch.tensor([1.0, 0, 0, 0]).repeat(n, 1))
    assert torch.equal(op, DR.inverse_sigmoid(0.1 * torch.ones(n, 1)))
    assert ref.exist_since_iter.dtype == torch.int32 and not ref.exist_since_iter.any() and ref.max_radii2D.shape == (n,)
    assert all(p.requires_grad for p in ref.params())


def test_libtorch_adam_of_training_setup_equals_c_oracle_and_python_adam(RM):
    """trainingSetup's seven groups (reference :483-518: eps 1e-15, per-group rates, xyz scaled by spatial_lr_scale,
    features_rest at feature_lr / 20) stepped by libtorch's own Adam, against (1) the C oracle's Adam (oracle/lgs_oracle.c,
    what the CUDA kernel is held to bit for bit) and (2) torch.optim.Adam with the same groups: parameters within 1e-6
    relative, moments likewise."""
    import oracle as O
    P, steps = 700, 4
    t, exist, g = make_params(P, seed=77)
    ref = RM.GaussianModel(3)
    ref.set_state(t, exist, 2.5)
    ref.training_setup(**OPT)
    lrs = ref.lrs()
    want_lrs = [np.float32(OPT["position_lr_init"]) * np.float32(2.5), np.float32(OPT["feature_lr"]),
                float(np.float32(OPT["feature_lr"])) / 20.0, np.float32(OPT["language_feature_lr"]), np.float32(OPT["opacity_lr"]),
                np.float32(OPT["scaling_lr"]), np.float32(OPT["rotation_lr"])]
    assert lrs == [float(x) for x in want_lrs]
    py_p = [torch.nn.Parameter(x.clone()) for x in t]
    py = torch.optim.Adam([dict(params=[p], lr=lr) for p, lr in zip(py_p, lrs)], lr=0.0, eps=1e-15, foreach=False, fused=False)
    c_p = [x.numpy().copy().reshape(-1) for x in t]
    c_m = [np.zeros_like(x) for x in c_p]
    c_v = [np.zeros_like(x) for x in c_p]
    for s in range(1, steps + 1):
        grads = [torch.randn(*x.shape, generator=g) * 0.01 for x in t]
        ref.set_grads(grads)
        ref.step()
        for p, gr in zip(py_p, grads):
            p.grad = gr.clone()
        py.step()
        for i in range(7):
            O.adam(c_p[i], grads[i].numpy().reshape(-1), c_m[i], c_v[i], lrs[i], step=s)
    for i in range(7):
        r = ref.params()[i].detach()
        st = ref.moments(i)
        assert st[0] == steps
        for a, b in ((st[1], py.state[py_p[i]]["exp_avg"]), (st[2], py.state[py_p[i]]["exp_avg_sq"])):
            assert float((a - b).abs().max()) <= 1e-6 * float(b.abs().max())   # mul_ + add_ there, lerp_ here: ulps apart
        scale = float(r.abs().max())
        assert float((r - py_p[i].detach()).abs().max()) <= 1e-6 * scale
        assert float(np.abs(r.numpy().reshape(-1) - c_p[i]).max()) <= 1e-6 * scale
        assert float(np.abs(st[1].numpy().reshape(-1) - c_m[i]).max()) <= 1e-6 * float(st[1].abs().max())
        assert float(np.abs(st[2].numpy().reshape(-1) - c_v[i]).max()) <= 1e-6 * float(st[2].abs().max())


def test_learning_rate_schedule_and_setters_equal_reference_model(RM):
    """updateLearningRate / exponLrFunc (reference :520-531, :1143-1157) against the mapper's schedule, float for float, over
    whole runs; the per-group setters (:543-565) against Mapper's (position scaled by spatial_lr_scale, features_rest =
    feature / 20).  The reference holds its settings as float, so the rates handed to the mapper are float-rounded first; its
    setters then agree to the last bit of the float the Adam kernel receives."""
    from leg_slam_b200 import mapper as M
    from leg_slam_b200 import synthetic
    f32 = lambda x: float(np.float32(x))  # noqa: E731
    t, exist, _ = make_params(10, seed=1)
    mp = M.Mapper(synthetic.make_scene(10, seed=3), sh_degree=3)
    lr_of = lambda k: mp.optimizer.param_groups[M.PARAM_ORDER.index(k)]["lr"]  # noqa: E731
    for scale, max_steps in ((2.5, 30000), (1.0, 500), (6.0, 30000)):
        ref = RM.GaussianModel(3)
        ref.set_state(t, exist, scale)
        ref.training_setup(**dict(OPT, position_lr_max_steps=max_steps))
        mp.set_position_lr_schedule(f32(OPT["position_lr_init"]), f32(OPT["position_lr_final"]),
                                    f32(OPT["position_lr_delay_mult"]), max_steps, spatial_lr_scale=scale)
        assert f32(lr_of("xyz")) == ref.lrs()[0]
        steps = list(range(0, 40)) + list(range(40, max_steps + 2000, 97)) + [max_steps - 1, max_steps, max_steps + 1]
        for s in steps:
            want = ref.update_learning_rate(s)
            got = mp.update_learning_rate(s)
            assert got == want, (scale, max_steps, s, got, want)
            assert ref.lrs()[0] == want and lr_of("xyz") == want
    ref.set_position_learning_rate(0.0002)
    ref.set_feature_learning_rate(0.003)
    ref.set_language_feature_learning_rate(0.002)
    ref.set_opacity_learning_rate(0.06)
    ref.set_scaling_learning_rate(0.004)
    ref.set_rotation_learning_rate(0.0015)
    mp.set_position_learning_rate(f32(0.0002))
    mp.set_feature_learning_rate(f32(0.003))
    mp.set_language_feature_learning_rate(f32(0.002))
    mp.set_opacity_learning_rate(f32(0.06))
    mp.set_scaling_learning_rate(f32(0.004))
    mp.set_rotation_learning_rate(f32(0.0015))
    assert [f32(lr_of(k)) for k in M.PARAM_ORDER] == [f32(x) for x in ref.lrs()]


def test_sh_degree_and_activations_of_reference_model(RM):
    """setShDegree / oneUpShDegree (reference :101-108) against Mapper's host logic, and the activations GaussianRenderer::render
    feeds the rasterizer (:46-99) against the torch statements the package's renderer tests use."""
    t, exist, _ = make_params(200, seed=2)
    ref = RM.GaussianModel(3)
    ref.set_state(t, exist, 1.0)
    assert ref.active_sh_degree() == 0
    for want in (1, 2, 3, 3):
        ref.one_up_sh_degree()
        assert ref.active_sh_degree() == want
    ref.set_sh_degree(7)
    assert ref.active_sh_degree() == 3
    ref.set_sh_degree(1)
    assert ref.active_sh_degree() == 1
    xyz, f_dc, f_rest, lf, op, sc, rot = t
    assert torch.equal(ref.get_scaling_activation(), torch.exp(sc))
    assert torch.equal(ref.get_opacity_activation(), torch.sigmoid(op))
    assert torch.equal(ref.get_rotation_activation(), torch.nn.functional.normalize(rot))
    assert torch.equal(ref.get_features(), torch.cat([f_dc, f_rest], dim=1))
    assert torch.equal(ref.get_language_features(), lf)
    # getCovarianceActivation: L = R(q / |q|) diag(s), Sigma = L L^T, upper triangle
    cov = ref.get_covariance_activation(1)
    R = DR.build_rotation(rot)
    L = R * torch.exp(sc)[:, None, :]
    S = L @ L.transpose(1, 2)
    want = torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], dim=1)
    assert float((cov.detach() - want).abs().max()) <= 1e-6 * float(want.abs().max())


def test_scaled_transformations_of_reference_model(RM):
    """applyScaledTransformation (reference :387-405) and scaledTransformVisiblePointsOfKeyframe (:422-481) with the CUDA
    operators supplied by their numpy restatements (oracle/ingest_ref.py, held bit-identical to the compiled operators:
    tests/golden/geometry.npz): which tensors are replaced, what their Adam state becomes, which rows count as unstable --
    the behaviour tests/test_densify.py holds the mapper's methods to on a B200."""
    def transform_points(points, T):
        points.copy_(torch.from_numpy(IR.transform_points(points.numpy(), T.numpy())))

    seen = {}

    def scale_and_transform(points, rots, not_transformed, unstable, T, view, proj, scale):
        seen["unstable"] = unstable.clone()
        seen["rots"] = rots.clone()
        p, r, f, n = IR.scale_and_transform_then_mark_visible(points.numpy(), rots.numpy(), not_transformed.numpy(),
                                                              unstable.numpy(), T.numpy(), view.numpy(), scale)
        points.copy_(torch.from_numpy(p))
        rots.copy_(torch.from_numpy(r))
        not_transformed.copy_(torch.from_numpy(f))
        return int(n)

    RM.set_transform_points(transform_points)
    RM.set_scale_and_transform(scale_and_transform)
    try:
        P = 600
        ref, ours, g = make_pair(RM, P, seed=13, steps=3)
        th = 0.3
        Rm = torch.tensor([[np.cos(th), -np.sin(th), 0.0], [np.sin(th), np.cos(th), 0.0], [0.0, 0.0, 1.0]], dtype=torch.float32)
        tv = torch.tensor([0.5, -0.25, 0.125])
        xyz0, sc0 = ours.p["xyz"].clone(), ours.p["scaling"].clone()
        keep = [ref.moments(i)[1].clone() for i in range(7)]
        with torch.no_grad():
            ref.apply_scaled_transformation(1.25, Rm, tv)
        T = torch.eye(4)
        T[:3, :3], T[:3, 3] = Rm, tv
        Tt = T.t().contiguous()     # the tensor transformPoints receives is the transposed pose (:393-394)
        want = torch.from_numpy(IR.transform_points((xyz0 * 1.25).numpy(), Tt.numpy()))
        assert torch.equal(ref.params()[0].detach(), want)
        assert torch.equal(ref.params()[5].detach(), sc0 * 1.25)
        for i in range(7):
            st = ref.moments(i)
            assert st[0] == 3
            if i in (0, 5):
                assert not st[1].any() and not st[2].any()
            else:
                assert torch.equal(st[1], keep[i])
        assert ref.state_size() == 7
        # the loop-closure correction of one keyframe
        ref.exist_since_iter = torch.randint(0, 40, (P,), generator=g, dtype=torch.int32)
        flags = torch.rand(P, generator=g) > 0.2
        flags0 = flags.clone()
        view = torch.eye(4)
        view[3, 2] = 4.0            # transposed storage: the translation sits in the last row
        proj = torch.eye(4)
        rot_before = ref.params()[6].detach().clone()
        xyz_before = ref.params()[0].detach().clone()
        n = ref.scaled_transform_visible_points_of_keyframe(flags, Tt, view, proj, 17, 15, 1.03)
        assert torch.equal(seen["unstable"], torch.abs(ref.exist_since_iter - 17) < 15)
        assert torch.equal(seen["rots"], torch.nn.functional.normalize(rot_before))   # the ACTIVATED rotations go in ...
        p, r, f, n_want = IR.scale_and_transform_then_mark_visible(xyz_before.numpy(), seen["rots"].numpy(), flags0.numpy(),
                                                                   seen["unstable"].numpy(), Tt.numpy(), view.numpy(), 1.03)
        assert n == int(n_want) and 0 < n < P
        assert torch.equal(flags, torch.from_numpy(f))
        assert torch.equal(ref.params()[0].detach(), torch.from_numpy(p))
        assert torch.equal(ref.params()[6].detach(), torch.from_numpy(r))             # ... and come back as the parameter
        for i in range(7):
            st = ref.moments(i)
            assert st[0] == 3
            if i in (0, 5, 6):
                assert not st[1].any() and not st[2].any()
            else:
                assert torch.equal(st[1], keep[i])
    finally:
        RM.set_transform_points(None)
        RM.set_scale_and_transform(None)


def test_ply_restatement_equals_reference_model_save_and_load(RM, tmp_path):
    """savePly / loadPly of the reference class itself (reference :854-1075): the file oracle/ply_ref.py writes for the same
    tensors is byte-identical to the reference's; the restatement reads the reference's file back, and the reference's loader
    parses the restatement's (it ignores lf_*)."""
    P = 257
    t, exist, _ = make_params(P, seed=21)
    ref = RM.GaussianModel(3)
    ref.set_state(t, exist, 1.0)
    a, b = str(tmp_path / "ref.ply"), str(tmp_path / "ours.ply")
    ref.save_ply(a)
    xyz, f_dc, f_rest, lf, op, sc, rot = [x.numpy() for x in t]
    ply_ref.write_ply(b, xyz, f_dc, f_rest, lf, op, sc, rot)
    assert open(a, "rb").read() == open(b, "rb").read()
    back = ply_ref.read_ply(a, max_sh_degree=3)
    for k, v in zip(DR.PARAMS, (xyz, f_dc, f_rest, lf, op, sc, rot)):
        assert np.array_equal(np.asarray(back[k]).reshape(v.shape), v), k
    # loadPly with a CPU device: `from_blob(vector.data()).to(device_type_)` (:943-966) does not copy when the device is the
    # CPU, so every tensor but features_rest (which .contiguous() after the transpose does copy) aliases a local std::vector
    # that is gone when the function returns -- harmless on the CUDA device the reference runs on, unusable here.  What can
    # be held on CPU: the reference's property requests parse the restatement's file, the element count, features_rest and
    # the SH degree it activates.
    ld = RM.GaussianModel(3)
    ld.load_ply(b)
    got = ld.params()
    assert [tuple(x.shape) for x in got[:3]] == [(P, 3), (P, 1, 3), (P, 15, 3)]
    assert torch.equal(got[2].detach(), t[2])
    assert ld.active_sh_degree() == 3


def test_cpp_model_of_the_package_side_by_side_with_the_reference_class(RM):
    """The package's C++ GaussianModel (include/gaussian_model.h, leg_slam_b200/_L2.so -- the reference's member names on
    liblgs) next to the reference's class on the same CPU tensors, for everything of it that is plain libtorch: trainingSetup's
    seven rates, updateLearningRate over a whole schedule (the same float at every step), the six setters, the SH-degree
    bookkeeping, the activations, and the optimizer-state surgery of resetOpacity and prunePoints (what needs a kernel refuses
    CPU tensors there: tests/test_host_cpp.py; it is held to the restatement on a B200: tests/test_host_cpp_gpu.py)."""
    from leg_slam_b200 import build_host
    build_host.build()
    from leg_slam_b200 import _L2
    n = 40
    t, exist, g = make_params(n, seed=8)
    ref = RM.GaussianModel(3)
    ref.set_state(t, exist, 5.3)
    ref.training_setup(**OPT)
    ours = _L2.GaussianModel(3)
    for name, x in zip(("xyz_", "features_dc_", "features_rest_", "language_features_", "opacity_", "scaling_", "rotation_"), t):
        setattr(ours, name, x.clone().requires_grad_())
    ours.exist_since_iter_ = exist.clone()
    ours.max_radii2D_ = torch.zeros(n)
    ours.spatial_lr_scale_ = 5.3
    a = _L2.GaussianOptimizationParams()   # the defaults of include/gaussian_parameters.h = OPT
    ours.trainingSetup(a)
    lrs = lambda: [ours.learning_rate(i) for i in range(7)]  # noqa: E731
    assert lrs() == ref.lrs() and ours.percentDense() == ref.percent_dense()
    for s in list(range(0, 30)) + list(range(30, 32000, 61)) + [29999, 30000, 30001, -1]:
        assert ours.updateLearningRate(s) == ref.update_learning_rate(s), s
        assert lrs() == ref.lrs()
    for ro, rr, v in ((ours.setPositionLearningRate, ref.set_position_learning_rate, 2e-4),
                      (ours.setFeatureLearningRate, ref.set_feature_learning_rate, 3e-3),
                      (ours.setLanguageFeatureLearningRate, ref.set_language_feature_learning_rate, 2e-3),
                      (ours.setOpacityLearningRate, ref.set_opacity_learning_rate, 0.06),
                      (ours.setScalingLearningRate, ref.set_scaling_learning_rate, 4e-3),
                      (ours.setRotationLearningRate, ref.set_rotation_learning_rate, 1.5e-3)):
        ro(v), rr(v)
        assert lrs() == ref.lrs()
    for _ in range(5):
        ours.oneUpShDegree(), ref.one_up_sh_degree()
        assert ours.active_sh_degree_ == ref.active_sh_degree()
    for sh in (7, 1, 0, 3):
        ours.setShDegree(sh), ref.set_sh_degree(sh)
        assert ours.active_sh_degree_ == ref.active_sh_degree()
    assert torch.equal(ours.getScalingActivation(), ref.get_scaling_activation())
    assert torch.equal(ours.getRotationActivation(), ref.get_rotation_activation())
    assert torch.equal(ours.getOpacityActivation(), ref.get_opacity_activation())
    assert torch.equal(ours.getFeatures(), ref.get_features())
    assert torch.equal(ours.getLanguageFeatures(), ref.get_language_features())
    co, cr = ours.getCovarianceActivation(1).detach(), ref.get_covariance_activation(1).detach()
    assert float((co - cr).abs().max()) <= 1e-6 * float(cr.abs().max())

    def same_params():
        names = ("xyz_", "features_dc_", "features_rest_", "language_features_", "opacity_", "scaling_", "rotation_")
        for name, r in zip(names, ref.params()):
            assert torch.equal(getattr(ours, name).detach(), r.detach()), name
        assert ours.params_are_the_optimizers()
        assert torch.equal(ours.exist_since_iter_, ref.exist_since_iter)
        assert torch.equal(ours.denom_, ref.denom) and torch.equal(ours.xyz_gradient_accum_, ref.xyz_gradient_accum)
        assert torch.equal(ours.max_radii2D_, ref.max_radii2D)

    same_params()
    # the reference's replaceTensorToOptimizer dereferences the optimizer's state entry of the tensor it replaces (:580-581),
    # which exists only after a first Adam::step -- resetOpacity before any step is a null dereference there.  A step on zero
    # gradients creates the entries and moves nothing (m = v = 0 -> update 0); the package's class needs no such entry.
    ref.set_grads([torch.zeros_like(x) for x in t])
    ref.step()
    same_params()
    ours.resetOpacity(), ref.reset_opacity()
    same_params()
    mask = torch.rand(n, generator=g) < 0.3
    ours.prunePoints(mask), ref.prune_points(mask)
    same_params()
    assert ours.xyz_.shape[0] == n - int(mask.sum())


def test_restatements_match_the_reference_model_golden():
    """tests/golden/model.npz -- inputs and outputs of the UNMODIFIED reference class (tests/golden/make_model_golden.py; stored
    so that the pin survives where oracle/_ref/ref_model.so cannot be built): the C oracle's Adam on trainingSetup's rates
    reproduces its two steps (1e-6), oracle/densify_ref.py reproduces addDensificationStats x 3 -> densifyAndPrune ->
    resetOpacity bit for bit from the stored state, and the mapper's schedule returns the stored float at every step."""
    import oracle as O
    from conftest import golden
    from leg_slam_b200 import mapper as M
    G = golden("model")
    for case in ("a", "b"):
        lrs = G[f"{case}_lrs"]
        for i, k in enumerate(DR.PARAMS):
            p = G[f"{case}_init_{k}"].copy().reshape(-1)
            m, v = np.zeros_like(p), np.zeros_like(p)
            for s in range(2):
                O.adam(p, G[f"{case}_grad{s}_{k}"].reshape(-1), m, v, float(lrs[i]), step=s + 1)
            for got, want in ((p, G[f"{case}_stepped_p_{k}"]), (m, G[f"{case}_stepped_m_{k}"]), (v, G[f"{case}_stepped_v_{k}"])):
                assert float(np.abs(got - want.reshape(-1)).max()) <= 1e-6 * float(np.abs(want).max()), (case, k)
            assert int(G[f"{case}_stepped_step_{k}"]) == 2
        t = torch.from_numpy
        ours = DR.Model({k: t(G[f"{case}_stepped_p_{k}"]) for k in DR.PARAMS}, percent_dense=OPT["percent_dense"])
        for k in DR.PARAMS:
            ours.m[k], ours.v[k] = t(G[f"{case}_stepped_m_{k}"]).clone(), t(G[f"{case}_stepped_v_{k}"]).clone()
        ours.exist_since_iter = t(G[f"{case}_stepped_exist"]).clone()
        for vw in range(3):
            ours.add_stats(t(G[f"{case}_view{vw}_radii"]), t(G[f"{case}_view{vw}_grad"]))
        for got, key in ((ours.xyz_gradient_accum, "accum"), (ours.denom, "denom"), (ours.max_radii2D, "max_radii")):
            assert torch.equal(got, t(G[f"{case}_stats_{key}"])), (case, key)
        max_grad, min_opacity, extent, mss = G[f"{case}_args"]
        ours.densify_and_prune(float(max_grad), float(min_opacity), float(extent), int(mss),
                               seeded_normal01(int(G[f"{case}_normal_seed"])))
        for k in DR.PARAMS:
            assert torch.equal(ours.p[k], t(G[f"{case}_densified_p_{k}"])), (case, k)
            assert torch.equal(ours.m[k], t(G[f"{case}_densified_m_{k}"])) and torch.equal(ours.v[k], t(G[f"{case}_densified_v_{k}"]))
        assert ours.p["xyz"].shape[0] != G[f"{case}_stepped_p_xyz"].shape[0]
        assert torch.equal(ours.exist_since_iter, t(G[f"{case}_densified_exist"]))
        assert not ours.denom.any() and not G[f"{case}_densified_denom"].any()
        ours.reset_opacity()
        assert torch.equal(ours.p["opacity"], t(G[f"{case}_reset_p_opacity"]))
        assert not ours.m["opacity"].any() and not G[f"{case}_reset_m_opacity"].any() and int(G[f"{case}_reset_step_opacity"]) == 2
    f32 = lambda x: float(np.float32(x))  # noqa: E731
    for i in range(3):
        scale, max_steps = G[f"lr{i}_cfg"]
        sched = M.ExponLr(f32(OPT["position_lr_init"]) * float(scale), f32(OPT["position_lr_final"]) * float(scale),
                          f32(OPT["position_lr_delay_mult"]), int(max_steps))
        got = np.array([sched(int(s)) for s in G[f"lr{i}_steps"]])
        assert np.array_equal(got, G[f"lr{i}_values"]), i


@pytest.mark.parametrize("what", ["nothing_selected", "everything_cloned", "everything_split", "everything_pruned", "unseen"])
def test_densify_edge_cases_equal_reference_model(RM, what):
    """The corners of densifyAndPrune (reference :729-832): no Gaussian over the gradient threshold (empty clone and split
    sets), every Gaussian cloned, every Gaussian split, every Gaussian pruned (the model ends with zero rows), and statistics
    of Gaussians no view has seen (0 / 0 -> nan -> 0, :811-812) -- the restatement follows the reference class through each."""
    P = 64
    ref, ours, g = make_pair(RM, P, seed=200)
    accum, denom = torch.rand(P, 1, generator=g) * 4e-4 + 3e-4, torch.ones(P, 1)
    args = dict(max_grad=2e-4, min_opacity=0.005, extent=4.0, mss=0)
    if what == "nothing_selected":
        args["max_grad"] = 1.0
    elif what == "everything_cloned":
        args["extent"] = 1e6            # every scale is under percent_dense * extent
    elif what == "everything_split":
        args["extent"] = 1e-6           # every scale is over it (and the world-size prune stays off: mss = 0)
    elif what == "everything_pruned":
        args["min_opacity"] = 1.1
    elif what == "unseen":
        denom = torch.zeros(P, 1)
        accum = torch.zeros(P, 1)
        denom[::3], accum[::3] = 2.0, 9e-4
    for m in (ref, ours):
        m.xyz_gradient_accum, m.denom = accum.clone(), denom.clone()
    mr = torch.rand(P, generator=g) * 40
    ref.max_radii2D = mr.clone()
    ours.max_radii2D = mr.clone()
    torch.manual_seed(17)
    ref.densify_and_prune(args["max_grad"], args["min_opacity"], args["extent"], args["mss"])
    ours.densify_and_prune(args["max_grad"], args["min_opacity"], args["extent"], args["mss"], seeded_normal01(17))
    assert_same(ref, ours, 2)
    n = ours.p["xyz"].shape[0]
    if what == "everything_pruned":
        assert n == 0
    elif what == "nothing_selected":
        assert n <= P
    elif what == "everything_cloned":
        assert n > P
    elif what == "everything_split":
        assert n > P and not torch.equal(ours.p["scaling"][:1], ours.p["scaling"][:1] * 0)


def _adam_like_libtorch(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-15):
    """torch::optim::Adam::step for one tensor, statement for statement (torch/csrc/api/src/optim/adam.cpp): in place."""
    import math
    bc1 = 1 - math.pow(beta1, step)
    bc2 = 1 - math.pow(beta2, step)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


@pytest.mark.parametrize("white_background", [False, True])
def test_density_control_lines_of_the_mapper_over_a_run(RM, white_background):
    """The density-control block and the optimizer step of GaussianMapper::trainForOneIteration AS THEY STAND (reference
    src/gaussian_mapper.cpp:737-761, 793-797, cut out of the file at build time and compiled into a harness that carries the
    mapper's member names, oracle/ref_model_wrap.cpp RefDensityControl) drive the reference's own GaussianModel over a whole
    run with gradients arriving every iteration.  Beside it, leg_slam_b200.mapper.density_control_actions decides what
    oracle/densify_ref.py does to its model: the same statistics, the same densifications with the reference's size threshold,
    the same opacity resets, and -- as in the reference -- the optimizer step after them, which skips every tensor a
    densification or a reset has just replaced (it carries no gradient).  State bit-identical after every iteration."""
    from leg_slam_b200 import mapper as M
    P = 200
    cfg = M.DensityControlParams(densification_interval=10, opacity_reset_interval=30, densify_from_iter=20, densify_until_iter=75,
                                 densify_grad_threshold=2e-4, densify_min_opacity=0.05, prune_big_point_after_iter=40,
                                 white_background=white_background)
    n_iter, extent = 90, 4.0
    ref, ours, g = make_pair(RM, P, seed=300, steps=2)
    ref.zero_grad()
    dc = RM.DensityControl(ref, iterations=n_iter, densification_interval=cfg.densification_interval,
                           opacity_reset_interval=cfg.opacity_reset_interval, densify_from_iter=cfg.densify_from_iter,
                           densify_until_iter=cfg.densify_until_iter, densify_grad_threshold=cfg.densify_grad_threshold,
                           densify_min_opacity=cfg.densify_min_opacity, prune_big_point_after_iter=cfg.prune_big_point_after_iter,
                           white_background=white_background, cameras_extent=extent)
    steps = {k: 2 for k in DR.PARAMS}
    lrs = ref.lrs()
    seen = dict(densify=0, reset=0, big=0, stats=0)
    for it in range(1, n_iter + 1):
        n = ours.p["xyz"].shape[0]
        assert n > 0
        grads = [torch.randn(*ours.p[k].shape, generator=g) * 0.01 for k in DR.PARAMS]
        radii = torch.randint(-5, 40, (n,), generator=g, dtype=torch.int32).clamp_min(0)
        vs_grad = torch.randn(n, 3, generator=g) * 3e-4
        visible = radii > 0
        # the reference: backward has left gradients on the parameters; then its own lines
        ref.set_grads(grads)
        torch.manual_seed(1000 + it)
        dc.run(it, vs_grad, visible, radii)
        # the restatement under the mapper's decision function
        act = M.density_control_actions(it, cfg)
        stepped = set(M.optimizer_step_groups(act))
        assert set(M.PARAM_ORDER) == set(DR.PARAMS)
        if act["update_stats"]:
            ours.add_stats(radii, vs_grad)
            seen["stats"] += 1
        if act["densify"]:
            ours.densify_and_prune(cfg.densify_grad_threshold, cfg.densify_min_opacity, extent, act["size_threshold"],
                                   seeded_normal01(1000 + it))
            assert not stepped                   # all seven tensors were rebuilt: no gradient on them
            seen["densify"] += 1
            seen["big"] += act["size_threshold"] == 20
        if act["reset_opacity"]:
            ours.reset_opacity()
            assert "opacity" not in stepped
            seen["reset"] += 1
        if it < n_iter:                          # :793-797
            for i, k in enumerate(DR.PARAMS):
                if k in stepped:
                    steps[k] += 1
                    _adam_like_libtorch(ours.p[k], grads[i], ours.m[k], ours.v[k], steps[k], lrs[i])
        for i, k in enumerate(DR.PARAMS):
            assert torch.equal(ref.params()[i].detach(), ours.p[k]), (it, k)
            st = ref.moments(i)
            assert st[0] == steps[k], (it, k, st[0], steps[k])
            assert torch.equal(st[1], ours.m[k]) and torch.equal(st[2], ours.v[k]), (it, k)
        assert torch.equal(ref.exist_since_iter, ours.exist_since_iter)
        for a, b in ((ref.xyz_gradient_accum, ours.xyz_gradient_accum), (ref.denom, ours.denom), (ref.max_radii2D, ours.max_radii2D)):
            assert torch.equal(a, b), it
    assert seen["stats"] == cfg.densify_until_iter - 1 and seen["densify"] == 5 and seen["big"] == 3
    assert seen["reset"] == (3 if white_background else 2)      # 30, 60 (+ densify_from_iter = 20 on a white background)
    assert ours.p["xyz"].shape[0] != P


def test_model_golden_is_what_its_generator_writes(RM, tmp_path):
    """tests/golden/model.npz is reproducible: tests/golden/make_model_golden.py, run again on the compiled reference class,
    writes the same arrays (so the committed fixture is the reference's output, not an edited one)."""
    import importlib.util
    from conftest import golden
    spec = importlib.util.spec_from_file_location("make_model_golden", os.path.join(HERE, "golden", "make_model_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    out = str(tmp_path / "model.npz")
    gen.main(out)
    a, b = golden("model"), np.load(out)
    assert sorted(a.files) == sorted(b.files) and len(a.files) > 200
    for k in a.files:
        assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), k
