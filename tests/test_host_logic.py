"""CPU: host-side logic of the reference-shaped interface (no kernels run)."""
import math
import os
import sys

import numpy as np
import pytest
import torch

import cases
from leg_slam_b200 import GaussianRasterizationSettings, GaussianRasterizer, synthetic
from leg_slam_b200 import _lib
from leg_slam_b200 import rasterize_points as rp


def _settings(cs):
    return GaussianRasterizationSettings(cs["H"], cs["W"], cs["tanfovx"], cs["tanfovy"], cs["bg"], 1.0, cs["viewmatrix"],
                                         cs["projmatrix"], cs["degree"], cs["campos"], False, True)


def test_rasterizer_argument_validation_matches_reference():
    """GaussianRasterizer::forward throws on ambiguous inputs (src/gaussian_rasterizer.cpp:196-206)."""
    cs = cases.make_case("sh3_lf")
    r = GaussianRasterizer(_settings(cs))
    m2d = torch.zeros_like(cs["means3D"])
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(cs["means3D"], m2d, cs["opacities"], scales=cs["scales"], rotations=cs["rotations"])
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(cs["means3D"], m2d, cs["opacities"], shs=cs["shs"], colors_precomp=cs["means3D"], scales=cs["scales"],
          rotations=cs["rotations"])
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(cs["means3D"], m2d, cs["opacities"], shs=cs["shs"], scales=cs["scales"])
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(cs["means3D"], m2d, cs["opacities"], shs=cs["shs"], scales=cs["scales"], rotations=cs["rotations"],
          cov3D_precomp=torch.zeros(cs["P"], 6))


def test_no_cpu_fallback():
    """CPU tensors are refused loudly: the hot path is CUDA only."""
    cs = cases.make_case("sh3_lf")
    with pytest.raises(_lib.LgsError, match="no CPU path"):
        rp.rasterize_gaussians(*cases.fwd_args(cs))
    with pytest.raises(ValueError, match="num_points, 3"):
        bad = dict(cs, means3D=cs["means3D"][:, :2])
        rp.rasterize_gaussians(*cases.fwd_args(bad))
    from leg_slam_b200 import cosine_query, FusedAdam
    with pytest.raises(_lib.LgsError):
        cosine_query(torch.zeros(4, 64), torch.zeros(64))
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.zeros(4)
    with pytest.raises(_lib.LgsError):
        FusedAdam([p], lr=1e-3).step()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "liblgs.so"))
    with pytest.raises(_lib.LgsError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_camera_conventions_match_reference():
    """viewmatrix = W2C^T, projmatrix = viewmatrix @ P^T, campos = inverse(viewmatrix)[3,:3]
    (src/gaussian_keyframe.cpp:111-193); the kernels read both column-major."""
    cam = synthetic.make_cameras(3, 640, 480, seed=5)[1]
    assert math.isclose(cam.tanfovx, 1.0, rel_tol=1e-6) and math.isclose(cam.tanfovy, 0.75, rel_tol=1e-6)
    w2c = cam.viewmatrix.t()
    R = w2c[:3, :3]
    assert torch.allclose(R @ R.t(), torch.eye(3), atol=1e-5) and torch.det(R) > 0
    assert torch.allclose(w2c[3], torch.tensor([0.0, 0.0, 0.0, 1.0]))
    # camera centre maps to the view-space origin
    c = torch.cat([cam.campos, torch.ones(1)])
    assert torch.allclose(w2c @ c, torch.tensor([0.0, 0.0, 0.0, 1.0]), atol=1e-5)
    # a point 2 m straight ahead projects to the image centre with w = z_view
    fwd = R[2]  # third row of R_w2c = camera z axis in world coordinates
    pt = torch.cat([cam.campos + 2.0 * fwd, torch.ones(1)])
    hom = pt @ cam.projmatrix  # row-vector convention of the stored (transposed) matrix
    assert abs(float(hom[0] / hom[3])) < 1e-5 and abs(float(hom[1] / hom[3])) < 1e-5
    assert math.isclose(float(hom[3]), 2.0, rel_tol=1e-5)
    P = synthetic.projection_matrix(0.01, 100.0, 2 * math.atan(1.0), 2 * math.atan(0.75))
    assert math.isclose(float(P[0, 0]), 1.0, rel_tol=1e-6) and math.isclose(float(P[1, 1]), 1 / 0.75, rel_tol=1e-6)
    assert float(P[3, 2]) == 1.0 and math.isclose(float(P[2, 2]), 100.0 / 99.99, rel_tol=1e-6)


def test_scene_generator_is_deterministic_and_shaped():
    a = synthetic.make_scene(5000, seed=3)
    b = synthetic.make_scene(5000, seed=3)
    c = synthetic.make_scene(5000, seed=4)
    for k in a:
        assert torch.equal(a[k], b[k])
    assert not torch.equal(a["xyz"], c["xyz"])
    assert a["features_dc"].shape == (5000, 1, 3) and a["features_rest"].shape == (5000, 15, 3)
    assert a["lang_feat"].shape == (5000, 64) and a["rotation"].shape == (5000, 4)
    act = synthetic.activate(a)
    assert act["shs"].shape == (5000, 16, 3)
    assert torch.allclose(act["rotations"].norm(dim=1), torch.ones(5000), atol=1e-5)
    assert (act["opacities"] > 0).all() and (act["opacities"] < 1).all() and (act["scales"] > 0).all()
    assert (a["xyz"].min(0).values > -0.2).all() and (a["xyz"].max(0).values < torch.tensor([6.2, 4.2, 3.0])).all()


def test_cov3d_precomp_helper_matches_oracle(oracle_mod):
    """tests/cases.covariance_from_scale_rot restates forward.cu:118-152; the oracle's cov3D agrees."""
    cs = cases.make_case("sh3_lf")
    f = cases.oracle_forward(cs, oracle_mod)
    vis = f["radii"] > 0
    cov = cases.covariance_from_scale_rot(cs["scales"], cs["rotations"]).numpy()
    assert cases.rel_err(cov[vis], f["cov3D"][vis]) <= 1e-5


def test_dp_range_partition_covers_the_flat_index_space_once():
    """leg_slam_b200.dp._Range: the pieces of the flat index space (other tensors | early / late part of the language
    features) sharded over the ranks are disjoint, 16-byte aligned and cover every index exactly once."""
    from leg_slam_b200.dp import _Range
    dev = torch.device("cpu")
    for world in (1, 2, 3, 4, 8):
        for pieces in ([(0, 123 * 40, 0)], [(0, 51 * 40, 0), (51 * 40, 51 * 40 + 1024, 1), (51 * 40 + 1024, 115 * 40, 2), (115 * 40, 123 * 40, 0)],
                       [(0, 8, 0), (8, 12, 2), (12, 16, 0)]):
            n = pieces[-1][1]
            seen = np.zeros(n, np.int32)
            for rank in range(world):
                for b, e, ph in pieces:
                    r = _Range(b, e, ph, world, rank, dev)
                    assert r.sb % 4 == 0 and r.se % 4 == 0 and b <= r.sb <= r.se <= e
                    assert r.exp_avg.numel() >= max(r.se - r.sb, 4) and r.exp_avg_sq.numel() == r.exp_avg.numel()
                    seen[r.sb:r.se] += 1
            assert (seen == 1).all(), (world, pieces)


def test_position_lr_schedule_matches_the_c_statement(oracle_mod):
    """Mapper's ExponLr against the float / libm statement of GaussianModel::exponLrFunc (reference
    src/gaussian_model.cpp:1143-1157) in the C oracle, plus the closed-form ends: lr_init at step 0, lr_final from max_steps
    on, the geometric mean half way, 0 for a negative step, and the delay factor's ends when it is enabled."""
    from leg_slam_b200.mapper import ExponLr
    init, final = 1.6e-4 * 5.3, 1.6e-6 * 5.3
    e = ExponLr(init, final, 0.01, 30_000)
    for step in (0, 1, 17, 999, 15_000, 29_999, 30_000, 31_000, 10 ** 6):
        ref = oracle_mod.expon_lr(step, init, final, 0.01, 0, 30_000)
        assert abs(e(step) - ref) <= 1.2e-7 * ref, (step, e(step), ref)  # one float ulp: numpy's and libm's logf / expf
    assert abs(e(0) - init) <= 3e-7 * init and abs(e(30_000) - final) <= 3e-7 * final and e(45_000) == e(30_000)
    assert abs(e(15_000) - math.sqrt(init * final)) <= 1e-6 * math.sqrt(init * final)
    assert e(-1) == 0.0 and ExponLr(0.0, 0.0)(5) == 0.0
    vals = [e(s) for s in range(0, 30_001, 1000)]
    assert all(a > b for a, b in zip(vals, vals[1:]))
    d = ExponLr(1e-2, 1e-4, 0.01, 1000, lr_delay_steps=100)
    for step in (0, 1, 50, 99, 100, 500):
        ref = oracle_mod.expon_lr(step, 1e-2, 1e-4, 0.01, 100, 1000)
        assert abs(d(step) - ref) <= 2.4e-7 * ref, (step, d(step), ref)
    assert abs(d(0) - 1e-4) <= 1e-10


def test_mapper_per_iteration_settings():
    """What trainForOneIteration sets before it renders (reference src/gaussian_mapper.cpp:662-683 on
    src/gaussian_model.cpp:100-107,520-565): learning rates reach the optimizer's groups (feature_rest at a twentieth, xyz
    times spatial_lr_scale), the xyz schedule, and the active SH degree clamps at the model's."""
    from leg_slam_b200 import mapper as M
    sc = synthetic.make_scene(50, seed=3)
    mp = M.Mapper(sc, sh_degree=3)
    lr_of = lambda k: mp.optimizer.param_groups[M.PARAM_ORDER.index(k)]["lr"]  # noqa: E731
    assert lr_of("xyz") == M.DEFAULT_LRS["xyz"]
    mp.set_position_lr_schedule(1.6e-4, 1.6e-6, 0.01, 30_000, spatial_lr_scale=5.0)
    assert abs(lr_of("xyz") - 8e-4) < 1e-12
    assert mp.update_learning_rate(30_000) == lr_of("xyz") and abs(lr_of("xyz") - 8e-6) < 1e-11
    mp.set_position_learning_rate(2e-4)
    assert abs(lr_of("xyz") - 1e-3) < 1e-12 and mp.learning_rate("xyz") == lr_of("xyz")
    mp.set_feature_learning_rate(2.5e-3)
    assert lr_of("features_dc") == 2.5e-3 and lr_of("features_rest") == 2.5e-3 / 20.0
    mp.set_language_feature_learning_rate(1e-3), mp.set_opacity_learning_rate(0.04), mp.set_scaling_learning_rate(4e-3)
    mp.set_rotation_learning_rate(2e-3)
    assert [lr_of(k) for k in ("lang_feat", "opacity", "scaling", "rotation")] == [1e-3, 0.04, 4e-3, 2e-3]
    with pytest.raises(ValueError, match="set_position_lr_schedule"):
        M.Mapper(sc, sh_degree=3).update_learning_rate(1)
    mp.set_sh_degree(0)
    assert mp.sh_degree == 0
    for want in (1, 2, 3, 3):
        mp.one_up_sh_degree()
        assert mp.sh_degree == want
    mp.set_sh_degree(7)
    assert mp.sh_degree == 3 and mp.max_sh_degree == 3


def test_density_control_cadence_matches_reference_conditions():
    """density_control_actions against the conditions of trainForOneIteration written out (reference
    src/gaussian_mapper.cpp:737-761), over every iteration of a run, for the default and the Replica settings; and
    Mapper.density_control calls densify_and_prune / reset_opacity exactly then, with the reference's arguments."""
    from leg_slam_b200 import mapper as M
    replica = M.DensityControlParams(densification_interval=100, opacity_reset_interval=0, densify_from_iter=600, densify_until_iter=15_000,
                                     densify_grad_threshold=0.001, densify_min_opacity=0.02, prune_big_point_after_iter=30_000)
    white = M.DensityControlParams(white_background=True)
    for p in (M.DensityControlParams(), replica, white):
        for it in list(range(0, 1300)) + list(range(2900, 3200)) + list(range(14_890, 15_110)):
            a = M.density_control_actions(it, p)
            active = it < p.densify_until_iter
            densify = active and it > p.densify_from_iter and it % p.densification_interval == 0
            reset = active and bool(p.opacity_reset_interval) and (it % p.opacity_reset_interval == 0 or
                                                                   (p.white_background and it == p.densify_from_iter))
            assert a == dict(update_stats=active, densify=densify, size_threshold=(20 if it > p.prune_big_point_after_iter else 0) if densify else 0,
                             reset_opacity=reset), (it, p)
    assert M.density_control_actions(700, replica)["densify"] and M.density_control_actions(700, replica)["size_threshold"] == 0
    assert not M.density_control_actions(600, replica)["densify"] and not M.density_control_actions(15_000, replica)["update_stats"]
    assert M.density_control_actions(3000, M.DensityControlParams()) == dict(update_stats=True, densify=True, size_threshold=20, reset_opacity=True)
    assert M.density_control_actions(500, white)["reset_opacity"] and not M.density_control_actions(500, M.DensityControlParams())["reset_opacity"]

    calls = []

    class Recorder(M.Mapper):
        def densify_and_prune(self, *a, **kw):
            calls.append(("densify", a, kw))
            return dict(new_P=1)

        def reset_opacity(self):
            calls.append(("reset",))
    mp = Recorder(synthetic.make_scene(20, seed=1), sh_degree=3)
    gen = torch.Generator()
    assert mp.density_control(650, replica, 4.5) == dict(update_stats=True, densify=False, size_threshold=0, reset_opacity=False) and not calls
    act = mp.density_control(700, replica, 4.5, generator=gen)
    assert act["densify"] and act["info"] == dict(new_P=1)
    assert calls == [("densify", (0.001, 0.02, 4.5, 0), dict(generator=gen))]
    calls.clear()
    mp.density_control(3000, M.DensityControlParams(), 2.0)
    assert [c[0] for c in calls] == ["densify", "reset"] and calls[0][1] == (0.0002, 0.005, 2.0, 20)


REF_EVAL = "/root/reference/eval"


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF_EVAL, "utils.py")), reason="reference tree not present")
def test_synthetic_cameras_equal_the_reference_minicam(monkeypatch):
    """The camera tensors every test and bench renders with (leg_slam_b200.synthetic.camera_from_pose) against the reference's
    own Python camera -- eval/utils.py: get_world2view (:67-81), focal2fov (:83-84), MiniCam (:10-48: transposed view matrix,
    projection from the FoV only, full_proj = view @ P^T, centre from the inverse), imported unmodified (its unrelated imports
    -- clip, cv2, torchvision -- replaced by empty modules, `.cuda()` by the identity: there is no GPU here)."""
    import importlib.util
    import types
    for name in ("clip", "cv2", "torchvision", "torchvision.transforms"):
        if name not in sys.modules:
            monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    spec = importlib.util.spec_from_file_location("ref_eval_utils", os.path.join(REF_EVAL, "utils.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for (W, H, fx, fy, seed) in ((640, 480, 320.0, 320.0, 5), (1296, 968, 1169.7, 1169.7, 6), (320, 240, 160.0, 171.5, 7)):
        g = torch.Generator().manual_seed(seed)
        c = torch.rand(3, generator=g, dtype=torch.float64) * 2 + 1
        tgt = torch.rand(3, generator=g, dtype=torch.float64) * 4
        R = synthetic.look_at(tuple(c.tolist()), tuple(tgt.tolist()))
        cam = synthetic.camera_from_pose(R, c, W, H, fx, fy)
        fovx, fovy = ref.focal2fov(fx, W), ref.focal2fov(fy, H)
        mini = ref.MiniCam(W, H, fovx, fovy, ref.get_world2view(R.numpy(), c.numpy()))
        assert cam.tanfovx == math.tan(fovx * 0.5) and cam.tanfovy == math.tan(fovy * 0.5)   # what eval/render.py:29-30 passes
        # the projection has no pose in it: bit-identical
        P_ours = synthetic.projection_matrix(0.01, 100.0, fovx, fovy)
        assert torch.equal(P_ours, ref.MiniCam.get_projection_matrix(0.01, 100.0, fovx, fovy))
        # the view matrix: R^T and -R^T c written out here, a 4x4 inverse in double there -- equal after rounding to float
        # up to the last bit
        for ours, theirs in ((cam.viewmatrix, mini.world_view_transform), (cam.projmatrix, mini.full_proj_transform),
                             (cam.campos, mini.camera_center)):
            assert ours.shape == theirs.shape and ours.dtype == theirs.dtype == torch.float32
            assert float((ours - theirs).abs().max()) <= 1e-6 * max(1.0, float(theirs.abs().max()))
        assert float((cam.campos.double() - c).abs().max()) <= 1e-6 * float(c.abs().max())
