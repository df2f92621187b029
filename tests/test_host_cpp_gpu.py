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
