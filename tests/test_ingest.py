"""Keyframe-ingest geometry: reprojectDepthPinhole / transformPoints / distCUDA2 on
liblgs.so against the numpy restatement (oracle/ingest_ref.py) and the compiled, unmodified reference simple-knn
(oracle/_ref/ref_simple_knn.so)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import ingest_ref as IR  # noqa: E402


def test_restatement_small_cases_cpu():
    # collinear points: the three other points are the three nearest
    p = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [5, 0, 0]], np.float32)
    d = IR.knn_mean_dist2(p)
    np.testing.assert_allclose(d, [(1 + 4 + 25) / 3, (1 + 1 + 16) / 3, (1 + 4 + 9) / 3, (9 + 16 + 25) / 3], rtol=1e-6)
    pts = IR.reproject_depth_pinhole(np.array([2.0, 0.0, 4.0, 1.0], np.float32), [True, False, True, True], (2.0, 4.0, 0.5, 0.5), 2)
    np.testing.assert_allclose(pts, [[-0.5, -0.25, 2.0], [0, 0, 0], [-1.0, 0.5, 4.0], [0.25, 0.125, 1.0]])
    T = np.eye(4, dtype=np.float32)
    T[3, :3] = [1, 2, 3]  # stored transposed: translation in the last ROW (elements 12..14)
    np.testing.assert_allclose(IR.transform_points(np.array([[1, 1, 1]], np.float32), T), [[2, 3, 4]])


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.gpu
def test_reproject_and_transform_match_restatement(dev):
    from leg_slam_b200 import ingest
    g = torch.Generator().manual_seed(3)
    W, H = 64, 48
    depth = torch.rand(W * H, generator=g) * 5 + 0.1
    mask = torch.rand(W * H, generator=g) > 0.3
    intr = (60.0, 61.5, 31.5, 23.5)
    pts = ingest.reprojectDepthPinhole(depth.to(dev), mask.to(dev), intr, W)
    ref = IR.reproject_depth_pinhole(depth.numpy(), mask.numpy(), intr, W)
    np.testing.assert_array_equal(pts.cpu().numpy(), ref)  # same IEEE operations in the same order
    q = torch.randn(4, generator=g)
    q = q / q.norm()
    r, x, y, z = q.tolist()
    R = torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)],
                      [2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)],
                      [2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]])
    T = torch.eye(4)
    T[:3, :3] = R
    T[:3, 3] = torch.tensor([0.3, -1.2, 2.0])
    Tt = T.t().contiguous()  # the reference stores its pose tensors transposed
    out = ingest.transformPoints(pts, Tt.to(dev))
    ref2 = IR.transform_points(ref, Tt.numpy())
    np.testing.assert_allclose(out.cpu().numpy(), ref2, rtol=2e-6, atol=2e-6)
    with pytest.raises(ValueError):
        ingest.transformPoints(torch.zeros(5, 2, device=dev), Tt.to(dev))
    with pytest.raises(ValueError):
        ingest.reprojectDepthPinhole(torch.zeros(5, 2, device=dev), mask.to(dev), intr, W)


@pytest.mark.gpu
@pytest.mark.parametrize("P", [4, 37, 1500])
def test_knn_matches_bruteforce(P, dev):
    from leg_slam_b200 import ingest
    g = torch.Generator().manual_seed(P)
    pts = torch.randn(P, 3, generator=g) * torch.tensor([3.0, 2.0, 1.4])
    d = ingest.distCUDA2(pts.to(dev)).cpu().numpy()
    ref = IR.knn_mean_dist2(pts.numpy())
    np.testing.assert_allclose(d, ref, rtol=3e-7, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("P", [5000, 300_000])
def test_knn_bit_exact_vs_compiled_reference(P, dev):
    """distCUDA2 vs the unmodified reference simple-knn on the same device buffer: bit-identical (the mean of the
    three smallest squared distances does not depend on the search order)."""
    import build_ref
    from leg_slam_b200 import ingest
    try:
        lib = build_ref.load_knn()
    except FileNotFoundError as ex:
        pytest.skip(str(ex))
    g = torch.Generator().manual_seed(P + 1)
    # room-like: points on surfaces + duplicates (zero distances)
    pts = torch.rand(P, 3, generator=g) * torch.tensor([6.0, 4.0, 2.8])
    pts[: P // 3, 2] = 0.0
    pts[P // 3: P // 3 + 50] = pts[:50]
    pts = pts.to(dev).contiguous()
    ours = ingest.distCUDA2(pts)
    ref = torch.zeros(P, device=dev)
    torch.cuda.synchronize()
    assert lib.ref_simple_knn(P, pts.data_ptr(), ref.data_ptr()) == 0
    assert torch.equal(ours, ref)


# ---- the two neighbouring operators: loop-closure correction, inactive-geometry densification ---------------------------
def _pose(g, t=(0.3, -1.2, 2.0)):
    """A rigid transform as the reference stores its pose tensors: transposed (translation in the last row)."""
    q = torch.randn(4, generator=g)
    q = q / q.norm()
    r, x, y, z = q.tolist()
    R = torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)],
                      [2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)],
                      [2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]])
    T = torch.eye(4)
    T[:3, :3] = R
    T[:3, 3] = torch.tensor(t)
    return T.t().contiguous()


def _loop_closure_case(P, seed):
    g = torch.Generator().manual_seed(seed)
    pts = (torch.rand(P, 3, generator=g) - 0.5) * torch.tensor([6.0, 4.0, 6.0])
    rots = torch.randn(P, 4, generator=g)
    rots = rots / rots.norm(dim=1, keepdim=True)
    n4 = min(P, 40)
    # rows that take the three trace <= 0 branches of the matrix -> quaternion conversion: half turns about x / y / z
    # (combined with a near-identity correction T below)
    rots[:n4] = torch.tensor([[0.02, 1.0, 0.01, 0.03], [0.01, 0.02, 1.0, 0.03], [0.03, 0.01, 0.02, 1.0], [1.0, 0.0, 0.0, 0.0]]).repeat(n4 // 4 + 1, 1)[:n4]
    nt = torch.rand(P, generator=g) > 0.3
    un = torch.rand(P, generator=g) > 0.3
    return pts.contiguous(), rots.contiguous(), nt, un


def _quat_close(a, b, tol):
    """Quaternion rows equal up to `tol` per component, relative to the row's largest component."""
    if a.size == 0:
        return a.shape == b.shape
    scale = np.maximum(np.abs(b).max(axis=1, keepdims=True), 1e-30)
    return float((np.abs(a - b) / scale).max()) <= tol


def test_loop_closure_restatement_cpu():
    """Known answers: identity correction leaves a unit quaternion and the point where they were (up to the shipped
    (w, x, z, 0) row layout); a quarter turn about z composes as quaternions do; culled / unflagged rows are untouched."""
    pts = np.array([[0, 0, 1.0], [0, 0, 2.0], [0, 0, -1.0], [1, 2, 3.0]], np.float32)
    rots = np.array([[1, 0, 0, 0], [0.5, 0.5, 0.5, 0.5], [1, 0, 0, 0], [0, 0, 0, 1.0]], np.float32)
    nt = np.array([True, True, True, False])
    un = np.array([True, True, True, True])
    I = np.eye(4, dtype=np.float32)
    p2, r2, nt2, n = IR.scale_and_transform_then_mark_visible(pts, rots, nt, un, I, I, scale=2.0)
    assert n == 2 and list(nt2) == [False, False, True, False]
    np.testing.assert_allclose(p2, [[0, 0, 2], [0, 0, 4], [0, 0, -1], [1, 2, 3]])
    np.testing.assert_allclose(r2, [[1, 0, 0, 0], [0.5, 0.5, 0.5, 0], [1, 0, 0, 0], [0, 0, 0, 1]], atol=1e-6)
    c = np.float32(np.sqrt(0.5))
    Rz = np.array([[0, -1, 0, 0], [1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32).T.copy()  # stored transposed
    q = IR.quaternion_through_matrix(np.array([[1, 0, 0, 0], [0, 1, 0, 0]], np.float32), Rz)
    np.testing.assert_allclose(q, [[c, 0, 0, c], [0, c, c, 0]], atol=1e-6)


def test_inactive_geo_restatement_cpu():
    px = np.array([[2, 1], [5, 1], [3, 1], [9, 9], [0, 0]], np.float32)
    has = np.array([True, True, False, False, False])
    p3 = np.array([[0.1, 0.2, 2.0], [0.3, 0.4, 4.0], [9, 9, 9], [9, 9, 9], [9, 9, 9]], np.float32)
    colors = np.arange(200, dtype=np.float32)
    pts, col = IR.inactive_geo_densify(px, has, p3, colors, 8.0, (2.0, 4.0, 0.5, 0.5), 10)
    # keypoint 2 is 1 px from keypoint 0 (depth 2) and 2 px from keypoint 1; keypoint 3 has nothing within sqrt(8) px and is
    # dropped; keypoint 4 is at squared distance 5 from keypoint 0
    np.testing.assert_allclose(pts, [[0.1, 0.2, 2.0], [0.3, 0.4, 4.0], [(3 - 0.5) * 2 / 2, (1 - 0.5) * 2 / 4, 2.0],
                                     [(0 - 0.5) * 2 / 2, (0 - 0.5) * 2 / 4, 2.0]])
    np.testing.assert_allclose(col, [[12, 13, 14], [15, 16, 17], [13, 14, 15], [0, 1, 2]])


def test_restatement_matches_reference_golden_cpu():
    """The numpy restatement against outputs of the UNMODIFIED reference operators (tests/golden/geometry.npz, written on a
    B200 by tests/golden/make_geometry_golden.py from oracle/_ref/ref_geometry.so): depth reprojection bit-exact, point
    transform within 2 ulp-scale (the restatement evaluates in float64), the loop-closure selection / count / flags exact with
    points and rotation rows within 1e-5, the inactive-geometry densification's selection exact (same rows, same colours)
    with reprojected points within 1e-6 -- for fractional pixels and for integer pixels with distance ties."""
    G = np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))
    pts = IR.reproject_depth_pinhole(G["rp_depth"], G["rp_mask"], G["rp_intr"], int(G["rp_width"]))
    np.testing.assert_array_equal(pts, G["rp_points"])
    np.testing.assert_allclose(IR.transform_points(G["rp_points"], G["tp_T"]), G["tp_out"], rtol=2e-6, atol=2e-6)
    p2, r2, nt2, num = IR.scale_and_transform_then_mark_visible(G["lc_points"], G["lc_rots"], G["lc_not_transformed"], G["lc_unstable"],
                                                               G["lc_T"], G["lc_view"], float(G["lc_scale"]))
    assert num == int(G["lc_num"]) and 0 < num < G["lc_points"].shape[0]
    np.testing.assert_array_equal(nt2, G["lc_out_not_transformed"])
    np.testing.assert_allclose(p2, G["lc_out_points"], rtol=3e-6, atol=3e-6)
    sel = G["lc_not_transformed"] & ~G["lc_out_not_transformed"]
    assert sel.sum() == num and np.all(G["lc_out_rots"][sel][:, 3] == 0)          # the shipped (w, x, z, 0) rows
    np.testing.assert_array_equal(r2[~sel], G["lc_out_rots"][~sel])
    assert _quat_close(r2[sel], G["lc_out_rots"][sel], 1e-5)
    for tag in ("ig", "igi"):
        qp, qc = IR.inactive_geo_densify(G[f"{tag}_pixels"], G[f"{tag}_has3D"], G[f"{tag}_points"], G[f"{tag}_colors"],
                                         float(G[f"{tag}_maxd"]), G[f"{tag}_intr"], int(G[f"{tag}_width"]))
        assert qp.shape == G[f"{tag}_out_points"].shape and 0 < qp.shape[0] < G[f"{tag}_pixels"].shape[0], tag
        np.testing.assert_array_equal(qc, G[f"{tag}_out_colors"])
        np.testing.assert_allclose(qp, G[f"{tag}_out_points"], rtol=1e-6, atol=1e-7)
        np.testing.assert_array_equal(qp[:, 2], G[f"{tag}_out_points"][:, 2])  # borrowed depths: the same neighbours were chosen


@pytest.fixture(scope="module")
def ref_geo(dev):
    import build_ref
    try:
        return build_ref.load_geometry()
    except FileNotFoundError as ex:
        pytest.skip(str(ex))


@pytest.mark.gpu
def test_reproject_and_transform_bit_exact_vs_compiled_reference(dev, ref_geo):
    """SURVEY.md 8f row 4 pinned by the unmodified src/stereo_vision.cu / src/operate_points.cu (these two had
    only a numpy restatement behind them)."""
    from leg_slam_b200 import ingest
    g = torch.Generator().manual_seed(11)
    for (W, H) in ((64, 48), (1296, 968)):
        depth = (torch.rand(W * H, generator=g) * 5 + 0.1).to(dev)
        mask = (torch.rand(W * H, generator=g) > 0.3).to(dev)
        intr = [584.87 * W / 648.0, 585.1 * W / 648.0, W / 2 - 0.5, H / 2 - 0.5]
        ours = ingest.reprojectDepthPinhole(depth, mask, intr, W)
        ref = ref_geo.reproject_depth_pinhole(depth, mask, intr, W)
        assert torch.equal(ours, ref)
        Tt = _pose(g).to(dev)
        assert torch.equal(ingest.transformPoints(ours, Tt), ref_geo.transform_points(ref.clone(), Tt))


@pytest.mark.gpu
@pytest.mark.parametrize("P,scale", [(1, 1.0), (257, 1.0), (20000, 1.0), (500000, 1.07)])
def test_loop_closure_correction_vs_compiled_reference(P, scale, dev, ref_geo):
    """scaleAndTransformThenMarkVisiblePoints: selection, count, points and the cleared flags bit-identical to the unmodified
    reference, and so are the corrected rotation rows (w, x, z, 0) -- the kernel spells out the fused multiply-adds of the
    reference build; untouched rows untouched; and the (w, x, y, z) variant against the numpy restatement."""
    from leg_slam_b200 import ingest
    pts, rots, nt, un = _loop_closure_case(P, 100 + P)
    g = torch.Generator().manual_seed(7)
    # a small correction (loop closure moves a keyframe by centimetres and a fraction of a degree): keeps the half-turn
    # rows on the trace <= 0 branches; the big case uses a generic pose
    T = _pose(g) if P > 20000 else (torch.eye(4) + 0.01 * torch.randn(4, 4, generator=g)).contiguous()
    T[:, 3] = torch.tensor([0.0, 0.0, 0.0, 1.0])
    view = _pose(g, t=(0.1, 0.2, 0.5))
    proj = torch.eye(4)
    a = [t.clone().to(dev) for t in (pts, rots, nt, un)]
    b = [t.clone().to(dev) for t in (pts, rots, nt, un)]
    n_ref = ref_geo.scale_and_transform_then_mark_visible(b[0], b[1], b[2], b[3], T.to(dev), view.to(dev), proj.to(dev), 5, scale)
    n_ours = ingest.scaleAndTransformThenMarkVisiblePoints(a[0], a[1], a[2], a[3], T.to(dev), view.to(dev), proj.to(dev), 5, scale)
    assert n_ours == n_ref
    assert torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
    assert torch.equal(a[0], b[0])
    sel = (nt.to(dev) & ~a[2])
    assert int(sel.sum()) == n_ref - 5
    if P > 1:
        assert 0 < int(sel.sum()) < P
    assert torch.equal(a[1][~sel], rots.to(dev)[~sel])
    ro, rr = a[1][sel].cpu().numpy(), b[1][sel].cpu().numpy()
    assert np.all(ro[:, 3] == 0) and np.all(rr[:, 3] == 0)  # the shipped row layout
    assert np.array_equal(ro, rr), float(np.abs(ro - rr).max())
    # restatement (CPU, float64 inside) on a sample, both rotation layouts
    k = min(P, 3000)
    p2, r2, nt2, n2 = IR.scale_and_transform_then_mark_visible(pts[:k].numpy(), rots[:k].numpy(), nt[:k].numpy(), un[:k].numpy(),
                                                              T.numpy(), view.numpy(), scale)
    assert np.array_equal(nt2, a[2][:k].cpu().numpy())
    np.testing.assert_allclose(a[0][:k].cpu().numpy(), p2, rtol=3e-6, atol=3e-6)
    assert _quat_close(a[1][:k].cpu().numpy(), r2, 2e-5)
    c = [t.clone().to(dev) for t in (pts[:k], rots[:k], nt[:k], un[:k])]
    ingest.scaleAndTransformThenMarkVisiblePoints(c[0], c[1], c[2], c[3], T.to(dev), view.to(dev), proj.to(dev), 0, scale,
                                                  faithful_rot_store=False)
    s = sel[:k].cpu().numpy()
    full = IR.quaternion_through_matrix(rots[:k].numpy()[s], T.numpy())
    assert _quat_close(c[1].cpu().numpy()[s], full, 2e-5)


@pytest.mark.gpu
def test_loop_closure_argument_checks(dev):
    from leg_slam_b200 import ingest
    I = torch.eye(4, device=dev)
    with pytest.raises(ValueError, match="points must have dimensions"):
        ingest.scaleAndTransformThenMarkVisiblePoints(torch.zeros(5, 2, device=dev), torch.zeros(5, 4, device=dev),
                                                      torch.ones(5, dtype=torch.bool, device=dev),
                                                      torch.ones(5, dtype=torch.bool, device=dev), I, I, I)
    with pytest.raises(ValueError, match="points_mask must have dimensions"):
        ingest.scaleAndTransformThenMarkVisiblePoints(torch.zeros(5, 3, device=dev), torch.zeros(5, 4, device=dev),
                                                      torch.ones(4, dtype=torch.bool, device=dev),
                                                      torch.ones(5, dtype=torch.bool, device=dev), I, I, I)
    assert ingest.scaleAndTransformThenMarkVisiblePoints(torch.zeros(0, 3, device=dev), torch.zeros(0, 4, device=dev),
                                                         torch.ones(0, dtype=torch.bool, device=dev),
                                                         torch.ones(0, dtype=torch.bool, device=dev), I, I, I, 3) == 3


def _keypoint_case(N, W, H, seed, frac3d=0.4, integer_pixels=False):
    g = torch.Generator().manual_seed(seed)
    px = torch.rand(N, 2, generator=g) * torch.tensor([W - 1.0, H - 1.0])
    if integer_pixels:
        px = px.floor()
    has = torch.rand(N, generator=g) < frac3d
    p3 = torch.randn(N, 3, generator=g)
    p3[:, 2] = torch.rand(N, generator=g) * 6 - 0.5   # some keypoints carry a non-positive depth
    colors = torch.rand(H * W * 3, generator=g)
    return px.contiguous(), has, p3.contiguous(), colors


@pytest.mark.gpu
@pytest.mark.parametrize("N,W,H,maxd,ints", [(1, 64, 48, 100.0, False), (300, 64, 48, 30.0, True), (1000, 640, 480, 400.0, False),
                                             (6000, 1296, 968, 900.0, False), (2500, 640, 480, 1e9, True), (50, 64, 48, 0.0, True)])
def test_inactive_geo_densify_bit_exact_vs_compiled_reference(N, W, H, maxd, ints, dev, ref_geo):
    """monocularPinholeInactiveGeoDensify...: returned points and colours bit-identical to the unmodified reference,
    including ties between equidistant keypoints (integer pixels), the threshold on the squared distance and keypoints
    that find nothing (dropped)."""
    from leg_slam_b200 import ingest
    px, has, p3, colors = _keypoint_case(N, W, H, 300 + N, integer_pixels=ints)
    intr = [0.9 * W, 0.91 * W, W / 2 - 0.5, H / 2 - 0.5]
    args = [t.to(dev) for t in (px, has, p3, colors)]
    op, oc = ingest.monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints(*args, maxd, intr, W)
    rp_, rc_ = ref_geo.inactive_geo_densify(*args, maxd, intr, W)
    assert op.shape == rp_.shape and oc.shape == rc_.shape
    assert torch.equal(op, rp_) and torch.equal(oc, rc_)
    if N >= 300 and maxd < 1e8:
        assert 0 < op.shape[0] < N
    if N <= 1000:  # the numpy restatement (Python loop)
        qp, qc = IR.inactive_geo_densify(px.numpy(), has.numpy(), p3.numpy(), colors.numpy(), maxd, intr, W)
        assert qp.shape == tuple(op.shape)
        np.testing.assert_allclose(op.cpu().numpy(), qp, rtol=1e-6, atol=1e-7)
        np.testing.assert_array_equal(oc.cpu().numpy(), qc)


@pytest.mark.gpu
def test_inactive_geo_densify_argument_checks(dev):
    from leg_slam_b200 import ingest
    f = ingest.monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints
    z = torch.zeros
    with pytest.raises(ValueError, match="kps_pixel must have dimensions"):
        f(z(4, 3, device=dev), z(4, dtype=torch.bool, device=dev), z(4, 3, device=dev), z(100, device=dev), 1.0, [1, 1, 0, 0], 8)
    with pytest.raises(ValueError, match="kps_point_local must have dimensions"):
        f(z(4, 2, device=dev), z(4, dtype=torch.bool, device=dev), z(4, 2, device=dev), z(100, device=dev), 1.0, [1, 1, 0, 0], 8)
    p, c = f(z(0, 2, device=dev), z(0, dtype=torch.bool, device=dev), z(0, 3, device=dev), z(100, device=dev), 1.0, [1, 1, 0, 0], 8)
    assert p.numel() == 0 and c.numel() == 0
    # no keypoint has a 3D point: nothing comes back
    p, c = f(torch.rand(10, 2, device=dev), z(10, dtype=torch.bool, device=dev), z(10, 3, device=dev), z(100, device=dev), 50.0,
             [1, 1, 0, 0], 8)
    assert p.shape == (0, 3) and c.shape == (0, 3)
