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
