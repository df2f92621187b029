"""TEST INFRASTRUCTURE -- numpy restatement of the reference's keyframe-ingest geometry.

reproject_depth_pinhole: /root/reference/cuda_rasterizer/stereo_vision.h:41-55 + src/stereo_vision.cu:40-61;
transform_points: cuda_rasterizer/auxiliary.h:58-66 (transformPoint4x3) via src/operate_points.cu:39-50;
knn_mean_dist2: third_party/simple-knn/simple_knn.cu:133-183 -- the mean of the three smallest squared distances to
the OTHER points, here by brute force (the reference's Morton/box search only prunes candidates).  Small inputs only.
Pinned against the compiled reference simple-knn (oracle/_ref/ref_simple_knn.so, oracle/build_ref.py) on the GPU
(tests/test_ingest.py); the two closed-form functions have nothing to pin beyond their formula."""
import numpy as np

FLT_MAX = np.float32(3.4028234663852886e38)


def reproject_depth_pinhole(depth, mask, intr, width):
    depth = np.asarray(depth, np.float32)
    P = depth.shape[0]
    fx, fy, cx, cy = (np.float32(v) for v in intr[:4])
    idx = np.arange(P)
    v = (idx // width).astype(np.float32)
    u = (idx - (idx // width) * width).astype(np.float32)
    pts = np.zeros((P, 3), np.float32)
    m = np.asarray(mask, bool)
    pts[m, 0] = ((u[m] - cx) * depth[m] / fx).astype(np.float32)
    pts[m, 1] = ((v[m] - cy) * depth[m] / fy).astype(np.float32)
    pts[m, 2] = depth[m]
    return pts


def transform_points(points, T):
    p = np.asarray(points, np.float64)
    m = np.asarray(T, np.float64).reshape(16)
    out = np.empty_like(p)
    for r in range(3):
        out[:, r] = m[r] * p[:, 0] + m[4 + r] * p[:, 1] + m[8 + r] * p[:, 2] + m[12 + r]
    return out.astype(np.float32)


def knn_mean_dist2(points):
    p = np.asarray(points, np.float32)
    P = p.shape[0]
    out = np.empty(P, np.float32)
    for i in range(P):
        d = p - p[i]
        # same float sequence as the kernels: fma(dz, dz, fma(dy, dy, dx*dx)) evaluated in double then rounded is within
        # half an ulp of it; the comparison in the tests allows 1 ulp
        d2 = (d[:, 0].astype(np.float64) ** 2 + d[:, 1].astype(np.float64) ** 2 + d[:, 2].astype(np.float64) ** 2).astype(np.float32)
        d2[i] = FLT_MAX
        best = np.sort(d2)[:3] if P >= 3 else np.concatenate([np.sort(d2), np.full(3 - P, FLT_MAX, np.float32)])
        with np.errstate(over="ignore"):
            out[i] = (np.float32(best[0]) + np.float32(best[1]) + np.float32(best[2])) / np.float32(3.0)
    return out
