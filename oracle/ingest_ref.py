"""TEST INFRASTRUCTURE -- numpy restatement of the reference's keyframe-ingest geometry.

reproject_depth_pinhole: /root/reference/cuda_rasterizer/stereo_vision.h:41-55 + src/stereo_vision.cu:40-61;
transform_points: cuda_rasterizer/auxiliary.h:58-66 (transformPoint4x3) via src/operate_points.cu:39-50;
knn_mean_dist2: third_party/simple-knn/simple_knn.cu:133-183 -- the mean of the three smallest squared distances to
the OTHER points, here by brute force (the reference's Morton/box search only prunes candidates).  Small inputs only.
scale_and_transform_then_mark_visible: src/operate_points.cu:52-70,96-140 + cuda_rasterizer/operate_points.h:54-179
(loop-closure correction of the Gaussians a keyframe sees: a caller of markVisible);
inactive_geo_densify: src/stereo_vision.cu:63-133,164-212 (monocular keypoints without depth borrow the depth of the
nearest keypoint that has a 3D point).
Pinned on the GPU (tests/test_ingest.py) against the compiled, unmodified reference: simple-knn
(oracle/_ref/ref_simple_knn.so) and src/stereo_vision.cu + src/operate_points.cu (oracle/_ref/ref_geometry.so), both built
by oracle/build_ref.py."""
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


def _f32(x):
    return np.asarray(x, np.float32)


def mark_visible(points, viewmatrix):
    """in_frustum as the reference ships it (auxiliary.h:139-164): only the view-space z <= 0.2 test is live."""
    p = _f32(points).astype(np.float64)
    m = _f32(viewmatrix).reshape(16).astype(np.float64)
    z = m[2] * p[:, 0] + m[6] * p[:, 1] + m[10] * p[:, 2] + m[14]
    return z.astype(np.float32) > np.float32(0.2)


def quaternion_through_matrix(q, T):
    """(w, x, y, z) rows -> quaternion of T[:3,:3] * R(q) by Shoemake's branches (operate_points.h:54-150), evaluated in
    float64 from float32 inputs (the kernels run float32 with fused multiply-adds: tests compare at 1e-5)."""
    q = _f32(q).astype(np.float64)
    m = _f32(T).reshape(16).astype(np.float64)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    R0 = np.stack([np.stack([1 - (tyy + tzz), txy - twz, txz + twy], -1),
                   np.stack([txy + twz, 1 - (txx + tzz), tyz - twx], -1),
                   np.stack([txz - twy, tyz + twx, 1 - (txx + tyy)], -1)], 1)  # [n, 3, 3]
    A = np.array([[m[0], m[4], m[8]], [m[1], m[5], m[9]], [m[2], m[6], m[10]]])  # stored transposed
    R = np.einsum("ij,njk->nik", A, R0)
    out = np.zeros((q.shape[0], 4))  # w, x, y, z
    for n in range(q.shape[0]):
        r = R[n]
        t = r[0, 0] + r[1, 1] + r[2, 2]
        if t > 0:
            t = np.sqrt(t + 1.0)
            out[n, 0] = 0.5 * t
            t = 0.5 / t
            out[n, 1:] = [(r[2, 1] - r[1, 2]) * t, (r[0, 2] - r[2, 0]) * t, (r[1, 0] - r[0, 1]) * t]
        else:
            i = 0
            if r[1, 1] > r[0, 0]:
                i = 1
            if r[2, 2] > r[i, i]:
                i = 2
            j = (i + 1) % 3
            k = (j + 1) % 3
            t = np.sqrt(r[i, i] - r[j, j] - r[k, k] + 1.0)
            xyz = [0.0, 0.0, 0.0]
            xyz[i] = 0.5 * t
            t = 0.5 / t
            out[n, 0] = (r[k, j] - r[j, k]) * t
            xyz[j] = (r[j, i] + r[i, j]) * t
            xyz[k] = (r[k, i] + r[i, k]) * t
            out[n, 1:] = xyz
    return out.astype(np.float32)


def scale_and_transform_then_mark_visible(points, rots, not_transformed, unstable, T, viewmatrix, scale=1.0):
    """Returns (points, rots, not_transformed, num_transformed) after the reference's in-place update.  The rotation rows
    come back as (w, x, z, 0): insert_rot_to_rots stores z at offset 2 twice and never writes offset 3
    (operate_points.h:169-178) into a zero-filled tensor, and the caller copies whole rows (operate_points.cu:134-135)."""
    pts = _f32(points).copy()
    rt = _f32(rots).copy()
    nt = np.asarray(not_transformed, bool).copy()
    final = nt & np.asarray(unstable, bool) & mark_visible(pts, viewmatrix)
    if final.any():
        sc = (pts[final] * np.float32(scale)).astype(np.float32)
        pts[final] = transform_points(sc, T)
        qn = quaternion_through_matrix(rt[final], T)
        rt[final] = np.stack([qn[:, 0], qn[:, 1], qn[:, 3], np.zeros_like(qn[:, 0])], -1)
        nt[final] = False
    return pts, rt, nt, int(final.sum())


def inactive_geo_densify(kps_pixel, kps_has3D, kps_point_local, colors, max_pixel_dist, intr, width):
    """(points, colours) of the keypoints that end with a positive depth (stereo_vision.cu:63-133,198-209).  As shipped:
    `max_pixel_dist` bounds the SQUARED pixel distance (:105-108), ties keep the lowest index (:108), the pixel offset
    v * width + u is evaluated in float and truncated (:81), the colour triple is read at that offset WITHOUT a factor of
    three (:88-90,125-127), and u, v are truncated to integers before the reprojection (stereo_vision.h:41-55)."""
    px = _f32(kps_pixel)
    has = np.asarray(kps_has3D, bool)
    p3 = _f32(kps_point_local)
    col = _f32(colors).reshape(-1)
    fx, fy, cx, cy = (np.float32(v) for v in intr[:4])
    N = px.shape[0]
    res_p = np.zeros((N, 3), np.float32)
    res_c = np.zeros((N, 3), np.float32)
    idx3d = np.nonzero(has)[0]
    for n in range(N):
        u, v = px[n]
        off = int(np.float32(np.float64(v) * width + np.float64(u)))  # one rounding: the reference compiles to fma(v, width, u)
        if has[n]:
            res_p[n] = p3[n]
            res_c[n] = col[off:off + 3]
            continue
        depth = np.float32(-1.0)
        cand = idx3d[idx3d != n]
        if cand.size:
            du = (u - px[cand, 0]).astype(np.float32)
            dv = (v - px[cand, 1]).astype(np.float32)
            # the kernel's dist is fma(du, du, dv * dv) or its mirror; in double then rounded it differs from either by at
            # most half an ulp, which only matters at an exact tie or exactly at the threshold (the tests avoid neither:
            # integer-valued pixel coordinates make every term exact)
            d = (du.astype(np.float64) ** 2 + dv.astype(np.float64) ** 2).astype(np.float32)
            ok = ~(d > np.float32(max_pixel_dist))
            if ok.any():
                dm = np.where(ok, d, np.float32(np.inf))
                depth = p3[cand[int(np.argmin(dm))], 2]  # argmin returns the first minimum = lowest index
        if depth > 0:
            ui, vi = int(u), int(v)
            res_p[n] = [np.float32(np.float32(np.float32(ui) - cx) * depth) / fx,
                        np.float32(np.float32(np.float32(vi) - cy) * depth) / fy, depth]
            res_c[n] = col[off:off + 3]
        else:
            res_p[n, 2] = -1.0
    keep = res_p[:, 2] > 0
    return res_p[keep], res_c[keep]
