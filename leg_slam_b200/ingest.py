"""Keyframe-ingest geometry on liblgs.so: the reference's operators that turn a new
RGB-D keyframe into Gaussians, just before the mapping hot path --
    reprojectDepthPinhole   include/stereo_vision.h, src/stereo_vision.cu:135-162
    transformPoints         include/operate_points.h, src/operate_points.cu:59-78
    distCUDA2               third_party/simple-knn/spatial.cu:15-27 (used for the initial scales,
                            src/gaussian_model.cpp:157,242,331)
and two neighbours of the path that the reference also runs as CUDA operators --
    scaleAndTransformThenMarkVisiblePoints   include/operate_points.h, src/operate_points.cu:96-140 (a caller of
                            markVisible: loop-closure correction of the Gaussians a keyframe sees)
    monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints
                            include/stereo_vision.h, src/stereo_vision.cu:164-212 (keypoints without depth)
-- same names, argument meaning and error behaviour.  No CPU path."""
import torch

from . import _lib
from ._lib import check


def _s(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _cuda(t):
    if not t.is_cuda:
        raise _lib.LgsError("leg_slam_b200 has no CPU path: tensors must live on a CUDA device")
    return t


def reprojectDepthPinhole(depth, mask, intr, width):
    """depth [P] float32 (row-major image, P = H*W), mask [P] bool, intr = (fx, fy, cx, cy) -> points [P,3]
    (zeros where the mask is false)."""
    if depth.dim() != 1:
        raise ValueError("points must have dimensions (num_points)")  # AT_ERROR, stereo_vision.cu:140-142
    P = int(depth.shape[0])
    if P == 0:
        return torch.empty(0)  # the reference returns an undefined tensor
    depth = _cuda(depth).contiguous().float()
    mask = mask.contiguous().to(torch.bool)
    points = torch.empty(P, 3, dtype=torch.float32, device=depth.device)
    fx, fy, cx, cy = (float(v) for v in intr[:4])
    with torch.cuda.device(depth.device):
        check(_lib.lib().lgs_reproject_depth_pinhole(P, int(width), fx, fy, cx, cy, depth.data_ptr(), mask.data_ptr(),
                                                     points.data_ptr(), _s(depth)), "lgs_reproject_depth_pinhole")
    return points


def transformPoints(points, transformmatrix):
    """points [P,3], transformmatrix [4,4] stored transposed like the reference's Twc tensors -> transformed copy
    (the reference rebinds its `points` argument to the result)."""
    if points.dim() != 2 or points.shape[1] != 3:
        raise ValueError("points must have dimensions (num_points, 3)")  # AT_ERROR, operate_points.cu:62-64
    P = int(points.shape[0])
    if P == 0:
        return points
    points = _cuda(points).contiguous().float()
    T = transformmatrix.contiguous().float().to(points.device)
    out = torch.empty_like(points)
    with torch.cuda.device(points.device):
        check(_lib.lib().lgs_transform_points(P, points.data_ptr(), T.data_ptr(), out.data_ptr(), _s(points)),
              "lgs_transform_points")
    return out


def distCUDA2(points):
    """points [P,3] -> [P] mean squared distance to the 3 nearest neighbours."""
    P = int(points.shape[0])
    points = _cuda(points).contiguous().float()
    means = torch.zeros(P, dtype=torch.float32, device=points.device)
    if P == 0:
        return means
    L = _lib.lib()
    with torch.cuda.device(points.device):
        scratch = torch.empty(L.lgs_knn_scratch_bytes(P), dtype=torch.uint8, device=points.device)
        check(L.lgs_knn_mean_dist2(P, points.data_ptr(), means.data_ptr(), scratch.data_ptr(), _s(points)), "lgs_knn_mean_dist2")
    return means


def scaleAndTransformThenMarkVisiblePoints(points, rots, point_not_transformed_mask, point_unstable_mask, transformmatrix,
                                           viewmatrix, projmatrix, num_transformed=0, scale=1.0, faithful_rot_store=True):
    """In place, like the reference (src/operate_points.cu:96-140): the rows that are visible from `viewmatrix`
    (markVisible: view-space z > 0.2), still flagged in `point_not_transformed_mask` and flagged in `point_unstable_mask`
    get  points <- T (scale * points),  rots <- quaternion of T[:3,:3] R(rots),  and their not-transformed flag cleared.
    Returns num_transformed + the number of such rows (the reference's `int &num_transformed`; one host read-back, as the
    reference's `.item<int>()`).  One kernel instead of the reference's markVisible + temporaries + boolean-index copies.
    faithful_rot_store=True keeps what the reference leaves in a corrected rotation row, (w, x, z, 0)
    (cuda_rasterizer/operate_points.h:169-178 stores z twice and never the fourth element); False stores (w, x, y, z)."""
    if points.dim() != 2 or points.shape[1] != 3:
        raise ValueError("points must have dimensions (num_points, 3)")  # AT_ERROR, operate_points.cu:106-108
    P = int(points.shape[0])
    if point_not_transformed_mask.shape[0] != P or point_unstable_mask.shape[0] != P:
        raise ValueError("points_mask must have dimensions (num_points)")  # :116-118
    if P == 0:
        return int(num_transformed)
    _cuda(points)
    for t, dt in ((points, torch.float32), (rots, torch.float32), (point_not_transformed_mask, torch.bool)):
        if not t.is_contiguous() or t.dtype != dt or t.device != points.device:
            raise ValueError("points / rots (float32) and point_not_transformed_mask (bool) are updated in place: "
                             "contiguous tensors on one CUDA device")
    if rots.dim() != 2 or rots.shape[0] != P or rots.shape[1] != 4:
        raise ValueError("rots must have dimensions (num_points, 4)")
    dev = points.device
    unstable = point_unstable_mask.to(device=dev, dtype=torch.bool).contiguous()
    T = transformmatrix.to(device=dev, dtype=torch.float32).contiguous()
    V = viewmatrix.to(device=dev, dtype=torch.float32).contiguous()
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().lgs_scale_transform_mark_visible(P, float(scale), points.data_ptr(), rots.data_ptr(),
                                                          point_not_transformed_mask.data_ptr(), unstable.data_ptr(),
                                                          T.data_ptr(), V.data_ptr(), 1 if faithful_rot_store else 0,
                                                          count.data_ptr(), _s(points)), "lgs_scale_transform_mark_visible")
    return int(num_transformed) + int(count.item())


def monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints(kps_pixel, kps_has3D, kps_point_local, colors,
                                                                       max_pixel_dist, intr, width):
    """kps_pixel [N,2], kps_has3D [N] bool, kps_point_local [N,3], colors = the image buffer the reference indexes at
    trunc(v * width + u) + {0,1,2} -> (points [n,3], colours [n,3]) of the keypoints that end with a positive depth, in
    keypoint order (src/stereo_vision.cu:164-212).  max_pixel_dist bounds the SQUARED pixel distance, as shipped."""
    if kps_pixel.dim() != 2 or kps_pixel.shape[1] != 2:
        raise ValueError("kps_pixel must have dimensions (num_points, 2)")       # AT_ERROR, stereo_vision.cu:175-180
    if kps_has3D.dim() != 1:
        raise ValueError("kps_has3D must have dimensions (num_points)")
    if kps_point_local.dim() != 2 or kps_point_local.shape[1] != 3:
        raise ValueError("kps_point_local must have dimensions (num_points, 3)")
    N = int(kps_pixel.shape[0])
    if N == 0:
        return torch.empty(0), torch.empty(0)  # the reference returns two undefined tensors
    px = _cuda(kps_pixel).contiguous().float()
    dev = px.device
    has = kps_has3D.to(device=dev, dtype=torch.bool).contiguous()
    p3 = kps_point_local.to(device=dev, dtype=torch.float32).contiguous()
    col = colors.to(device=dev, dtype=torch.float32).contiguous()
    fx, fy, cx, cy = (float(v) for v in intr[:4])
    out_p = torch.empty(N, 3, dtype=torch.float32, device=dev)
    out_c = torch.empty(N, 3, dtype=torch.float32, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        scratch = torch.empty(L.lgs_inactive_geo_scratch_bytes(N), dtype=torch.uint8, device=dev)
        check(L.lgs_inactive_geo_densify(N, int(width), fx, fy, cx, cy, float(max_pixel_dist), px.data_ptr(), has.data_ptr(),
                                         p3.data_ptr(), col.data_ptr(), int(col.numel()), out_p.data_ptr(), out_c.data_ptr(),
                                         count.data_ptr(), scratch.data_ptr(), _s(px)), "lgs_inactive_geo_densify")
    n = int(count.item())
    return out_p[:n], out_c[:n]
