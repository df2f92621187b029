"""Keyframe-ingest geometry on liblgs.so: the reference's operators that turn a new
RGB-D keyframe into Gaussians, just before the mapping hot path --
    reprojectDepthPinhole   include/stereo_vision.h, src/stereo_vision.cu:135-162
    transformPoints         include/operate_points.h, src/operate_points.cu:59-78
    distCUDA2               third_party/simple-knn/spatial.cu:15-27 (used for the initial scales,
                            src/gaussian_model.cpp:157,242,331)
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
