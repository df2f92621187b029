// geometry_ops.cpp -- the reference's libtorch geometry operators (include/operate_points.h, include/stereo_vision.h,
// include/spatial.h: same signatures, same error strings, same in-place / rebinding behaviour) on the C ABI of liblgs
// (include/lgs.h).  Reference implementations: src/operate_points.cu:72-140, src/stereo_vision.cu:135-212,
// third_party/simple-knn/spatial.cu:15-27.
//
// Differences callers cannot observe: work goes to torch's current CUDA stream (the reference launches on the legacy
// default stream); scaleAndTransformThenMarkVisiblePoints is one kernel updating the caller's tensors in place instead
// of markVisible + temporaries + boolean-index copies; the inactive-geometry densification compacts on the device.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "lgs.h"
#include "operate_points.h"
#include "spatial.h"
#include "stereo_vision.h"

namespace {
void check(int status, const char* what) {
    TORCH_CHECK(status == LGS_OK, what, ": ", lgs_status_string(status), " (cudaError ", lgs_last_cuda_error(), ")");
}
void need_cuda(const torch::Tensor& t) {
    TORCH_CHECK(t.is_cuda(), "leg_slam_b200 has no CPU path: tensors must live on a CUDA device");
}
void* stream() { return (void*)at::cuda::getCurrentCUDAStream().stream(); }
torch::Tensor f32(const torch::Tensor& t, const torch::Device& dev) { return t.to(dev, torch::kFloat32).contiguous(); }
}  // namespace

void transformPoints(torch::Tensor& points, torch::Tensor& transformmatrix) {
    if (points.ndimension() != 2 || points.size(1) != 3) {
        AT_ERROR("points must have dimensions (num_points, 3)");
    }
    const int P = points.size(0);
    if (P == 0) return;
    need_cuda(points);
    const c10::cuda::CUDAGuard guard(points.device());
    torch::Tensor in = f32(points, points.device()), T = f32(transformmatrix, points.device());
    torch::Tensor out = torch::empty_like(in);
    check(lgs_transform_points(P, in.data_ptr<float>(), T.data_ptr<float>(), out.data_ptr<float>(), stream()), "lgs_transform_points");
    points = out;
}

void scaleAndTransformThenMarkVisiblePoints(torch::Tensor& points, torch::Tensor& rots, torch::Tensor& point_not_transformed_mask,
                                            torch::Tensor& point_unstable_mask, torch::Tensor& transformmatrix,
                                            torch::Tensor& viewmatrix, torch::Tensor& projmatrix, int& num_transformed,
                                            const float scale) {
    (void)projmatrix;  // markVisible never uses it (auxiliary.h:150-154)
    if (points.ndimension() != 2 || points.size(1) != 3) {
        AT_ERROR("points must have dimensions (num_points, 3)");
    }
    const int64_t P = points.size(0);
    if (point_not_transformed_mask.size(0) != P || point_unstable_mask.size(0) != P) {
        AT_ERROR("points_mask must have dimensions (num_points)");
    }
    if (P == 0) return;
    need_cuda(points);
    const c10::cuda::CUDAGuard guard(points.device());
    const auto dev = points.device();
    // The kernel works in place.  A tensor that is not already a contiguous float32 / bool CUDA tensor is staged and
    // copied back, so that the caller's tensor object is updated either way (the reference uses index_put_).
    torch::Tensor p = f32(points, dev), r = f32(rots, dev);
    torch::Tensor nt = point_not_transformed_mask.to(dev, torch::kBool).contiguous();
    torch::Tensor un = point_unstable_mask.to(dev, torch::kBool).contiguous();
    torch::Tensor T = f32(transformmatrix, dev), V = f32(viewmatrix, dev);
    TORCH_CHECK(r.ndimension() == 2 && r.size(0) == P && r.size(1) == 4, "rots must have dimensions (num_points, 4)");
    torch::Tensor count = torch::zeros({1}, points.options().dtype(torch::kInt32));
    check(lgs_scale_transform_mark_visible((int)P, scale, p.data_ptr<float>(), r.data_ptr<float>(), (unsigned char*)nt.data_ptr<bool>(),
                                           (const unsigned char*)un.data_ptr<bool>(), T.data_ptr<float>(), V.data_ptr<float>(),
                                           /*faithful_rot_store=*/1, count.data_ptr<int>(), stream()),
          "lgs_scale_transform_mark_visible");
    if (!p.is_same(points)) points.copy_(p);
    if (!r.is_same(rots)) rots.copy_(r);
    if (!nt.is_same(point_not_transformed_mask)) point_not_transformed_mask.copy_(nt);
    num_transformed += count.item<int>();  // the reference's final_mask.sum().item<int>()
}

torch::Tensor reprojectDepthPinhole(torch::Tensor& depth, torch::Tensor& mask, std::vector<float>& intr, int width) {
    if (depth.ndimension() != 1) {
        AT_ERROR("points must have dimensions (num_points)");
    }
    const int P = depth.size(0);
    torch::Tensor points;  // undefined for an empty image, like the reference
    if (P == 0) return points;
    need_cuda(depth);
    TORCH_CHECK(intr.size() >= 4, "intr must hold fx, fy, cx, cy");
    const c10::cuda::CUDAGuard guard(depth.device());
    torch::Tensor d = f32(depth, depth.device());
    torch::Tensor m = mask.to(depth.device(), torch::kBool).contiguous();
    points = torch::empty({P, 3}, d.options());
    check(lgs_reproject_depth_pinhole(P, width, intr[0], intr[1], intr[2], intr[3], d.data_ptr<float>(),
                                      (const unsigned char*)m.data_ptr<bool>(), points.data_ptr<float>(), stream()),
          "lgs_reproject_depth_pinhole");
    return points;
}

std::tuple<torch::Tensor, torch::Tensor>
monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints(torch::Tensor& kps_pixel, torch::Tensor& kps_has3D,
                                                                   torch::Tensor& kps_point_local, torch::Tensor& colors,
                                                                   float max_pixel_dist, std::vector<float>& intr, int width) {
    if (kps_pixel.ndimension() != 2 || kps_pixel.size(1) != 2) AT_ERROR("kps_pixel must have dimensions (num_points, 2)");
    if (kps_has3D.ndimension() != 1) AT_ERROR("kps_has3D must have dimensions (num_points)");
    if (kps_point_local.ndimension() != 2 || kps_point_local.size(1) != 3)
        AT_ERROR("kps_point_local must have dimensions (num_points, 3)");
    const int N = kps_pixel.size(0);
    torch::Tensor result_pt, result_color;  // undefined without keypoints, like the reference
    if (N == 0) return std::make_tuple(result_pt, result_color);
    need_cuda(kps_pixel);
    TORCH_CHECK(intr.size() >= 4, "intr must hold fx, fy, cx, cy");
    const c10::cuda::CUDAGuard guard(kps_pixel.device());
    const auto dev = kps_pixel.device();
    torch::Tensor px = f32(kps_pixel, dev), p3 = f32(kps_point_local, dev), col = f32(colors, dev);
    torch::Tensor has = kps_has3D.to(dev, torch::kBool).contiguous();
    torch::Tensor out_p = torch::empty({N, 3}, px.options()), out_c = torch::empty({N, 3}, px.options());
    torch::Tensor count = torch::zeros({1}, px.options().dtype(torch::kInt32));
    torch::Tensor scratch = torch::empty({(long long)lgs_inactive_geo_scratch_bytes(N)}, px.options().dtype(torch::kByte));
    check(lgs_inactive_geo_densify(N, width, intr[0], intr[1], intr[2], intr[3], max_pixel_dist, px.data_ptr<float>(),
                                   (const unsigned char*)has.data_ptr<bool>(), p3.data_ptr<float>(), col.data_ptr<float>(),
                                   (long long)col.numel(), out_p.data_ptr<float>(), out_c.data_ptr<float>(), count.data_ptr<int>(),
                                   (char*)scratch.data_ptr<uint8_t>(), stream()),
          "lgs_inactive_geo_densify");
    const int n = count.item<int>();  // the reference's boolean index synchronises as well
    return std::make_tuple(out_p.narrow(0, 0, n), out_c.narrow(0, 0, n));
}

torch::Tensor distCUDA2(const torch::Tensor& points) {
    const int P = points.size(0);
    auto float_opts = points.options().dtype(torch::kFloat32);
    torch::Tensor means = torch::zeros({P}, float_opts);
    if (P == 0) return means;
    need_cuda(points);
    const c10::cuda::CUDAGuard guard(points.device());
    torch::Tensor p = f32(points, points.device());
    torch::Tensor scratch = torch::empty({(long long)lgs_knn_scratch_bytes(P)}, points.options().dtype(torch::kByte));
    check(lgs_knn_mean_dist2(P, p.data_ptr<float>(), means.data_ptr<float>(), (char*)scratch.data_ptr<uint8_t>(), stream()),
          "lgs_knn_mean_dist2");
    return means;
}
