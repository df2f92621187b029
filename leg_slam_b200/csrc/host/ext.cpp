// ext.cpp -- pybind module `_C` with the three names the reference's Python package binds
// (eval/submodules/diff-gaussian-rasterization-legs-slam/ext.cpp:14-18), so
// diff_gaussian_rasterization_legs_slam/__init__.py and eval/render.py work on it unchanged -- plus, for the tests, the
// libtorch geometry operators of include/operate_points.h / stereo_vision.h / spatial.h (the reference has no Python binding
// for those; reference parameters come back as return values here).
#include <torch/extension.h>

#include "operate_points.h"
#include "rasterize_points.h"
#include "spatial.h"
#include "stereo_vision.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("rasterize_gaussians", &RasterizeGaussiansCUDA);
    m.def("rasterize_gaussians_backward", &RasterizeGaussiansBackwardCUDA);
    m.def("mark_visible", &markVisible);
    m.def("transform_points", [](torch::Tensor points, torch::Tensor transformmatrix) {
        transformPoints(points, transformmatrix);
        return points;  // the rebound tensor
    });
    m.def("scale_and_transform_then_mark_visible",
          [](torch::Tensor points, torch::Tensor rots, torch::Tensor point_not_transformed_mask, torch::Tensor point_unstable_mask,
             torch::Tensor transformmatrix, torch::Tensor viewmatrix, torch::Tensor projmatrix, int num_transformed, float scale) {
              scaleAndTransformThenMarkVisiblePoints(points, rots, point_not_transformed_mask, point_unstable_mask, transformmatrix,
                                                     viewmatrix, projmatrix, num_transformed, scale);
              return num_transformed;  // tensors are updated in place
          });
    m.def("reproject_depth_pinhole", [](torch::Tensor depth, torch::Tensor mask, std::vector<float> intr, int width) {
        return reprojectDepthPinhole(depth, mask, intr, width);
    });
    m.def("inactive_geo_densify", [](torch::Tensor kps_pixel, torch::Tensor kps_has3D, torch::Tensor kps_point_local,
                                     torch::Tensor colors, float max_pixel_dist, std::vector<float> intr, int width) {
        return monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints(kps_pixel, kps_has3D, kps_point_local, colors,
                                                                                  max_pixel_dist, intr, width);
    });
    m.def("dist_cuda2", &distCUDA2);
}
