// ref_geometry_wrap.cpp -- pybind entry points around the UNMODIFIED reference geometry operators (TEST INFRASTRUCTURE,
// oracle/_ref/ref_geometry.so).  The sources are compiled from where they lie (/root/reference/src/stereo_vision.cu,
// /root/reference/src/operate_points.cu; headers via -I/root/reference on the command line of oracle/build_ref.py) and
// linked with the reference rasterizer objects that build_ref.build() already made (markVisible lives in
// rasterize_points.cu); nothing of them is copied.  The wrappers only adapt reference-parameter signatures to Python.
#include <torch/extension.h>

#include <tuple>
#include <vector>

#include "include/operate_points.h"
#include "include/stereo_vision.h"

static torch::Tensor reproject_depth_pinhole(torch::Tensor depth, torch::Tensor mask, std::vector<float> intr, int width) {
    return reprojectDepthPinhole(depth, mask, intr, width);
}

// the reference rebinds its `points` argument to the transformed copy
static torch::Tensor transform_points(torch::Tensor points, torch::Tensor transformmatrix) {
    transformPoints(points, transformmatrix);
    return points;
}

// points / rots / point_not_transformed_mask are updated in place (index_put_); returns the updated counter
static int scale_and_transform_then_mark_visible(torch::Tensor points, torch::Tensor rots, torch::Tensor point_not_transformed_mask,
                                                 torch::Tensor point_unstable_mask, torch::Tensor transformmatrix,
                                                 torch::Tensor viewmatrix, torch::Tensor projmatrix, int num_transformed, float scale) {
    scaleAndTransformThenMarkVisiblePoints(points, rots, point_not_transformed_mask, point_unstable_mask, transformmatrix,
                                           viewmatrix, projmatrix, num_transformed, scale);
    return num_transformed;
}

static std::tuple<torch::Tensor, torch::Tensor> inactive_geo_densify(torch::Tensor kps_pixel, torch::Tensor kps_has3D,
                                                                     torch::Tensor kps_point_local, torch::Tensor colors,
                                                                     float max_pixel_dist, std::vector<float> intr, int width) {
    return monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints(kps_pixel, kps_has3D, kps_point_local, colors,
                                                                              max_pixel_dist, intr, width);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("reproject_depth_pinhole", &reproject_depth_pinhole);
    m.def("transform_points", &transform_points);
    m.def("scale_and_transform_then_mark_visible", &scale_and_transform_then_mark_visible);
    m.def("inactive_geo_densify", &inactive_geo_densify);
}
