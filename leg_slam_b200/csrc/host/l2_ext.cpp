// l2_ext.cpp -- pybind module `_L2`: the C++ L2 layer (include/gaussian_rasterizer.h) made callable from Python so that the
// test-suite can hold it to the Python twin (leg_slam_b200/rasterizer.py).  Not part of the reference's Python package.
#include <torch/extension.h>

#include "gaussian_rasterizer.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    pybind11::class_<GaussianRasterizationSettings>(m, "GaussianRasterizationSettings")
        .def(pybind11::init<int, int, float, float, torch::Tensor&, float, torch::Tensor&, torch::Tensor&, int, torch::Tensor&,
                            bool, bool>());
    pybind11::class_<GaussianRasterizer, std::shared_ptr<GaussianRasterizer>>(m, "GaussianRasterizer")
        .def(pybind11::init<GaussianRasterizationSettings&>())
        .def("forward", &GaussianRasterizer::forward)
        .def("markVisibleGaussians", &GaussianRasterizer::markVisibleGaussians);
}
