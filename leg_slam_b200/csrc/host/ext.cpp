// ext.cpp -- pybind module `_C` with the three names the reference's Python package binds
// (eval/submodules/diff-gaussian-rasterization-legs-slam/ext.cpp:14-18), so
// diff_gaussian_rasterization_legs_slam/__init__.py and eval/render.py work on it unchanged.
#include <torch/extension.h>

#include "rasterize_points.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("rasterize_gaussians", &RasterizeGaussiansCUDA);
    m.def("rasterize_gaussians_backward", &RasterizeGaussiansBackwardCUDA);
    m.def("mark_visible", &markVisible);
}
