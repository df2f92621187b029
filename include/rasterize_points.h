/*
 * rasterize_points.h -- the reference's L1 libtorch interface, re-declared with identical
 * signatures (reference include/rasterize_points.h:19-72; implementation
 * src/rasterize_points.cu:37-228) and implemented on liblgs (include/lgs.h) in
 * leg_slam_b200/csrc/host/rasterize_points.cpp.  Callers -- GaussianRasterizerFunction
 * (src/gaussian_rasterizer.cpp:53,148), markVisibleGaussians (:18-25), the pybind module `_C`
 * (eval/submodules/diff-gaussian-rasterization-legs-slam/ext.cpp:14-18) -- compile unchanged.
 */
#pragma once
#include <torch/extension.h>

#include <tuple>

std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors,
                       const torch::Tensor& lang_feat, const torch::Tensor& opacity, const torch::Tensor& scales,
                       const torch::Tensor& rotations, const float scale_modifier, const torch::Tensor& cov3D_precomp,
                       const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, const float tan_fovx,
                       const float tan_fovy, const int image_height, const int image_width, const torch::Tensor& sh,
                       const int degree, const torch::Tensor& campos, const bool prefiltered,
                       const bool include_lang_feat);

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor, torch::Tensor>
RasterizeGaussiansBackwardCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
                               const torch::Tensor& colors, const torch::Tensor& lang_feat, const torch::Tensor& scales,
                               const torch::Tensor& rotations, const float scale_modifier,
                               const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                               const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                               const torch::Tensor& dL_dout_color, const torch::Tensor& dL_dout_lang_feat,
                               const torch::Tensor& dL_dout_depth, const torch::Tensor& sh, const int degree,
                               const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
                               const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer,
                               const bool include_lang_feat);

torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix);
