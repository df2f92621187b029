/*
 * gaussian_rasterizer.h -- the reference's L2 interface (autograd wrapper + module), re-declared with the same type and
 * member names so that gaussian_renderer.cpp / gaussian_trainer.cpp compile against it unchanged:
 *     GaussianRasterizationSettings      reference include/gaussian_rasterizer.h:25-61
 *     GaussianRasterizerFunction         :63-84,  src/gaussian_rasterizer.cpp:27-176
 *     rasterizeGaussians                 :86-110
 *     GaussianRasterizer                 :112-141, src/gaussian_rasterizer.cpp:18-25,178-236
 * Unlike the reference's header it does not pull gaussian_model.h (Eigen / OpenCV / Sophus): the rasterizer needs none of
 * it.  Implemented in leg_slam_b200/csrc/host/gaussian_rasterizer.cpp on rasterize_points.h (L1) -> lgs.h (C ABI).
 */
#pragma once
#include <torch/torch.h>

#include <tuple>

#include "rasterize_points.h"

struct GaussianRasterizationSettings {
    GaussianRasterizationSettings(int image_height, int image_width, float tanfovx, float tanfovy, torch::Tensor& bg,
                                  float scale_modifier, torch::Tensor& viewmatrix, torch::Tensor& projmatrix, int sh_degree,
                                  torch::Tensor& campos, bool prefiltered, bool include_language_features)
        : image_height_(image_height), image_width_(image_width), tanfovx_(tanfovx), tanfovy_(tanfovy), bg_(bg),
          scale_modifier_(scale_modifier), viewmatrix_(viewmatrix), projmatrix_(projmatrix), sh_degree_(sh_degree),
          campos_(campos), prefiltered_(prefiltered), include_language_features_(include_language_features) {}

    int image_height_, image_width_;
    float tanfovx_, tanfovy_;
    torch::Tensor bg_;
    float scale_modifier_;
    torch::Tensor viewmatrix_, projmatrix_;
    int sh_degree_;
    torch::Tensor campos_;
    bool prefiltered_, include_language_features_;
};

class GaussianRasterizerFunction : public torch::autograd::Function<GaussianRasterizerFunction> {
public:
    // returns {color [3,H,W], language feature [64,H,W], depth [1,H,W], radii [P] (int32, not differentiable)}
    static torch::autograd::tensor_list forward(torch::autograd::AutogradContext* ctx, torch::Tensor means3D,
                                                torch::Tensor means2D, torch::Tensor sh, torch::Tensor colors_precomp,
                                                torch::Tensor lang_feats, torch::Tensor opacities, torch::Tensor scales,
                                                torch::Tensor rotations, torch::Tensor cov3Ds_precomp,
                                                GaussianRasterizationSettings raster_settings);
    // gradients in input order: means3D, means2D, sh, colors_precomp, lang_feats, opacities, scales, rotations,
    // cov3Ds_precomp, (settings)
    static torch::autograd::tensor_list backward(torch::autograd::AutogradContext* ctx,
                                                 torch::autograd::tensor_list grad_outputs);
};

inline torch::autograd::tensor_list rasterizeGaussians(torch::Tensor& means3D, torch::Tensor& means2D, torch::Tensor& sh,
                                                       torch::Tensor& colors_precomp, torch::Tensor& lang_feat,
                                                       torch::Tensor& opacities, torch::Tensor& scales,
                                                       torch::Tensor& rotations, torch::Tensor& cov3Ds_precomp,
                                                       GaussianRasterizationSettings& raster_settings) {
    return GaussianRasterizerFunction::apply(means3D, means2D, sh, colors_precomp, lang_feat, opacities, scales, rotations,
                                             cov3Ds_precomp, raster_settings);
}

class GaussianRasterizer : public torch::nn::Module {
public:
    explicit GaussianRasterizer(GaussianRasterizationSettings& raster_settings) : raster_settings_(raster_settings) {}

    torch::Tensor markVisibleGaussians(torch::Tensor& positions);

    // the has_* flags stand for the Python wrapper's `is not None` tests (the reference passes both)
    std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> forward(
        torch::Tensor means3D, torch::Tensor means2D, torch::Tensor opacities, bool has_shs, bool has_colors_precomp,
        bool has_lang_feat, bool has_scales, bool has_rotations, bool has_cov3D_precomp, torch::Tensor shs,
        torch::Tensor colors_precomp, torch::Tensor lang_feat, torch::Tensor scales, torch::Tensor rotations,
        torch::Tensor cov3D_precomp);

    GaussianRasterizationSettings raster_settings_;
};
