/*
 * gaussian_renderer.h -- the reference's GaussianRenderer::render (include/gaussian_renderer.h:29-43,
 * src/gaussian_renderer.cpp:24-160) with the same signature, on the C++ GaussianModel / GaussianRasterizer of this repo, and the
 * rasterizer-facing part of GaussianMapper::trainForOneIteration (src/gaussian_mapper.cpp:686-796) as two functions.
 * Implemented in leg_slam_b200/csrc/host/gaussian_renderer.cpp.
 */
#pragma once
#include <torch/torch.h>

#include <memory>
#include <tuple>

#include "gaussian_keyframe.h"
#include "gaussian_model.h"
#include "gaussian_rasterizer.h"

/* reference include/gaussian_parameters.h:43-53 */
class GaussianPipelineParams {
public:
    GaussianPipelineParams(bool convert_SHs = false, bool compute_cov3D = false) : convert_SHs_(convert_SHs), compute_cov3D_(compute_cov3D) {}
    bool convert_SHs_;
    bool compute_cov3D_;
};

class GaussianRenderer {
public:
    /* -> (render, render_lf, render_depth, viewspace_points, visibility_filter, radii) */
    static std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> render(
        std::shared_ptr<GaussianKeyframe> viewpoint_camera, int image_height, int image_width, std::shared_ptr<GaussianModel> gaussians,
        GaussianPipelineParams &pipe, torch::Tensor &bg_color, torch::Tensor &override_color, float scaling_modifier = 1.0f,
        bool has_override_color = false, bool include_language_features = false);
};

/* One view of one mapping iteration up to the optimizer step (src/gaussian_mapper.cpp:686-745): render with language features,
 * the mapper's loss -- (1 - l) L1 + l (1 - SSIM) + mean cosine similarity (ADDED, as shipped) + depth L1 on the masked images,
 * the keyframe's feature map resized nearest -- and its gradient w.r.t. the three images in one fused call (lgs_mapping_loss),
 * backward through the rasterizer into the model's .grad tensors, and, when `update_densification_stats`, max_radii2D_ and
 * addDensificationStats for the visible Gaussians.  mask [3,H,W] or an empty tensor.  Returns the loss (0-dim, on the device). */
torch::Tensor mappingIterationBackward(std::shared_ptr<GaussianModel> gaussians, std::shared_ptr<GaussianKeyframe> viewpoint_cam,
                                       GaussianPipelineParams &pipe, torch::Tensor &background, torch::Tensor &gt_image,
                                       torch::Tensor &gt_depth, torch::Tensor &mask, float lambda_dssim,
                                       bool update_densification_stats);

/* optimizer_->step(); optimizer_->zero_grad(true)  (src/gaussian_mapper.cpp:793-796) */
void mappingIterationStep(std::shared_ptr<GaussianModel> gaussians);
