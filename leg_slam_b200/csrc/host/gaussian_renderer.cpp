// gaussian_renderer.cpp -- GaussianRenderer::render and the rasterizer-facing part of one mapping iteration
// (include/gaussian_renderer.h; reference src/gaussian_renderer.cpp:24-160, src/gaussian_mapper.cpp:686-796) in C++ on this
// repo's GaussianModel, GaussianRasterizer (autograd) and the fused loss of liblgs.
#include "gaussian_renderer.h"

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <cmath>
#include <vector>

#include "lgs.h"

namespace {
// real spherical harmonics up to degree 3 (include/sh_utils.h:32-130): sh [P,3,K], dirs [P,3] unit -> [P,3]
const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
const float C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f, 0.5462742152960396f};
const float C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                     -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

torch::Tensor eval_sh(int deg, const torch::Tensor& sh, const torch::Tensor& dirs) {
    auto c = [&](int k) { return sh.select(-1, k); };
    torch::Tensor result = C0 * c(0);
    if (deg > 0) {
        auto x = dirs.slice(-1, 0, 1), y = dirs.slice(-1, 1, 2), z = dirs.slice(-1, 2, 3);
        result = result - C1 * y * c(1) + C1 * z * c(2) - C1 * x * c(3);
        if (deg > 1) {
            auto xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            result = result + C2[0] * xy * c(4) + C2[1] * yz * c(5) + C2[2] * (2.0f * zz - xx - yy) * c(6) + C2[3] * xz * c(7) +
                     C2[4] * (xx - yy) * c(8);
            if (deg > 2) {
                result = result + C3[0] * y * (3 * xx - yy) * c(9) + C3[1] * xy * z * c(10) + C3[2] * y * (4 * zz - xx - yy) * c(11) +
                         C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * c(12) + C3[4] * x * (4 * zz - xx - yy) * c(13) +
                         C3[5] * z * (xx - yy) * c(14) + C3[6] * x * (xx - 3 * yy) * c(15);
            }
        }
    }
    return result;
}
}  // namespace

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> GaussianRenderer::render(
    std::shared_ptr<GaussianKeyframe> viewpoint_camera, int image_height, int image_width, std::shared_ptr<GaussianModel> pc,
    GaussianPipelineParams& pipe, torch::Tensor& bg_color, torch::Tensor& override_color, float scaling_modifier,
    bool use_override_color, bool include_language_features) {
    // a zero leaf whose gradient autograd fills with the gradient of the 2D (screen-space) means (:41-48)
    torch::Tensor screenspace_points = torch::zeros_like(pc->getXYZ()).requires_grad_(true);
    const float tanfovx = std::tan(viewpoint_camera->FoVx_ * 0.5f);
    const float tanfovy = std::tan(viewpoint_camera->FoVy_ * 0.5f);
    GaussianRasterizationSettings raster_settings(image_height, image_width, tanfovx, tanfovy, bg_color, scaling_modifier,
                                                  viewpoint_camera->world_view_transform_, viewpoint_camera->full_proj_transform_,
                                                  pc->active_sh_degree_, viewpoint_camera->camera_center_, false,
                                                  include_language_features);
    GaussianRasterizer rasterizer(raster_settings);
    torch::Tensor means3D = pc->getXYZ(), means2D = screenspace_points, opacity = pc->getOpacityActivation();
    // a precomputed 3D covariance, or scaling / rotation for the rasterizer to build it from (:76-90)
    bool has_scales = false, has_rotations = false, has_cov3D_precomp = false;
    torch::Tensor scales, rotations, cov3D_precomp;
    if (pipe.compute_cov3D_) {
        cov3D_precomp = pc->getCovarianceActivation();
        has_cov3D_precomp = true;
    } else {
        scales = pc->getScalingActivation();
        rotations = pc->getRotationActivation();
        has_scales = has_rotations = true;
    }
    // override colours, SH -> RGB here, or SHs for the rasterizer to convert (:95-116)
    bool has_shs = false, has_color_precomp = false;
    torch::Tensor shs, colors_precomp;
    if (use_override_color) {
        colors_precomp = override_color;
        has_color_precomp = true;
    } else if (pipe.convert_SHs_) {
        const int64_t n = (int64_t)(pc->max_sh_degree_ + 1) * (pc->max_sh_degree_ + 1);
        torch::Tensor shs_view = pc->getFeatures().transpose(1, 2).reshape({-1, 3, n});
        torch::Tensor dir_pp = pc->getXYZ() - viewpoint_camera->camera_center_.unsqueeze(0);
        dir_pp = dir_pp / dir_pp.norm(2, std::vector<int64_t>{1}, /*keepdim=*/true);
        colors_precomp = torch::clamp_min(eval_sh(pc->active_sh_degree_, shs_view, dir_pp) + 0.5, 0.0);
        has_color_precomp = true;
    } else {
        shs = pc->getFeatures();
        has_shs = true;
    }
    bool has_lang_feat = false;
    torch::Tensor lang_feat;
    if (include_language_features) {
        has_lang_feat = true;
        lang_feat = pc->getLanguageFeatures();
    }
    torch::Tensor none = torch::empty({0}, pc->getXYZ().options());  // the reference's default-constructed tensors
    auto pick = [&](const torch::Tensor& t) { return t.defined() ? t : none; };
    auto out = rasterizer.forward(means3D, means2D, opacity, has_shs, has_color_precomp, has_lang_feat, has_scales, has_rotations,
                                  has_cov3D_precomp, pick(shs), pick(colors_precomp), pick(lang_feat), pick(scales), pick(rotations),
                                  pick(cov3D_precomp));
    torch::Tensor radii = std::get<3>(out);
    // Gaussians that were frustum culled or had a radius of 0 were not visible (:147-159)
    return std::make_tuple(std::get<0>(out), std::get<1>(out), std::get<2>(out), screenspace_points, radii > 0, radii);
}

torch::Tensor mappingIterationBackward(std::shared_ptr<GaussianModel> gaussians, std::shared_ptr<GaussianKeyframe> cam,
                                       GaussianPipelineParams& pipe, torch::Tensor& background, torch::Tensor& gt_image,
                                       torch::Tensor& gt_depth, torch::Tensor& mask, float lambda_dssim,
                                       bool update_densification_stats) {
    torch::Tensor override_color;
    const int H = cam->image_height_, W = cam->image_width_;
    auto pkg = GaussianRenderer::render(cam, H, W, gaussians, pipe, background, override_color, 1.0f, false, true);
    torch::Tensor image = std::get<0>(pkg), lf = std::get<1>(pkg), depth = std::get<2>(pkg);
    torch::Tensor viewspace_point_tensor = std::get<3>(pkg), visibility_filter = std::get<4>(pkg), radii = std::get<5>(pkg);
    TORCH_CHECK(image.is_cuda(), "leg_slam_b200 has no CPU path: tensors must live on a CUDA device");
    const c10::cuda::CUDAGuard guard(image.device());
    auto f = image.options();
    torch::Tensor gi = torch::empty_like(image), gl = torch::empty_like(lf), gd = torch::empty_like(depth);
    torch::Tensor loss_out = torch::empty({8}, f);
    torch::Tensor scratch = torch::empty({(int64_t)lgs_mapping_loss_scratch_bytes(W, H)}, f.dtype(torch::kByte));
    torch::Tensor gt_lf = cam->language_features_.contiguous(), gti = gt_image.contiguous(), gtd = gt_depth.contiguous();
    torch::Tensor m = mask.numel() ? mask.contiguous() : mask;
    {
        torch::NoGradGuard no_grad;
        torch::Tensor ic = image.contiguous(), lc = lf.contiguous(), dc = depth.contiguous();
        const int st = lgs_mapping_loss(W, H, (int)gt_lf.size(2), (int)gt_lf.size(1), ic.data_ptr<float>(), lc.data_ptr<float>(),
                                        dc.data_ptr<float>(), gti.data_ptr<float>(), gt_lf.data_ptr<float>(), gtd.data_ptr<float>(),
                                        m.numel() ? m.data_ptr<float>() : nullptr, lambda_dssim, /*cos_sign=*/+1, gi.data_ptr<float>(),
                                        gl.data_ptr<float>(), gd.data_ptr<float>(), loss_out.data_ptr<float>(),
                                        (char*)scratch.data_ptr<uint8_t>(), (void*)at::cuda::getCurrentCUDAStream().stream());
        TORCH_CHECK(st == LGS_OK, "lgs_mapping_loss: ", lgs_status_string(st), " (cudaError ", lgs_last_cuda_error(), ")");
    }
    torch::autograd::backward({image, lf, depth}, {gi, gl, gd});  // loss.backward() of the reference, from the fused gradients
    if (update_densification_stats) {
        torch::NoGradGuard no_grad;
        gaussians->max_radii2D_.index_put_({visibility_filter},
                                           torch::max(gaussians->max_radii2D_.index({visibility_filter}),
                                                      radii.index({visibility_filter}).to(gaussians->max_radii2D_.dtype())));
        gaussians->addDensificationStats(viewspace_point_tensor, visibility_filter);
    }
    return loss_out[0].clone();
}

void mappingIterationStep(std::shared_ptr<GaussianModel> gaussians) {
    gaussians->optimizer_->step();
    gaussians->optimizer_->zero_grad(true);
}
