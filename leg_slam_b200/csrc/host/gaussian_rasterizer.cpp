// gaussian_rasterizer.cpp -- L2 of the boundary in C++: the autograd node and the module the reference's renderer holds
// (reference src/gaussian_rasterizer.cpp:18-236), on our L1 functions (rasterize_points.cpp -> liblgs).
//
// Behaviour kept: argument validation and its two messages (:196-206), empty-tensor sentinels for absent inputs
// (:208-220), gradient order (:175), radii returned as a fourth, non-differentiable output.
// Ours: the per-call scalars travel as one IValue tuple held by the node instead of six map entries; undefined
// upstream gradients (an output the loss never touched) become zeros, which the reference would pass on as undefined
// tensors and crash on; the device of the sentinels follows means3D instead of "cuda:current".
#include "gaussian_rasterizer.h"

#include <stdexcept>

namespace {
// the per-call scalars, kept by the node as ONE IValue tuple in this order
enum Call { NUM_RENDERED, SH_DEGREE, IMG_H, IMG_W, SCALE_MODIFIER, TANFOVX, TANFOVY, INCLUDE_LF };
enum Saved { BG, VIEW, PROJ, CAMPOS, COLORS, LANG, MEANS3D, SCALES, ROTS, COV3D, RADII, SH, GEOM, BINNING, IMG };
}  // namespace

torch::Tensor GaussianRasterizer::markVisibleGaussians(torch::Tensor& positions) {
    torch::NoGradGuard no_grad;
    return markVisible(positions, raster_settings_.viewmatrix_, raster_settings_.projmatrix_);
}

torch::autograd::tensor_list GaussianRasterizerFunction::forward(
    torch::autograd::AutogradContext* ctx, torch::Tensor means3D, torch::Tensor means2D, torch::Tensor sh,
    torch::Tensor colors_precomp, torch::Tensor lang_feats, torch::Tensor opacities, torch::Tensor scales,
    torch::Tensor rotations, torch::Tensor cov3Ds_precomp, GaussianRasterizationSettings rs) {
    (void)means2D;  // only its gradient slot is used (screen-space points, gaussian_renderer.cpp:41-48)
    auto [num_rendered, color, lf, depth, radii, geom, binning, img] = RasterizeGaussiansCUDA(
        rs.bg_, means3D, colors_precomp, lang_feats, opacities, scales, rotations, rs.scale_modifier_, cov3Ds_precomp,
        rs.viewmatrix_, rs.projmatrix_, rs.tanfovx_, rs.tanfovy_, rs.image_height_, rs.image_width_, sh, rs.sh_degree_,
        rs.campos_, rs.prefiltered_, rs.include_language_features_);
    ctx->saved_data["call"] = c10::ivalue::Tuple::create(
        {c10::IValue((int64_t)num_rendered), c10::IValue((int64_t)rs.sh_degree_), c10::IValue((int64_t)rs.image_height_),
         c10::IValue((int64_t)rs.image_width_), c10::IValue((double)rs.scale_modifier_), c10::IValue((double)rs.tanfovx_),
         c10::IValue((double)rs.tanfovy_), c10::IValue(rs.include_language_features_)});
    ctx->save_for_backward({rs.bg_, rs.viewmatrix_, rs.projmatrix_, rs.campos_, colors_precomp, lang_feats, means3D, scales,
                            rotations, cov3Ds_precomp, radii, sh, geom, binning, img});
    ctx->mark_non_differentiable({radii});
    return {color, lf, depth, radii};
}

torch::autograd::tensor_list GaussianRasterizerFunction::backward(torch::autograd::AutogradContext* ctx,
                                                                  torch::autograd::tensor_list grad_outputs) {
    const auto call = ctx->saved_data["call"].toTupleRef().elements();
    const int64_t H = call[IMG_H].toInt(), W = call[IMG_W].toInt();
    const auto s = ctx->get_saved_variables();
    const auto fopt = s[MEANS3D].options().dtype(torch::kFloat32);
    auto upstream = [&](size_t i, int64_t channels) {
        if (i < grad_outputs.size() && grad_outputs[i].defined()) return grad_outputs[i].contiguous();
        return torch::zeros({channels, H, W}, fopt);
    };
    const torch::Tensor g_color = upstream(0, 3), g_lf = upstream(1, 64), g_depth = upstream(2, 1);
    auto [dL_dmeans2D, dL_dcolors, dL_dlang_feat, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations] =
        RasterizeGaussiansBackwardCUDA(s[BG], s[MEANS3D], s[RADII], s[COLORS], s[LANG], s[SCALES], s[ROTS],
                                       (float)call[SCALE_MODIFIER].toDouble(), s[COV3D], s[VIEW], s[PROJ],
                                       (float)call[TANFOVX].toDouble(), (float)call[TANFOVY].toDouble(), g_color, g_lf,
                                       g_depth, s[SH], (int)call[SH_DEGREE].toInt(), s[CAMPOS], s[GEOM],
                                       (int)call[NUM_RENDERED].toInt(), s[BINNING], s[IMG], call[INCLUDE_LF].toBool());
    return {dL_dmeans3D, dL_dmeans2D, dL_dsh, dL_dcolors, dL_dlang_feat, dL_dopacity, dL_dscales, dL_drotations, dL_dcov3D,
            torch::Tensor()};
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> GaussianRasterizer::forward(
    torch::Tensor means3D, torch::Tensor means2D, torch::Tensor opacities, bool has_shs, bool has_colors_precomp,
    bool has_lang_feat, bool has_scales, bool has_rotations, bool has_cov3D_precomp, torch::Tensor shs,
    torch::Tensor colors_precomp, torch::Tensor lang_feats, torch::Tensor scales, torch::Tensor rotations,
    torch::Tensor cov3D_precomp) {
    if (has_shs == has_colors_precomp)
        throw std::runtime_error("Please provide excatly one of either SHs or precomputed colors!");
    const bool pair = has_scales && has_rotations, either = has_scales || has_rotations;
    if ((!pair && !has_cov3D_precomp) || (either && has_cov3D_precomp))
        throw std::runtime_error("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
    const torch::Tensor none = torch::empty({0}, means3D.options().dtype(torch::kFloat32));  // data_ptr() == nullptr
    if (!has_shs) shs = none;
    if (!has_colors_precomp) colors_precomp = none;
    if (!has_scales) scales = none;
    if (!has_rotations) rotations = none;
    if (!has_cov3D_precomp) cov3D_precomp = none;
    if (!has_lang_feat) lang_feats = none;
    auto out = rasterizeGaussians(means3D, means2D, shs, colors_precomp, lang_feats, opacities, scales, rotations,
                                  cov3D_precomp, raster_settings_);
    return std::make_tuple(out[0], out[1], out[2], out[3]);
}
