// ref_loss_wrap.cpp -- pybind entry points around the UNMODIFIED reference loss (TEST INFRASTRUCTURE, oracle/_ref/ref_loss.so).
// The header is included from where it lies (/root/reference/include/loss_utils.h, -I on the command line of
// oracle/build_ref.py); nothing of it is copied.  mapping_loss() chains its functions exactly as
// GaussianMapper::trainForOneIteration does (reference src/gaussian_mapper.cpp:707-721): nearest resize of the ground-truth
// feature map, the undistortion mask, (1 - l) L1 + l (1 - SSIM) + mean cosine similarity (ADDED, appendix A.11) + depth L1.
#include <torch/extension.h>

#include <vector>

#include "loss_utils.h"

static torch::Tensor ref_l1(torch::Tensor a, torch::Tensor b) { return loss_utils::l1_loss(a, b); }
static torch::Tensor ref_psnr(torch::Tensor a, torch::Tensor b) { return loss_utils::psnr(a, b); }
static torch::Tensor ref_cos(torch::Tensor a, torch::Tensor b) { return loss_utils::cosine_similarity(a, b); }
static torch::Tensor ref_ssim(torch::Tensor a, torch::Tensor b) {
    return loss_utils::ssim(a, b, a.is_cuda() ? torch::kCUDA : torch::kCPU);
}

static torch::Tensor mapping_loss(torch::Tensor rendered_image, torch::Tensor rendered_lf, torch::Tensor rendered_depth,
                                  torch::Tensor gt_image, torch::Tensor language_features, torch::Tensor gt_depth, torch::Tensor mask,
                                  double lambda_dssim) {
    auto gt_lf = torch::squeeze(torch::nn::functional::interpolate(
        torch::unsqueeze(language_features, 0),
        torch::nn::functional::InterpolateFuncOptions().size(std::vector<int64_t>({rendered_lf.size(1), rendered_lf.size(2)}))));
    torch::Tensor masked_image = rendered_image * mask;
    torch::Tensor masked_lf = rendered_lf * torch::unsqueeze(mask[0], 0);
    torch::Tensor masked_depth = rendered_depth * torch::unsqueeze(mask[0], 0);
    auto Ll1 = loss_utils::l1_loss(masked_image, gt_image);
    auto similarity_lf = loss_utils::cosine_similarity(masked_lf, gt_lf);
    auto Ll1_depth = loss_utils::l1_loss(masked_depth, gt_depth);
    auto dev = rendered_image.is_cuda() ? torch::kCUDA : torch::kCPU;
    return (1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - loss_utils::ssim(masked_image, gt_image, dev)) + similarity_lf + Ll1_depth;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("l1_loss", &ref_l1);
    m.def("psnr", &ref_psnr);
    m.def("cosine_similarity", &ref_cos);
    m.def("ssim", &ref_ssim);
    m.def("mapping_loss", &mapping_loss);
}
