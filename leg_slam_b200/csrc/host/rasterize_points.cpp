// rasterize_points.cpp -- RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA / markVisible with
// the reference's libtorch signatures (include/rasterize_points.h; reference
// src/rasterize_points.cu:37-228) on the C ABI of liblgs (include/lgs.h).
//
// Differences from the reference's implementation that callers cannot observe:
//   * outputs the kernels fully overwrite are torch::empty, not zero-filled (the reference memsets
//     68*H*W floats per forward and 140 floats/Gaussian per backward);
//   * work goes to torch's current CUDA stream instead of the legacy default stream;
//   * the backward's hand-off scratch is a torch allocation (caching allocator) passed down.
#include "rasterize_points.h"

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <string>

#include "lgs.h"

namespace {
constexpr int NUM_CHANNELS = LGS_NUM_CHANNELS;
constexpr int LF_NUM_CHANNELS = LGS_LF_DIM;

void check(int status, const char* what) {
    TORCH_CHECK(status == LGS_OK, what, ": ", lgs_status_string(status), " (cudaError ", lgs_last_cuda_error(), ")");
}
// data pointer, or nullptr for the reference's empty-tensor sentinels (torch::tensor({}))
template <typename T>
T* ptr(const torch::Tensor& t) {
    return t.numel() == 0 ? nullptr : t.data_ptr<T>();
}
}  // namespace

std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors,
                       const torch::Tensor& lang_feat, const torch::Tensor& opacity, const torch::Tensor& scales,
                       const torch::Tensor& rotations, const float scale_modifier, const torch::Tensor& cov3D_precomp,
                       const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, const float tan_fovx,
                       const float tan_fovy, const int image_height, const int image_width, const torch::Tensor& sh,
                       const int degree, const torch::Tensor& campos, const bool prefiltered,
                       const bool include_lang_feat) {
    if (means3D.ndimension() != 2 || means3D.size(1) != 3) {
        AT_ERROR("means3D must have dimensions (num_points, 3)");
    }
    TORCH_CHECK(means3D.is_cuda(), "leg_slam_b200 has no CPU path: tensors must live on a CUDA device");
    const c10::cuda::CUDAGuard guard(means3D.device());
    const int P = means3D.size(0), H = image_height, W = image_width;
    auto fopt = means3D.options().dtype(torch::kFloat32);
    auto bopt = means3D.options().dtype(torch::kByte);
    if (P == 0) {
        return std::make_tuple(0, torch::zeros({NUM_CHANNELS, H, W}, fopt), torch::zeros({LF_NUM_CHANNELS, H, W}, fopt),
                               torch::zeros({1, H, W}, fopt), torch::zeros({0}, means3D.options().dtype(torch::kInt32)),
                               torch::empty({0}, bopt), torch::empty({0}, bopt), torch::empty({0}, bopt));
    }
    torch::Tensor out_color = torch::empty({NUM_CHANNELS, H, W}, fopt);
    torch::Tensor out_lang_feat = include_lang_feat ? torch::empty({LF_NUM_CHANNELS, H, W}, fopt)
                                                    : torch::zeros({LF_NUM_CHANNELS, H, W}, fopt);
    torch::Tensor out_depth = torch::empty({1, H, W}, fopt);
    torch::Tensor radii = torch::empty({P}, means3D.options().dtype(torch::kInt32));
    torch::Tensor geomBuffer = torch::empty({(long long)lgs_geom_bytes(P)}, bopt);
    torch::Tensor imgBuffer = torch::empty({(long long)lgs_image_bytes(W, H)}, bopt);

    const auto bg = background.contiguous(), m3 = means3D.contiguous(), col = colors.contiguous(),
               lf = lang_feat.contiguous(), op = opacity.contiguous(), sc = scales.contiguous(),
               rot = rotations.contiguous(), cov = cov3D_precomp.contiguous(), vm = viewmatrix.contiguous(),
               pm = projmatrix.contiguous(), shc = sh.contiguous(), cp = campos.contiguous();
    const int M = sh.size(0) != 0 ? (int)sh.size(1) : 0;
    void* stream = at::cuda::getCurrentCUDAStream().stream();

    int rendered = 0;
    check(lgs_forward_stage1(P, degree, M, W, H, ptr<float>(m3), ptr<float>(shc), ptr<float>(col), ptr<float>(op),
                             ptr<float>(sc), scale_modifier, ptr<float>(rot), ptr<float>(cov), ptr<float>(vm),
                             ptr<float>(pm), ptr<float>(cp), tan_fovx, tan_fovy, prefiltered ? 1 : 0,
                             reinterpret_cast<char*>(geomBuffer.data_ptr()), radii.data_ptr<int>(), &rendered, stream),
          "RasterizeGaussiansCUDA (preprocess)");
    // capacity rounded up so that consecutive iterations reuse the same cached allocator block
    const int cap = ((rendered + (1 << 18) - 1) >> 18) << 18;
    torch::Tensor binningBuffer = torch::empty({(long long)lgs_binning_bytes(cap)}, bopt);
    check(lgs_forward_stage2(P, W, H, rendered, ptr<float>(bg), include_lang_feat ? ptr<float>(lf) : nullptr,
                             reinterpret_cast<char*>(geomBuffer.data_ptr()),
                             reinterpret_cast<char*>(binningBuffer.data_ptr()),
                             reinterpret_cast<char*>(imgBuffer.data_ptr()), out_color.data_ptr<float>(),
                             out_lang_feat.data_ptr<float>(), out_depth.data_ptr<float>(), include_lang_feat ? 1 : 0,
                             stream),
          "RasterizeGaussiansCUDA (binning + render)");
    return std::make_tuple(rendered, out_color, out_lang_feat, out_depth, radii, geomBuffer, binningBuffer, imgBuffer);
}

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
                               const bool include_lang_feat) {
    const int P = means3D.size(0);
    const int H = dL_dout_color.size(1), W = dL_dout_color.size(2);
    const int M = sh.size(0) != 0 ? (int)sh.size(1) : 0;
    auto o = means3D.options();
    const bool has_sh = M != 0 && sh.numel() != 0, has_scales = scales.numel() != 0;
    auto mk = [&](std::initializer_list<int64_t> s, bool written) {
        return (P != 0 && written) ? torch::empty(s, o) : torch::zeros(s, o);
    };
    torch::Tensor dL_dmeans3D = mk({P, 3}, true), dL_dmeans2D = mk({P, 3}, true), dL_dcolors = mk({P, NUM_CHANNELS}, true);
    torch::Tensor dL_dlang_feats = mk({P, LF_NUM_CHANNELS}, include_lang_feat);
    torch::Tensor dL_dconic = mk({P, 2, 2}, true), dL_dopacity = mk({P, 1}, true), dL_dcov3D = mk({P, 6}, true);
    torch::Tensor dL_dsh = mk({P, M, 3}, has_sh), dL_dscales = mk({P, 3}, has_scales), dL_drotations = mk({P, 4}, has_scales);
    if (P != 0) {
        TORCH_CHECK(means3D.is_cuda(), "leg_slam_b200 has no CPU path");
        const c10::cuda::CUDAGuard guard(means3D.device());
        const auto bg = background.contiguous(), m3 = means3D.contiguous(), col = colors.contiguous(),
                   lf = lang_feat.contiguous(), sc = scales.contiguous(), rot = rotations.contiguous(),
                   cov = cov3D_precomp.contiguous(), vm = viewmatrix.contiguous(), pm = projmatrix.contiguous(),
                   shc = sh.contiguous(), cp = campos.contiguous(), rad = radii.contiguous(),
                   gc = dL_dout_color.contiguous(), gl = dL_dout_lang_feat.contiguous(), gd = dL_dout_depth.contiguous();
        torch::Tensor scratch = torch::empty({(long long)lgs_backward_scratch_bytes(R, W, H)}, o.dtype(torch::kByte));
        check(lgs_backward(P, degree, M, R, W, H, ptr<float>(bg), ptr<float>(m3), ptr<float>(shc), ptr<float>(col),
                           include_lang_feat ? ptr<float>(lf) : nullptr, ptr<float>(sc), scale_modifier, ptr<float>(rot),
                           ptr<float>(cov), ptr<float>(vm), ptr<float>(pm), ptr<float>(cp), tan_fovx, tan_fovy,
                           rad.data_ptr<int>(), reinterpret_cast<const char*>(geomBuffer.data_ptr()),
                           reinterpret_cast<const char*>(binningBuffer.data_ptr()),
                           reinterpret_cast<const char*>(imageBuffer.data_ptr()), ptr<float>(gc),
                           include_lang_feat ? ptr<float>(gl) : nullptr, ptr<float>(gd), dL_dmeans2D.data_ptr<float>(),
                           dL_dconic.data_ptr<float>(), dL_dopacity.data_ptr<float>(), dL_dcolors.data_ptr<float>(),
                           dL_dlang_feats.data_ptr<float>(), nullptr, dL_dmeans3D.data_ptr<float>(),
                           dL_dcov3D.data_ptr<float>(), ptr<float>(dL_dsh), ptr<float>(dL_dscales),
                           ptr<float>(dL_drotations), include_lang_feat ? 1 : 0, /*zero_outputs=*/1,
                           reinterpret_cast<char*>(scratch.data_ptr()), at::cuda::getCurrentCUDAStream().stream()),
              "RasterizeGaussiansBackwardCUDA");
    }
    return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dlang_feats, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh,
                           dL_dscales, dL_drotations);
}

torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix) {
    const int P = means3D.size(0);
    torch::Tensor present = torch::full({P}, false, means3D.options().dtype(at::kBool));
    if (P != 0) {
        TORCH_CHECK(means3D.is_cuda(), "leg_slam_b200 has no CPU path");
        const c10::cuda::CUDAGuard guard(means3D.device());
        const auto m3 = means3D.contiguous(), vm = viewmatrix.contiguous(), pm = projmatrix.contiguous();
        check(lgs_mark_visible(P, m3.data_ptr<float>(), vm.data_ptr<float>(), pm.data_ptr<float>(),
                               reinterpret_cast<unsigned char*>(present.data_ptr<bool>()),
                               at::cuda::getCurrentCUDAStream().stream()),
              "markVisible");
    }
    return present;
}
