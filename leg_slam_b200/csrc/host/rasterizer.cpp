// rasterizer.cpp -- CudaRasterizer::Rasterizer::{forward, backward, markVisible} with the
// reference's exact signatures (include/cuda_rasterizer/rasterizer.h; reference
// cuda_rasterizer/rasterizer.h:20-93, rasterizer_impl.cu:141-153,198-343,347-453), as a thin
// C++ shim over the C ABI of liblgs (include/lgs.h).  No CUDA code here: plain g++.
#include "cuda_rasterizer/rasterizer.h"

#include <stdexcept>
#include <string>

#include "lgs.h"

namespace {
thread_local void* g_stream = nullptr;

void check(int status, const char* what) {
    if (status == LGS_OK) return;
    throw std::runtime_error(std::string(what) + ": " + lgs_status_string(status) + " (cudaError " +
                             std::to_string(lgs_last_cuda_error()) + ")");
}
}  // namespace

extern "C" void lgs_host_set_stream(void* stream) { g_stream = stream; }

namespace CudaRasterizer {

void Rasterizer::markVisible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present) {
    check(lgs_mark_visible(P, means3D, viewmatrix, projmatrix, reinterpret_cast<unsigned char*>(present), g_stream),
          "Rasterizer::markVisible");
}

int Rasterizer::forward(std::function<char*(size_t)> geometryBuffer, std::function<char*(size_t)> binningBuffer,
                        std::function<char*(size_t)> imageBuffer, const int P, int D, int M,
                        const float* background, const int width, int height, const float* means3D,
                        const float* shs, const float* colors_precomp, const float* lang_feat,
                        const float* opacities, const float* scales, const float scale_modifier,
                        const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                        const float* projmatrix, const float* cam_pos, const float tan_fovx, float tan_fovy,
                        const bool prefiltered, float* out_color, float* out_lang_feat, float* out_depth,
                        int* radii, bool include_lang_feat) {
    // same order of callback invocations as the reference: geometry, image, (readback), binning
    char* geom = geometryBuffer(lgs_geom_bytes(P));
    char* img = imageBuffer(lgs_image_bytes(width, height));
    if (shs == nullptr && colors_precomp == nullptr)  // NUM_CHANNELS == 3 always; the reference's analogue at :243-245
        throw std::runtime_error("For non-RGB, provide precomputed Gaussian colors!");
    int num_rendered = 0;
    check(lgs_forward_stage1(P, D, M, width, height, means3D, shs, colors_precomp, opacities, scales, scale_modifier,
                             rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy,
                             prefiltered ? 1 : 0, geom, radii, &num_rendered, g_stream),
          "Rasterizer::forward (preprocess)");
    char* binning = binningBuffer(lgs_binning_bytes(num_rendered));
    check(lgs_forward_stage2(P, width, height, num_rendered, background, include_lang_feat ? lang_feat : nullptr, geom,
                             binning, img, out_color, out_lang_feat, out_depth, include_lang_feat ? 1 : 0, g_stream),
          "Rasterizer::forward (binning + render)");
    return num_rendered;
}

void Rasterizer::backward(const int P, int D, int M, int R, const float* background, const int width, int height,
                          const float* means3D, const float* shs, const float* colors_precomp,
                          const float* lang_feat, const float* scales, const float scale_modifier,
                          const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                          const float* projmatrix, const float* campos, const float tan_fovx, float tan_fovy,
                          const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer,
                          const float* dL_dpix, const float* dL_dpixlf, const float* dL_dpix_depth,
                          float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                          float* dL_dlang_feat, float* dL_ddepth, float* dL_dmean3D, float* dL_dcov3D,
                          float* dL_dsh, float* dL_dscale, float* dL_drot, bool include_lang_feat) {
    // zero_outputs = 0: the caller pre-zeroed every output, exactly as the reference requires;
    // scratch = NULL: taken from the CUDA stream-ordered pool for the duration of the call.
    check(lgs_backward(P, D, M, R, width, height, background, means3D, shs, colors_precomp,
                       include_lang_feat ? lang_feat : nullptr, scales, scale_modifier, rotations, cov3D_precomp,
                       viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, radii, geom_buffer, binning_buffer,
                       image_buffer, dL_dpix, include_lang_feat ? dL_dpixlf : nullptr, dL_dpix_depth, dL_dmean2D,
                       dL_dconic, dL_dopacity, dL_dcolor, dL_dlang_feat, dL_ddepth, dL_dmean3D, dL_dcov3D, dL_dsh,
                       dL_dscale, dL_drot, include_lang_feat ? 1 : 0, /*zero_outputs=*/0, /*bwd_scratch=*/nullptr,
                       g_stream),
          "Rasterizer::backward");
}

}  // namespace CudaRasterizer
