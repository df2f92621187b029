/*
 * cuda_rasterizer/rasterizer.h -- the reference's L0 raw-pointer interface, re-declared with
 * identical signatures (reference cuda_rasterizer/rasterizer.h:20-93) and implemented on the
 * sm_100a kernels of liblgs through the C ABI of ../lgs.h.
 *
 * A build of LEG-SLAM that includes this header instead of its own and links liblgs_host.so +
 * liblgs.so gets the B200 path with no source change in src/rasterize_points.cu,
 * src/gaussian_rasterizer.cpp or anything above them.
 *
 * Contract kept from the reference:
 *   - every pointer is a device pointer owned by the caller; NULL shs / colors_precomp / scales /
 *     rotations / cov3D_precomp / lang_feat selects the alternative path;
 *   - the three buffer callbacks are each invoked once, synchronously, with the required size in
 *     bytes, and must return device memory that stays valid until backward();
 *   - backward() ACCUMULATES into dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dlang_feat,
 *     dL_ddepth (caller pre-zeroes them) and overwrites the rows of visible Gaussians in
 *     dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot;
 *   - forward() returns num_rendered and throws std::runtime_error where the reference does;
 *   - work is issued on the legacy default stream unless lgs_host_set_stream() was called
 *     (rasterize_points.cpp sets torch's current stream around each call).
 */
#ifndef CUDA_RASTERIZER_H_INCLUDED
#define CUDA_RASTERIZER_H_INCLUDED

#include <functional>
#include <vector>

namespace CudaRasterizer {
class Rasterizer {
public:
    static void markVisible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present);

    static int forward(std::function<char*(size_t)> geometryBuffer, std::function<char*(size_t)> binningBuffer,
                       std::function<char*(size_t)> imageBuffer, const int P, int D, int M, const float* background,
                       const int width, int height, const float* means3D, const float* shs,
                       const float* colors_precomp, const float* lang_feat, const float* opacities,
                       const float* scales, const float scale_modifier, const float* rotations,
                       const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                       const float* cam_pos, const float tan_fovx, float tan_fovy, const bool prefiltered,
                       float* out_color, float* out_lang_feat, float* out_depth, int* radii = nullptr,
                       bool include_lang_feat = false);

    static void backward(const int P, int D, int M, int R, const float* background, const int width, int height,
                         const float* means3D, const float* shs, const float* colors_precomp,
                         const float* lang_feat, const float* scales, const float scale_modifier,
                         const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                         const float* projmatrix, const float* campos, const float tan_fovx, float tan_fovy,
                         const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer,
                         const float* dL_dpix, const float* dL_dpixlf, const float* dL_dpix_depth,
                         float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                         float* dL_dlang_feat, float* dL_ddepth, float* dL_dmean3D, float* dL_dcov3D,
                         float* dL_dsh, float* dL_dscale, float* dL_drot, bool include_lang_feat);
};
}  // namespace CudaRasterizer

/* Stream on which the calls above issue their work (thread-local; default: legacy default stream,
 * like the reference).  Takes a cudaStream_t as void*. */
extern "C" void lgs_host_set_stream(void* stream);

#endif
