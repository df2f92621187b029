// common.cuh -- shared definitions for the sm_100a kernels of liblgs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "lgs.h"

namespace lgs {

constexpr int TILE = LGS_TILE;          // 8x8 pixel tiles (reference config.h:17-18)
constexpr int TILE_PIX = TILE * TILE;   // 64
constexpr int LF = LGS_LF_DIM;          // 64 language-feature channels
constexpr int NCH = LGS_NUM_CHANNELS;   // 3 colour channels
constexpr int REC_FLOATS = 12;          // per-Gaussian render record, 48 B (3 x float4)
constexpr int HREC_FLOATS = 68;         // render-backward half-record: {gx - cx, gy - cy, 0, id} | w[32] | t[32]  (272 B)

// ---- per-Gaussian render record (geometry buffer) --------------------------------
// One 48-byte record replaces the reference's separate means2D / depths /
// conic_opacity / rgb arrays (rasterizer_impl.h:33-48): the render kernels move
// one contiguous 48 B chunk per (tile, Gaussian) instance instead of four gathers.
//   q0 = (x, y, depth, Gaussian index bits)   q1 = (conic.a, conic.b, conic.c, opacity)   q2 = (r, g, b, 0)
struct __align__(16) GaussRec {
    float4 q0, q1, q2;
};
static_assert(sizeof(GaussRec) == 48, "record must be 48 bytes");

// ---- carved views of the three opaque buffers ------------------------------------
// geometry-buffer header: everything the later kernels of a forward need to know about the frame, kept ON THE DEVICE
// (zeroed by stage1 before preprocess; stage2 and the backward read it there, never through the host)
enum GeomHeader {
    HDR_R = 0,          // number of (Gaussian, tile) instances = sum of tiles_touched (accumulated by preprocess)
    HDR_MAX_DEPTH = 1,  // largest depth bit pattern of a rendered Gaussian (bounds the sort keys)
    HDR_OVERFLOW = 2,   // set by stage2 when HDR_R exceeds the capacity of the binning buffer it was given
    HDR_ERROR = 3,      // set when a sort look-back gave up (never observed; the frame's lists are then incomplete)
    HDR_CURSOR = 4,     // key emission: next free slot
    HDR_TICKET = 5,     // [3] tile tickets of the three radix passes
    HDR_WORDS = 16
};
constexpr int RS_BITS = 11;                 // radix digit width: passes over bits 0-10, 11-21, 22-31 of the 32-bit key
constexpr int RS_BINS = 1 << RS_BITS;       // 2048
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 20;                // pairs per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 5120 pairs per tile
constexpr int RS_GROUP = 16;                // tiles per look-back group

struct GeomState {
    GaussRec* rec;            // [P]
    float* cov3D;             // [P*6]
    uint32_t* tiles_touched;  // [P]
    int32_t* internal_radii;  // [P]
    uint8_t* clamped;         // [P] bit c = colour channel c clamped at 0
    uint32_t* hdr;            // [HDR_WORDS] GeomHeader
    uint32_t* hist;           // [3][RS_BINS] digit histograms of the sort keys (filled by key emission)
    char* end;
};
struct ImageState {
    uint2* ranges;        // [tiles]
    float* final_T;       // [H*W]
    uint32_t* n_contrib;  // [H*W]
    uint32_t* tile_last;  // [tiles] max n_contrib over the tile's pixels (bwd skips the rest)
    char* end;
};
struct BinningState {
    uint32_t* keys[2];      // [R] ping-pong: emitted keys in [0]; sorted keys in [1] after the three passes
    uint32_t* vals[2];      // [R] ping-pong Gaussian indices
    uint32_t* point_list;   // = vals[1]: sorted Gaussian indices
    uint32_t* grp_stat;     // [3][groups][RS_BINS] look-back status of the tile groups (flag << 30 | count)
    uint16_t* tile_agg;     // [3][tiles][RS_BINS]  per-tile digit counts (bit 15 = published)
    uint32_t n_tiles_cap, n_groups_cap;
    size_t status_bytes;
    char* end;
};
struct DebugKeys {            // lgs_debug_reference_keys: the reference's arrays (rasterizer_impl.h:50-63)
    uint64_t* keys_unsorted;  // [R]
    uint64_t* keys_sorted;    // [R]
    uint32_t* vals_unsorted;  // [R]
    uint32_t* offsets;        // [P] inclusive scan of tiles_touched (the reference's point_offsets)
    char* end;
};

template <typename T>
static inline void carve(char*& p, T*& out, size_t count, size_t align = 256) {
    uintptr_t a = (reinterpret_cast<uintptr_t>(p) + align - 1) & ~(uintptr_t)(align - 1);
    out = reinterpret_cast<T*>(a);
    p = reinterpret_cast<char*>(out + count);
}

GeomState geom_from_chunk(char* chunk, int P);
ImageState image_from_chunk(char* chunk, int W, int H);
BinningState binning_from_chunk(char* chunk, int R);
DebugKeys debug_keys_from_chunk(char* chunk, int P, int R);

// ---- error plumbing ----------------------------------------------------------------
void set_last_cuda_error(cudaError_t e);
#define LGS_CUDA_TRY(expr)                                   \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) {                             \
            ::lgs::set_last_cuda_error(_e);                  \
            return LGS_ERR_CUDA;                             \
        }                                                    \
    } while (0)
#define LGS_LAUNCH_CHECK() LGS_CUDA_TRY(cudaGetLastError())

// ---- optional per-stage timing (lgs_profile_enable): CUDA events recorded on the launch stream
enum ProfMark { PM_S1_BEGIN = 0, PM_PREPROCESS, PM_S2_BEGIN, PM_EMIT, PM_SORT, PM_RANGES, PM_RENDER_FWD,
                PM_BWD_BEGIN, PM_ZERO, PM_RENDER_BWD_PIX, PM_RENDER_BWD, PM_PREPROCESS_BWD, PM_COUNT };
void prof_mark(int id, cudaStream_t s);

// ---- kernel launchers (one per .cu) ------------------------------------------------
struct ViewParams {
    float view[16];
    float proj[16];
    float cam[3];
    float tan_fovx, tan_fovy, focal_x, focal_y;
    int W, H, tiles_x, tiles_y;
};

int launch_preprocess(int P, int D, int M, const float* means3D, const float* shs, const float* shs_rest,
                      const float* colors_precomp, const float* opacities, const float* scales,
                      float scale_modifier, const float* rotations, const float* cov3D_precomp,
                      const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                      int W, int H, float tan_fovx, float tan_fovy, int prefiltered,
                      GeomState& g, int* radii, cudaStream_t s);
int launch_mark_visible(int P, const float* means3D, const float* viewmatrix,
                        unsigned char* present, cudaStream_t s);
// R: the capacity the binning buffer was carved for (>= the frame's instance count, which the kernels read from g.hdr)
int launch_binning(int P, int R, int W, int H, const GeomState& g, const int* radii,
                   BinningState& b, ImageState& im, cudaStream_t s);
int launch_debug_reference_keys(int P, int R, int W, int H, const GeomState& g, const BinningState& b, const ImageState& im,
                                DebugKeys& d, cudaStream_t s);
int launch_render_fwd(int W, int H, int R, const GeomState& g, const BinningState& b,
                      ImageState& im, const float* background, const float* lang_feat,
                      float* out_color, float* out_lang_feat, float* out_depth,
                      bool include_lf, cudaStream_t s);
int launch_render_fwd_tc(int W, int H, const GeomState& g, const BinningState& b, ImageState& im,
                         const float* background, const float* lang_feat, float* out_color,
                         float* out_lang_feat, float* out_depth, cudaStream_t s);
int launch_render_bwd(int P, int W, int H, int R, const GeomState& g, const BinningState& b,
                      const ImageState& im, const float* background, const float* lang_feat,
                      const float* dL_dpix, const float* dL_dpix_lf, const float* dL_dpix_depth,
                      float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                      float* dL_dlang_feat, float* dL_ddepth, bool include_lf, char* scratch, bool zero_outputs, cudaStream_t s);
size_t render_bwd_scratch_bytes(int R, int W, int H);
int launch_render_bwd_chan_tc(int W, int H, const ImageState& im, const float* dL_dpix, const float* dL_dpix_lf,
                              const float* dL_dpix_depth, const float* hrec, const uint32_t* hcount, uint32_t* work_counter,
                              float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                              float* dL_dlang_feat, float* dL_ddepth, cudaStream_t s);
int launch_preprocess_bwd(int P, int D, int M, const float* means3D, const int* radii,
                          const float* shs, const float* shs_rest, const float* scales, const float* rotations,
                          float scale_modifier, const float* cov3D, const float* viewmatrix,
                          const float* projmatrix, const float* cam_pos, int W, int H,
                          float tan_fovx, float tan_fovy, const GeomState& g,
                          float* dL_dmean2D, float* dL_dconic, float* dL_dmean3D,
                          const float* dL_dcolor, float* dL_dcov3D, float* dL_dsh, float* dL_dsh_rest,
                          bool accumulate_sh, float* dL_dscale, float* dL_drot, bool write_zeros, cudaStream_t s);
int launch_zero_grads(int P, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                      float* dL_dcolor, float* dL_dlang_feat, float* dL_ddepth, bool include_lf,
                      cudaStream_t s);

#ifdef __CUDACC__
// ---- per-fragment falloff shared by the forward and backward blend kernels -----------------
// power = -0.5*(a*dx*dx + c*dy*dy) - b*dx*dy with the exact association the reference's
// renderCUDA compiles to on sm_100 (forward.cu:337-341 -> SASS: s = fma(dx, dx*a, dy*(dy*c));
// power = fma(s, -0.5, -(dy*(dx*b)))), so that the alpha tests, T and n_contrib are decided on
// bit-identical values.
__device__ __forceinline__ float eval_power(float gx, float gy, float px, float py, float a, float b, float c,
                                            float& dx, float& dy) {
    dx = __fsub_rn(gx, px);
    dy = __fsub_rn(gy, py);
    const float s = __fmaf_rn(dx, __fmul_rn(dx, a), __fmul_rn(dy, __fmul_rn(dy, c)));
    return __fmaf_rn(s, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, b)));
}

// Conservative footprint test of one Gaussian against a pixel rectangle [x0,x1] x [y0,y1] (pixel centres): can ANY pixel in
// it pass the blend kernels' tests power <= 0 and alpha = o*exp(power) >= 1/255?  alpha >= 1/255 needs 0.5 * d^T Q d <= tau =
// ln(255 o).  The minimum of the quadratic form over the rectangle is 0 when the centre lies inside, otherwise the smallest
// of the four edge minima, each a clamped 1-D quadratic.  tau is inflated by 1 % + 0.01 and the comparison gets a slack
// proportional to the magnitude of the terms, so that the rounding of this evaluation and of the per-pixel one can never
// reject a pair the exact per-pixel test accepts; degenerate conics are reported as touching.  (Round 1 used the bounding box
// of the ellipse; this test lets 17 % fewer (Gaussian, 32-pixel half) pairs through at cfgB -- tools/analyze_workload.py --
// with bit-identical images and gradients, profiles/r02_experimental_switches.json.)
__device__ __forceinline__ bool footprint_touches(float gx, float gy, float a, float b, float c, float o, float x0,
                                                  float x1, float y0, float y1) {
    const float tau = __logf(255.0f * o) * 1.01f + 0.01f;
    if (!(tau > 0.f)) return !(tau <= 0.f);  // o < 1/255: nothing can pass; NaN: stay conservative
    const float det = a * c - b * b;
    if (!(det > 0.f) || !(a > 0.f) || !(c > 0.f)) return true;
    const float X0 = x0 - gx, X1 = x1 - gx, Y0 = y0 - gy, Y1 = y1 - gy;  // the rectangle relative to the centre
    if (X0 <= 0.f && X1 >= 0.f && Y0 <= 0.f && Y1 >= 0.f) return true;
    const float nb_c = -b / c, nb_a = -b / a;
    float q = 3.0e38f;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const float X = e ? X1 : X0, Y = e ? Y1 : Y0;
        const float dy = fminf(fmaxf(nb_c * X, Y0), Y1);  // edge dx = X
        q = fminf(q, a * X * X + 2.0f * b * X * dy + c * dy * dy);
        const float dx = fminf(fmaxf(nb_a * Y, X0), X1);  // edge dy = Y
        q = fminf(q, a * dx * dx + 2.0f * b * dx * Y + c * Y * Y);
    }
    if (!(q == q)) return true;
    const float mx = fmaxf(fabsf(X0), fabsf(X1)), my = fmaxf(fabsf(Y0), fabsf(Y1));
    const float slack = 2.0e-6f * (a * mx * mx + 2.0f * fabsf(b) * mx * my + c * my * my);
    return 0.5f * q <= tau + slack;
}
#endif

}  // namespace lgs
