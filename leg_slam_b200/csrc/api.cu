// api.cu -- the extern "C" surface of liblgs.so (declared in include/lgs.h).
//
// Host orchestration that replaces CudaRasterizer::Rasterizer::{forward,backward,
// markVisible} (reference rasterizer_impl.cu:141-153,198-343,347-453).  Nothing here
// allocates device memory and nothing throws: every failure becomes an lgs_status.
#include <cstdio>
#include <cstring>
#include "common.cuh"

namespace lgs {

static thread_local int g_last_cuda_error = 0;
void set_last_cuda_error(cudaError_t e) { g_last_cuda_error = (int)e; }

// optional per-stage timing of the CALLING HOST THREAD (lgs_profile_enable): events recorded on the launch stream
struct ProfState {
    bool on = false, made = false;
    cudaEvent_t ev[PM_COUNT];
    bool set[PM_COUNT];
};
static thread_local ProfState g_prof;
void prof_mark(int id, cudaStream_t s) {
    ProfState& p = g_prof;
    if (!p.on) return;
    if (!p.made) {
        for (int i = 0; i < PM_COUNT; ++i) { cudaEventCreate(&p.ev[i]); p.set[i] = false; }
        p.made = true;
    }
    cudaEventRecord(p.ev[id], s);
    p.set[id] = true;
}

// Stream hooks of the calling host thread (lgs_stream_hooks): lets a data-parallel caller overlap the exchange of the
// language-feature parameters / gradients (52 % of the payload) with the stages that do not touch them.
struct StreamHooks {
    cudaEvent_t wait_before_render_fwd = nullptr;   // stage2 waits for it after binning, before the render kernel
    cudaEvent_t record_after_render_bwd = nullptr;  // backward records it once dL_dlang_feat is final
};
static thread_local StreamHooks g_hooks;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace lgs

using namespace lgs;

extern "C" {

const char* lgs_status_string(int status) {
    switch (status) {
        case LGS_OK: return "ok";
        case LGS_ERR_INVALID_ARG: return "invalid argument";
        case LGS_ERR_NO_COLOR: return "neither shs nor colors_precomp given";
        case LGS_ERR_NO_COV: return "neither scales+rotations nor cov3D_precomp given";
        case LGS_ERR_CUDA: return "CUDA runtime error (see lgs_last_cuda_error)";
        case LGS_ERR_ALIGNMENT: return "pointer not 16-byte aligned";
        case LGS_ERR_PREFILTERED: return "prefiltered contract violated";
        default: return "unknown status";
    }
}

int lgs_last_cuda_error(void) { return g_last_cuda_error; }
int lgs_abi_version(void) { return 1; }

// ---- buffer sizes ----------------------------------------------------------------------
// Same trick as the reference's required<T>() (rasterizer_impl.h:66-73): carve from a null
// base and read how far the cursor moved, plus slack for the base pointer's alignment.
size_t lgs_geom_bytes(int P) {
    if (P < 0) return 0;
    GeomState g = geom_from_chunk(nullptr, P);
    return (size_t)reinterpret_cast<uintptr_t>(g.end) + 256;
}
size_t lgs_image_bytes(int W, int H) {
    if (W < 0 || H < 0) return 0;
    ImageState im = image_from_chunk(nullptr, W, H);
    return (size_t)reinterpret_cast<uintptr_t>(im.end) + 256;
}
size_t lgs_backward_scratch_bytes(int R, int W, int H) {
    return (R < 0 || W <= 0 || H <= 0) ? 0 : render_bwd_scratch_bytes(R, W, H);
}
size_t lgs_binning_bytes(int R) {
    if (R < 0) return 0;
    BinningState b = binning_from_chunk(nullptr, R);
    return (size_t)reinterpret_cast<uintptr_t>(b.end) + 256;
}
size_t lgs_debug_keys_bytes(int P, int R) {
    if (P < 0 || R < 0) return 0;
    DebugKeys d = debug_keys_from_chunk(nullptr, P, R);
    return (size_t)reinterpret_cast<uintptr_t>(d.end) + 256;
}

// ---- forward ---------------------------------------------------------------------------
static int forward_stage1_impl(int P, int D, int M, int W, int H, const float* means3D, const float* shs,
                               const float* shs_rest, const float* colors_precomp, const float* opacities,
                               const float* scales, float scale_modifier, const float* rotations,
                               const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                               const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered, char* geom_buffer,
                               int* radii, int* num_rendered_host, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (num_rendered_host) *num_rendered_host = 0;
    if (P < 0 || W <= 0 || H <= 0 || D < 0 || D > 3) return LGS_ERR_INVALID_ARG;
    if (P == 0) return LGS_OK;
    if (!means3D || !opacities || !viewmatrix || !projmatrix || !geom_buffer) return LGS_ERR_INVALID_ARG;
    if (!shs && !colors_precomp) return LGS_ERR_NO_COLOR;
    if (!colors_precomp && (!cam_pos || M < (D + 1) * (D + 1))) return LGS_ERR_INVALID_ARG;
    if (!cov3D_precomp && (!scales || !rotations)) return LGS_ERR_NO_COV;
    if (rotations && !aligned16(rotations)) return LGS_ERR_ALIGNMENT;
    if (shs && !shs_rest && !aligned16(shs)) return LGS_ERR_ALIGNMENT;

    GeomState g = geom_from_chunk(geom_buffer, P);
    if (!radii) radii = g.internal_radii;
    prof_mark(PM_S1_BEGIN, s);
    int st = launch_preprocess(P, D, M, means3D, shs, shs_rest, colors_precomp, opacities, scales, scale_modifier,
                               rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, W, H, tan_fovx,
                               tan_fovy, prefiltered, g, radii, s);
    if (st != LGS_OK) return st;
    prof_mark(PM_PREPROCESS, s);
    if (!num_rendered_host) return LGS_OK;  // the caller keeps R on the device (lgs.h): no read-back, no synchronisation
    // the one readback of the forward (reference rasterizer_impl.cu:281-282); R is accumulated by preprocess itself
    uint32_t R = 0;
    LGS_CUDA_TRY(cudaMemcpyAsync(&R, g.hdr + HDR_R, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    LGS_CUDA_TRY(cudaStreamSynchronize(s));
    *num_rendered_host = (int)R;
    return LGS_OK;
}

int lgs_forward_stage1(int P, int D, int M, int W, int H, const float* means3D, const float* shs,
                       const float* colors_precomp, const float* opacities, const float* scales,
                       float scale_modifier, const float* rotations, const float* cov3D_precomp,
                       const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                       float tan_fovx, float tan_fovy, int prefiltered, char* geom_buffer, int* radii,
                       int* num_rendered_host, void* stream) {
    return forward_stage1_impl(P, D, M, W, H, means3D, shs, nullptr, colors_precomp, opacities, scales, scale_modifier,
                               rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy, prefiltered,
                               geom_buffer, radii, num_rendered_host, stream);
}

int lgs_forward_stage1_split_sh(int P, int D, int M, int W, int H, const float* means3D, const float* features_dc,
                                const float* features_rest, const float* opacities, const float* scales,
                                float scale_modifier, const float* rotations, const float* cov3D_precomp,
                                const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                                float tan_fovy, int prefiltered, char* geom_buffer, int* radii, int* num_rendered_host,
                                void* stream) {
    if (!features_dc || M < 1 || M > 16 || (M > 1 && !features_rest)) return LGS_ERR_INVALID_ARG;
    static const float no_rest[4] = {0.f, 0.f, 0.f, 0.f};  // M == 1: the rest pointer only selects the split layout
    return forward_stage1_impl(P, D, M, W, H, means3D, features_dc, M > 1 ? features_rest : no_rest, nullptr, opacities,
                               scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx,
                               tan_fovy, prefiltered, geom_buffer, radii, num_rendered_host, stream);
}

int lgs_forward_stage2(int P, int W, int H, int R, const float* background, const float* lang_feat,
                       char* geom_buffer, char* binning_buffer, char* image_buffer, float* out_color,
                       float* out_lang_feat, float* out_depth, int include_lang_feat, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P < 0 || W <= 0 || H <= 0 || R < 0) return LGS_ERR_INVALID_ARG;
    if (!background || !image_buffer || !out_color || !out_depth) return LGS_ERR_INVALID_ARG;
    if (include_lang_feat && (!lang_feat || !out_lang_feat)) return LGS_ERR_INVALID_ARG;
    if (include_lang_feat && !aligned16(lang_feat)) return LGS_ERR_ALIGNMENT;
    if (P > 0 && !geom_buffer) return LGS_ERR_INVALID_ARG;
    if (R > 0 && !binning_buffer) return LGS_ERR_INVALID_ARG;

    GeomState g = geom_from_chunk(geom_buffer, P);
    BinningState b = binning_from_chunk(binning_buffer, R);
    ImageState im = image_from_chunk(image_buffer, W, H);
    // radii for key emission: the internal copy is always written by stage1 when the
    // caller passed NULL; otherwise stage1 wrote the caller's array and mirrors it here.
    prof_mark(PM_S2_BEGIN, s);
    int st = launch_binning(P, R, W, H, g, g.internal_radii, b, im, s);
    if (st != LGS_OK) return st;
    if (g_hooks.wait_before_render_fwd) LGS_CUDA_TRY(cudaStreamWaitEvent(s, g_hooks.wait_before_render_fwd, 0));
    st = launch_render_fwd(W, H, R, g, b, im, background, lang_feat, out_color, out_lang_feat, out_depth,
                           include_lang_feat != 0, s);
    prof_mark(PM_RENDER_FWD, s);
    return st;
}

// ---- markVisible -----------------------------------------------------------------------
int lgs_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     unsigned char* present, void* stream) {
    (void)projmatrix;  // the reference computes p_proj and never uses it (auxiliary.h:150-154)
    if (P < 0) return LGS_ERR_INVALID_ARG;
    if (P == 0) return LGS_OK;
    if (!means3D || !viewmatrix || !present) return LGS_ERR_INVALID_ARG;
    return launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
}

// ---- backward --------------------------------------------------------------------------
static int backward_impl(int P, int D, int M, int R, int W, int H, const float* background, const float* means3D,
                 const float* shs, const float* shs_rest, const float* colors_precomp, const float* lang_feat, const float* scales,
                 float scale_modifier, const float* rotations, const float* cov3D_precomp,
                 const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                 float tan_fovy, const int* radii, const char* geom_buffer, const char* binning_buffer,
                 const char* image_buffer, const float* dL_dpix, const float* dL_dpix_lf,
                 const float* dL_dpix_depth, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                 float* dL_dcolor, float* dL_dlang_feat, float* dL_ddepth, float* dL_dmean3D,
                 float* dL_dcov3D, float* dL_dsh, float* dL_dsh_rest, int accumulate_sh, float* dL_dscale, float* dL_drot,
                 int include_lang_feat, int zero_outputs, char* bwd_scratch, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P < 0 || W <= 0 || H <= 0 || R < 0 || D < 0 || D > 3) return LGS_ERR_INVALID_ARG;
    if (P == 0) return LGS_OK;
    if (!background || !means3D || !viewmatrix || !projmatrix || !geom_buffer || !image_buffer)
        return LGS_ERR_INVALID_ARG;
    if (R > 0 && !binning_buffer) return LGS_ERR_INVALID_ARG;
    if (!dL_dpix || !dL_dpix_depth || !dL_dmean2D || !dL_dconic || !dL_dopacity || !dL_dcolor || !dL_dmean3D ||
        !dL_dcov3D)
        return LGS_ERR_INVALID_ARG;
    if (include_lang_feat && (!lang_feat || !dL_dpix_lf || !dL_dlang_feat)) return LGS_ERR_INVALID_ARG;
    if (shs && (!dL_dsh || !cam_pos)) return LGS_ERR_INVALID_ARG;
    if (scales && (!rotations || !dL_dscale || !dL_drot)) return LGS_ERR_INVALID_ARG;
    if (!cov3D_precomp && !scales) return LGS_ERR_NO_COV;
    if ((rotations && !aligned16(rotations)) || (dL_drot && !aligned16(dL_drot)) || !aligned16(dL_dconic) ||
        (include_lang_feat && (!aligned16(lang_feat) || !aligned16(dL_dlang_feat))))
        return LGS_ERR_ALIGNMENT;

    GeomState g = geom_from_chunk(const_cast<char*>(geom_buffer), P);
    BinningState b = binning_from_chunk(const_cast<char*>(binning_buffer), R);
    ImageState im = image_from_chunk(const_cast<char*>(image_buffer), W, H);
    if (!radii) radii = g.internal_radii;

    int st;
    prof_mark(PM_BWD_BEGIN, s);
    // zero_outputs: the accumulated-into arrays are cleared by the render backward's pixel kernel as a side job (no separate
    // memset launch); without instances there is no render backward, so clear them here
    if (zero_outputs && R <= 0) {
        st = launch_zero_grads(P, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dlang_feat, dL_ddepth,
                               include_lang_feat != 0, s);
        if (st != LGS_OK) return st;
    }
    prof_mark(PM_ZERO, s);
    if (R > 0) {
        // the pixel -> channel hand-off records; callers that keep the reference's signature pass no
        // scratch, then it comes from the stream-ordered pool (no synchronisation)
        char* scratch = bwd_scratch;
        if (!scratch) LGS_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&scratch), render_bwd_scratch_bytes(R, W, H), s));
        st = launch_render_bwd(P, W, H, R, g, b, im, background, lang_feat, dL_dpix, dL_dpix_lf, dL_dpix_depth,
                               dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dlang_feat, dL_ddepth,
                               include_lang_feat != 0, scratch, zero_outputs != 0, s);
        if (!bwd_scratch) {
            cudaError_t e = cudaFreeAsync(scratch, s);
            if (st == LGS_OK && e != cudaSuccess) { set_last_cuda_error(e); st = LGS_ERR_CUDA; }
        }
        if (st != LGS_OK) return st;
    }
    prof_mark(PM_RENDER_BWD, s);
    if (g_hooks.record_after_render_bwd) LGS_CUDA_TRY(cudaEventRecord(g_hooks.record_after_render_bwd, s));
    const float* cov3D = cov3D_precomp ? cov3D_precomp : g.cov3D;
    st = launch_preprocess_bwd(P, D, M, means3D, radii, shs, shs_rest, scales, rotations, scale_modifier, cov3D,
                               viewmatrix, projmatrix, cam_pos, W, H, tan_fovx, tan_fovy, g, dL_dmean2D,
                               dL_dconic, dL_dmean3D, dL_dcolor, dL_dcov3D, dL_dsh, dL_dsh_rest, accumulate_sh != 0,
                               dL_dscale, dL_drot, zero_outputs != 0, s);
    prof_mark(PM_PREPROCESS_BWD, s);
    return st;
}

int lgs_backward(int P, int D, int M, int R, int W, int H, const float* background, const float* means3D,
                 const float* shs, const float* colors_precomp, const float* lang_feat, const float* scales,
                 float scale_modifier, const float* rotations, const float* cov3D_precomp,
                 const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                 float tan_fovy, const int* radii, const char* geom_buffer, const char* binning_buffer,
                 const char* image_buffer, const float* dL_dpix, const float* dL_dpix_lf,
                 const float* dL_dpix_depth, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                 float* dL_dcolor, float* dL_dlang_feat, float* dL_ddepth, float* dL_dmean3D,
                 float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, int include_lang_feat,
                 int zero_outputs, char* bwd_scratch, void* stream) {
    return backward_impl(P, D, M, R, W, H, background, means3D, shs, nullptr, colors_precomp, lang_feat, scales,
                         scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy, radii,
                         geom_buffer, binning_buffer, image_buffer, dL_dpix, dL_dpix_lf, dL_dpix_depth, dL_dmean2D, dL_dconic,
                         dL_dopacity, dL_dcolor, dL_dlang_feat, dL_ddepth, dL_dmean3D, dL_dcov3D, dL_dsh, nullptr, 0,
                         dL_dscale, dL_drot, include_lang_feat, zero_outputs, bwd_scratch, stream);
}

int lgs_backward_split_sh(int P, int D, int M, int R, int W, int H, const float* background, const float* means3D,
                          const float* features_dc, const float* features_rest, const float* lang_feat,
                          const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
                          const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                          float tan_fovy, const int* radii, const char* geom_buffer, const char* binning_buffer,
                          const char* image_buffer, const float* dL_dpix, const float* dL_dpix_lf,
                          const float* dL_dpix_depth, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                          float* dL_dcolor, float* dL_dlang_feat, float* dL_ddepth, float* dL_dmean3D, float* dL_dcov3D,
                          float* dL_dfeatures_dc, float* dL_dfeatures_rest, int accumulate_sh, float* dL_dscale,
                          float* dL_drot, int include_lang_feat, int zero_outputs, char* bwd_scratch, void* stream) {
    if (!features_dc || !dL_dfeatures_dc || M < 1 || M > 16 || (M > 1 && (!features_rest || !dL_dfeatures_rest)))
        return LGS_ERR_INVALID_ARG;
    static const float no_rest[4] = {0.f, 0.f, 0.f, 0.f};
    // M == 1: any non-NULL rest pointers select the split layout; they are never dereferenced (row = 3)
    return backward_impl(P, D, M, R, W, H, background, means3D, features_dc, M > 1 ? features_rest : no_rest, nullptr,
                         lang_feat, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos,
                         tan_fovx, tan_fovy, radii, geom_buffer, binning_buffer, image_buffer, dL_dpix, dL_dpix_lf,
                         dL_dpix_depth, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dlang_feat, dL_ddepth, dL_dmean3D,
                         dL_dcov3D, dL_dfeatures_dc, M > 1 ? dL_dfeatures_rest : dL_dfeatures_dc, accumulate_sh, dL_dscale,
                         dL_drot, include_lang_feat, zero_outputs, bwd_scratch, stream);
}

// ---- stream hooks ----------------------------------------------------------------------
int lgs_stream_hooks(void* wait_before_render_fwd, void* record_after_render_bwd) {
    g_hooks.wait_before_render_fwd = (cudaEvent_t)wait_before_render_fwd;
    g_hooks.record_after_render_bwd = (cudaEvent_t)record_after_render_bwd;
    return LGS_OK;
}

// ---- per-stage timing ------------------------------------------------------------------
int lgs_profile_enable(int on) {
    g_prof.on = on != 0;
    return LGS_OK;
}
int lgs_profile_read(float* ms, int n) {
    // ms[0..8] = preprocess, emit_keys, sort, tile_ranges, render_fwd, zero_grads, render_bwd_pix, render_bwd_chan,
    //            preprocess_bwd of this thread's most recent forward + backward; the caller must have synchronised the
    //            stream.  Entries whose stage did not run are -1.
    static const int a[9] = {PM_S1_BEGIN, PM_S2_BEGIN, PM_EMIT, PM_SORT, PM_RANGES, PM_BWD_BEGIN, PM_ZERO, PM_RENDER_BWD_PIX,
                             PM_RENDER_BWD};
    static const int b[9] = {PM_PREPROCESS, PM_EMIT, PM_SORT, PM_RANGES, PM_RENDER_FWD, PM_ZERO, PM_RENDER_BWD_PIX, PM_RENDER_BWD,
                             PM_PREPROCESS_BWD};
    if (!ms || n < 9) return LGS_ERR_INVALID_ARG;
    const ProfState& p = g_prof;
    for (int i = 0; i < 9; ++i) {
        ms[i] = -1.f;
        if (p.made && p.set[a[i]] && p.set[b[i]]) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, p.ev[a[i]], p.ev[b[i]]) == cudaSuccess) ms[i] = t;
            else (void)cudaGetLastError();
        }
    }
    return LGS_OK;
}

// ---- frame status (device-side header of the geometry buffer) ---------------------------
int lgs_forward_status(const char* geom_buffer, int P, unsigned int* status_host, void* stream) {
    if (!geom_buffer || !status_host || P < 0) return LGS_ERR_INVALID_ARG;
    GeomState g = geom_from_chunk(const_cast<char*>(geom_buffer), P);
    LGS_CUDA_TRY(cudaMemcpyAsync(status_host, g.hdr, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return LGS_OK;
}

// ---- introspection ---------------------------------------------------------------------
int lgs_view_binning(const char* binning_buffer, int R, lgs_binning_view* out) {
    if (!out || R < 0) return LGS_ERR_INVALID_ARG;
    BinningState b = binning_from_chunk(const_cast<char*>(binning_buffer), R);
    out->point_list = b.point_list;
    out->keys_sorted32 = b.keys[1];
    return LGS_OK;
}
int lgs_debug_reference_keys(int P, int R, int W, int H, const char* geom_buffer, const char* binning_buffer,
                             const char* image_buffer, char* debug_buffer, lgs_reference_keys_view* out, void* stream) {
    if (!out || P < 0 || R < 0 || W <= 0 || H <= 0 || !debug_buffer) return LGS_ERR_INVALID_ARG;
    if ((P > 0 && !geom_buffer) || (R > 0 && !binning_buffer) || !image_buffer) return LGS_ERR_INVALID_ARG;
    GeomState g = geom_from_chunk(const_cast<char*>(geom_buffer), P);
    BinningState b = binning_from_chunk(const_cast<char*>(binning_buffer), R);
    ImageState im = image_from_chunk(const_cast<char*>(image_buffer), W, H);
    DebugKeys d = debug_keys_from_chunk(debug_buffer, P, R);
    out->keys_unsorted = d.keys_unsorted;
    out->values_unsorted = d.vals_unsorted;
    out->keys_sorted = d.keys_sorted;
    out->point_offsets = d.offsets;
    return launch_debug_reference_keys(P, R, W, H, g, b, im, d, (cudaStream_t)stream);
}
int lgs_view_image(const char* image_buffer, int W, int H, lgs_image_view* out) {
    if (!out || W <= 0 || H <= 0) return LGS_ERR_INVALID_ARG;
    ImageState im = image_from_chunk(const_cast<char*>(image_buffer), W, H);
    out->ranges = reinterpret_cast<const uint32_t*>(im.ranges);
    out->final_T = im.final_T;
    out->n_contrib = im.n_contrib;
    return LGS_OK;
}
int lgs_view_geom(const char* geom_buffer, int P, lgs_geom_view* out) {
    if (!out || P < 0) return LGS_ERR_INVALID_ARG;
    GeomState g = geom_from_chunk(const_cast<char*>(geom_buffer), P);
    out->records = reinterpret_cast<const float*>(g.rec);
    out->cov3D = g.cov3D;
    out->tiles_touched = g.tiles_touched;
    out->internal_radii = g.internal_radii;
    out->clamped = g.clamped;
    return LGS_OK;
}

}  // extern "C"
