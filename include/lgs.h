/*
 * lgs.h -- C ABI of the B200-native LEG-SLAM mapping hot path (liblgs.so).
 *
 * This is the drop-in boundary *beneath* the reference's raw-pointer interface
 *   CudaRasterizer::Rasterizer::{forward,backward,markVisible}
 *   (reference: cuda_rasterizer/rasterizer.h:24-92, impl rasterizer_impl.cu:141-453)
 * and beneath the libtorch functions
 *   RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA / markVisible
 *   (reference: include/rasterize_points.h:19-72, impl src/rasterize_points.cu:37-228).
 * include/cuda_rasterizer/rasterizer.h and include/rasterize_points.h in this repo
 * re-declare those two levels with identical signatures; both are thin shims
 * over the functions below.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless it says "host";
 *   - nothing here allocates device memory; the caller sizes the three opaque
 *     work buffers with lgs_*_bytes() (replaces required<GeometryState/ImageState/
 *     BinningState>(), reference rasterizer_impl.h:66-73);
 *   - every launch goes to `stream` (a cudaStream_t passed as void*); the
 *     reference uses the legacy default stream;
 *   - return value: 0 = LGS_OK, otherwise an lgs_status (no exceptions cross the ABI);
 *   - a NULL shs / colors_precomp / scales / rotations / cov3D_precomp / lang_feat
 *     selects the alternative path exactly like the reference's nullptr tests
 *     (forward.cu:205,241; backward.cu:390,394);
 *   - matrices are 4x4 column-major as the reference kernels index them
 *     (auxiliary.h:58-77), i.e. torch row-major of the transposed matrix.
 *   - LGS_LF_DIM (64) language-feature channels, 3 colour channels, 8x8 tiles
 *     (reference config.h:15-18, CMakeLists.txt:4).
 */
#ifndef LGS_H_INCLUDED
#define LGS_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGS_LF_DIM 64
#define LGS_NUM_CHANNELS 3
#define LGS_TILE 8

typedef enum lgs_status {
    LGS_OK = 0,
    LGS_ERR_INVALID_ARG = 1,   /* bad sizes / NULL where a pointer is required        */
    LGS_ERR_NO_COLOR = 2,      /* neither shs nor colors_precomp (rasterizer_impl.cu:243-245 analogue) */
    LGS_ERR_NO_COV = 3,        /* neither scales+rotations nor cov3D_precomp           */
    LGS_ERR_CUDA = 4,          /* a CUDA runtime call failed; see lgs_last_cuda_error  */
    LGS_ERR_ALIGNMENT = 5,     /* a pointer that must be 16-byte aligned is not        */
    LGS_ERR_PREFILTERED = 6    /* reserved: prefiltered contract violated              */
} lgs_status;

const char* lgs_status_string(int status);
/* cudaError_t value of the most recent failing CUDA call on this host thread (0 if none). */
int lgs_last_cuda_error(void);
int lgs_abi_version(void);

/* ---- opaque work-buffer sizes (bytes) -------------------------------------------- */
size_t lgs_geom_bytes(int P);          /* replaces required<GeometryState>(P)          */
size_t lgs_image_bytes(int W, int H);  /* replaces required<ImageState>(W*H)           */
size_t lgs_binning_bytes(int R);       /* replaces required<BinningState>(R)           */

/* ---- forward ---------------------------------------------------------------------
 * Split in two because the reference sizes the binning buffer from the number of
 * (Gaussian, tile) instances R = num_rendered, read back once per forward
 * (rasterizer_impl.cu:281-286).
 *
 * stage1 = FORWARD::preprocess (rasterizer_impl.cu:248-275):
 *   writes radii[P] (int32; pass NULL to use an internal array), fills geom_buffer -- including a small device-side frame
 *   header: R (accumulated by preprocess itself, so there is no scan over Gaussians), the frame's largest depth, status
 *   flags -- then writes *num_rendered_host and synchronises `stream` once (the :282 readback).
 *   num_rendered_host == NULL skips the read-back AND the synchronisation: R stays on the device, every later kernel
 *   reads it there, and the caller hands stage2 / the backward the CAPACITY its binning buffer was sized for instead
 *   (see lgs_forward_stage2; lgs_forward_status fetches R and the overflow flag whenever the caller next synchronises).
 * stage2 = duplicateWithKeys + the tile|depth radix sort + identifyTileRanges + renderCUDA (rasterizer_impl.cu:290-340),
 *   every kernel hand-written (csrc/binning.cu):
 *   out_color [3,H,W], out_lang_feat [64,H,W] (written only if include_lang_feat),
 *   out_depth [1,H,W]. Every pixel of the written outputs is written (no pre-zero
 *   needed).
 *   R is the number of instances binning_buffer was sized for with lgs_binning_bytes(R): the value stage1 returned, or
 *   any upper bound of it.  The same R must be passed to lgs_backward.  If the frame has more instances than R, the
 *   surplus is dropped (memory-safe, images incomplete) and the overflow flag of lgs_forward_status is set.
 *   Call stage2 once per stage1 (it consumes counters stage1 zeroed in the geometry buffer).
 */
int lgs_forward_stage1(
    int P, int D, int M, int W, int H,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* opacities, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* cam_pos,
    float tan_fovx, float tan_fovy, int prefiltered,
    char* geom_buffer, int* radii, int* num_rendered_host, void* stream);

/* stage1 with the SH coefficients given as the reference's two parameter tensors features_dc [P,1,3] and
 * features_rest [P,M-1,3] (gaussian_model.h; M <= 16) instead of their concatenation: the caller skips the
 * torch::cat the reference runs every iteration (src/gaussian_model.cpp:58-62).  Otherwise identical. */
int lgs_forward_stage1_split_sh(
    int P, int D, int M, int W, int H,
    const float* means3D, const float* features_dc, const float* features_rest,
    const float* opacities, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* cam_pos,
    float tan_fovx, float tan_fovy, int prefiltered,
    char* geom_buffer, int* radii, int* num_rendered_host, void* stream);

int lgs_forward_stage2(
    int P, int W, int H, int R,
    const float* background, const float* lang_feat,
    char* geom_buffer, char* binning_buffer, char* image_buffer,
    float* out_color, float* out_lang_feat, float* out_depth,
    int include_lang_feat, void* stream);

/* ---- backward  (Rasterizer::backward, rasterizer_impl.cu:347-453) -----------------
 * dL_dmean2D [P,3], dL_dconic [P,2,2] (slots 0,1,3 used), dL_dopacity [P,1],
 * dL_dcolor [P,3], dL_dlang_feat [P,64], dL_ddepth [P,1] are ACCUMULATED into and must
 * be zero on entry unless zero_outputs != 0, in which case this call zeroes them
 * itself (one fused memset kernel).  dL_dmean3D [P,3], dL_dcov3D [P,6],
 * dL_dsh [P,M,3], dL_dscale [P,3], dL_drot [P,4] are fully overwritten (zeros for
 * culled Gaussians) when zero_outputs != 0; with zero_outputs == 0 only visible
 * Gaussians are written, exactly like the reference.
 * dL_ddepth may be NULL (the reference computes and discards it,
 * rasterize_points.cu:161,207).
 * bwd_scratch: lgs_backward_scratch_bytes(R, W, H) bytes of device memory for the render backward's
 * pixel->channel hand-off, or NULL to take it from the CUDA stream-ordered pool
 * (cudaMallocAsync/cudaFreeAsync on `stream`; no synchronisation).
 */
int lgs_backward(
    int P, int D, int M, int R, int W, int H,
    const float* background,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* lang_feat, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* cam_pos,
    float tan_fovx, float tan_fovy, const int* radii,
    const char* geom_buffer, const char* binning_buffer, const char* image_buffer,
    const float* dL_dpix, const float* dL_dpix_lf, const float* dL_dpix_depth,
    float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
    float* dL_dlang_feat, float* dL_ddepth, float* dL_dmean3D, float* dL_dcov3D,
    float* dL_dsh, float* dL_dscale, float* dL_drot,
    int include_lang_feat, int zero_outputs, char* bwd_scratch, void* stream);
/* backward for the split SH layout of lgs_forward_stage1_split_sh: dL/dSH goes to dL_dfeatures_dc [P,1,3] and
 * dL_dfeatures_rest [P,M-1,3] (what autograd's cat-backward would produce).  accumulate_sh != 0 ADDS the visible
 * Gaussians' rows to those two arrays (several views per iteration) and leaves culled rows untouched. */
int lgs_backward_split_sh(
    int P, int D, int M, int R, int W, int H,
    const float* background,
    const float* means3D, const float* features_dc, const float* features_rest,
    const float* lang_feat, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* cam_pos,
    float tan_fovx, float tan_fovy, const int* radii,
    const char* geom_buffer, const char* binning_buffer, const char* image_buffer,
    const float* dL_dpix, const float* dL_dpix_lf, const float* dL_dpix_depth,
    float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
    float* dL_dlang_feat, float* dL_ddepth, float* dL_dmean3D, float* dL_dcov3D,
    float* dL_dfeatures_dc, float* dL_dfeatures_rest, int accumulate_sh, float* dL_dscale, float* dL_drot,
    int include_lang_feat, int zero_outputs, char* bwd_scratch, void* stream);
/* bytes of bwd_scratch for a backward over R instances of a WxH image (544 B/instance + 8 B/tile) */
size_t lgs_backward_scratch_bytes(int R, int W, int H);

/* ---- markVisible  (rasterizer_impl.cu:54-66,141-153): present[i] = view-z > 0.2 --- */
int lgs_mark_visible(int P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, unsigned char* present, void* stream);

/* ---- introspection of the opaque buffers (parity tests: bit-exact keys / order /
 *      ranges; reference layouts rasterizer_impl.h:33-63).  Returned pointers alias the
 *      caller's buffers. */
typedef struct lgs_binning_view {
    const uint32_t* point_list;      /* [R] sorted Gaussian indices (the reference's point_list)             */
    const uint32_t* keys_sorted32;   /* [R] the production sort keys: tile << q | top depth bits             */
} lgs_binning_view;
typedef struct lgs_reference_keys_view {   /* the reference's BinningState arrays (rasterizer_impl.h:50-63)  */
    const uint64_t* keys_unsorted;   /* [R] tile<<32 | depth bits, the reference's emission order            */
    const uint32_t* values_unsorted; /* [R] Gaussian index                                                   */
    const uint64_t* keys_sorted;     /* [R] read off point_list + ranges as this library computed them       */
    const uint32_t* point_offsets;   /* [P] inclusive scan of tiles_touched (GeometryState::point_offsets)   */
} lgs_reference_keys_view;
typedef struct lgs_image_view {
    const uint32_t* ranges;          /* [tiles][2] (start,end) into point_list          */
    const float*    final_T;         /* [H*W]                                           */
    const uint32_t* n_contrib;       /* [H*W]                                           */
} lgs_image_view;
typedef struct lgs_geom_view {
    const float*    records;         /* [P][12]: x,y,depth,idx bits | conic a,b,c, opacity | r,g,b,0 */
    const float*    cov3D;           /* [P][6]                                          */
    const uint32_t* tiles_touched;   /* [P]                                             */
    const int32_t*  internal_radii;  /* [P]                                             */
    const uint8_t*  clamped;         /* [P] bit c set if colour channel c was clamped   */
} lgs_geom_view;
int lgs_view_binning(const char* binning_buffer, int R, lgs_binning_view* out);
/* Parity-test aid: re-express what a forward computed as the reference's key arrays.  The production path sorts 32-bit
 * keys (tile id | the top depth bits, bounded by the frame's largest depth) emitted in no particular order and repairs the
 * runs of equal keys afterwards; this call writes, into debug_buffer (lgs_debug_keys_bytes(P, R) bytes), the pairs
 * duplicateWithKeys would have emitted in ITS order (Gaussian-major, y, x; rasterizer_impl.cu:70-111) and the sorted 64-bit
 * key array that corresponds to this library's point_list and ranges.  If those equal the reference's arrays bit for bit,
 * the lists do.  Never on the mapping path; no process-wide switch is involved. */
size_t lgs_debug_keys_bytes(int P, int R);
int lgs_debug_reference_keys(int P, int R, int W, int H, const char* geom_buffer, const char* binning_buffer,
                             const char* image_buffer, char* debug_buffer, lgs_reference_keys_view* out, void* stream);
/* The frame header of a geometry buffer, copied asynchronously on `stream` into status_host[0..3] (host memory, pinned for a
 * truly asynchronous copy): [0] = R (instances of the frame), [1] = largest depth bit pattern of a rendered Gaussian,
 * [2] = 1 if R exceeded the capacity stage2 was given, [3] = 1 if the sort gave up (never observed).  Valid once the
 * caller has synchronised `stream`. */
int lgs_forward_status(const char* geom_buffer, int P, unsigned int* status_host, void* stream);
int lgs_view_image(const char* image_buffer, int W, int H, lgs_image_view* out);
int lgs_view_geom(const char* geom_buffer, int P, lgs_geom_view* out);

/* ---- optional per-stage timing (bench.py's roofline) of the CALLING HOST THREAD: when enabled, CUDA events are recorded
 *      on the launch stream between the kernels of this thread's stage1 / stage2 / backward calls.  After the caller
 *      has synchronised, lgs_profile_read fills ms[0..8] = preprocess, emit_keys, sort (3 radix passes), tile_ranges,
 *      render_fwd, zero_grads, render_bwd_pix, render_bwd_chan, preprocess_bwd (milliseconds, -1 if the stage did not
 *      run).  State is thread-local: nothing here is shared between host threads.  Off by default. */
int lgs_profile_enable(int on);
int lgs_profile_read(float* ms, int n);

/* FP32 FMA micro-benchmark for bench.py (issues blocks*256*iters*256 flops of dependent FFMA
 * chains, 8 per thread); not part of the mapping path. */
int lgs_bench_fma(int blocks, int iters, float* sink, void* stream);

/* ---- fused multi-tensor Adam  (torch::optim::Adam, reference
 *      src/gaussian_model.cpp:483-518, step at src/gaussian_mapper.cpp:793-796) -----
 * One launch over n <= 16 tensors (host arrays of device pointers).  step is the 1-based
 * step count AFTER increment, as torch uses it for bias correction.  lr / betas / eps are
 * double like libtorch's AdamOptions (bias corrections are evaluated in double, then cast).
 *   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g
 *   p -= (lr/(1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 */
int lgs_adam_multi(int n_tensors, float* const* params, const float* const* grads,
                   float* const* exp_avg, float* const* exp_avg_sq,
                   const int64_t* numel, const double* lr,
                   double beta1, double beta2, double eps, int step, void* stream);

/* ---- data-parallel exchange fused with Adam, over NVLink peer memory (new) -------
 * One launch per rank: for flat indices [shard_begin, shard_end) (multiples of 4) it sums the G ranks'
 * gradients (multimem.ld_reduce on grads_mc when non-NULL, else peer loads of grads_peers[0..G) in rank
 * order), applies Adam (state arrays cover the shard only), and writes the updated parameters to every
 * rank (multimem.st on params_mc when non-NULL, else peer stores to params_peers[0..G)).
 * seg_start[0..n_seg] are the flat offsets of the parameter tensors (multiples of 4), lr[t] their rates.
 * grads_peers / params_peers are HOST arrays of device pointers into symmetric (peer-mapped) buffers; the
 * caller synchronises the ranks before (gradients complete) and after (parameters landed) the launch.
 * max_ctas > 0 bounds the grid (a launch meant to run underneath other kernels on a side stream); 0 = whole GPU. */
int lgs_dp_adam_shard(int n_seg, const int64_t* seg_start, const double* lr, int world, int rank,
                      const float* const* grads_peers, float* const* params_peers, const float* grads_mc,
                      float* params_mc, int64_t shard_begin, int64_t shard_end, float* exp_avg_shard,
                      float* exp_avg_sq_shard, double beta1, double beta2, double eps, int step, int max_ctas,
                      void* stream);

/* Sparse variant of the P2P exchange (no multicast): every tensor t is [n_rows, row_len[t]] (one row per Gaussian) and
 * row_mask[g] has bit r set when rank r's gradient row of Gaussian g can be non-zero, i.e. when r rendered g in one of its
 * views -- a culled Gaussian's row is exactly +0.0 there (reference backward.cu / rasterize_points.cu:157-167 zero-fill), 60 % of
 * the rows with one view per rank at cfgB.  A rank's copy of a 16-byte vector is loaded only if a Gaussian it touches has its
 * bit set: same sums, bit for bit, fewer bytes over NVLink.  The masks come from the three helpers below: each rank marks
 * its rendered Gaussians from `radii` (accumulating over its views), publishes the bytes to slot [rank] of every rank's table
 * (symmetric memory, lgs_dp_rows_table_bytes each; peer stores), and -- after the caller's cross-rank barrier -- combines
 * its local table into the bit mask. */
int lgs_dp_adam_shard_sparse(int n_seg, const int64_t* seg_start, const double* lr, const int* row_len, int64_t n_rows,
                             const uint16_t* row_mask, int world, int rank, const float* const* grads_peers,
                             float* const* params_peers, int64_t shard_begin, int64_t shard_end, float* exp_avg_shard,
                             float* exp_avg_sq_shard, double beta1, double beta2, double eps, int step, int max_ctas,
                             void* stream);
int lgs_dp_rows_mark(int P, const int* radii, unsigned char* vis, int accumulate, void* stream);
size_t lgs_dp_rows_table_bytes(int P, int world);
int lgs_dp_rows_publish(int P, int world, int rank, const unsigned char* vis, unsigned char* const* tables_peers, void* stream);
int lgs_dp_rows_combine(int P, int world, const unsigned char* table, unsigned short* row_mask, void* stream);

/* Stream hooks of the CALLING HOST THREAD, for callers that overlap the exchange of the language-feature tensors (64 of
 * the 123 floats per Gaussian) with the stages that never touch them.  Both are cudaEvent_t handles owned by the caller;
 * NULL clears a hook.  While set: lgs_forward_stage2 makes its stream wait for `wait_before_render_fwd` after binning
 * and before the render kernel (the only forward stage that reads lang_feat; reference forward.cu:261-392), and
 * lgs_backward* records `record_after_render_bwd` on its stream after the render backward, when dL_dlang_feat is final
 * and lang_feat is no longer read (preprocess backward follows; reference rasterizer_impl.cu:404-453). */
int lgs_stream_hooks(void* wait_before_render_fwd, void* record_after_render_bwd);

/* ---- activations  (reference src/gaussian_model.cpp:46-68) ---------------
 * forward : scales = exp(scaling) [P,3], rotations = normalize(rotation) [P,4], opacities =
 *           sigmoid(opacity) [P,1], shs = cat(features_dc [P,1,3], features_rest [P,n_rest,3]) [P,1+n_rest,3]
 *           (shs == NULL skips the cat: callers of the *_split_sh entry points do not need it)
 * backward: raw-parameter gradients from the gradients w.r.t. those four tensors (what autograd
 *           computes through exp / F.normalize / sigmoid / cat); accumulate != 0 adds to the outputs;
 *           dL_dshs == NULL skips the cat backward. */
int lgs_activations_fwd(int P, int n_rest, const float* scaling, const float* rotation, const float* opacity,
                        const float* features_dc, const float* features_rest, float* scales, float* rotations,
                        float* opacities, float* shs, void* stream);
int lgs_activations_bwd(int P, int n_rest, int accumulate, const float* rotation, const float* scales,
                        const float* opacities, const float* dL_dscales, const float* dL_drotations,
                        const float* dL_dopacities, const float* dL_dshs, float* dL_dscaling, float* dL_drotation,
                        float* dL_dopacity, float* dL_dfeatures_dc, float* dL_dfeatures_rest, void* stream);

/* ---- mapping loss + its gradient w.r.t. the rendered images, fused (reference
 *      src/gaussian_mapper.cpp:707-724, include/loss_utils.h:27-131) --------
 * loss = (1-l)*L1(image*mask, gt_image) + l*(1 - SSIM(image*mask, gt_image))
 *        + (cos_sign >= 0 ? +mean_px cos(lf*mask0, up(gt_lf)) : 1 - mean_px cos(...))
 *        + L1(depth*mask0, gt_depth)
 * image [3,H,W], lf [64,H,W], depth [1,H,W]; gt_lf [64,lf_h,lf_w] is up-sampled nearest on the fly;
 * mask [3,H,W] or NULL.  Writes dL/dimage, dL/dlf, dL/ddepth (fully overwritten) and loss_out[0..4] =
 * total, L1, SSIM, mean cosine, depth L1 (device memory).  The reference ADDS the mean cosine
 * (cos_sign = +1). */
size_t lgs_mapping_loss_scratch_bytes(int W, int H);
int lgs_mapping_loss(int W, int H, int lf_w, int lf_h, const float* image, const float* lf, const float* depth,
                     const float* gt_image, const float* gt_lf, const float* gt_depth, const float* mask,
                     float lambda_dssim, int cos_sign, float* dL_dimage, float* dL_dlf, float* dL_ddepth,
                     float* loss_out, char* scratch, void* stream);

/* ---- adaptive density control with optimizer-state surgery, fused -------------------
 * lgs_densify_stats: GaussianModel::addDensificationStats (src/gaussian_model.cpp:834-847) + the max_radii2D update
 *   (src/gaussian_mapper.cpp:739-742) for the Gaussians with radii > 0: xyz_gradient_accum += |dL_dmeans2D.xy|,
 *   denom += 1, max_radii2D = max(max_radii2D, radii).  All [P] float32 except radii (int32), dL_dmeans2D [P,3].
 * lgs_densify_plan / lgs_densify_apply: GaussianModel::densifyAndPrune (src/gaussian_model.cpp:806-824 = clone :775-804,
 *   split into N = 2 :729-773, prune :597-651, each with the Adam-state surgery of densificationPostfix :653-727) as ONE
 *   classification pass, four scans and ONE gather of all tensors into their final rows.
 *   plan   classifies every Gaussian (scaling [P,3] and opacity [P,1] are the RAW parameters), scans, and returns
 *          totals_host[4] = { originals kept, clones kept, split parents whose 2 children are kept, split parents };
 *          the new point count is totals[0] + totals[1] + 2*totals[2].  Synchronises `stream` once.
 *          plan_scratch: lgs_densify_plan_bytes(P) bytes, handed unchanged to apply.
 *   apply  gathers n_tensors <= 24 row-major float32 (or any 4-byte element) tensors src[t] [P,row_floats[t]] into
 *          dst[t] [new P,row_floats[t]] in the reference's order (kept originals, clones, children copy 1, copy 2).
 *          modes[t]: 0 copy the parent's row; 1 copy for originals, ZERO for appended points (Adam moments);
 *          2 xyz: children get R(q/|q|) * (samples * exp(scaling)) + xyz; 3 scaling: children get log(exp(s)/1.6).
 *          samples [2*totals[3],3]: standard-normal draws, row k*totals[3]+m for copy k of the m-th split parent
 *          (the reference's at::normal(0, stds) consumes the generator identically).  map_scratch: 2*newP uint32.
 *   src/dst/row_floats/modes are HOST arrays; the tensors they point to are device memory. */
int lgs_densify_stats(int P, const int* radii, const float* dL_dmeans2D, float* xyz_gradient_accum, float* denom,
                      float* max_radii2D, void* stream);
size_t lgs_densify_plan_bytes(int P);
int lgs_densify_plan(int P, const float* xyz_gradient_accum, const float* denom, const float* scaling, const float* opacity,
                     float max_grad, float min_opacity, float extent, float percent_dense, int max_screen_size,
                     char* plan_scratch, int* totals_host, void* stream);
int lgs_densify_apply(int P, const char* plan_scratch, const int* totals, int n_tensors, const float* const* src,
                      float* const* dst, const int* row_floats, const int* modes, const float* scaling,
                      const float* rotation, const float* samples, uint32_t* map_scratch, void* stream);

/* ---- keyframe ingest geometry (the step before the path) --------------------------------
 * lgs_reproject_depth_pinhole: reprojectDepthPinhole (src/stereo_vision.cu:40-61,135-162): points[i] =
 *   ((u-cx)*d/fx, (v-cy)*d/fy, d) for pixel i = v*width+u where mask[i] (bool bytes), zeros elsewhere.  points [P,3].
 * lgs_transform_points: transformPoints (src/operate_points.cu:39-94): out = transformPoint4x3(points, matrix), matrix
 *   4x4 column-major as the reference indexes it.  out must not alias points.
 * lgs_knn_mean_dist2: distCUDA2 (third_party/simple-knn/simple_knn.cu:185-220): mean squared distance of every point
 *   to its 3 nearest neighbours (FLT_MAX terms when P < 4, like the reference).  scratch: lgs_knn_scratch_bytes(P). */
int lgs_reproject_depth_pinhole(int P, int width, float fx, float fy, float cx, float cy, const float* depth,
                                const unsigned char* mask, float* points, void* stream);
int lgs_transform_points(int P, const float* points, const float* transformmatrix, float* out, void* stream);
size_t lgs_knn_scratch_bytes(int P);
int lgs_knn_mean_dist2(int P, const float* points, float* mean_dist2, char* scratch, void* stream);

/* ---- neighbours of the path that call into it or feed it (widening per SURVEY.md 8b "who calls it") -----------
 * lgs_scale_transform_mark_visible: scaleAndTransformThenMarkVisiblePoints (src/operate_points.cu:52-70,96-140; a caller of
 *   markVisible, used by the loop-closure correction): rows i with not_transformed_mask[i] && unstable_mask[i] && view-z of
 *   points[i] > 0.2 get  points[i] <- T * (scale * points[i]),  rots[i] <- quaternion of T[:3,:3] * R(rots[i]) (w,x,y,z; not
 *   normalised),  not_transformed_mask[i] <- 0, IN PLACE; *num_transformed (device int) is incremented by the number of such
 *   rows.  faithful_rot_store != 0 stores the rotation row as the reference does, (w, x, z, 0)
 *   (cuda_rasterizer/operate_points.h:169-178 writes z to offset 2 twice, never offset 3); 0 stores (w, x, y, z).
 *   Matrices 4x4 as the reference indexes them (pose tensors stored transposed).  rots 16-byte aligned.
 * lgs_inactive_geo_densify: monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints (src/stereo_vision.cu:63-133,
 *   164-212): N keypoints (kps_pixel [N,2], kps_has3D bool bytes, kps_point_local [N,3]); a keypoint without a 3D point takes the
 *   depth of the nearest keypoint that has one (squared pixel distance <= max_pixel_dist, ties to the lowest index) and is
 *   reprojected with it; rows ending with z > 0 are written in order to out_points / out_colors [<= N,3], their number to
 *   *out_count (device int).  colors: the image buffer the reference indexes at trunc(v * width + u) + {0,1,2}; offsets
 *   outside [0, colors_len) read as zero.  scratch: lgs_inactive_geo_scratch_bytes(N).  kps_pixel 8-byte aligned. */
int lgs_scale_transform_mark_visible(int P, float scale, float* points, float* rots, unsigned char* not_transformed_mask,
                                     const unsigned char* unstable_mask, const float* transformmatrix, const float* viewmatrix,
                                     int faithful_rot_store, int* num_transformed, void* stream);
size_t lgs_inactive_geo_scratch_bytes(int N);
int lgs_inactive_geo_densify(int N, int width, float fx, float fy, float cx, float cy, float max_pixel_dist,
                             const float* kps_pixel, const unsigned char* kps_has3D, const float* kps_point_local,
                             const float* colors, long long colors_len, float* out_points, float* out_colors, int* out_count,
                             char* scratch, void* stream);

/* ---- .ply checkpoint records (reference GaussianModel::savePly / loadPly,
 *      src/gaussian_model.cpp:854-1075) ------------------------------------------------------------------
 * block is the [P][C] float32 vertex block exactly as it lies in a binary_little_endian .ply.  Column c holds element
 * col_elem[c] of the row of tensor col_tensor[c] (device int arrays; -1 = zero column when packing, skipped when
 * unpacking).  tensors / row_floats are HOST arrays of n_tensors <= 24 device tensors [P,row_floats[t]]. */
int lgs_ply_pack(long long P, int C, const int* col_tensor, const int* col_elem, int n_tensors,
                 const float* const* tensors, const int* row_floats, float* block, void* stream);
int lgs_ply_unpack(long long P, int C, const int* col_tensor, const int* col_elem, int n_tensors, float* const* tensors,
                   const int* row_floats, const float* block, void* stream);

/* ---- semantic query (reference eval/find_objects_gaussians.py:160-175) ------------
 * sim[p,q] = <f_p/|f_p|, t_q/|t_q|> for feats [P,64] and text [Q,64] (both row-major,
 * un-normalised; eps 1e-12 like F.normalize).  out is [P,Q] row-major.
 * lgs_minmax_invert turns one similarity column into the reference's score
 *   1 - (s-min)/(max-min)   (find_objects_gaussians.py:173-175), in place.
 */
int lgs_cosine_query(int P, int Q, const float* feats, const float* text, float* out,
                     void* stream);       /* tcgen05 tensor cores, 3xTF32 (fp32-level accuracy), TMEM accumulators */
int lgs_cosine_query_simt(int P, int Q, const float* feats, const float* text, float* out,
                          void* stream);  /* fp32 SIMT cross-check of the same contraction */
int lgs_minmax_invert(int64_t n, float* scores, float* scratch2, void* stream);
/* Per-pixel variant (reference eval/find_objects_gaussians.py:323: F.cosine_similarity(rendered_lf, text[:,None,None], dim=0)):
 * image is a rendered planar feature image [64,H,W] (HW = H*W), text [Q,64]; out [Q,H,W] with
 * out[q][px] = <image[:,px], text_q> / (max(|image[:,px]|, 1e-8) * max(|text_q|, 1e-8))  (torch's cosine_similarity eps). */
int lgs_cosine_image(int64_t HW, int Q, const float* image, const float* text, float* out, void* stream);
/* Heat colours for "query, then heat-map render" (BASELINE.json configs[4]): colors[p] = blue -> red ramp of
 * scores[p * stride] clamped to [0,1] (stride = Q selects one column of a [P,Q] score matrix); feed them to the forward as
 * colors_precomp with include_lang_feat = 0, like the reference's recolouring of the queried Gaussians (:186-189). */
int lgs_heat_colors(int P, const float* scores, int stride, float* colors, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LGS_H_INCLUDED */
