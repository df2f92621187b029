// render_fwd_tc.cu -- per-tile front-to-back alpha compositing of RGB + depth + 64-D language feature with the
// 68-channel accumulation on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.  Same inputs and outputs
// as render_fwd_kernel (render_fwd.cu); replaces FORWARD::render / renderCUDA<3,64> (reference forward.cu:261-392).
//
// The blend of one tile is   out[px][ch] = sum_j w[px][j] * v[j][ch],   w = alpha * T   (forward.cu:360-369),
// a [64 px x n] by [n x 68] product whose left factor is produced by a strictly sequential per-pixel chain
// (alpha test, T update, early termination).  The SIMT kernel spends 68 FFMA + 17 LDS.128 per (warp, Gaussian)
// on the product; here the chain stays on the CUDA cores, bit-identical to the reference (same float sequence,
// so final_T, n_contrib and the termination decisions are unchanged), and the product runs as tcgen05.mma
// kind::tf32 with both operands split into TF32 hi + lo parts so the result keeps fp32-level accuracy:
//
//   D[128 x 64] += A[128 x 32] * B^T[32 x 64]  per batch of 32 Gaussians and per B in {V_hi, V_lo}
//     A rows 0-63 = w_hi of pixel r, rows 64-127 = w_lo of pixel r - 64   (K = Gaussian within the batch)
//     B row c = feature channel c                                         (K-major: one row per channel)
//   out[px][ch] = D[px][ch] + D[64 + px][ch]  =  (w_hi + w_lo) * (v_hi + v_lo), summed in fp32 in TMEM.
// The 3 colour channels and depth stay on the CUDA cores (4 FFMA per fragment, exact fp32): with N = 64 the
// accumulator takes 64 TMEM columns, so 6 CTAs fit an SM instead of 4.
//
// One CTA (4 warps) per tile.  Warps 0-1 own one pixel per thread and run the alpha/T chain, writing w as hi/lo
// rows of A.  Warps 2-3 own one feature channel per thread: they gather the batch's feature values straight from
// global memory (one coalesced 128-byte LDG per warp and Gaussian, issued one batch ahead into registers) and
// write the K-major B tiles with one conflict-free STS.128 per 4 Gaussians; they also stage the next batch's
// 48-byte records for the pixel warps and issue the batch's 8 MMAs.  (A TMA bulk copy per Gaussian, as in the
// SIMT kernels, costs ~7 issue slots per copy on the uniform datapath -- 9 % of all instructions in the first
// version of this kernel -- and the rows would still have to be re-read from shared memory for the transposition.)
// The accumulators never leave TMEM until the tile is finished; the epilogue adds the hi and lo row halves through
// shared memory and writes the planar images.
//
// Operand split: kind::tf32 ignores the low 13 mantissa bits, so hi = x & 0xffffe000, lo = x - hi (tc.cuh).
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"
#include "tc.cuh"

namespace lgs {

constexpr int TF_B = 32;                       // Gaussians per batch (= K of one accumulation round)
constexpr int TF_N = LF;                       // B rows: the 64 feature channels
constexpr int TF_THREADS = 128;
constexpr int TF_REC = TF_B * 48;              // 1536: one batch of render records
constexpr int TF_LBO_A = 128 * 16;             // bytes between consecutive 16-byte K-chunks of A
constexpr int TF_A_BYTES = (TF_B / 4) * TF_LBO_A;   // 16384
constexpr int TF_LBO_B = TF_N * 16;            // 1024
constexpr int TF_B_BYTES = (TF_B / 4) * TF_LBO_B;   // 8192
constexpr int TF_SM_A = 0;
constexpr int TF_SM_BHI = TF_SM_A + TF_A_BYTES;
constexpr int TF_SM_BLO = TF_SM_BHI + TF_B_BYTES;
constexpr int TF_SM_REC = TF_SM_BLO + TF_B_BYTES;    // [2][32] records
constexpr int TF_SM_IDS = TF_SM_REC + 2 * TF_REC;    // [2 producer warps][32] Gaussian ids of the batch being gathered
constexpr int TF_SM_TOTAL = TF_SM_IDS + 2 * 128;     // 36096
constexpr int TF_TMEM_COLS = 64;
constexpr int TF_CTAS = 6;                     // resident CTAs per SM (shared memory and TMEM allow 6, registers <= 85)

__global__ void __launch_bounds__(TF_THREADS, TF_CTAS)
render_fwd_tc_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
                     const GaussRec* __restrict__ rec, const float* __restrict__ lang_feat,
                     const float* __restrict__ bg, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
                     uint32_t* __restrict__ tile_last, float* __restrict__ out_color, float* __restrict__ out_lf,
                     float* __restrict__ out_depth) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t mma_done;
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t s_tile_last[2];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile_id = blockIdx.y * gridDim.x + blockIdx.x;
    const int px = tid & 63;  // warps 2-3 finish the epilogue for pixel tid - 64
    const uint32_t pxi = blockIdx.x * TILE + (px & 7);
    const uint32_t pyi = blockIdx.y * TILE + (px >> 3);
    const bool inside = pxi < (uint32_t)W && pyi < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * pyi + pxi;
    const size_t HW = (size_t)H * W;

    const uint2 range = ranges[tile_id];
    const int n = (int)(range.y - range.x);
    if (n <= 0) {  // empty tile: background only (forward.cu:377-391 with T = 1)
        if (inside) {
            if (warp < 2) {
                final_T[pix_id] = 1.0f;
                n_contrib[pix_id] = 0;
                out_color[0 * HW + pix_id] = bg[0];
                out_color[1 * HW + pix_id] = bg[1];
                out_color[2 * HW + pix_id] = bg[2];
                out_depth[pix_id] = 0.f;
#pragma unroll 4
                for (int k = 0; k < LF / 2; ++k) out_lf[(size_t)k * HW + pix_id] = 0.f;
            } else {
#pragma unroll 4
                for (int k = LF / 2; k < LF; ++k) out_lf[(size_t)k * HW + pix_id] = 0.f;
            }
        }
        if (tid == 0) tile_last[tile_id] = 0;
        return;
    }
    const int nb = (n + TF_B - 1) / TF_B;
    const uint32_t sb = smem_u32(smem);

    if (tid == 0) {
        mbar_init(&mma_done, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, TF_TMEM_COLS);
    // ---- gather state of the producer warps: thread t owns feature channel t.  v[4*i + u] = channel value of the
    // (4*i + u)-th Gaussian of the NEXT batch to be converted; list positions past the end repeat the last Gaussian
    // (their w is 0).
    const int t = tid - 64;
    float v[TF_B];
    uint32_t id_next = 0;  // lane j: id of Gaussian j of the batch after the one in v[]
    const uint32_t ids_sm = sb + TF_SM_IDS + (warp & 1) * 128;
    auto load_ids = [&](int b) {  // lane j <- id of the j-th Gaussian of batch b
        id_next = point_list[range.x + min(b * TF_B + lane, n - 1)];
    };
    auto gather = [&]() {  // ids of the batch in id_next -> v[] (features), and its records -> record buffer `slot`
        __syncwarp();
        sts32(ids_sm + lane * 4, id_next);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < TF_B / 4; ++i) {
            const uint4 id4 = lds128u(ids_sm + i * 16);
            v[4 * i + 0] = __ldg(lang_feat + (size_t)id4.x * LF + t);
            v[4 * i + 1] = __ldg(lang_feat + (size_t)id4.y * LF + t);
            v[4 * i + 2] = __ldg(lang_feat + (size_t)id4.z * LF + t);
            v[4 * i + 3] = __ldg(lang_feat + (size_t)id4.w * LF + t);
        }
    };
    auto stage_records = [&](int slot) {  // 32 records = 96 float4, by the 64 producer threads, of the batch in id_next
        const float4* r4 = reinterpret_cast<const float4*>(rec);
        const int j0 = t / 3, q0 = t - 3 * j0;            // float4 #t
        const int k1 = 64 + t, j1 = k1 / 3, q1 = k1 - 3 * j1;  // float4 #(64 + t), t < 32
        const uint32_t ida = __shfl_sync(0xffffffffu, id_next, j0 & 31);
        const uint32_t idb = __shfl_sync(0xffffffffu, id_next, j1 & 31);
        // every producer warp holds the batch's ids in its lanes (lane j = Gaussian j)
        constexpr int NF4 = TF_B * 3;
        if (t < NF4) sts128(sb + TF_SM_REC + slot * TF_REC + t * 16, __ldg(r4 + (size_t)ida * 3 + q0));
        if (64 + t < NF4) sts128(sb + TF_SM_REC + slot * TF_REC + (64 + t) * 16, __ldg(r4 + (size_t)idb * 3 + q1));
    };
    if (warp >= 2) {
        load_ids(0);
        stage_records(0);
        gather();
        if (nb > 1) load_ids(1);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = make_idesc_tf32(128, TF_N);

    // ---- per-pixel state (warps 0-1)
    const float pxf = (float)pxi, pyf = (float)pyi;
    const float wx0 = (float)(blockIdx.x * TILE), wy0 = (float)(blockIdx.y * TILE + 4 * (warp & 1));  // this warp's 8x4 pixels
    bool done = !inside;
    float T = 1.0f;
    uint32_t last_contributor = 0;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f, Dacc = 0.f;

    int b_last = nb - 1;
    for (int b = 0; b < nb; ++b) {
        const int cnt = min(TF_B, n - b * TF_B);
        const uint32_t recs = sb + TF_SM_REC + (b & 1) * TF_REC;

        uint32_t vis = 0;
        if (warp < 2) {
            // lane j tests Gaussian j's alpha >= 1/255 ellipse against this warp's 8x4 pixels (common.cuh)
            bool touch = false;
            if (lane < cnt) {
                const float4 t0 = lds128(recs + lane * 48), t1 = lds128(recs + lane * 48 + 16);
                touch = footprint_touches(t0.x, t0.y, t1.x, t1.y, t1.z, t1.w, wx0, wx0 + 7.0f, wy0, wy0 + 3.0f);
            }
            vis = __ballot_sync(0xffffffffu, touch);
        }
        if (b > 0) {  // the previous batch's MMAs have read A and B: the tiles may be overwritten
            mbar_wait(&mma_done, (uint32_t)((b - 1) & 1));
            tc_fence_after();
        }

        if (warp < 2) {
            // ---- alpha / T chain, 4 Gaussians (one 16-byte K-chunk of A) at a time
            const uint32_t arow = sb + TF_SM_A + px * 16;
#pragma unroll 1
            for (int c = 0; c < TF_B / 4; ++c) {
                const uint32_t m = (vis >> (4 * c)) & 15u;
                float wv[4] = {0.f, 0.f, 0.f, 0.f};
                // one warp-uniform block per Gaussian that touches this warp's pixels: alpha test, T update, colour and
                // depth on the CUDA cores (forward.cu:337-369; same comparisons, same float ops)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (m & (1u << u)) {
                        const float4 q0 = lds128(recs + (4 * c + u) * 48);       // x, y, depth, id
                        const float4 q1 = lds128(recs + (4 * c + u) * 48 + 16);  // conic a,b,c, opacity
                        const float4 q2 = lds128(recs + (4 * c + u) * 48 + 32);  // r, g, b
                        float dx, dy;
                        const float power = eval_power(q0.x, q0.y, pxf, pyf, q1.x, q1.y, q1.z, dx, dy);
                        const float alpha = fminf(0.99f, __fmul_rn(q1.w, expf(power)));
                        const bool ok = !(power > 0.0f) && !(alpha < 1.0f / 255.0f) && !done;
                        const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                        const bool term = ok && test_T < 0.0001f;
                        done = done || term;
                        const bool act = ok && !term;
                        const float w = act ? __fmul_rn(alpha, T) : 0.0f;
                        wv[u] = w;
                        C0 = fmaf(w, q2.x, C0);
                        C1 = fmaf(w, q2.y, C1);
                        C2 = fmaf(w, q2.z, C2);
                        Dacc = fmaf(w, q0.z, Dacc);
                        T = act ? test_T : T;
                        last_contributor = act ? (uint32_t)(b * TF_B + 4 * c + u + 1) : last_contributor;  // 1-based list position
                    }
                }
                const float4 w4 = make_float4(wv[0], wv[1], wv[2], wv[3]);
                float4 h, l;
                split_trunc4(w4, h, l);
                sts128(arow + c * TF_LBO_A, h);
                sts128(arow + c * TF_LBO_A + 64 * 16, l);
            }
        } else {
            // ---- B tiles from the gathered registers: row t, one K-chunk per STS.128
            const int r = t;
            const uint32_t brow = sb + TF_SM_BHI + (r >> 3) * 128 + (r & 7) * 16;
#pragma unroll
            for (int i = 0; i < TF_B / 4; ++i) {
                float4 h, l;
                split_trunc4(make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]), h, l);
                sts128(brow + i * TF_LBO_B, h);
                sts128(brow + i * TF_LBO_B + TF_B_BYTES, l);
            }
            if (b + 1 < nb) {  // next batch: records into the other slot (batch b-1's, no longer read), features into v[]
                stage_records((b + 1) & 1);
                gather();
                if (b + 2 < nb) load_ids(b + 2);
            }
        }
        fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
        tc_fence_before();
        // operands complete; also the tile-wide termination vote (forward.cu:315 __syncthreads_count)
        const int all_done = __syncthreads_and(warp >= 2 || done);
        if (tid == 64) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < TF_B / 8; ++ks) {  // K = 8 Gaussians per instruction = two 16-byte chunks
                const uint64_t dA = make_desc(sb + TF_SM_A + ks * 2 * TF_LBO_A, TF_LBO_A, 128);
                const uint64_t dBh = make_desc(sb + TF_SM_BHI + ks * 2 * TF_LBO_B, TF_LBO_B, 128);
                const uint64_t dBl = make_desc(sb + TF_SM_BLO + ks * 2 * TF_LBO_B, TF_LBO_B, 128);
                umma_tf32(tmem, dA, dBh, idesc, (b > 0 || ks > 0) ? 1u : 0u);
                umma_tf32(tmem, dA, dBl, idesc, 1u);
            }
            umma_commit(&mma_done);
        }
        if (all_done) {
            b_last = b;
            break;
        }
    }
    mbar_wait(&mma_done, (uint32_t)(b_last & 1));
    tc_fence_after();

    // ---- epilogue (forward.cu:377-391).  TMEM lane = A row: warps 0-1 hold the w_hi sums of pixel px, warps 2-3 the
    // w_lo sums.  Each side keeps 32 of the 64 columns and hands the other 32 over through shared memory.
    static_assert(TF_A_BYTES + 2 * TF_B_BYTES >= 64 * 64 * 4, "the exchange buffer overlays the operand tiles");
    float* xch = reinterpret_cast<float*>(smem + TF_SM_A);  // [64][64] floats over the A and B tiles (all MMAs have completed)
    const uint32_t tb = tmem + ((uint32_t)(warp * 32) << 16);
    const int keep0 = warp < 2 ? 0 : 32;  // first column this thread finishes
    {
        uint32_t give[32];
        tmem_ld32(tb + (32 - keep0), give);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) xch[(32 - keep0 + k) * 64 + px] = __uint_as_float(give[k]);
    }
    if (warp < 2) {
        const uint32_t wmax = __reduce_max_sync(0xffffffffu, last_contributor);
        if (lane == 0) s_tile_last[warp] = wmax;
    }
    uint32_t keep[32];
    tmem_ld32(tb + keep0, keep);
    tmem_ld_wait();
    tc_fence_before();
    __syncthreads();
    if (inside) {
        if (warp < 2) {
            final_T[pix_id] = T;
            n_contrib[pix_id] = last_contributor;
            out_color[0 * HW + pix_id] = fmaf(T, bg[0], C0);
            out_color[1 * HW + pix_id] = fmaf(T, bg[1], C1);
            out_color[2 * HW + pix_id] = fmaf(T, bg[2], C2);
            out_depth[pix_id] = Dacc;
        }
#pragma unroll
        for (int k = 0; k < 32; ++k)
            out_lf[(size_t)(keep0 + k) * HW + pix_id] = __uint_as_float(keep[k]) + xch[(keep0 + k) * 64 + px];
    }
    if (tid == 0) tile_last[tile_id] = max(s_tile_last[0], s_tile_last[1]);
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, TF_TMEM_COLS);
    }
}

int launch_render_fwd_tc(int W, int H, const GeomState& g, const BinningState& b, ImageState& im,
                         const float* background, const float* lang_feat, float* out_color, float* out_lang_feat,
                         float* out_depth, cudaStream_t s) {
    // function attributes are per device: set them on every launch (a cheap driver call) instead of caching "configured"
    // in a process-wide flag, which would be wrong for the second GPU of a process
    LGS_CUDA_TRY(cudaFuncSetAttribute(render_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SM_TOTAL));
    LGS_CUDA_TRY(cudaFuncSetAttribute(render_fwd_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    const dim3 grid((W + TILE - 1) / TILE, (H + TILE - 1) / TILE, 1);
    render_fwd_tc_kernel<<<grid, TF_THREADS, TF_SM_TOTAL, s>>>(im.ranges, b.point_list, W, H, g.rec, lang_feat, background, im.final_T,
                                                               im.n_contrib, im.tile_last, out_color, out_lang_feat, out_depth);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

}  // namespace lgs
