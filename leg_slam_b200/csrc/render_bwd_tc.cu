// render_bwd_tc.cu -- the per-Gaussian reduction half of the render backward on the 5th-generation
// tensor cores (tcgen05 + TMEM), sm_100a.  Same inputs, same outputs as render_bwd_chan_kernel
// (render_bwd.cu), i.e. the reference's 74 atomicAdds per fragment per pixel (backward.cu:557-609).
//
// Per (tile, pixel-warp half) the pixel kernel leaves a stream of 272-byte half-records
// {gx-cx, gy-cy, 0, id | w[32] | t[32]}.  The per-Gaussian gradients are
//     dL/dfeature[id][ch] += sum_px w[rec][px] * g[px][ch]                       (64 channels)
//     dL/d{colour, depth}  += sum_px w[rec][px] * g_{r,g,b,d}[px]                 (4 columns)
//     raw moments          += sum_px t[rec][px] * {1, u, v, u^2, uv, v^2}[px]     (6 columns)
// a [records x 32] by [32 x 74] product per stream.  The SIMT kernel spends 64 FFMA + 16 LDS per record and
// column (301 M warp instructions at cfgB); here the products run as tcgen05.mma kind::tf32 with the fp32
// operands split into hi + lo parts (hi = the 19 bits kind::tf32 reads, lo = the exact remainder) so the result
// keeps fp32-level accuracy:
//
//   main  D[128 x 64]  = A_main[128 x 32] * W^T[32 x 64]      rows 0-63 = g_hi[ch], rows 64-127 = g_lo[ch];
//                                                             two MMAs per k-step (W_hi, W_lo) give all four
//                                                             hi/lo products; TMEM lane = channel, so one
//                                                             warp's red.global.add covers 128 contiguous bytes
//   aux   D[64 x 8]    = W[64 x 32] * B_aux^T[32 x 8]         columns = {r,g,b,d}_hi, {r,g,b,d}_lo
//   mom   D[64 x 8]    = T[64 x 32] * B_mom^T[32 x 8]         columns = 1,u,v,u^2,uv,v^2 (exact in TF32)
//
// Persistent CTAs (128 threads, 3 per SM) take (tile, pixel-warp half) work items from a counter.  Per batch of 64
// records: one TMA bulk copy brings the raw records (re-issued for the next batch as soon as this one has been
// split, so it flies during the MMAs and the epilogue), all threads split w/t into hi/lo tiles in the canonical
// K-major layout, one thread issues 24 MMAs, and the four warps drain TMEM: hi-row and lo-row warps swap halves of
// their record columns through shared memory (each sum is ONE red per record and channel, 32 reds per thread),
// 16 lanes per warp finish colour / depth / moments.
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"
#include "tc.cuh"

namespace lgs {

constexpr int TB = 64;             // half-records per batch
constexpr int TC_THREADS = 128;
constexpr int HREC_BYTES = HREC_FLOATS * 4;
constexpr int RAW_BYTES = TB * HREC_BYTES;         // 17408; also the relay [64 ch][68 floats]
constexpr int WT_TILE = TB * 32 * 4;               // 8192: one [64 rec x 32 px] TF32 tile
constexpr int AMAIN_HALF = 128 * 32 * 4;           // 16384
constexpr int SM_RAW = 0;
constexpr int SM_WT = SM_RAW + RAW_BYTES;          // W_hi, W_lo, T_hi, T_lo
constexpr int SM_AMAIN = SM_WT + 4 * WT_TILE;      // [128 x 32]: the work item's half
constexpr int SM_BAUX = SM_AMAIN + AMAIN_HALF;     // [8 x 32]
constexpr int SM_BMOM = SM_BAUX + 1024;            // [2 halves][8 x 32]
constexpr int SM_HDR = SM_BMOM + 2 * 1024;         // [64] float4 record headers of the current batch
constexpr int SM_TOTAL = SM_HDR + TB * 16;         // 70656: three CTAs per SM
constexpr int TC_CTAS = 3;
constexpr int TMEM_COLS = 128;                     // main 0-63, aux 64-71, mom 72-79
constexpr int RELAY_PITCH = 36;                    // floats per thread in the epilogue relay (16-byte aligned, spreads banks)
static_assert(TC_THREADS * RELAY_PITCH * 4 <= 4 * WT_TILE, "relay must fit the operand tiles");

// hi = x with the 13 low mantissa bits cleared (what kind::tf32 reads anyway), lo = x - hi (tc.cuh split_trunc4)
__device__ __forceinline__ void split_store(const float4 x, uint8_t* hi, uint8_t* lo) {
    float4 h, l;
    split_trunc4(x, h, l);
    *reinterpret_cast<float4*>(hi) = h;
    *reinterpret_cast<float4*>(lo) = l;
}

// 8 consecutive pixels of one image row (zeros outside the image)
__device__ __forceinline__ void load_row8(const float* __restrict__ plane, int W, int H, uint32_t x0, uint32_t y, float4& a, float4& b) {
    a = make_float4(0.f, 0.f, 0.f, 0.f);
    b = a;
    if (y >= (uint32_t)H) return;
    const float* p = plane + (size_t)y * W + x0;
    if (x0 + 8 <= (uint32_t)W && ((W & 3) == 0)) {
        a = __ldg(reinterpret_cast<const float4*>(p));
        b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    } else {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (x0 + i < (uint32_t)W) ? __ldg(p + i) : 0.f;
        a = make_float4(v[0], v[1], v[2], v[3]);
        b = make_float4(v[4], v[5], v[6], v[7]);
    }
}

__global__ void __launch_bounds__(TC_THREADS, TC_CTAS)
render_bwd_chan_tc_kernel(const uint2* __restrict__ ranges, int W, int H, int tiles_x, int n_tiles,
                          const float* __restrict__ dL_dpix, const float* __restrict__ dL_dpix_lf,
                          const float* __restrict__ dL_dpix_depth, const float* __restrict__ hrec_buf,
                          const uint32_t* __restrict__ hrec_count, uint32_t* __restrict__ work_counter,
                          float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconic, float* __restrict__ dL_dopacity,
                          float* __restrict__ dL_dcolor, float* __restrict__ dL_dlang_feat, float* __restrict__ dL_ddepth) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t raw_full;
    __shared__ __align__(8) uint64_t mma_done;
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_tile;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t HW = (size_t)H * W;

    if (tid == 0) {
        mbar_init(&raw_full, 1);
        mbar_init(&mma_done, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, TMEM_COLS);
    // moment operand: rows 1, u, v, u^2, uv, v^2, 0, 0 over the half's 32 pixels (u, v = pixel - tile centre)
    for (int idx = tid; idx < 2 * 8 * 32; idx += TC_THREADS) {
        const int h = idx >> 8, n = (idx >> 5) & 7, k = idx & 31;
        const float u = (float)(k & 7) - 3.5f, v = (float)(4 * h + (k >> 3)) - 3.5f;
        const float x = n == 0 ? 1.f : n == 1 ? u : n == 2 ? v : n == 3 ? u * u : n == 4 ? u * v : n == 5 ? v * v : 0.f;
        *reinterpret_cast<float*>(smem + SM_BMOM + h * 1024 + canon_off(n, k >> 2, 8) + (k & 3) * 4) = x;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc_main = make_idesc_tf32(128, TB);
    const uint32_t idesc_aux = make_idesc_tf32(64, 8);
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t dA_base = make_desc(smem_base + SM_AMAIN, 128 * 16, 128);
    const uint64_t dWh_base = make_desc(smem_base + SM_WT, TB * 16, 128), dWl_base = make_desc(smem_base + SM_WT + WT_TILE, TB * 16, 128);
    const uint64_t dTh_base = make_desc(smem_base + SM_WT + 2 * WT_TILE, TB * 16, 128);
    const uint64_t dTl_base = make_desc(smem_base + SM_WT + 3 * WT_TILE, TB * 16, 128);
    const uint64_t dBa_base = make_desc(smem_base + SM_BAUX, 128, 128), dBm_base = make_desc(smem_base + SM_BMOM, 128, 128);

    uint32_t raw_phase = 0, mma_phase = 0;

    // A work item's inputs -- its record count and range, and the half's upstream gradient rows -- are fetched ONE ITEM
    // AHEAD into registers, so their global-memory round trips overlap the previous item's batches (they were 20 % of
    // all warp samples when loaded at the start of the item, profiles/r01_chan_tc_v2_ncu.txt).
    struct Item {
        int id, cnt;
        uint2 range;
        float4 a[2][2];  // A_main rows of tasks tid, tid + 128
        float4 b[2];     // B_aux row (threads 0..15)
    };
    auto fetch = [&](int id, Item& it) {
        it.id = id;
        it.cnt = 0;
        if (id >= 2 * n_tiles) return;
        const int tile = id >> 1, half = id & 1;
        it.cnt = (int)hrec_count[id];
        it.range = ranges[tile];
        const uint32_t tx0 = (uint32_t)(tile % tiles_x) * TILE, ty0 = (uint32_t)(tile / tiles_x) * TILE + 4 * half;
#pragma unroll
        for (int i = 0; i < 2; ++i) {  // A_main: task = (row y, channel ch); consecutive lanes = consecutive channels
            const int task = tid + TC_THREADS * i;
            load_row8(dL_dpix_lf + (size_t)(task & 63) * HW, W, H, tx0, ty0 + (task >> 6), it.a[i][0], it.a[i][1]);
        }
        if (tid < 16) {  // B_aux: rows {r,g,b,d}
            const int c = tid & 3;
            load_row8(c < 3 ? dL_dpix + (size_t)c * HW : dL_dpix_depth, W, H, tx0, ty0 + (tid >> 2), it.b[0], it.b[1]);
        }
    };
    if (tid == 0) s_tile = (int)atomicAdd(work_counter, 1u);
    __syncthreads();
    Item cur, nxt;
    fetch(s_tile, cur);

    for (;;) {
        if (cur.id >= 2 * n_tiles) break;
        __syncthreads();  // everyone has read s_tile (and finished the previous item)
        if (tid == 0) s_tile = (int)atomicAdd(work_counter, 1u);
        const int item = cur.id, tile = item >> 1, half = item & 1;
        const int cnt_all = cur.cnt;
        const uint2 range = cur.range;
        const int n_all = (int)(range.y - range.x);
        const int nbt = (cnt_all + TB - 1) / TB;
        const float* stream = hrec_buf + ((size_t)2 * range.x + (size_t)half * n_all) * HREC_FLOATS;
        (void)tile;

        auto issue = [&](int q) {  // thread 0: bulk copy of the item's q-th batch into the raw buffer
            const int cnt = min(TB, cnt_all - q * TB);
            const uint32_t bytes = (uint32_t)cnt * HREC_BYTES;
            mbar_arrive_expect_tx(&raw_full, bytes);
            tma_bulk_g2s(smem + SM_RAW, stream + (size_t)q * TB * HREC_FLOATS, bytes, &raw_full);
        };
        if (tid == 0 && cnt_all > 0) issue(0);

        // ---- the half's upstream gradients as MMA operands (all earlier MMAs have completed: mma_done was waited on)
        if (cnt_all > 0) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int task = tid + TC_THREADS * i;
                const int y = task >> 6, ch = task & 63;
                uint8_t* base = smem + SM_AMAIN;
                const int kc = y * 2;
                split_store(cur.a[i][0], base + canon_off(ch, kc, 128), base + canon_off(64 + ch, kc, 128));
                split_store(cur.a[i][1], base + canon_off(ch, kc + 1, 128), base + canon_off(64 + ch, kc + 1, 128));
            }
            if (tid < 16) {  // B_aux: rows {r,g,b,d}_hi, {r,g,b,d}_lo
                const int c = tid & 3, y = tid >> 2;
                uint8_t* base = smem + SM_BAUX;
                const int kc = y * 2;
                split_store(cur.b[0], base + canon_off(c, kc, 8), base + canon_off(4 + c, kc, 8));
                split_store(cur.b[1], base + canon_off(c, kc + 1, 8), base + canon_off(4 + c, kc + 1, 8));
            }
        }
        __syncthreads();  // s_tile (the next item) is visible
        fetch(s_tile, nxt);  // in flight during this item's batches

        for (int q = 0; q < nbt; ++q) {
            const int cnt = min(TB, cnt_all - q * TB);
            uint8_t* raw = smem + SM_RAW;
            mbar_wait(&raw_full, raw_phase);
            raw_phase ^= 1u;

            // ---- split w / t into TF32 hi / lo tiles, canonical K-major [64 rec][32 px]
            {
                const int r = tid & 63, cb = tid >> 6;
                if (tid < TB) *reinterpret_cast<float4*>(smem + SM_HDR + tid * 16) = *reinterpret_cast<const float4*>(raw + tid * HREC_BYTES);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c16 = cb + 2 * i;  // 0-7: w chunks, 8-15: t chunks
                    const float4 x = *reinterpret_cast<const float4*>(raw + r * HREC_BYTES + 16 + c16 * 16);
                    uint8_t* t = smem + SM_WT + (c16 >> 3) * 2 * WT_TILE + (c16 & 7) * (TB * 16) + r * 16;
                    split_store(x, t, t + WT_TILE);
                }
            }
            fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
            __syncthreads();

            if (tid == 0) {
                if (q + 1 < nbt) issue(q + 1);  // every thread has read the raw records: refill during the MMAs + epilogue
                tc_fence_after();
                // descriptors differ between k-steps (and halves) only in the 14-bit start-address field (16-byte units)
                const uint64_t dA0 = dA_base, dBa0 = dBa_base;
                const uint64_t dBm0 = dBm_base + (uint64_t)(half * (1024 >> 4));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {  // K = 8 pixels per instruction = two 16-byte chunks
                    const uint32_t acc = ks > 0 ? 1u : 0u;
                    const uint64_t kA = (uint64_t)(ks * ((2 * 128 * 16) >> 4)), kW = (uint64_t)(ks * ((2 * TB * 16) >> 4));
                    const uint64_t kB = (uint64_t)(ks * ((2 * 128) >> 4));
                    umma_tf32(tmem + 0, dA0 + kA, dWh_base + kW, idesc_main, acc);
                    umma_tf32(tmem + 0, dA0 + kA, dWl_base + kW, idesc_main, 1u);
                    umma_tf32(tmem + 64, dWh_base + kW, dBa0 + kB, idesc_aux, acc);
                    umma_tf32(tmem + 64, dWl_base + kW, dBa0 + kB, idesc_aux, 1u);
                    umma_tf32(tmem + 72, dTh_base + kW, dBm0 + kB, idesc_aux, acc);
                    umma_tf32(tmem + 72, dTl_base + kW, dBm0 + kB, idesc_aux, 1u);
                }
                umma_commit(&mma_done);
            }
            mbar_wait(&mma_done, mma_phase);
            mma_phase ^= 1u;
            tc_fence_after();

            // ---- drain TMEM.  Warp w reads TMEM lanes 32*(w%4) .. +31.
            const uint32_t tb = tmem + ((uint32_t)(warp * 32) << 16);
            uint32_t va[32], vb[32], ax[8], mo[8];
            tmem_ld32(tb + 0, va);
            tmem_ld32(tb + 32, vb);
            tmem_ld8(tb + 64, ax);
            tmem_ld8(tb + 72, mo);
            tmem_ld_wait();
            // Channel ch's sum is (g_hi row, warps 0-1) + (g_lo row, warps 2-3).  Each side hands the other half of its
            // 64 record columns over through shared memory and finishes its own half: warps 0-1 records 0-31, warps
            // 2-3 records 32-63 -- 32 red.global.add per thread, all four warps busy.  The W/T operand tiles are free
            // (the MMAs have completed) and serve as the relay: [128 threads][36 floats].
            float* relay = reinterpret_cast<float*>(smem + SM_WT);
            {
                float* rl = relay + tid * RELAY_PITCH;
                if (warp < 2) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) *reinterpret_cast<uint4*>(rl + 4 * k) = make_uint4(vb[4 * k], vb[4 * k + 1], vb[4 * k + 2], vb[4 * k + 3]);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) *reinterpret_cast<uint4*>(rl + 4 * k) = make_uint4(va[4 * k], va[4 * k + 1], va[4 * k + 2], va[4 * k + 3]);
                }
            }
            // colour / depth / moments: the M = 64 accumulators keep record 16*w + l on lane l < 16 of warp w
            if (lane < 16) {
                const int r = 16 * warp + lane;
                if (r < cnt) {
                    const float4 hd = *reinterpret_cast<const float4*>(smem + SM_HDR + r * 16);
                    const size_t id = (size_t)__float_as_uint(hd.w);
                    const float gx = hd.x, gy = hd.y;
                    red_add_f32(dL_dcolor + id * 3 + 0, __uint_as_float(ax[0]) + __uint_as_float(ax[4]));
                    red_add_f32(dL_dcolor + id * 3 + 1, __uint_as_float(ax[1]) + __uint_as_float(ax[5]));
                    red_add_f32(dL_dcolor + id * 3 + 2, __uint_as_float(ax[2]) + __uint_as_float(ax[6]));
                    if (dL_ddepth != nullptr) red_add_f32(dL_ddepth + id, __uint_as_float(ax[3]) + __uint_as_float(ax[7]));
                    const float S0 = __uint_as_float(mo[0]), Su = __uint_as_float(mo[1]), Sv = __uint_as_float(mo[2]);
                    const float Suu = __uint_as_float(mo[3]), Suv = __uint_as_float(mo[4]), Svv = __uint_as_float(mo[5]);
                    // dx = gx - u, dy = gy - v (Gaussian centre minus pixel, both relative to the tile centre)
                    red_add_f32(dL_dopacity + id, S0);
                    red_add_f32(dL_dmean2D + id * 3 + 0, gx * S0 - Su);
                    red_add_f32(dL_dmean2D + id * 3 + 1, gy * S0 - Sv);
                    red_add_v4_f32(dL_dconic + id * 4, Suu + gx * (gx * S0 - 2.f * Su), Suv + gx * gy * S0 - gy * Su - gx * Sv, 0.f,
                                   Svv + gy * (gy * S0 - 2.f * Sv));  // slot 2 of the [2,2] conic gradient is unused: + 0
                }
            }
            __syncthreads();
            {
                const int ch = (warp & 1) * 32 + lane;
                const float* rl = relay + (tid ^ 64) * RELAY_PITCH;  // the partner thread holds the other half of channel ch
                const int c0 = warp < 2 ? 0 : 32;  // first record column this thread finishes
                float fin[32];                     // channel ch of records c0 .. c0 + 31
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 o = *reinterpret_cast<const float4*>(rl + 4 * k);
                    const float o4[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) fin[4 * k + e] = __uint_as_float(warp < 2 ? va[4 * k + e] : vb[4 * k + e]) + o4[e];
                }
                // Two channels per reduction: neighbouring lanes (channels 2i, 2i + 1) trade halves of their 32 records, so that
                // the even lane finishes records 0-15 and the odd lane records 16-31 of the PAIR with red.global.add.v2 -- a warp
                // instruction still covers two contiguous 128-byte rows, and the L2 takes twice the floats per request
                // (tools/micro/red_bench.cu: 1.46 against 0.75 T float-adds/s for this access pattern).
                const bool odd = lane & 1;
                float* outp = dL_dlang_feat + (ch & ~1);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float send = odd ? fin[c] : fin[16 + c];
                    const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
                    const int rc = c0 + (odd ? 16 + c : c);
                    if (rc < cnt) {
                        const uint32_t id = __float_as_uint(*reinterpret_cast<const float*>(smem + SM_HDR + rc * 16 + 12));
                        red_add_v2_f32(outp + (size_t)id * LF, odd ? recv : fin[c], odd ? fin[16 + c] : recv);
                    }
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();  // TMEM, the operand tiles and this raw buffer are free again
        }
        cur = nxt;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, TMEM_COLS);
    }
}

int launch_render_bwd_chan_tc(int W, int H, const ImageState& im, const float* dL_dpix, const float* dL_dpix_lf,
                              const float* dL_dpix_depth, const float* hrec, const uint32_t* hcount, uint32_t* work_counter,
                              float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                              float* dL_dlang_feat, float* dL_ddepth, cudaStream_t s) {
    // per-device facts, looked up on every launch (function attributes and the SM count belong to the current device, so a
    // process-wide "configured" flag would be wrong for the second GPU of a process)
    int dev = 0, n_sm = 0;
    LGS_CUDA_TRY(cudaGetDevice(&dev));
    LGS_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    LGS_CUDA_TRY(cudaFuncSetAttribute(render_bwd_chan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    const int tiles_x = (W + TILE - 1) / TILE, tiles_y = (H + TILE - 1) / TILE;
    const int n_tiles = tiles_x * tiles_y;
    const int grid = 2 * n_tiles < TC_CTAS * n_sm ? 2 * n_tiles : TC_CTAS * n_sm;
    render_bwd_chan_tc_kernel<<<grid, TC_THREADS, SM_TOTAL, s>>>(im.ranges, W, H, tiles_x, n_tiles, dL_dpix, dL_dpix_lf, dL_dpix_depth,
                                                                 hrec, hcount, work_counter, dL_dmean2D, dL_dconic, dL_dopacity,
                                                                 dL_dcolor, dL_dlang_feat, dL_ddepth);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

}  // namespace lgs
