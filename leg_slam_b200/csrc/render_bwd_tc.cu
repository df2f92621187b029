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
//   main  D[128 x 64]  = A_main[128 x 32] * W^T[32 x 64]      rows 0-63 = g_hi[ch], rows 64-127 = g_lo[ch]; A_main is
//                                                             constant over a work item and lives in TENSOR MEMORY
//                                                             (written once per item with tcgen05.st);
//                                                             two MMAs per k-step (W_hi, W_lo) give all four
//                                                             hi/lo products; TMEM lane = channel, so one
//                                                             warp's red.global.add covers 128 contiguous bytes
//   aux   D[64 x 8]    = W[64 x 32] * B_aux^T[32 x 8]         columns = {r,g,b,d}_hi, {r,g,b,d}_lo
//   mom   D[64 x 8]    = T[64 x 32] * B_mom^T[32 x 8]         columns = 1,u,v,u^2,uv,v^2 (exact in TF32)
//
// Persistent CTAs of 256 threads, 2 per SM, take (tile, pixel-warp half) work items from a counter and run a two-stage
// pipeline over batches of 64 records:
//   front (warps 0-3)  TMA bulk copy of the raw records (two stages, requested one batch ahead -- across work items too),
//                      split of w / t into hi / lo tiles in the canonical K-major layout, 24 MMAs issued by three threads
//                      (main / colour + depth / moments) into one of two TMEM accumulator sets;
//   drain (warps 4-7)  TMEM -> registers, hi-row and lo-row warps swap halves of their record columns through shared memory
//                      (each sum is ONE red per record and channel pair, 16 red.v2 per thread), 16 lanes per warp finish
//                      colour / depth / moments.
// The drain of batch q runs under the copy + split + MMAs of batch q + 1 (mbarriers: raw_full, meta_ready, mma_done,
// slot_free; every wait is bounded, ptx.cuh); the earlier kernel ran the two halves back to back in every CTA (21 % issue-slot
// utilisation).  What bounds it now is the SM's L1TEX data pipe: shared-memory traffic of the split and the relay, the tensor
// core's operand reads and the global reductions all pass through it (profiles/r02_step_ncu.txt).
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"
#include "tc.cuh"

namespace lgs {

constexpr int TB = 64;             // half-records per batch
constexpr int TC_THREADS = 256;
constexpr int TC_FRONT = 128;      // threads of the front half (and of the drain half)
constexpr int HREC_BYTES = HREC_FLOATS * 4;
constexpr int RAW_BYTES = TB * HREC_BYTES;         // 17408
constexpr int WT_TILE = TB * 32 * 4;               // 8192: one [64 rec x 32 px] TF32 tile
constexpr int RELAY_PITCH = 36;                    // floats per thread in the drain's relay (16-byte aligned, spreads banks)
constexpr int SM_RAW = 0;                          // two stages
constexpr int SM_WT = SM_RAW + 2 * RAW_BYTES;      // W_hi, W_lo, T_hi, T_lo
constexpr int SM_BAUX = SM_WT + 4 * WT_TILE;       // [8 x 32]
constexpr int SM_BMOM = SM_BAUX + 1024;            // [2 halves][8 x 32]
constexpr int SM_HDR = SM_BMOM + 2 * 1024;         // [2 slots][64] float4 record headers
constexpr int SM_RELAY = SM_HDR + 2 * TB * 16;     // [128 drain threads][36 floats]
constexpr int SM_TOTAL = SM_RELAY + TC_FRONT * RELAY_PITCH * 4;  // 91136: two CTAs per SM
constexpr int TC_CTAS = 2;
// tensor memory, 256 columns: two accumulator sets (main [128 x 64] + colour/depth [64 x 8] + moments [64 x 8]) and the work
// item's A operand [128 x 32] (rows 0-63 g_hi[ch], rows 64-127 g_lo[ch]; one column per pixel of the half)
constexpr int TM_MAIN = 0;      // + 64 * set
constexpr int TM_AUX = 128;     // + 16 * set
constexpr int TM_MOM = 136;     // + 16 * set
constexpr int TM_A = 192;
constexpr int TMEM_COLS = 256;

__device__ __forceinline__ void bar_named(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

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
    __shared__ __align__(8) uint64_t raw_full[2];    // the bulk copy of a raw stage has landed
    __shared__ __align__(8) uint64_t meta_ready[2];  // headers + record count of a slot are written
    __shared__ __align__(8) uint64_t mma_done[2];    // the MMAs of a slot have completed
    __shared__ __align__(8) uint64_t slot_free[2];   // the drain is done with a slot (TMEM set, headers, count)
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_pop[2];                         // work-item ids popped two items ahead
    __shared__ int s_cnt[2];
    __shared__ __align__(16) uint32_t s_ids[2][TB];  // Gaussian ids of a slot's records

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t HW = (size_t)H * W;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            mbar_init(&raw_full[i], 1);
            mbar_init(&meta_ready[i], 1);
            mbar_init(&mma_done[i], 3);   // one commit per issuing thread (main / colour+depth / moments)
            mbar_init(&slot_free[i], 4);  // one arrival per drain warp
        }
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

    if (warp < 4) {
        // ============================================================ front: copy, split, MMA issue
        const uint32_t idesc_main = make_idesc_tf32(128, TB);
        const uint32_t idesc_aux = make_idesc_tf32(64, 8);
        const uint32_t smem_base = smem_u32(smem);
        const uint64_t dWh_base = make_desc(smem_base + SM_WT, TB * 16, 128), dWl_base = make_desc(smem_base + SM_WT + WT_TILE, TB * 16, 128);
        const uint64_t dTh_base = make_desc(smem_base + SM_WT + 2 * WT_TILE, TB * 16, 128);
        const uint64_t dTl_base = make_desc(smem_base + SM_WT + 3 * WT_TILE, TB * 16, 128);
        const uint64_t dBa_base = make_desc(smem_base + SM_BAUX, 128, 128), dBm_base = make_desc(smem_base + SM_BMOM, 128, 128);

        // A work item's inputs are fetched AHEAD into registers, so that their global-memory round trips overlap earlier items'
        // batches (items are short: 116 records = 2.3 batches on average at cfgB): record count and range two items ahead
        // (the count decides the bulk copy that is requested during the previous item's last batch), the half's upstream
        // gradient rows one item ahead.
        struct Meta {
            int id, cnt;
            uint2 range;
        };
        struct Rows {
            float4 a[4][2];  // this thread's A_main row: channel tid & 63 over the half's 4 x 8 pixels
            float4 b[2];     // B_aux row (threads 0..15)
        };
        auto fetch_meta = [&](int id, Meta& m) {
            m.id = id;
            m.cnt = 0;
            m.range = make_uint2(0u, 0u);
            if (id >= 2 * n_tiles) return;
            m.cnt = (int)hrec_count[id];
            m.range = ranges[id >> 1];
        };
        auto fetch_rows = [&](int id, Rows& it) {
            if (id >= 2 * n_tiles) return;
            const int tile = id >> 1, half = id & 1;
            const uint32_t tx0 = (uint32_t)(tile % tiles_x) * TILE, ty0 = (uint32_t)(tile / tiles_x) * TILE + 4 * half;
#pragma unroll
            for (int y = 0; y < 4; ++y)  // consecutive lanes = consecutive channels (planes); 32 contiguous bytes per lane and row
                load_row8(dL_dpix_lf + (size_t)(tid & 63) * HW, W, H, tx0, ty0 + y, it.a[y][0], it.a[y][1]);
            if (tid < 16) {  // B_aux: rows {r,g,b,d}
                const int c = tid & 3;
                load_row8(c < 3 ? dL_dpix + (size_t)c * HW : dL_dpix_depth, W, H, tx0, ty0 + (tid >> 2), it.b[0], it.b[1]);
            }
        };
        auto stream_of = [&](const Meta& it) {  // the item's half-record stream: the halves of a tile lie back to back
            const size_t n_all = (size_t)(it.range.y - it.range.x);
            return hrec_buf + ((size_t)2 * it.range.x + (size_t)(it.id & 1) * n_all) * HREC_FLOATS;
        };
        auto issue = [&](const float* src, int left, uint32_t stage) {  // thread 0: bulk copy of the next <= 64 records
            const uint32_t bytes = (uint32_t)min(TB, left) * HREC_BYTES;
            mbar_arrive_expect_tx(&raw_full[stage], bytes);
            tma_bulk_g2s(smem + SM_RAW + stage * RAW_BYTES, src, bytes, &raw_full[stage]);
        };
        int popped = 0;  // thread 0: the id popped at the start of an item, published at its end (the atomic's round trip
                         // then overlaps the item's batches instead of holding the other front warps at the barrier)
        if (tid == 0) {
            const int first = (int)atomicAdd(work_counter, 3u);
            s_pop[0] = first;
            s_pop[1] = first + 1;
            popped = first + 2;
        }
        bar_named(1, TC_FRONT);
        Meta cur, nxt, nn;  // items i, i + 1, i + 2
        Rows cur_rows, nxt_rows;
        fetch_meta(s_pop[0], cur);
        fetch_rows(cur.id, cur_rows);
        fetch_meta(s_pop[1], nxt);
        bar_named(1, TC_FRONT);        // both ids have been read: the ring may be rewritten
        if (tid == 0) s_pop[0] = popped;
        uint32_t q = 0;                // batches this CTA has started: batch q uses raw stage / slot q & 1
        bool first_in_flight = false;  // the item's first batch was requested during the previous item's last batch
        int pop_slot = 0;

        for (;;) {
            if (cur.id >= 2 * n_tiles) break;
            if (tid == 0) popped = (int)atomicAdd(work_counter, 1u);  // item i + 3, used from the end of this item on
            const int half = cur.id & 1;
            const int cnt_all = cur.cnt;
            const int nbt = (cnt_all + TB - 1) / TB;
            const float* stream = stream_of(cur);
            // the stage of batch q was last read by the split of batch q - 2: free
            if (tid == 0 && cnt_all > 0 && !first_in_flight) issue(stream, cnt_all, q & 1u);
            bar_named(1, TC_FRONT);  // the id published at the end of the previous item is visible
            fetch_meta(s_pop[pop_slot], nn);  // item i + 2: in flight during this item's and the next item's batches
            fetch_rows(nxt.id, nxt_rows);     // in flight during this item's batches
            pop_slot ^= 1;

            for (int j = 0; j < nbt; ++j, ++q) {
                const uint32_t s = q & 1u, use = q >> 1;
                const int cnt = min(TB, cnt_all - j * TB);
                // ---- request the batch after this one (of this item, or the first of the next): its stage was read by the
                // split of batch q - 1, which every front thread has left (barrier below)
                if (tid == 0) {
                    if (j + 1 < nbt) issue(stream + (size_t)(j + 1) * TB * HREC_FLOATS, cnt_all - (j + 1) * TB, s ^ 1u);
                    else if (nxt.cnt > 0) issue(stream_of(nxt), nxt.cnt, s ^ 1u);
                }
                // ---- the operand tiles are free once the previous batch's MMAs have completed
                if (q > 0) {
                    mbar_wait(&mma_done[s ^ 1u], ((q - 1) >> 1) & 1u);
                    tc_fence_after();
                }
                if (j == 0) {  // the half's upstream gradients as MMA operands
                    // A_main into tensor memory: this thread's row (TMEM lane tid) is channel tid & 63, hi part for rows 0-63,
                    // lo part for rows 64-127; column k = pixel k of the half
                    uint32_t arow[32];
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        float4 h0, l0, h1, l1;
                        split_trunc4(cur_rows.a[y][0], h0, l0);
                        split_trunc4(cur_rows.a[y][1], h1, l1);
                        const float4 p0 = tid < 64 ? h0 : l0, p1 = tid < 64 ? h1 : l1;
                        arow[8 * y + 0] = __float_as_uint(p0.x); arow[8 * y + 1] = __float_as_uint(p0.y);
                        arow[8 * y + 2] = __float_as_uint(p0.z); arow[8 * y + 3] = __float_as_uint(p0.w);
                        arow[8 * y + 4] = __float_as_uint(p1.x); arow[8 * y + 5] = __float_as_uint(p1.y);
                        arow[8 * y + 6] = __float_as_uint(p1.z); arow[8 * y + 7] = __float_as_uint(p1.w);
                    }
                    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + TM_A, arow);
                    tmem_st_wait();
                    tc_fence_before();  // ordered before the MMAs through the barrier below
                    if (tid < 16) {  // B_aux: rows {r,g,b,d}_hi, {r,g,b,d}_lo
                        const int c = tid & 3, y = tid >> 2;
                        uint8_t* base = smem + SM_BAUX;
                        const int kc = y * 2;
                        split_store(cur_rows.b[0], base + canon_off(c, kc, 8), base + canon_off(4 + c, kc, 8));
                        split_store(cur_rows.b[1], base + canon_off(c, kc + 1, 8), base + canon_off(4 + c, kc + 1, 8));
                    }
                }
                // ---- slot s (TMEM set, headers, count) is free once the drain has finished batch q - 2
                if (q >= 2) mbar_wait(&slot_free[s], (use - 1) & 1u);
                mbar_wait(&raw_full[s], use & 1u);

                // ---- split w / t into TF32 hi / lo tiles, canonical K-major [64 rec][32 px]
                {
                    const uint8_t* raw = smem + SM_RAW + s * RAW_BYTES;
                    const int r = tid & 63, cb = tid >> 6;
                    if (tid < TB) *reinterpret_cast<float4*>(smem + SM_HDR + s * (TB * 16) + tid * 16) = *reinterpret_cast<const float4*>(raw + tid * HREC_BYTES);
                    if (tid < TB) s_ids[s][tid] = *reinterpret_cast<const uint32_t*>(raw + tid * HREC_BYTES + 12);
                    if (tid == 0) s_cnt[s] = cnt;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int c16 = cb + 2 * i;  // 0-7: w chunks, 8-15: t chunks
                        const float4 x = *reinterpret_cast<const float4*>(raw + r * HREC_BYTES + 16 + c16 * 16);
                        uint8_t* t = smem + SM_WT + (c16 >> 3) * 2 * WT_TILE + (c16 & 7) * (TB * 16) + r * 16;
                        split_store(x, t, t + WT_TILE);
                    }
                }
                fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
                bar_named(1, TC_FRONT);

                // Three threads (lane 0 of warps 0-2) issue the batch's 24 MMAs -- eight each, one commit each: the issue is a
                // few hundred single-thread instructions (descriptors through uniform registers) on the front's critical path
                if (lane == 0 && warp < 3) {
                    if (warp == 0) mbar_arrive(&meta_ready[s]);  // headers + count of slot s (release; the barrier above made them this thread's)
                    tc_fence_after();
                    // descriptors differ between k-steps (and halves) only in the 14-bit start-address field (16-byte units)
                    const uint64_t dBm0 = dBm_base + (uint64_t)(half * (1024 >> 4));
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {  // K = 8 pixels per instruction = two 16-byte chunks = 8 TMEM columns of A
                        const uint32_t acc = ks > 0 ? 1u : 0u;
                        const uint64_t kW = (uint64_t)(ks * ((2 * TB * 16) >> 4));
                        const uint64_t kB = (uint64_t)(ks * ((2 * 128) >> 4));
                        if (warp == 0) {
                            umma_tf32_ts(tmem + TM_MAIN + 64 * s, tmem + TM_A + 8 * ks, dWh_base + kW, idesc_main, acc);
                            umma_tf32_ts(tmem + TM_MAIN + 64 * s, tmem + TM_A + 8 * ks, dWl_base + kW, idesc_main, 1u);
                        } else if (warp == 1) {
                            umma_tf32(tmem + TM_AUX + 16 * s, dWh_base + kW, dBa_base + kB, idesc_aux, acc);
                            umma_tf32(tmem + TM_AUX + 16 * s, dWl_base + kW, dBa_base + kB, idesc_aux, 1u);
                        } else {
                            umma_tf32(tmem + TM_MOM + 16 * s, dTh_base + kW, dBm0 + kB, idesc_aux, acc);
                            umma_tf32(tmem + TM_MOM + 16 * s, dTl_base + kW, dBm0 + kB, idesc_aux, 1u);
                        }
                    }
                    umma_commit(&mma_done[s]);
                }
            }
            first_in_flight = nbt > 0 && nxt.cnt > 0;
            cur = nxt;
            cur_rows = nxt_rows;
            nxt = nn;
            // s_pop[pop_slot] was last read after the barrier at the start of the PREVIOUS item, which every front thread
            // has left (it passed this item's barrier); the next item's barrier publishes the new id
            if (tid == 0) s_pop[pop_slot] = popped;
        }
        // ---- no more work: tell the drain (a count of -1 in the next slot, once the drain has left it)
        if (tid == 0) {
            const uint32_t s = q & 1u;
            if (q >= 2) mbar_wait(&slot_free[s], ((q >> 1) - 1) & 1u);
            s_cnt[s] = -1;
            mbar_arrive(&meta_ready[s]);
        }
    } else {
        // ============================================================ drain: TMEM -> reductions
        const int dt = tid - TC_FRONT, dw = warp - 4;  // dw = TMEM lane quadrant (warp % 4)
        float* relay = reinterpret_cast<float*>(smem + SM_RELAY);
        for (uint32_t q = 0;; ++q) {
            const uint32_t s = q & 1u, par = (q >> 1) & 1u;
            mbar_wait(&meta_ready[s], par);
            const int cnt = *reinterpret_cast<volatile int*>(&s_cnt[s]);
            if (cnt < 0) break;
            const uint8_t* hdr = smem + SM_HDR + s * (TB * 16);
            mbar_wait(&mma_done[s], par);
            tc_fence_after();

            // Warp dw reads TMEM lanes 32 * dw .. + 31 of accumulator set s.
            const uint32_t tb = tmem + ((uint32_t)(dw * 32) << 16);
            uint32_t va[32], vb[32], ax[8], mo[8];
            tmem_ld32(tb + TM_MAIN + 64 * s, va);
            tmem_ld32(tb + TM_MAIN + 64 * s + 32, vb);
            tmem_ld8(tb + TM_AUX + 16 * s, ax);
            tmem_ld8(tb + TM_MOM + 16 * s, mo);
            tmem_ld_wait();
            // Channel ch's sum is (g_hi row, warps 0-1) + (g_lo row, warps 2-3).  Each side hands the other half of its
            // 64 record columns over through shared memory and finishes its own half: warps 0-1 records 0-31, warps
            // 2-3 records 32-63.  Partners are thread dt and dt ^ 64: the two warps of a pair meet on their own barrier.
            {
                float* rl = relay + dt * RELAY_PITCH;
                if (dw < 2) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) *reinterpret_cast<uint4*>(rl + 4 * k) = make_uint4(vb[4 * k], vb[4 * k + 1], vb[4 * k + 2], vb[4 * k + 3]);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) *reinterpret_cast<uint4*>(rl + 4 * k) = make_uint4(va[4 * k], va[4 * k + 1], va[4 * k + 2], va[4 * k + 3]);
                }
            }
            bar_named(2 + (dw & 1), 64);
            // colour / depth / moments: the M = 64 accumulators keep record 16*w + l on lane l < 16 of warp w
            if (lane < 16) {
                const int r = 16 * dw + lane;
                if (r < cnt) {
                    const float4 hd = *reinterpret_cast<const float4*>(hdr + r * 16);
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
            {
                const int ch = (dw & 1) * 32 + lane;
                const float* rl = relay + (dt ^ 64) * RELAY_PITCH;  // the partner thread holds the other half of channel ch
                const int c0 = dw < 2 ? 0 : 32;    // first record column this thread finishes
                float fin[32];                     // channel ch of records c0 .. c0 + 31
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 o = *reinterpret_cast<const float4*>(rl + 4 * k);
                    const float o4[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) fin[4 * k + e] = __uint_as_float(dw < 2 ? va[4 * k + e] : vb[4 * k + e]) + o4[e];
                }
                bar_named(2 + (dw & 1), 64);  // both partners have read: the relay rows may be rewritten (next batch)
                // Two channels per reduction: neighbouring lanes (channels 2i, 2i + 1) trade halves of their 32 records, so that
                // the even lane finishes records 0-15 and the odd lane records 16-31 of the PAIR with red.global.add.v2 -- a warp
                // instruction still covers two contiguous 128-byte rows, and the L2 takes twice the floats per request
                // (tools/micro/red_bench.cu: 1.46 against 0.75 T float-adds/s for this access pattern).
                const bool odd = lane & 1;
                float* outp = dL_dlang_feat + (ch & ~1);
                const int rc0 = c0 + (odd ? 16 : 0);
                // the 16 ids up front (the reductions are compiler barriers: a load between them would wait out its latency)
                uint32_t idq[16];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint4 v = *reinterpret_cast<const uint4*>(&s_ids[s][rc0 + 4 * k]);
                    idq[4 * k] = v.x; idq[4 * k + 1] = v.y; idq[4 * k + 2] = v.z; idq[4 * k + 3] = v.w;
                }
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float send = odd ? fin[c] : fin[16 + c];
                    const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
                    if (rc0 + c < cnt) red_add_v2_f32(outp + (size_t)idq[c] * LF, odd ? recv : fin[c], odd ? fin[16 + c] : recv);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&slot_free[s]);  // this warp is done with TMEM set s, its headers and its count
        }
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
