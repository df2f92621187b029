// binning.cu -- tile binning for sm_100a: which Gaussians fall on which 8x8 tile, each tile's list in the
// reference's order; plus the carving of the three opaque work buffers.  No library call: every kernel is ours.
//
// Replaces (reference rasterizer_impl.cu): InclusiveSum :277, duplicateWithKeys :70-111, getHigherMsb :35-50,
// cub::DeviceRadixSort::SortPairs :304-309, cudaMemset + identifyTileRanges :116-138,311-320, and
// GeometryState / ImageState / BinningState::fromChunk :155-194.
//
// The reference sorts all R (Gaussian, tile) instances by the 64-bit key tile << 32 | depth bits with a stable radix
// sort over 32 + msb(tiles) bits: six 8-bit passes over 12 B per instance.  The order it produces is, per tile,
// ascending (depth bits, Gaussian index): a Gaussian appears at most once per tile and instances are emitted
// Gaussian-major, so the stable sort breaks depth ties by index.  That order is unique, which leaves the emission
// order and the key width free:
//
//   emit_keys_kernel     persistent CTAs.  A warp takes 32 Gaussians, scans their tile counts, reserves a block of
//                        slots with ONE atomic per CTA round and writes the (key, index) pairs with coalesced stores
//                        (lane i writes instance i of the warp's flattened list) -- no per-Gaussian offsets, hence no
//                        scan over Gaussians.  key = tile << q | (depth bits - bits(0.2f)) >> shift: every rendered
//                        Gaussian has depth > 0.2 (the near cull, auxiliary.h:154) and the frame's largest depth is
//                        accumulated by preprocess, so the ordered depth range needs ~26 bits of which the top
//                        q = 32 - bits(tiles) go into a 32-bit key.  The kernel also builds the three 11-bit digit
//                        histograms of the sort (shared-memory atomics, one flush per CTA) and clears the sort's status
//                        arrays and the tile ranges.
//   radix_pass_kernel    x3 (digits 11 + 11 + 10 bits, least significant first): single-pass ("onesweep") stable
//                        scatter.  A CTA takes tiles of 5120 pairs from a ticket counter, ranks them per digit with
//                        warp-private counters (match.any), publishes its per-digit counts, and obtains its global
//                        offsets by a TWO-LEVEL decoupled look-back: tiles are grouped by 16; inside a group a tile
//                        sums its (at most 15) predecessors' counts, the last tile of a group publishes the group's
//                        aggregate and then its inclusive prefix, and every tile walks back over the groups until it
//                        meets an inclusive prefix.  With ~280 tiles in flight at R = 1.4 M the classic one-level chain
//                        would be ~280 dependent L2 round trips long; this one is ~3 + the number of groups without an
//                        inclusive prefix yet.
//   tile_ranges_fix_kernel  ranges[tile] = (start, end) (identifyTileRanges) and the repair of the key compaction:
//                        instances of one tile whose depths agree in the kept bits form a run of equal keys, in
//                        arbitrary order; every run is put into (depth bits, Gaussian index) order -- exactly the
//                        reference's order.  Short runs by their first thread, long runs (> 32) by the whole warp
//                        with a rank-counting sort through the idle half of the ping-pong arrays, so that quantised
//                        depths (thousands of equal keys in one tile) cost O(n^2 / 32) coalesced steps, not O(n^2)
//                        dependent loads of one thread.
//
// Everything data-dependent (R, the depth bound, hence q and shift) is read from the geometry buffer's header ON THE
// DEVICE: stage2 needs no value from the host beyond the capacity of the binning buffer, so a caller may skip the
// reference's blocking read-back of num_rendered (rasterizer_impl.cu:281-282) altogether (lgs.h, lgs_forward_stage1
// with num_rendered_host == NULL).  point_list and ranges are bit-identical to the reference's (GPU tests against
// the compiled reference, the golden fixtures and the CPU oracle; lgs_debug_reference_keys re-expresses the lists as
// the reference's 64-bit keys).
#include "common.cuh"

namespace lgs {

// ---- buffers -------------------------------------------------------------------------
GeomState geom_from_chunk(char* chunk, int P) {
    GeomState g;
    size_t n = (size_t)(P > 0 ? P : 1);
    carve(chunk, g.rec, n);
    carve(chunk, g.cov3D, n * 6);
    carve(chunk, g.tiles_touched, n);
    carve(chunk, g.internal_radii, n);
    carve(chunk, g.clamped, n);
    carve(chunk, g.hdr, (size_t)HDR_WORDS + 3 * RS_BINS);
    g.hist = g.hdr + HDR_WORDS;
    g.end = reinterpret_cast<char*>(g.hist + 3 * RS_BINS);
    return g;
}
ImageState image_from_chunk(char* chunk, int W, int H) {
    ImageState im;
    const size_t tiles = (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    const size_t npix = (size_t)W * H;
    carve(chunk, im.ranges, tiles > 0 ? tiles : 1);
    carve(chunk, im.final_T, npix > 0 ? npix : 1);
    carve(chunk, im.n_contrib, npix > 0 ? npix : 1);
    carve(chunk, im.tile_last, tiles > 0 ? tiles : 1);
    im.end = reinterpret_cast<char*>(im.tile_last + (tiles > 0 ? tiles : 1));
    return im;
}
BinningState binning_from_chunk(char* chunk, int R) {
    BinningState b;
    const size_t n = (size_t)(R > 0 ? R : 1);
    b.n_tiles_cap = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    b.n_groups_cap = (b.n_tiles_cap + RS_GROUP - 1) / RS_GROUP;
    carve(chunk, b.keys[0], n);
    carve(chunk, b.keys[1], n);
    carve(chunk, b.vals[0], n);
    carve(chunk, b.vals[1], n);
    b.point_list = b.vals[1];
    // the sort's status arrays, contiguous (emit_keys_kernel clears them in one sweep)
    carve(chunk, b.grp_stat, (size_t)3 * b.n_groups_cap * RS_BINS);
    b.tile_agg = reinterpret_cast<uint16_t*>(b.grp_stat + (size_t)3 * b.n_groups_cap * RS_BINS);
    b.status_bytes = (size_t)3 * b.n_groups_cap * RS_BINS * sizeof(uint32_t) + (size_t)3 * b.n_tiles_cap * RS_BINS * sizeof(uint16_t);
    b.end = reinterpret_cast<char*>(b.grp_stat) + b.status_bytes;
    return b;
}

// ---- key layout, derived on the device from the frame's largest depth ---------------------------------------------------
__host__ __device__ __forceinline__ int bits_to_write(uint32_t n) {
    // == getHigherMsb(n) of the reference for n >= 1 (rasterizer_impl.cu:35-50): number of bits needed to write n
    int bits = 1;
    while (bits < 32 && (n >> bits) != 0) ++bits;
    return bits;
}
constexpr uint32_t kNearBits = 0x3e4ccccdu;  // bit pattern of 0.2f: every rendered Gaussian lies beyond it
struct KeyLayout {
    int q;      // depth bits kept in the key (low q bits); the tile id sits above them
    int shift;  // low depth bits dropped: key depth = (bits - kNearBits) >> shift
};
__device__ __forceinline__ KeyLayout key_layout(uint32_t max_depth_bits, int tile_bits) {
    const uint32_t span = max_depth_bits >= kNearBits ? max_depth_bits - kNearBits : 0u;
    const int depth_bits = bits_to_write(span);
    KeyLayout k;
    k.q = min(depth_bits, 32 - tile_bits);
    k.shift = depth_bits - k.q;
    return k;
}

// tile rectangle of one Gaussian, the reference's getRect (auxiliary.h:46-56) from the stored centre and radius with the
// same float sequence as preprocess
__device__ __forceinline__ void tile_rect(const float4 q0, int rad, int tiles_x, int tiles_y, int& x0, int& y0, int& x1,
                                          int& y1) {
    const float rf = (float)rad;
    x0 = min(tiles_x, max(0, (int)__fmul_rn(__fsub_rn(q0.x, rf), 0.125f)));
    y0 = min(tiles_y, max(0, (int)__fmul_rn(__fsub_rn(q0.y, rf), 0.125f)));
    x1 = min(tiles_x, max(0, (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(q0.x, rf), 8.0f), -1.0f), 0.125f)));
    y1 = min(tiles_y, max(0, (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(q0.y, rf), 8.0f), -1.0f), 0.125f)));
}

// ---- key emission -----------------------------------------------------------------------
constexpr int EMIT_THREADS = 512;
constexpr int EMIT_WARPS = EMIT_THREADS / 32;

__global__ void __launch_bounds__(EMIT_THREADS)
emit_keys_kernel(int P, const GaussRec* __restrict__ rec, const int* __restrict__ radii, int tiles_x, int tiles_y, int tile_bits,
                 uint32_t* __restrict__ hdr, uint32_t* __restrict__ hist, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                 uint32_t cap, uint16_t* __restrict__ tile_agg, size_t agg_stride, uint32_t* __restrict__ grp_stat, size_t gst_stride,
                 uint2* __restrict__ ranges, int n_tiles) {
    __shared__ uint32_t s_hist[3][RS_BINS];
    __shared__ uint32_t s_warp_tot[EMIT_WARPS];
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const KeyLayout kl = key_layout(hdr[HDR_MAX_DEPTH], tile_bits);
    const uint32_t R_total = hdr[HDR_R];
    // clear what the later kernels of this forward expect to be zero: the part of the sort's status arrays this frame's
    // tiles will use (three passes each) and the tile ranges
    {
        const uint32_t R_eff = min(R_total, cap);
        const size_t st = (R_eff + RS_TILE - 1) / RS_TILE, sg = (st + RS_GROUP - 1) / RS_GROUP;
        const size_t na = st * (RS_BINS * sizeof(uint16_t) / 16), ng = sg * (RS_BINS * sizeof(uint32_t) / 16);  // uint4 per pass
        for (size_t i = (size_t)blockIdx.x * EMIT_THREADS + tid; i < 3 * (na + ng); i += (size_t)gridDim.x * EMIT_THREADS) {
            const size_t pass = i / (na + ng), k = i - pass * (na + ng);
            uint4* dst = k < na ? reinterpret_cast<uint4*>(tile_agg + pass * agg_stride) + k
                                : reinterpret_cast<uint4*>(grp_stat + pass * gst_stride) + (k - na);
            *dst = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    for (int i = blockIdx.x * EMIT_THREADS + tid; i < n_tiles; i += gridDim.x * EMIT_THREADS) ranges[i] = make_uint2(0u, 0u);
    for (int i = tid; i < 3 * RS_BINS; i += EMIT_THREADS) (&s_hist[0][0])[i] = 0u;
    if (blockIdx.x == 0 && tid == 0 && R_total > cap) hdr[HDR_OVERFLOW] = 1u;
    __syncthreads();

    const int rounds = (P + EMIT_THREADS - 1) / EMIT_THREADS;
    for (int round = blockIdx.x; round < rounds; round += gridDim.x) {
        const int idx = round * EMIT_THREADS + tid;
        uint32_t cnt = 0, xy = 0, w = 0, dq = 0;
        if (idx < P) {
            const int rad = radii[idx];
            if (rad > 0) {
                const float4 q0 = rec[idx].q0;  // x, y, depth
                int x0, y0, x1, y1;
                tile_rect(q0, rad, tiles_x, tiles_y, x0, y0, x1, y1);
                w = (uint32_t)(x1 - x0);
                cnt = w * (uint32_t)(y1 - y0);
                xy = (uint32_t)x0 | ((uint32_t)y0 << 16);
                dq = (__float_as_uint(q0.z) - kNearBits) >> kl.shift;
            }
        }
        // inclusive scan of the counts inside the warp, warp totals -> one slot reservation per CTA round
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) s_warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t t = lane < EMIT_WARPS ? s_warp_tot[lane] : 0u;
            uint32_t ti = t;
#pragma unroll
            for (int o = 1; o < EMIT_WARPS; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, ti, o);
                if (lane >= o) ti += y;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, ti, EMIT_WARPS - 1);
            uint32_t base = 0;
            if (lane == 0 && total != 0) base = atomicAdd(hdr + HDR_CURSOR, total);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (lane < EMIT_WARPS) s_warp_tot[lane] = base + ti - t;  // this warp's first slot
        }
        __syncthreads();
        const uint32_t warp_base = s_warp_tot[warp];
        const uint32_t T = __shfl_sync(0xffffffffu, incl, 31);
        // lane i writes instance i of the warp's flattened list: its owner is the first lane whose inclusive count exceeds i
        for (uint32_t i0 = 0; i0 < T; i0 += 32) {
            const uint32_t i = i0 + lane;
            int j = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const uint32_t v = __shfl_sync(0xffffffffu, incl, j + step - 1);
                if (v <= i) j += step;
            }
            j = min(j, 31);
            const uint32_t o_incl = __shfl_sync(0xffffffffu, incl, j);
            const uint32_t o_cnt = __shfl_sync(0xffffffffu, cnt, j);
            const uint32_t o_xy = __shfl_sync(0xffffffffu, xy, j);
            const uint32_t o_w = __shfl_sync(0xffffffffu, w, j);
            const uint32_t o_dq = __shfl_sync(0xffffffffu, dq, j);
            if (i < T) {
                const uint32_t local = i - (o_incl - o_cnt);
                const uint32_t ry = local / o_w, rx = local - ry * o_w;  // y-major, x-minor like the reference (irrelevant here)
                const uint32_t tile = ((o_xy >> 16) + ry) * (uint32_t)tiles_x + (o_xy & 0xffffu) + rx;
                const uint32_t key = (tile << kl.q) | o_dq;
                const uint32_t slot = warp_base + i;
                if (slot < cap) {
                    keys[slot] = key;
                    vals[slot] = (uint32_t)(round * EMIT_THREADS + warp * 32 + j);
                    atomicAdd(&s_hist[0][key & (RS_BINS - 1)], 1u);
                    atomicAdd(&s_hist[1][(key >> RS_BITS) & (RS_BINS - 1)], 1u);
                    atomicAdd(&s_hist[2][key >> (2 * RS_BITS)], 1u);
                }
            }
        }
        __syncthreads();  // s_warp_tot is rewritten by the next round
    }
    __syncthreads();
    for (int i = tid; i < 3 * RS_BINS; i += EMIT_THREADS) {
        const uint32_t c = (&s_hist[0][0])[i];
        if (c != 0) atomicAdd(hist + i, c);
    }
}

// ---- radix sort: one stable scatter pass ------------------------------------------------------------------------------
constexpr uint16_t AGG_READY = 0x8000u;        // tile_agg: bit 15 = published, bits 0-14 = count (<= RS_TILE)
constexpr uint32_t GS_AGG = 0x40000000u;       // grp_stat: bits 30-31 = 1 aggregate published, 2 inclusive prefix published
constexpr uint32_t GS_INC = 0x80000000u;
constexpr uint32_t GS_VAL = 0x3fffffffu;
constexpr int SPIN_LIMIT = 1 << 22;            // a look-back that never completes aborts (HDR_ERROR) instead of hanging

__device__ __forceinline__ uint4 ld_volatile_u4(const void* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u4(void* p, uint4 v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// this thread's 8 consecutive digits of up to 4 consecutive tiles' published counts, added to acc once all of them are there
// (the loads of one round are independent: one L2 round trip per 4 predecessors)
__device__ __forceinline__ bool wait_tile_agg4(const uint16_t* p, int n, uint32_t (&acc)[8]) {
    for (int spin = 0; spin < SPIN_LIMIT; ++spin) {
        uint4 v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) v[t] = t < n ? ld_volatile_u4(p + (size_t)t * RS_BINS) : make_uint4(0x80008000u, 0x80008000u, 0x80008000u, 0x80008000u);
        uint32_t ready = 0x80008000u;
#pragma unroll
        for (int t = 0; t < 4; ++t) ready &= v[t].x & v[t].y & v[t].z & v[t].w;
        if (ready == 0x80008000u) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const uint32_t w[4] = {v[t].x, v[t].y, v[t].z, v[t].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    acc[2 * k] += w[k] & 0x7fffu;
                    acc[2 * k + 1] += (w[k] >> 16) & 0x7fffu;
                }
            }
            return true;
        }
    }
    return false;
}
// this thread's 8 digits of one group's status, once all 8 are at least aggregates; returns the AND of the flag fields
__device__ __forceinline__ bool wait_group(const uint32_t* p, uint32_t (&c)[8], uint32_t& all_inclusive) {
    for (int spin = 0; spin < SPIN_LIMIT; ++spin) {
        const uint4 a = ld_volatile_u4(p), b = ld_volatile_u4(p + 4);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        bool ok = true;
        uint32_t inc = GS_INC;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            ok = ok && (w[k] & (GS_AGG | GS_INC)) != 0;
            inc &= w[k];
        }
        if (ok) {
#pragma unroll
            for (int k = 0; k < 8; ++k) c[k] = w[k];
            all_inclusive = inc & GS_INC;
            return true;
        }
    }
    return false;
}

template <int PASS>
__global__ void __launch_bounds__(RS_THREADS, 2)
radix_pass_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
                  uint32_t* __restrict__ vals_out, uint32_t* __restrict__ hdr, const uint32_t* __restrict__ hist, uint32_t cap,
                  uint16_t* __restrict__ tile_agg, uint32_t* __restrict__ grp_stat) {
    constexpr int SHIFT = RS_BITS * PASS;
    constexpr uint32_t MASK = RS_BINS - 1;
    constexpr int WARPS = RS_THREADS / 32;
    static_assert(RS_THREADS * 8 == RS_BINS, "one thread owns 8 consecutive digits");
    __shared__ __align__(16) uint16_t s_wcnt[WARPS][RS_BINS];  // per-warp digit counts, then exclusive prefixes over the warps
    __shared__ uint32_t s_dbase[RS_BINS];                      // first output slot of this tile's items of each digit
    __shared__ uint32_t s_part[WARPS];
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t R = min(hdr[HDR_R], cap);
    const uint32_t n_tiles = (R + RS_TILE - 1) / RS_TILE;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // global start of each of this thread's 8 digits: exclusive scan of the pass's histogram
    uint32_t gstart[8];
    {
        const uint4 a = *reinterpret_cast<const uint4*>(hist + tid * 8), b = *reinterpret_cast<const uint4*>(hist + tid * 8 + 4);
        const uint32_t h[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { gstart[k] = sum; sum += h[k]; }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) s_part[warp] = incl;
        __syncthreads();
        uint32_t before = incl - sum;
        for (int w2 = 0; w2 < warp; ++w2) before += s_part[w2];
#pragma unroll
        for (int k = 0; k < 8; ++k) gstart[k] += before;
    }

    for (;;) {
        __syncthreads();  // everyone is done with s_tile / s_wcnt / s_dbase of the previous tile
        if (tid == 0) s_tile = atomicAdd(hdr + HDR_TICKET + PASS, 1u);
        {   // clear the warp counters: 32 KB, 128 B per thread
            uint4* z = reinterpret_cast<uint4*>(&s_wcnt[0][0]);
#pragma unroll
            for (int k = 0; k < (WARPS * RS_BINS * 2) / (16 * RS_THREADS); ++k) z[tid + k * RS_THREADS] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= n_tiles) break;

        // ---- load: warp-striped, so that the order inside the tile is (warp, item, lane)
        uint32_t key[RS_ITEMS], val[RS_ITEMS];
        const uint32_t first = tile * RS_TILE + warp * (32 * RS_ITEMS) + lane;
#pragma unroll
        for (int k = 0; k < RS_ITEMS; ++k) {
            const uint32_t i = first + k * 32;
            const bool valid = i < R;
            key[k] = valid ? __ldcs(keys_in + i) : 0xffffffffu;
            val[k] = valid ? __ldcs(vals_in + i) : 0u;
        }
        // ---- rank inside the warp: running per-digit counts in shared memory, ties inside one row by lane.  The 20 warp
        // matches are independent of each other and of the counters: issue them all first (their latency overlaps), then walk
        // the rows for the counter updates, whose read-modify-write chain through shared memory is the only serial part.
        uint16_t rank[RS_ITEMS];
        uint16_t* my_cnt = &s_wcnt[warp][0];
        uint32_t peers[RS_ITEMS];
#pragma unroll
        for (int k = 0; k < RS_ITEMS; ++k) {
            const bool valid = first + k * 32 < R;
            const uint32_t d = valid ? ((key[k] >> SHIFT) & MASK) : (0x10000u + lane);  // invalid lanes match nobody
            peers[k] = __match_any_sync(0xffffffffu, d);
        }
#pragma unroll
        for (int k = 0; k < RS_ITEMS; ++k) {
            const bool valid = first + k * 32 < R;
            const uint32_t d = (key[k] >> SHIFT) & MASK;
            const int leader = __ffs(peers[k]) - 1;
            uint32_t old = 0;
            if (valid && lane == leader) {
                old = my_cnt[d];
                my_cnt[d] = (uint16_t)(old + __popc(peers[k]));
            }
            old = __shfl_sync(0xffffffffu, old, leader);
            rank[k] = (uint16_t)(old + __popc(peers[k] & lt_mask));
            __syncwarp();
        }
        __syncthreads();

        // ---- this thread's 8 digits: totals over the warps (the counters become exclusive prefixes over the warps)
        uint32_t own[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int w2 = 0; w2 < WARPS; ++w2) {
            uint4* p = reinterpret_cast<uint4*>(&s_wcnt[w2][tid * 8]);
            const uint4 v = *p;
            const uint32_t c[4] = {v.x, v.y, v.z, v.w};
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t lo = c[k] & 0xffffu, hi = c[k] >> 16;
                o[k] = own[2 * k] | (own[2 * k + 1] << 16);
                own[2 * k] += lo;
                own[2 * k + 1] += hi;
            }
            *p = make_uint4(o[0], o[1], o[2], o[3]);
        }
        // ---- publish this tile's counts
        {
            uint4 v;
            v.x = (own[0] | AGG_READY) | ((own[1] | AGG_READY) << 16);
            v.y = (own[2] | AGG_READY) | ((own[3] | AGG_READY) << 16);
            v.z = (own[4] | AGG_READY) | ((own[5] | AGG_READY) << 16);
            v.w = (own[6] | AGG_READY) | ((own[7] | AGG_READY) << 16);
            st_volatile_u4(tile_agg + (size_t)tile * RS_BINS + tid * 8, v);
        }
        // ---- two-level look-back
        const uint32_t grp = tile / RS_GROUP, t0 = grp * RS_GROUP;
        const bool last_of_group = (tile == t0 + RS_GROUP - 1) || (tile == n_tiles - 1);
        bool ok = true;
        uint32_t pre[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // items of each digit in earlier tiles
        for (uint32_t p = t0; p < tile && ok; p += 4)
            ok = wait_tile_agg4(tile_agg + (size_t)p * RS_BINS + tid * 8, (int)min(4u, tile - p), pre);
        uint32_t* my_stat = grp_stat + (size_t)grp * RS_BINS + tid * 8;
        if (last_of_group) {  // the group's aggregate is this tile's same-group prefix + its own counts
            st_volatile_u4(my_stat, make_uint4(GS_AGG | (pre[0] + own[0]), GS_AGG | (pre[1] + own[1]), GS_AGG | (pre[2] + own[2]),
                                               GS_AGG | (pre[3] + own[3])));
            st_volatile_u4(my_stat + 4, make_uint4(GS_AGG | (pre[4] + own[4]), GS_AGG | (pre[5] + own[5]), GS_AGG | (pre[6] + own[6]),
                                                   GS_AGG | (pre[7] + own[7])));
        }
        uint32_t gpre[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // items of each digit in earlier groups
        {
            bool done[8] = {false, false, false, false, false, false, false, false};
            for (int g2 = (int)grp - 1; g2 >= 0 && ok; --g2) {
                uint32_t c[8], all_inc;
                ok = wait_group(grp_stat + (size_t)g2 * RS_BINS + tid * 8, c, all_inc);
                bool all_done = true;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (!done[k]) {
                        gpre[k] += c[k] & GS_VAL;
                        done[k] = (c[k] & GS_INC) != 0;
                    }
                    all_done = all_done && done[k];
                }
                if (all_done) break;
            }
        }
        if (last_of_group) {
            st_volatile_u4(my_stat, make_uint4(GS_INC | (gpre[0] + pre[0] + own[0]), GS_INC | (gpre[1] + pre[1] + own[1]),
                                               GS_INC | (gpre[2] + pre[2] + own[2]), GS_INC | (gpre[3] + pre[3] + own[3])));
            st_volatile_u4(my_stat + 4, make_uint4(GS_INC | (gpre[4] + pre[4] + own[4]), GS_INC | (gpre[5] + pre[5] + own[5]),
                                                   GS_INC | (gpre[6] + pre[6] + own[6]), GS_INC | (gpre[7] + pre[7] + own[7])));
        }
        if (!ok) hdr[HDR_ERROR] = 1u;  // a predecessor never published: give up on this frame instead of hanging the GPU
#pragma unroll
        for (int k = 0; k < 8; ++k) s_dbase[tid * 8 + k] = gstart[k] + gpre[k] + pre[k];
        __syncthreads();

        // ---- scatter
#pragma unroll
        for (int k = 0; k < RS_ITEMS; ++k) {
            if (first + k * 32 < R) {
                const uint32_t d = (key[k] >> SHIFT) & MASK;
                const uint32_t pos = s_dbase[d] + my_cnt[d] + rank[k];
                if (pos < cap) {  // always true unless the frame was aborted
                    keys_out[pos] = key[k];
                    vals_out[pos] = val[k];
                }
            }
        }
    }
}

// ---- tile ranges + repair of the equal-key runs ------------------------------------------------------------------------
// order of two instances of one tile in the reference's list: (depth bits, Gaussian index)
__device__ __forceinline__ bool inst_less(uint32_t da, uint32_t ia, uint32_t db, uint32_t ib) { return da < db || (da == db && ia < ib); }

constexpr int RUN_SHORT = 32;  // runs up to this length are sorted by their first thread

__global__ void __launch_bounds__(256)
tile_ranges_fix_kernel(const uint32_t* __restrict__ hdr, uint32_t cap, const uint32_t* __restrict__ keys, uint32_t* __restrict__ point_list,
                       uint32_t* __restrict__ tmp_depth, uint32_t* __restrict__ tmp_ids, const GaussRec* __restrict__ rec,
                       uint2* __restrict__ ranges, int tile_bits) {
    const uint32_t L = min(hdr[HDR_R], cap);
    const KeyLayout kl = key_layout(hdr[HDR_MAX_DEPTH], tile_bits);
    const int q = kl.q;
    const int lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    // whole warps iterate together (the long-run repair below is warp-cooperative)
    const uint32_t rounds = (L + stride - 1) / stride;
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint32_t idx = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
        uint32_t run_len = 0;
        if (idx < L) {
            const uint32_t key = keys[idx];
            const uint32_t cur = key >> q;
            const uint32_t prev_key = idx > 0 ? keys[idx - 1] : ~key;
            if (idx == 0) {
                ranges[cur].x = 0;
            } else {
                const uint32_t prev = prev_key >> q;
                if (cur != prev) {
                    ranges[prev].y = idx;
                    ranges[cur].x = idx;
                }
            }
            if (idx == L - 1) ranges[cur].y = L;
            if (prev_key != key && idx + 1 < L && keys[idx + 1] == key) {
                // this thread owns the run of equal keys [idx, end): the emission order inside it is arbitrary
                uint32_t end = idx + 2;
                while (end < L && end - idx <= RUN_SHORT && keys[end] == key) ++end;
                run_len = end - idx;
                if (run_len <= RUN_SHORT) {  // insertion sort by (depth bits, Gaussian index)
                    for (uint32_t a = idx + 1; a < end; ++a) {
                        const uint32_t id_a = point_list[a];
                        const uint32_t d_a = __float_as_uint(rec[id_a].q0.z);
                        uint32_t b = a;
                        while (b > idx) {
                            const uint32_t id_b = point_list[b - 1];
                            const uint32_t d_b = __float_as_uint(rec[id_b].q0.z);
                            if (inst_less(d_b, id_b, d_a, id_a)) break;
                            point_list[b] = id_b;
                            --b;
                        }
                        point_list[b] = id_a;
                    }
                    run_len = 0;
                }
            }
        }
        // long runs: the warp sorts them one after the other by rank counting.  Phase 1 stages the depth bits (and the
        // ids) of the run in the idle ping-pong arrays, phase 2 counts for every element the elements that precede it
        // (32 candidates per coalesced load, shuffled through the warp), phase 3 writes the ids to their ranks.
        uint32_t pending = __ballot_sync(0xffffffffu, run_len != 0);
        while (pending) {
            const int src = __ffs(pending) - 1;
            pending &= pending - 1;
            const uint32_t start = __shfl_sync(0xffffffffu, idx, src);
            const uint32_t key = keys[start];
            uint32_t end = start + RUN_SHORT;  // known to be a run at least this long; find its end together
            for (;;) {
                const uint32_t i = end + lane;
                const uint32_t same = __ballot_sync(0xffffffffu, i < L && keys[i] == key);
                if (same == 0xffffffffu) { end += 32; continue; }
                end += __ffs(~same) - 1;
                break;
            }
            const uint32_t n = end - start;
            for (uint32_t i = lane; i < n; i += 32) {
                const uint32_t id = point_list[start + i];
                tmp_ids[start + i] = id;
                tmp_depth[start + i] = __float_as_uint(rec[id].q0.z);
            }
            __threadfence_block();
            __syncwarp();
            for (uint32_t i0 = 0; i0 < n; i0 += 32) {
                const uint32_t i = i0 + lane;
                const bool mine = i < n;
                const uint32_t di = mine ? tmp_depth[start + i] : 0u, ii = mine ? tmp_ids[start + i] : 0u;
                uint32_t rank = 0;
                for (uint32_t j0 = 0; j0 < n; j0 += 32) {
                    const uint32_t j = j0 + lane;
                    const uint32_t dj = j < n ? tmp_depth[start + j] : 0xffffffffu, ij = j < n ? tmp_ids[start + j] : 0xffffffffu;
#pragma unroll 8
                    for (int s = 0; s < 32; ++s) {
                        const uint32_t d2 = __shfl_sync(0xffffffffu, dj, s), i2 = __shfl_sync(0xffffffffu, ij, s);
                        rank += inst_less(d2, i2, di, ii) ? 1u : 0u;
                    }
                }
                if (mine) point_list[start + rank] = ii;
            }
            __syncwarp();
        }
    }
}

static int emit_grid() { return 148 * 2; }

int launch_binning(int P, int R, int W, int H, const GeomState& g, const int* radii, BinningState& b, ImageState& im, cudaStream_t s) {
    const int tiles_x = (W + TILE - 1) / TILE, tiles_y = (H + TILE - 1) / TILE;
    const int tiles = tiles_x * tiles_y;
    const int tile_bits = bits_to_write((uint32_t)tiles);
    if (tile_bits > 24) return LGS_ERR_INVALID_ARG;  // > 16 M tiles: not an image this library is meant for
    const uint32_t cap = (uint32_t)(R > 0 ? R : 0);
    const int rounds = (P + EMIT_THREADS - 1) / EMIT_THREADS;
    const int egrid = rounds < emit_grid() ? (rounds > 0 ? rounds : 1) : emit_grid();
    const size_t agg_stride = (size_t)b.n_tiles_cap * RS_BINS, gst_stride = (size_t)b.n_groups_cap * RS_BINS;
    emit_keys_kernel<<<egrid, EMIT_THREADS, 0, s>>>(P, g.rec, radii, tiles_x, tiles_y, tile_bits, g.hdr, g.hist, b.keys[0], b.vals[0], cap,
                                                    b.tile_agg, agg_stride, b.grp_stat, gst_stride, im.ranges, tiles);
    LGS_LAUNCH_CHECK();
    prof_mark(PM_EMIT, s);
    if (cap == 0) {
        prof_mark(PM_SORT, s);
        prof_mark(PM_RANGES, s);
        return LGS_OK;
    }
    const uint32_t want = b.n_tiles_cap;
    const int sgrid = (int)(want < 148u * 2u ? want : 148u * 2u);
    uint16_t* agg = b.tile_agg;
    uint32_t* gst = b.grp_stat;
    radix_pass_kernel<0><<<sgrid, RS_THREADS, 0, s>>>(b.keys[0], b.vals[0], b.keys[1], b.vals[1], g.hdr, g.hist, cap, agg, gst);
    radix_pass_kernel<1><<<sgrid, RS_THREADS, 0, s>>>(b.keys[1], b.vals[1], b.keys[0], b.vals[0], g.hdr, g.hist + RS_BINS, cap,
                                                      agg + agg_stride, gst + gst_stride);
    radix_pass_kernel<2><<<sgrid, RS_THREADS, 0, s>>>(b.keys[0], b.vals[0], b.keys[1], b.vals[1], g.hdr, g.hist + 2 * RS_BINS, cap,
                                                      agg + 2 * agg_stride, gst + 2 * gst_stride);
    LGS_LAUNCH_CHECK();
    prof_mark(PM_SORT, s);
    const uint32_t rb = (cap + 255) / 256;
    const int rgrid = (int)(rb < 148u * 8u ? rb : 148u * 8u);
    tile_ranges_fix_kernel<<<rgrid, 256, 0, s>>>(g.hdr, cap, b.keys[1], b.point_list, b.keys[0], b.vals[0], g.rec, im.ranges, tile_bits);
    LGS_LAUNCH_CHECK();
    prof_mark(PM_RANGES, s);
    return LGS_OK;
}

// ---- the reference's key arrays, for parity tests (lgs_debug_reference_keys) ---------------------------------------------
// Re-expresses what the production path computed in the reference's terms (rasterizer_impl.h:50-63): the emitted
// (key = tile << 32 | depth bits, value = Gaussian index) pairs in the reference's Gaussian-major order, and the sorted key
// array read off point_list + ranges.  Test infrastructure on the device side; never on the mapping path.
__global__ void __launch_bounds__(1024)
debug_scan_kernel(int P, const uint32_t* __restrict__ tiles_touched, uint32_t* __restrict__ offsets) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < P; base += 1024) {
        const int t = base + tid;
        uint32_t x = t < P ? tiles_touched[t] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[wrp] = x;
        __syncthreads();
        if (wrp == 0) {
            uint32_t w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const uint32_t incl = carry_s + x + (wrp > 0 ? warp_sums[wrp - 1] : 0u);
        if (t < P) offsets[t] = incl;
        __syncthreads();
        if (tid == 1023) carry_s = incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
debug_emit_keys64_kernel(int P, const GaussRec* __restrict__ rec, const uint32_t* __restrict__ offsets, const int* __restrict__ radii,
                         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int tiles_x, int tiles_y, uint32_t cap) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const int rad = radii[idx];
    if (rad <= 0) return;
    uint32_t off = (idx == 0) ? 0u : offsets[idx - 1];
    const float4 q0 = rec[idx].q0;
    int x0, y0, x1, y1;
    tile_rect(q0, rad, tiles_x, tiles_y, x0, y0, x1, y1);
    for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
            if (off < cap) {
                keys[off] = ((uint64_t)(uint32_t)(y * tiles_x + x) << 32) | __float_as_uint(q0.z);
                vals[off] = (uint32_t)idx;
            }
            ++off;
        }
}

__global__ void __launch_bounds__(256)
debug_sorted_keys_kernel(int tiles, const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                         const GaussRec* __restrict__ rec, uint64_t* __restrict__ keys_sorted) {
    const int tile = blockIdx.x;
    if (tile >= tiles) return;
    const uint2 rg = ranges[tile];
    for (uint32_t i = rg.x + threadIdx.x; i < rg.y; i += blockDim.x)
        keys_sorted[i] = ((uint64_t)(uint32_t)tile << 32) | __float_as_uint(rec[point_list[i]].q0.z);
}

DebugKeys debug_keys_from_chunk(char* chunk, int P, int R) {
    DebugKeys d;
    const size_t n = (size_t)(R > 0 ? R : 1), p = (size_t)(P > 0 ? P : 1);
    carve(chunk, d.keys_unsorted, n);
    carve(chunk, d.keys_sorted, n);
    carve(chunk, d.vals_unsorted, n);
    carve(chunk, d.offsets, p);
    d.end = reinterpret_cast<char*>(d.offsets + p);
    return d;
}

int launch_debug_reference_keys(int P, int R, int W, int H, const GeomState& g, const BinningState& b, const ImageState& im,
                                DebugKeys& d, cudaStream_t s) {
    const int tiles_x = (W + TILE - 1) / TILE, tiles_y = (H + TILE - 1) / TILE;
    if (P <= 0 || R <= 0) return LGS_OK;
    debug_scan_kernel<<<1, 1024, 0, s>>>(P, g.tiles_touched, d.offsets);
    debug_emit_keys64_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, g.rec, d.offsets, g.internal_radii, d.keys_unsorted, d.vals_unsorted,
                                                            tiles_x, tiles_y, (uint32_t)R);
    debug_sorted_keys_kernel<<<tiles_x * tiles_y, 256, 0, s>>>(tiles_x * tiles_y, im.ranges, b.point_list, g.rec, d.keys_sorted);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

}  // namespace lgs
