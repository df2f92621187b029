// binning.cu -- tile binning for sm_100a: which Gaussians fall on which 8x8 tile, each tile's list in the
// reference's order; plus the carving of the three opaque work buffers.
//
// Replaces (reference rasterizer_impl.cu): InclusiveSum :277, duplicateWithKeys :70-111,
// getHigherMsb :35-50, SortPairs :304-309, cudaMemset+identifyTileRanges :116-138,311-320,
// and GeometryState/ImageState/BinningState::fromChunk :155-194.
//
// The reference sorts all R (Gaussian, tile) instances by the 64-bit key tile<<32 | depth bits with a stable
// radix sort: six 8-bit passes over 12 B per instance, each a latency-bound device-wide kernel at these sizes
// (0.17 ms at R = 1.4 M).  The order it produces is, per tile, ascending (depth bits, Gaussian index): a
// Gaussian appears at most once per tile and instances are emitted Gaussian-major, so the stable sort breaks
// depth ties by index.  Default (mode 1): the same CUB sort over fewer key bits (see launch_binning).  The order is
// unique, so it can also be produced tile by tile ("tile-local", lgs_binning_mode(0), experimental):
//   1. count_tiles      instances per tile (one red.add per instance);
//   2. tile_scan        exclusive scan over the tiles -> ranges (the reference's identifyTileRanges result);
//   3. scatter          every instance to its tile's segment, any order (one atomic cursor per tile),
//                       as depth bits << 32 | Gaussian index;
//   4. tile_sort        one CTA per tile sorts its segment in shared memory (bitonic network on the unique
//                       64-bit values) and writes point_list; segments beyond the shared-memory capacity are
//                       ranked in global memory by big_tile_sort (a rare, slow but exact path).
// No device-wide sort, no scan over Gaussians (R is accumulated by preprocess).  point_list and ranges are
// bit-identical to the reference's in both modes (GPU tests against the compiled reference and the goldens).
// Measured at cfgB (R = 1.43 M): mode 1 0.20 ms, tile-local 0.25 ms (scatter atomics 0.10, bitonic sorts 0.08) --
// hence not the default.  CUB (CUDA toolkit library, as in the reference) does the scan and the sort of mode 1.
#include <cub/cub.cuh>
#include "common.cuh"

namespace lgs {

// ---- buffers -------------------------------------------------------------------------
size_t scan_temp_bytes(int P) {
    size_t n = 0;
    cub::DeviceScan::InclusiveSum(nullptr, n, (uint32_t*)nullptr, (uint32_t*)nullptr, P);
    return n;
}
size_t sort_temp_bytes(int R) {
    size_t n = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, n, (uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (uint32_t*)nullptr, (uint32_t*)nullptr, R);
    size_t m = 0;  // the 32-bit key variant (launch_binning)
    cub::DeviceRadixSort::SortPairs(nullptr, m, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (uint32_t*)nullptr, R);
    return n > m ? n : m;
}

static int g_binning_mode = 1;  // 1 = the reference's single (tile|depth) radix sort (default), 0 = tile-local
static int g_debug_keys = 0;    // tile-local mode: also materialise the 64-bit key arrays for lgs_view_binning
void set_binning_mode(int m) { g_binning_mode = m; }
int binning_mode() { return g_binning_mode; }
void set_debug_keys(int on) { g_debug_keys = on; }
int debug_keys_on() { return g_debug_keys; }

GeomState geom_from_chunk(char* chunk, int P) {
    GeomState g;
    size_t n = (size_t)(P > 0 ? P : 1);
    carve(chunk, g.rec, n);
    carve(chunk, g.cov3D, n * 6);
    carve(chunk, g.tiles_touched, n);
    carve(chunk, g.point_offsets, n);
    carve(chunk, g.internal_radii, n);
    carve(chunk, g.clamped, n);
    carve(chunk, g.total_touched, 2);
    g.scan_temp_bytes = scan_temp_bytes((int)n);
    carve(chunk, g.scan_temp, g.scan_temp_bytes);
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
    carve(chunk, im.tile_count, tiles > 0 ? tiles : 1);
    carve(chunk, im.tile_cursor, tiles > 0 ? tiles : 1);
    return im;
}
BinningState binning_from_chunk(char* chunk, int R) {
    BinningState b;
    size_t n = (size_t)(R > 0 ? R : 1);
    carve(chunk, b.keys_unsorted, n);
    carve(chunk, b.keys, n);
    carve(chunk, b.vals_unsorted, n);
    carve(chunk, b.point_list, n);
    b.sort_temp_bytes = sort_temp_bytes((int)n);
    carve(chunk, b.sort_temp, b.sort_temp_bytes);
    return b;
}

// ---- scan ------------------------------------------------------------------------------
int launch_scan(int P, GeomState& g, cudaStream_t s) {
    if (g_binning_mode != 1) return LGS_OK;  // tile-local binning needs no per-Gaussian offsets
    size_t n = g.scan_temp_bytes;
    LGS_CUDA_TRY(cub::DeviceScan::InclusiveSum(g.scan_temp, n, g.tiles_touched, g.point_offsets, P, s));
    return LGS_OK;
}

// ---- key emission -----------------------------------------------------------------------
// One thread per Gaussian walks its tile rectangle (y-major, x-minor) exactly like the
// reference so that equal keys keep the reference's emission order under the stable sort.
// Rectangles are recomputed from the stored pixel centre and radius with the same float
// sequence as preprocess (getRect, auxiliary.h:46-56).
__global__ void __launch_bounds__(256)
emit_keys_kernel(int P, const GaussRec* __restrict__ rec, const uint32_t* __restrict__ offsets,
                 const int* __restrict__ radii, uint64_t* __restrict__ keys,
                 uint32_t* __restrict__ vals, int tiles_x, int tiles_y, uint32_t depth_base, int depth_bits) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const int rad = radii[idx];
    if (rad <= 0) return;
    uint32_t off = (idx == 0) ? 0u : offsets[idx - 1];
    const float4 q0 = rec[idx].q0;  // x, y, depth
    const float rf = (float)rad;
    const int x0 = min(tiles_x, max(0, (int)__fmul_rn(__fsub_rn(q0.x, rf), 0.125f)));
    const int y0 = min(tiles_y, max(0, (int)__fmul_rn(__fsub_rn(q0.y, rf), 0.125f)));
    const int x1 = min(tiles_x, max(0, (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(q0.x, rf), 8.0f), -1.0f), 0.125f)));
    const int y1 = min(tiles_y, max(0, (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(q0.y, rf), 8.0f), -1.0f), 0.125f)));
    const uint32_t dbits = __float_as_uint(q0.z) - depth_base;
    for (int y = y0; y < y1; ++y) {
        for (int x = x0; x < x1; ++x) {
            const uint64_t key = ((uint64_t)(uint32_t)(y * tiles_x + x) << depth_bits) | dbits;
            keys[off] = key;
            vals[off] = (uint32_t)idx;
            ++off;
        }
    }
}

// ---- tile-local binning -------------------------------------------------------------------
// tile rectangle of one Gaussian, the reference's getRect (auxiliary.h:46-56) from the stored centre and radius
__device__ __forceinline__ void tile_rect(const float4 q0, int rad, int tiles_x, int tiles_y, int& x0, int& y0, int& x1,
                                          int& y1) {
    const float rf = (float)rad;
    x0 = min(tiles_x, max(0, (int)__fmul_rn(__fsub_rn(q0.x, rf), 0.125f)));
    y0 = min(tiles_y, max(0, (int)__fmul_rn(__fsub_rn(q0.y, rf), 0.125f)));
    x1 = min(tiles_x, max(0, (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(q0.x, rf), 8.0f), -1.0f), 0.125f)));
    y1 = min(tiles_y, max(0, (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(q0.y, rf), 8.0f), -1.0f), 0.125f)));
}

template <bool SCATTER>
__global__ void __launch_bounds__(256)
tile_count_scatter_kernel(int P, const GaussRec* __restrict__ rec, const int* __restrict__ radii, int tiles_x, int tiles_y,
                          uint32_t* __restrict__ tile_count, const uint2* __restrict__ ranges,
                          uint32_t* __restrict__ tile_cursor, uint64_t* __restrict__ inst) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const int rad = radii[idx];
    if (rad <= 0) return;
    const float4 q0 = rec[idx].q0;  // x, y, depth
    int x0, y0, x1, y1;
    tile_rect(q0, rad, tiles_x, tiles_y, x0, y0, x1, y1);
    const uint64_t v = ((uint64_t)__float_as_uint(q0.z) << 32) | (uint32_t)idx;
    const int w = x1 - x0, cnt = w * (y1 - y0);
    if (!SCATTER) {
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) atomicAdd(tile_count + y * tiles_x + x, 1u);  // result unused: RED
    } else {
        // four independent cursor atomics in flight per thread, then the four stores (the atomic's round trip to
        // L2 is the cost of this kernel)
        int x = x0, y = y0;
        for (int k = 0; k < cnt; k += 4) {
            int t[4];
            uint32_t pos[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                t[u] = (k + u < cnt) ? y * tiles_x + x : -1;
                if (++x == x1) { x = x0; ++y; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) pos[u] = t[u] >= 0 ? atomicAdd(tile_cursor + t[u], 1u) : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (t[u] >= 0) inst[ranges[t[u]].x + pos[u]] = v;
        }
    }
}

// exclusive scan of the per-tile counts by one CTA; ranges[t] = (start, end), (0, 0) for an empty tile like the
// reference's memset + identifyTileRanges; zeroes the scatter cursors
__global__ void __launch_bounds__(1024)
tile_scan_kernel(int tiles, const uint32_t* __restrict__ tile_count, uint2* __restrict__ ranges,
                 uint32_t* __restrict__ tile_cursor) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < tiles; base += 1024) {
        const int t = base + tid;
        const uint32_t c = t < tiles ? tile_count[t] : 0u;
        uint32_t x = c;
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
            warp_sums[lane] = w;  // inclusive
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t incl = carry + x + (wrp > 0 ? warp_sums[wrp - 1] : 0u);
        if (t < tiles) {
            ranges[t] = c ? make_uint2(incl - c, incl) : make_uint2(0u, 0u);
            tile_cursor[t] = 0u;
        }
        __syncthreads();
        if (tid == 1023) carry_s = incl;
        __syncthreads();
    }
}

constexpr int TS_CAP = 4096;     // largest segment sorted in shared memory (32 KB)
constexpr int TS_THREADS = 256;

// One CTA per tile: bitonic sort of the tile's (depth bits << 32 | index) values in shared memory.
// Two instantiations cover (0, 1024] (8 KB, most tiles) and (1024, TS_CAP] so that the common case keeps full occupancy.
template <int LO, int HI>
__global__ void __launch_bounds__(TS_THREADS)
tile_sort_kernel(const uint2* __restrict__ ranges, const uint64_t* __restrict__ inst, uint32_t* __restrict__ point_list,
                 uint64_t* __restrict__ keys_dbg) {
    __shared__ __align__(16) uint64_t sv[HI];
    const int tile = blockIdx.x, tid = threadIdx.x;
    const uint2 rg = ranges[tile];
    const int n = (int)(rg.y - rg.x);
    if (n <= LO || n > HI) return;
    int N = 32;
    while (N < n) N <<= 1;
    for (int i = tid; i < N; i += TS_THREADS) sv[i] = i < n ? inst[rg.x + i] : ~0ull;
    __syncthreads();
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (N >> 1); t += TS_THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const uint64_t a = sv[i], b = sv[l];
                const bool up = (i & k) == 0;
                if ((a > b) == up) {
                    sv[i] = b;
                    sv[l] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < n; i += TS_THREADS) {
        const uint64_t v = sv[i];
        point_list[rg.x + i] = (uint32_t)v;
        if (keys_dbg != nullptr) keys_dbg[rg.x + i] = ((uint64_t)(uint32_t)tile << 32) | (v >> 32);
    }
}

// Segments longer than TS_CAP: every CTA of the grid ranks a slice of the segment's elements against the whole
// segment (values are unique, so rank = final position).  O(n^2) reads of L2-resident data; exact; rare.
__global__ void __launch_bounds__(256)
big_tile_sort_kernel(int tiles, const uint2* __restrict__ ranges, const uint64_t* __restrict__ inst,
                     uint32_t* __restrict__ point_list, uint64_t* __restrict__ keys_dbg) {
    __shared__ uint64_t chunk[1024];
    __shared__ int s_next;
    int tile = 0;
    for (;;) {
        // next tile >= `tile` with a long segment: 256 tiles per step, every CTA finds the same sequence
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) s_next = 0x7fffffff;
            __syncthreads();
            const int t = tile + threadIdx.x;
            if (t < tiles) {
                const uint2 r = ranges[t];
                if ((int)(r.y - r.x) > TS_CAP) atomicMin(&s_next, t);
            }
            __syncthreads();
            if (s_next != 0x7fffffff || tile + 256 >= tiles) break;
            tile += 256;
        }
        if (s_next == 0x7fffffff) return;
        tile = s_next;
        const uint2 rg = ranges[tile];
        const int n = (int)(rg.y - rg.x);
        const uint64_t* seg = inst + rg.x;
        for (int i0 = blockIdx.x * 256; i0 < n; i0 += gridDim.x * 256) {  // uniform per CTA
            const int i = i0 + threadIdx.x;
            const uint64_t mine = i < n ? seg[i] : 0ull;
            uint32_t rank = 0;
            for (int c0 = 0; c0 < n; c0 += 1024) {
                __syncthreads();
                for (int k = threadIdx.x; k < 1024; k += 256) chunk[k] = (c0 + k < n) ? seg[c0 + k] : ~0ull;
                __syncthreads();
                const int m = min(1024, n - c0);
                for (int k = 0; k < m; ++k) rank += chunk[k] < mine ? 1u : 0u;
            }
            if (i < n) {
                point_list[rg.x + rank] = (uint32_t)mine;
                if (keys_dbg != nullptr) keys_dbg[rg.x + rank] = ((uint64_t)(uint32_t)tile << 32) | (mine >> 32);
            }
        }
        ++tile;
        if (tile >= tiles) return;
    }
}

// tile-local mode, debug only: the scattered instances as (tile << 32 | depth bits, index) pairs
__global__ void __launch_bounds__(256)
unsorted_keys_kernel(int tiles, const uint2* __restrict__ ranges, const uint64_t* __restrict__ inst, uint64_t* __restrict__ keys,
                     uint32_t* __restrict__ vals) {
    const int tile = blockIdx.x;
    if (tile >= tiles) return;
    const uint2 rg = ranges[tile];
    for (uint32_t i = rg.x + threadIdx.x; i < rg.y; i += blockDim.x) {
        const uint64_t v = inst[i];
        vals[i] = (uint32_t)v;
        keys[i] = ((uint64_t)(uint32_t)tile << 32) | (v >> 32);
    }
}

// ---- tile ranges -----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_ranges_kernel(int L, const uint64_t* __restrict__ keys, uint2* __restrict__ ranges, int depth_bits) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L) return;
    const uint32_t cur = (uint32_t)(keys[idx] >> depth_bits);
    if (idx == 0) {
        ranges[cur].x = 0;
    } else {
        const uint32_t prev = (uint32_t)(keys[idx - 1] >> depth_bits);
        if (cur != prev) {
            ranges[prev].y = idx;
            ranges[cur].x = idx;
        }
    }
    if (idx == L - 1) ranges[cur].y = L;
}

// ---- 32-bit keys (default when the depth bound is known) ---------------------------------------------------------
// key32 = tile << q | (depth bits - bits(0.2f)) >> shift, q = what is left of 32 bits after the tile id (19 at 640x480).
// Four radix passes over 8 B per instance instead of five over 12 B.  When shift > 0 the low depth bits are not in the
// key: instances of one tile whose depths agree in the kept bits form a run (still in index order -- the sort is
// stable); tile_ranges_fix_kernel puts every such run into (depth bits, index) order, which is exactly the order the
// reference's full key produces.  Runs are rare and short (depths within 2^shift ulps of each other inside one tile).
__global__ void __launch_bounds__(256)
emit_keys32_kernel(int P, const GaussRec* __restrict__ rec, const uint32_t* __restrict__ offsets, const int* __restrict__ radii,
                   uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, int tiles_x, int tiles_y, uint32_t depth_base,
                   int shift, int q) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const int rad = radii[idx];
    if (rad <= 0) return;
    uint32_t off = (idx == 0) ? 0u : offsets[idx - 1];
    const float4 q0 = rec[idx].q0;  // x, y, depth
    const float rf = (float)rad;
    const int x0 = min(tiles_x, max(0, (int)__fmul_rn(__fsub_rn(q0.x, rf), 0.125f)));
    const int y0 = min(tiles_y, max(0, (int)__fmul_rn(__fsub_rn(q0.y, rf), 0.125f)));
    const int x1 = min(tiles_x, max(0, (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(q0.x, rf), 8.0f), -1.0f), 0.125f)));
    const int y1 = min(tiles_y, max(0, (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(q0.y, rf), 8.0f), -1.0f), 0.125f)));
    const uint32_t dq = (__float_as_uint(q0.z) - depth_base) >> shift;
    for (int y = y0; y < y1; ++y) {
        for (int x = x0; x < x1; ++x) {
            keys[off] = ((uint32_t)(y * tiles_x + x) << q) | dq;
            vals[off] = (uint32_t)idx;
            ++off;
        }
    }
}

__global__ void __launch_bounds__(256)
tile_ranges_fix_kernel(int L, const uint32_t* __restrict__ keys, uint32_t* __restrict__ point_list,
                       const GaussRec* __restrict__ rec, uint2* __restrict__ ranges, int q, int shift) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L) return;
    const uint32_t key = keys[idx];
    const uint32_t cur = key >> q;
    const bool first = idx == 0 || keys[idx - 1] != key;
    if (idx == 0) {
        ranges[cur].x = 0;
    } else {
        const uint32_t prev = keys[idx - 1] >> q;
        if (cur != prev) {
            ranges[prev].y = idx;
            ranges[cur].x = idx;
        }
    }
    if (idx == L - 1) ranges[cur].y = L;
    if (shift > 0 && first && idx + 1 < L && keys[idx + 1] == key) {
        // this thread owns the run [idx, end): insertion sort by (depth bits, Gaussian index)
        int end = idx + 2;
        while (end < L && keys[end] == key) ++end;
        for (int a = idx + 1; a < end; ++a) {
            const uint32_t id_a = point_list[a];
            const uint32_t d_a = __float_as_uint(rec[id_a].q0.z);
            int b = a - 1;
            while (b >= idx) {
                const uint32_t id_b = point_list[b];
                const uint32_t d_b = __float_as_uint(rec[id_b].q0.z);
                if (d_b < d_a || (d_b == d_a && id_b < id_a)) break;
                point_list[b + 1] = id_b;
                --b;
            }
            point_list[b + 1] = id_a;
        }
    }
}

static int key_bits_for_tiles(uint32_t n) {
    // == getHigherMsb(n) of the reference for n >= 1: number of bits needed to write n
    int bits = 1;
    while ((n >> bits) != 0) ++bits;
    return bits;
}

int launch_binning(int P, int R, int W, int H, const GeomState& g, const int* radii,
                   BinningState& b, ImageState& im, uint32_t max_depth_bits, cudaStream_t s) {
    const int tiles_x = (W + TILE - 1) / TILE, tiles_y = (H + TILE - 1) / TILE;
    const int tiles = tiles_x * tiles_y;
    if (g_binning_mode != 1) {
        LGS_CUDA_TRY(cudaMemsetAsync(im.tile_count, 0, (size_t)tiles * sizeof(uint32_t), s));
        if (R > 0) {
            tile_count_scatter_kernel<false><<<(P + 255) / 256, 256, 0, s>>>(P, g.rec, radii, tiles_x, tiles_y, im.tile_count,
                                                                             nullptr, nullptr, nullptr);
            LGS_LAUNCH_CHECK();
        }
        tile_scan_kernel<<<1, 1024, 0, s>>>(tiles, im.tile_count, im.ranges, im.tile_cursor);
        LGS_LAUNCH_CHECK();
        if (R <= 0) return LGS_OK;
        uint64_t* inst = b.keys_unsorted;  // the scattered (depth bits << 32 | index) values
        tile_count_scatter_kernel<true><<<(P + 255) / 256, 256, 0, s>>>(P, g.rec, radii, tiles_x, tiles_y, nullptr, im.ranges,
                                                                        im.tile_cursor, inst);
        LGS_LAUNCH_CHECK();
        prof_mark(PM_EMIT, s);
        uint64_t* keys_dbg = g_debug_keys ? b.keys : nullptr;
        tile_sort_kernel<0, 1024><<<tiles, TS_THREADS, 0, s>>>(im.ranges, inst, b.point_list, keys_dbg);
        tile_sort_kernel<1024, TS_CAP><<<tiles, TS_THREADS, 0, s>>>(im.ranges, inst, b.point_list, keys_dbg);
        LGS_LAUNCH_CHECK();
        big_tile_sort_kernel<<<148, 256, 0, s>>>(tiles, im.ranges, inst, b.point_list, keys_dbg);
        LGS_LAUNCH_CHECK();
        prof_mark(PM_SORT, s);
        prof_mark(PM_RANGES, s);
        if (g_debug_keys) {  // after the sorts: rewrites `inst` in place as 64-bit keys (same element, same slot)
            unsorted_keys_kernel<<<tiles, 256, 0, s>>>(tiles, im.ranges, inst, b.keys_unsorted, b.vals_unsorted);
            LGS_LAUNCH_CHECK();
        }
        return LGS_OK;
    }
    LGS_CUDA_TRY(cudaMemsetAsync(im.ranges, 0, (size_t)tiles * sizeof(uint2), s));
    if (R <= 0) return LGS_OK;
    // The reference sorts bits [0, 32 + msb(tiles)) of tile << 32 | depth bits (rasterizer_impl.cu:301-309).  Every
    // rendered Gaussian has depth > 0.2 (auxiliary.h:154), so depth bits - bits(0.2f) orders identically and, with the
    // largest depth known from preprocess, needs fewer bits: 13 + 26 = 39 at 640x480 indoors.  Default: 32-bit keys with
    // the top depth bits + a fix-up of the rare equal-key runs (above); 64-bit compacted keys when the tile id needs
    // more than 20 bits.  With lgs_debug_keys(1) the reference's exact 64-bit keys are kept (parity tests compare them).
    uint32_t depth_base = 0;
    int depth_bits = 32;
    const uint32_t kNear = 0x3e4ccccdu;  // bits of 0.2f
    if (!g_debug_keys && max_depth_bits != 0xffffffffu && max_depth_bits >= kNear) {
        depth_base = kNear;
        depth_bits = key_bits_for_tiles(max_depth_bits - kNear);
    }
    const int tile_bits = key_bits_for_tiles((uint32_t)tiles);
    if (depth_base != 0 && tile_bits <= 20) {  // 32-bit keys: tile | as many depth bits as fit, runs fixed up afterwards
        const int q = min(depth_bits, 32 - tile_bits), shift = depth_bits - q;
        uint32_t* k32_unsorted = reinterpret_cast<uint32_t*>(b.keys_unsorted);  // the 64-bit key arrays hold the 32-bit keys
        uint32_t* k32 = reinterpret_cast<uint32_t*>(b.keys);
        emit_keys32_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, g.rec, g.point_offsets, radii, k32_unsorted, b.vals_unsorted,
                                                           tiles_x, tiles_y, depth_base, shift, q);
        LGS_LAUNCH_CHECK();
        prof_mark(PM_EMIT, s);
        size_t n32 = b.sort_temp_bytes;
        LGS_CUDA_TRY(cub::DeviceRadixSort::SortPairs(b.sort_temp, n32, k32_unsorted, k32, b.vals_unsorted, b.point_list, R, 0,
                                                     q + tile_bits, s));
        prof_mark(PM_SORT, s);
        tile_ranges_fix_kernel<<<(R + 255) / 256, 256, 0, s>>>(R, k32, b.point_list, g.rec, im.ranges, q, shift);
        LGS_LAUNCH_CHECK();
        prof_mark(PM_RANGES, s);
        return LGS_OK;
    }
    emit_keys_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, g.rec, g.point_offsets, radii,
                                                     b.keys_unsorted, b.vals_unsorted, tiles_x, tiles_y, depth_base, depth_bits);
    LGS_LAUNCH_CHECK();
    prof_mark(PM_EMIT, s);
    const int end_bit = depth_bits + tile_bits;
    size_t n = b.sort_temp_bytes;
    LGS_CUDA_TRY(cub::DeviceRadixSort::SortPairs(b.sort_temp, n, b.keys_unsorted, b.keys,
                                                 b.vals_unsorted, b.point_list, R, 0, end_bit, s));
    prof_mark(PM_SORT, s);
    tile_ranges_kernel<<<(R + 255) / 256, 256, 0, s>>>(R, b.keys, im.ranges, depth_bits);
    LGS_LAUNCH_CHECK();
    prof_mark(PM_RANGES, s);
    return LGS_OK;
}

}  // namespace lgs
