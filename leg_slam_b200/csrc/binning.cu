// binning.cu -- tile binning for sm_100a: prefix sum of tiles_touched, (tile|depth)
// key emission, 64-bit LSD radix sort, per-tile ranges; plus the carving of the three
// opaque work buffers.
//
// Replaces (reference rasterizer_impl.cu): InclusiveSum :277, duplicateWithKeys :70-111,
// getHigherMsb :35-50, SortPairs :304-309, cudaMemset+identifyTileRanges :116-138,311-320,
// and GeometryState/ImageState/BinningState::fromChunk :155-194.
//
// Parity contract: keys_unsorted, values_unsorted (emission order = Gaussian-major, then
// tile y, then tile x), the sorted key/value lists and ranges are bit-exact with the
// reference.  The sort is an LSD radix sort (stable) over key bits [0, 32+msb(tiles)),
// like the reference; CUB's DeviceRadixSort (CUDA toolkit library, as in the reference)
// instantiated here for sm_100a is used for the scan and the sort.
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
    return n;
}

GeomState geom_from_chunk(char* chunk, int P) {
    GeomState g;
    size_t n = (size_t)(P > 0 ? P : 1);
    carve(chunk, g.rec, n);
    carve(chunk, g.cov3D, n * 6);
    carve(chunk, g.tiles_touched, n);
    carve(chunk, g.point_offsets, n);
    carve(chunk, g.internal_radii, n);
    carve(chunk, g.clamped, n);
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
                 uint32_t* __restrict__ vals, int tiles_x, int tiles_y) {
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
    const uint32_t dbits = __float_as_uint(q0.z);
    for (int y = y0; y < y1; ++y) {
        for (int x = x0; x < x1; ++x) {
            const uint64_t key = ((uint64_t)(uint32_t)(y * tiles_x + x) << 32) | dbits;
            keys[off] = key;
            vals[off] = (uint32_t)idx;
            ++off;
        }
    }
}

// ---- tile ranges -----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_ranges_kernel(int L, const uint64_t* __restrict__ keys, uint2* __restrict__ ranges) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L) return;
    const uint32_t cur = (uint32_t)(keys[idx] >> 32);
    if (idx == 0) {
        ranges[cur].x = 0;
    } else {
        const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
        if (cur != prev) {
            ranges[prev].y = idx;
            ranges[cur].x = idx;
        }
    }
    if (idx == L - 1) ranges[cur].y = L;
}

static int key_bits_for_tiles(uint32_t n) {
    // == getHigherMsb(n) of the reference for n >= 1: number of bits needed to write n
    int bits = 1;
    while ((n >> bits) != 0) ++bits;
    return bits;
}

int launch_binning(int P, int R, int W, int H, const GeomState& g, const int* radii,
                   BinningState& b, ImageState& im, cudaStream_t s) {
    const int tiles_x = (W + TILE - 1) / TILE, tiles_y = (H + TILE - 1) / TILE;
    const int tiles = tiles_x * tiles_y;
    LGS_CUDA_TRY(cudaMemsetAsync(im.ranges, 0, (size_t)tiles * sizeof(uint2), s));
    if (R <= 0) return LGS_OK;
    emit_keys_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, g.rec, g.point_offsets, radii,
                                                     b.keys_unsorted, b.vals_unsorted, tiles_x, tiles_y);
    LGS_LAUNCH_CHECK();
    prof_mark(PM_EMIT, s);
    const int end_bit = 32 + key_bits_for_tiles((uint32_t)tiles);
    size_t n = b.sort_temp_bytes;
    LGS_CUDA_TRY(cub::DeviceRadixSort::SortPairs(b.sort_temp, n, b.keys_unsorted, b.keys,
                                                 b.vals_unsorted, b.point_list, R, 0, end_bit, s));
    prof_mark(PM_SORT, s);
    tile_ranges_kernel<<<(R + 255) / 256, 256, 0, s>>>(R, b.keys, im.ranges);
    LGS_LAUNCH_CHECK();
    prof_mark(PM_RANGES, s);
    return LGS_OK;
}

}  // namespace lgs
