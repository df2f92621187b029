// ingest.cu -- the geometry kernels that turn a new keyframe into Gaussians, the step BEFORE the mapping hot path
// (SURVEY.md section 8f row 4).  Same results as the reference's
//   reprojectDepthPinhole   src/stereo_vision.cu:40-61,135-162   (depth image -> camera-space points)
//   transformPoints         src/operate_points.cu:39-94          (points <- T * points, transformPoint4x3)
//   distCUDA2               third_party/simple-knn/simple_knn.cu:60-220, spatial.cu:15-27
//                           (mean squared distance to the 3 nearest neighbours; initial log-scales,
//                            src/gaussian_model.cpp:157,242,331)
// The k-NN keeps simple-knn's exact search (Morton order, 1024-point boxes, box rejection by the running 3rd-best
// distance) -- the result, the mean of the three smallest squared distances, does not depend on the search order --
// but runs without host round trips (the bounding box stays on the device; the reference reads it back twice),
// scans a Morton-ordered COPY of the points (coalesced, instead of an index indirection per candidate) and keeps the
// box table in shared memory.
#include <cfloat>
#include <cub/cub.cuh>
#include "common.cuh"

namespace lgs {

__global__ void __launch_bounds__(256)
reproject_depth_kernel(int P, int width, float fx, float fy, float cx, float cy, const float* __restrict__ depths,
                       const unsigned char* __restrict__ mask, float* __restrict__ points) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    float x = 0.f, y = 0.f, z = 0.f;  // the reference leaves its zero-initialised rows untouched where !mask
    if (mask[idx]) {
        const int v = idx / width, u = idx - v * width;
        const float d = depths[idx];
        x = __fdiv_rn(__fmul_rn(__fsub_rn((float)u, cx), d), fx);  // stereo_vision.h:51-53
        y = __fdiv_rn(__fmul_rn(__fsub_rn((float)v, cy), d), fy);
        z = d;
    }
    points[3 * (size_t)idx] = x;
    points[3 * (size_t)idx + 1] = y;
    points[3 * (size_t)idx + 2] = z;
}

__global__ void __launch_bounds__(256)
transform_points_kernel(int P, const float* __restrict__ in, const float* __restrict__ T, float* __restrict__ out) {
    __shared__ float m[16];
    if (threadIdx.x < 16) m[threadIdx.x] = T[threadIdx.x];
    __syncthreads();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float x = in[3 * (size_t)idx], y = in[3 * (size_t)idx + 1], z = in[3 * (size_t)idx + 2];
    // transformPoint4x3 (auxiliary.h:58-66) with the contraction nvcc gives the reference (see preprocess.cu)
#pragma unroll
    for (int r = 0; r < 3; ++r)
        out[3 * (size_t)idx + r] = __fadd_rn(m[12 + r], __fmaf_rn(z, m[8 + r], __fmaf_rn(x, m[r], __fmul_rn(y, m[4 + r]))));
}

// ---- k-NN -------------------------------------------------------------------------------------------------
constexpr int KNN_BOX = 1024;

// order-preserving float <-> uint mapping for atomicMin / atomicMax on floats
__device__ __forceinline__ uint32_t f2o(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float o2f(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// bbox[0..2] = min (ordered uint), bbox[3..5] = max.  simple-knn reduces with init {0,0,0} (simple_knn.cu:189-199),
// i.e. the box always contains the origin; the caller initialises bbox with f2o(0).
__global__ void __launch_bounds__(256)
knn_bbox_kernel(int P, const float* __restrict__ pts, uint32_t* __restrict__ bbox) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = pts[3 * (size_t)i + k];
            mn[k] = fminf(mn[k], v);
            mx[k] = fmaxf(mx[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(bbox + k, f2o(mn[k]));
            atomicMax(bbox + 3 + k, f2o(mx[k]));
        }
    }
}

__device__ __forceinline__ uint32_t prep_morton(uint32_t x) {  // simple_knn.cu:41-48
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}

__global__ void __launch_bounds__(256)
knn_morton_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ bbox, uint32_t* __restrict__ codes,
                  uint32_t* __restrict__ ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    uint32_t c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float mn = o2f(bbox[k]), mx = o2f(bbox[3 + k]);
        c[k] = prep_morton((uint32_t)(((pts[3 * (size_t)i + k] - mn) / (mx - mn)) * 1023.0f));  // :50-57
    }
    codes[i] = c[0] | (c[1] << 1) | (c[2] << 2);
    ids[i] = (uint32_t)i;
}

// Morton-ordered copy of the points (float4: x, y, z, original index bits) + per-box bounds
__global__ void __launch_bounds__(KNN_BOX)
knn_boxes_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ order, float4* __restrict__ sorted,
                 float* __restrict__ boxes) {
    const int i = blockIdx.x * KNN_BOX + threadIdx.x;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < P) {
        const uint32_t id = order[i];
        const float x = pts[3 * (size_t)id], y = pts[3 * (size_t)id + 1], z = pts[3 * (size_t)id + 2];
        sorted[i] = make_float4(x, y, z, __uint_as_float(id));
        mn[0] = mx[0] = x; mn[1] = mx[1] = y; mn[2] = mx[2] = z;
    }
    __shared__ float red[6][KNN_BOX / 32];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
        if ((threadIdx.x & 31) == 0) {
            red[k][threadIdx.x >> 5] = mn[k];
            red[3 + k][threadIdx.x >> 5] = mx[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = red[threadIdx.x][0];
        for (int w = 1; w < KNN_BOX / 32; ++w) v = threadIdx.x < 3 ? fminf(v, red[threadIdx.x][w]) : fmaxf(v, red[threadIdx.x][w]);
        boxes[6 * (size_t)blockIdx.x + threadIdx.x] = v;
    }
}

__device__ __forceinline__ void k_best3(float dist, float (&best)[3]) {  // updateKBest<3>, simple_knn.cu:133-147
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (best[j] > dist) {
            const float t = best[j];
            best[j] = dist;
            dist = t;
        }
    }
}
__device__ __forceinline__ float dist2(const float4 a, const float4 p) {
    const float dx = a.x - p.x, dy = a.y - p.y, dz = a.z - p.z;
    // d.x*d.x + d.y*d.y + d.z*d.z as nvcc contracts it in the reference build (SASS of boxMeanDist: FMUL on y, then
    // FFMA x, FFMA z) -- keeps the result bit-identical to simple-knn
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__global__ void __launch_bounds__(256)
knn_mean_dist_kernel(int P, int n_boxes, const float4* __restrict__ sorted, const float* __restrict__ boxes,
                     float* __restrict__ dists) {
    extern __shared__ float sbox[];  // [n_boxes][6]
    for (int k = threadIdx.x; k < 6 * n_boxes; k += blockDim.x) sbox[k] = boxes[k];
    __syncthreads();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float4 p = sorted[idx];
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    for (int i = max(0, idx - 3); i <= min(P - 1, idx + 3); ++i) {  // :157-162
        if (i == idx) continue;
        k_best3(dist2(sorted[i], p), best);
    }
    const float reject = best[2];
    best[0] = best[1] = best[2] = FLT_MAX;
    for (int b = 0; b < n_boxes; ++b) {  // :169-182
        const float* bx = sbox + 6 * b;
        float dx = 0.f, dy = 0.f, dz = 0.f;  // distBoxPoint :120-130
        if (p.x < bx[0] || p.x > bx[3]) dx = fminf(fabsf(p.x - bx[0]), fabsf(p.x - bx[3]));
        if (p.y < bx[1] || p.y > bx[4]) dy = fminf(fabsf(p.y - bx[1]), fabsf(p.y - bx[4]));
        if (p.z < bx[2] || p.z > bx[5]) dz = fminf(fabsf(p.z - bx[2]), fabsf(p.z - bx[5]));
        const float bd = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        if (bd > reject || bd > best[2]) continue;
        const int i1 = min(P, (b + 1) * KNN_BOX);
        for (int i = b * KNN_BOX; i < i1; ++i) {
            if (i == idx) continue;
            k_best3(dist2(sorted[i], p), best);
        }
    }
    dists[__float_as_uint(p.w)] = (best[0] + best[1] + best[2]) / 3.0f;  // :183
}

static size_t knn_sort_bytes(int P) {
    size_t n = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, n, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, P);
    return (n + 255) & ~(size_t)255;
}

}  // namespace lgs

using namespace lgs;

extern "C" int lgs_reproject_depth_pinhole(int P, int width, float fx, float fy, float cx, float cy, const float* depth,
                                           const unsigned char* mask, float* points, void* stream) {
    if (P < 0 || width <= 0) return LGS_ERR_INVALID_ARG;
    if (P == 0) return LGS_OK;
    if (!depth || !mask || !points) return LGS_ERR_INVALID_ARG;
    reproject_depth_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, width, fx, fy, cx, cy, depth, mask, points);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

extern "C" int lgs_transform_points(int P, const float* points, const float* transformmatrix, float* out, void* stream) {
    if (P < 0) return LGS_ERR_INVALID_ARG;
    if (P == 0) return LGS_OK;
    if (!points || !transformmatrix || !out || points == out) return LGS_ERR_INVALID_ARG;
    transform_points_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, points, transformmatrix, out);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

// scratch: bbox [6+2] u32 | codes [P] | codes_sorted [P] | ids [P] | order [P] | sorted [P] float4 | boxes [n_boxes][6] | CUB temp
extern "C" size_t lgs_knn_scratch_bytes(int P) {
    if (P <= 0) return 0;
    const size_t n = ((size_t)P + 63) & ~(size_t)63;
    const size_t nb = ((size_t)P + KNN_BOX - 1) / KNN_BOX;
    return 256 + 4 * n * 4 + n * 16 + ((nb * 24 + 255) & ~(size_t)255) + knn_sort_bytes(P) + 256;
}

extern "C" int lgs_knn_mean_dist2(int P, const float* points, float* mean_dist2, char* scratch, void* stream) {
    if (P < 0) return LGS_ERR_INVALID_ARG;
    if (P == 0) return LGS_OK;
    if (!points || !mean_dist2 || !scratch) return LGS_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = ((size_t)P + 63) & ~(size_t)63;
    const int n_boxes = (P + KNN_BOX - 1) / KNN_BOX;
    char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~(uintptr_t)255);
    uint32_t* bbox = reinterpret_cast<uint32_t*>(p); p += 256;
    uint32_t* codes = reinterpret_cast<uint32_t*>(p); p += n * 4;
    uint32_t* codes_sorted = reinterpret_cast<uint32_t*>(p); p += n * 4;
    uint32_t* ids = reinterpret_cast<uint32_t*>(p); p += n * 4;
    uint32_t* order = reinterpret_cast<uint32_t*>(p); p += n * 4;
    float4* sorted = reinterpret_cast<float4*>(p); p += n * 16;
    float* boxes = reinterpret_cast<float*>(p); p += ((size_t)n_boxes * 24 + 255) & ~(size_t)255;
    char* temp = p;
    // f2o(0.0f) = 0x80000000 in all six slots: the reference's reductions start from {0, 0, 0}
    static const uint32_t init[6] = {0x80000000u, 0x80000000u, 0x80000000u, 0x80000000u, 0x80000000u, 0x80000000u};
    LGS_CUDA_TRY(cudaMemcpyAsync(bbox, init, sizeof(init), cudaMemcpyHostToDevice, s));
    const int rb = P < 148 * 8 * 256 ? (P + 255) / 256 : 148 * 8;
    knn_bbox_kernel<<<rb, 256, 0, s>>>(P, points, bbox);
    LGS_LAUNCH_CHECK();
    knn_morton_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, points, bbox, codes, ids);
    LGS_LAUNCH_CHECK();
    size_t tb = knn_sort_bytes(P);
    LGS_CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp, tb, codes, codes_sorted, ids, order, P, 0, 32, s));
    knn_boxes_kernel<<<n_boxes, KNN_BOX, 0, s>>>(P, points, order, sorted, boxes);
    LGS_LAUNCH_CHECK();
    const size_t smem = (size_t)n_boxes * 24;
    if (smem > 200 * 1024) return LGS_ERR_INVALID_ARG;  // > 8.7 M points: outside what a keyframe ingest produces
    if (smem > 48 * 1024)  // per device and per launch: a cached "configured" size would be wrong on a process's second GPU
        LGS_CUDA_TRY(cudaFuncSetAttribute(knn_mean_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_mean_dist_kernel<<<(P + 255) / 256, 256, smem, s>>>(P, n_boxes, sorted, boxes, mean_dist2);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}
