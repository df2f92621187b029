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

namespace lgs {

// ================================ loop-closure correction of the visible Gaussians ====================================
// scaleAndTransformThenMarkVisiblePoints (reference src/operate_points.cu:52-70,96-140, device helpers
// cuda_rasterizer/operate_points.h:54-179): the Gaussians a corrected keyframe sees (markVisible), that are still
// flagged "not transformed" and "unstable", get   point <- T * (scale * point),   rotation <- quaternion of T[:3,:3] * R(q).
// The reference runs markVisible, two logical_ands, a sum().item(), two zero-filled temporaries, one kernel and six
// boolean-index gathers / scatters; here it is ONE kernel that updates the rows in place (every thread owns its row) and
// counts the selected rows with one atomic per warp.
//
// As shipped, the reference stores the new quaternion's z at offset 2 twice and never writes offset 3
// (operate_points.h:169-178) into a zero-filled temporary whose whole rows are then copied back: a corrected row reads
// (w, x, z, 0).  `faithful_rot_store` = 1 reproduces that ("do not fix silently"), 0 stores
// (w, x, y, z).
__global__ void __launch_bounds__(256)
scale_transform_visible_kernel(int P, float scale, float* __restrict__ points, float* __restrict__ rots,
                               unsigned char* __restrict__ not_transformed, const unsigned char* __restrict__ unstable,
                               const float* __restrict__ T, const float* __restrict__ view, int faithful_rot_store,
                               int* __restrict__ num_transformed) {
    __shared__ float m[16], vw[16];
    if (threadIdx.x < 16) {
        m[threadIdx.x] = T[threadIdx.x];
        vw[threadIdx.x] = view[threadIdx.x];
    }
    __syncthreads();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    bool sel = false;
    if (idx < P && not_transformed[idx] && unstable[idx]) {
        const float x = points[3 * (size_t)idx], y = points[3 * (size_t)idx + 1], z = points[3 * (size_t)idx + 2];
        // markVisible on the ORIGINAL point: view-space z > 0.2 (auxiliary.h:139-164; same expression as lgs_mark_visible)
        const float depth = __fadd_rn(vw[14], __fmaf_rn(z, vw[10], __fmaf_rn(x, vw[2], __fmul_rn(y, vw[6]))));
        sel = !(depth <= 0.2f);
        if (sel) {
            const float sx = __fmul_rn(x, scale), sy = __fmul_rn(y, scale), sz = __fmul_rn(z, scale);
#pragma unroll
            for (int r = 0; r < 3; ++r)  // transformPoint4x3 (auxiliary.h:58-66)
                points[3 * (size_t)idx + r] = __fadd_rn(m[12 + r], __fmaf_rn(sz, m[8 + r], __fmaf_rn(sx, m[r], __fmul_rn(sy, m[4 + r]))));
            // rotation matrix of the stored (w, x, y, z) quaternion, NOT normalised (operate_points.h:59-87)
            const float qw = rots[4 * (size_t)idx], qx = rots[4 * (size_t)idx + 1], qy = rots[4 * (size_t)idx + 2], qz = rots[4 * (size_t)idx + 3];
            // Every product sum below is spelled out with the fused multiply-adds nvcc gives the reference's statement of
            // these formulas (read off its SASS, like the transforms above): the rotation rows then come out bit-identical.
            const float tx = __fadd_rn(qx, qx), ty = __fadd_rn(qy, qy), tz = __fadd_rn(qz, qz);
            const float twx = __fmul_rn(tx, qw), twy = __fmul_rn(ty, qw), twz = __fmul_rn(tz, qw);
            const float tyy = __fmul_rn(ty, qy), tzz = __fmul_rn(tz, qz);
            float R0[3][3];
            R0[0][0] = __fsub_rn(1.0f, __fadd_rn(tyy, tzz));
            R0[1][1] = __fsub_rn(1.0f, __fmaf_rn(qx, tx, tzz));
            R0[2][2] = __fsub_rn(1.0f, __fmaf_rn(qx, tx, tyy));
            R0[0][1] = __fmaf_rn(qx, ty, -twz); R0[1][0] = __fmaf_rn(qx, ty, twz);
            R0[0][2] = __fmaf_rn(qx, tz, twy);  R0[2][0] = __fmaf_rn(qx, tz, -twy);
            R0[1][2] = __fmaf_rn(qy, tz, -twx); R0[2][1] = __fmaf_rn(qy, tz, twx);
            float R[3][3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)  // T[:3,:3] * R0 with T stored transposed (:90-99): middle product first ...
                    R[i][j] = __fmaf_rn(R0[2][j], m[8 + i], __fmaf_rn(R0[0][j], m[i], __fmul_rn(R0[1][j], m[4 + i])));
            // ... except the first element, which the reference build starts from its first product
            R[0][0] = __fmaf_rn(R0[2][0], m[8], __fmaf_rn(R0[1][0], m[4], __fmul_rn(R0[0][0], m[0])));
            // matrix -> quaternion, Shoemake's branches (:101-147)
            float ow, ox, oy, oz;
            float t = __fadd_rn(__fadd_rn(R[0][0], R[1][1]), R[2][2]);
            if (t > 0.0f) {
                t = __fsqrt_rn(__fadd_rn(t, 1.0f));
                ow = __fmul_rn(0.5f, t);
                t = __fdiv_rn(0.5f, t);
                ox = __fmul_rn(__fsub_rn(R[2][1], R[1][2]), t);
                oy = __fmul_rn(__fsub_rn(R[0][2], R[2][0]), t);
                oz = __fmul_rn(__fsub_rn(R[1][0], R[0][1]), t);
            } else {
                int i = 0;
                if (R[1][1] > R[0][0]) i = 1;
                if (R[2][2] > (i == 0 ? R[0][0] : R[1][1])) i = 2;
                // the three cyclic cases written out (register-resident matrix: no dynamic indexing)
                float rii, rjj, rkk, rkj, rjk, rji, rij, rki, rik;
                if (i == 0)      { rii = R[0][0]; rjj = R[1][1]; rkk = R[2][2]; rkj = R[2][1]; rjk = R[1][2]; rji = R[1][0]; rij = R[0][1]; rki = R[2][0]; rik = R[0][2]; }
                else if (i == 1) { rii = R[1][1]; rjj = R[2][2]; rkk = R[0][0]; rkj = R[0][2]; rjk = R[2][0]; rji = R[2][1]; rij = R[1][2]; rki = R[0][1]; rik = R[1][0]; }
                else             { rii = R[2][2]; rjj = R[0][0]; rkk = R[1][1]; rkj = R[1][0]; rjk = R[0][1]; rji = R[0][2]; rij = R[2][0]; rki = R[1][2]; rik = R[2][1]; }
                t = __fsqrt_rn(__fadd_rn(__fsub_rn(__fsub_rn(rii, rjj), rkk), 1.0f));
                const float ci = __fmul_rn(0.5f, t);
                t = __fdiv_rn(0.5f, t);
                ow = __fmul_rn(__fsub_rn(rkj, rjk), t);
                const float cj = __fmul_rn(__fadd_rn(rji, rij), t), ck = __fmul_rn(__fadd_rn(rki, rik), t);
                ox = i == 0 ? ci : (i == 1 ? ck : cj);
                oy = i == 0 ? cj : (i == 1 ? ci : ck);
                oz = i == 0 ? ck : (i == 1 ? cj : ci);
            }
            float4 o = faithful_rot_store ? make_float4(ow, ox, oz, 0.0f) : make_float4(ow, ox, oy, oz);
            *reinterpret_cast<float4*>(rots + 4 * (size_t)idx) = o;
            not_transformed[idx] = 0;
        }
    }
    const unsigned n = __popc(__ballot_sync(0xffffffffu, sel));
    if ((threadIdx.x & 31) == 0 && n > 0) atomicAdd(num_transformed, (int)n);
}

}  // namespace lgs

extern "C" int lgs_scale_transform_mark_visible(int P, float scale, float* points, float* rots, unsigned char* not_transformed_mask,
                                                const unsigned char* unstable_mask, const float* transformmatrix,
                                                const float* viewmatrix, int faithful_rot_store, int* num_transformed,
                                                void* stream) {
    if (P < 0) return LGS_ERR_INVALID_ARG;
    if (P == 0) return LGS_OK;
    if (!points || !rots || !not_transformed_mask || !unstable_mask || !transformmatrix || !viewmatrix || !num_transformed)
        return LGS_ERR_INVALID_ARG;
    if (reinterpret_cast<uintptr_t>(rots) & 15u) return LGS_ERR_INVALID_ARG;  // rows are stored as float4
    scale_transform_visible_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        P, scale, points, rots, not_transformed_mask, unstable_mask, transformmatrix, viewmatrix, faithful_rot_store, num_transformed);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

namespace lgs {

// ================================ inactive-geometry densification (monocular keypoints) ===============================
// monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints (reference src/stereo_vision.cu:63-133,164-212):
// a keypoint without a 3D point borrows the depth of the nearest keypoint (in pixels) that has one and is reprojected
// with it; keypoints that end with a positive depth are returned, in order, with their colours.  The reference runs one
// thread per keypoint over ALL keypoints in global memory and compacts with a boolean index on the host side; here the
// candidates stream through shared memory in tiles (visited in index order, so distance ties resolve to the lowest
// index exactly as in the sequential loop) and a second, single-CTA kernel compacts in order and leaves the count on the
// device.  Semantics kept as shipped: `max_pixel_dist` bounds the SQUARED distance (:105-108); the colour triple is read
// at float offset v * width + u, truncated, WITHOUT a factor of three (:81,88-90,125-127); u, v are truncated to integers
// for the reprojection (cuda_rasterizer/stereo_vision.h:41-55).  Offsets outside `colors` (undefined behaviour in the
// reference) read as zero.
constexpr int IGD_TILE = 256;

__global__ void __launch_bounds__(IGD_TILE)
inactive_geo_search_kernel(int N, int width, float fx, float fy, float cx, float cy, float max_pixel_dist,
                           const float2* __restrict__ pixels, const unsigned char* __restrict__ has3D,
                           const float* __restrict__ point3D, const float* __restrict__ colors, long long colors_len,
                           float* __restrict__ res_p, float* __restrict__ res_c, unsigned char* __restrict__ keep) {
    __shared__ float4 cand[IGD_TILE];  // u, v, z, has3D
    const int idx = blockIdx.x * IGD_TILE + threadIdx.x;
    const bool live = idx < N;
    float2 px = make_float2(0.f, 0.f);
    bool has = false;
    if (live) {
        px = pixels[idx];
        has = has3D[idx] != 0;
    }
    float min_dist = FLT_MAX, depth = -1.0f;
    for (int base = 0; base < N; base += IGD_TILE) {
        const int j = base + threadIdx.x;
        __syncthreads();
        if (j < N) {
            const float2 q = pixels[j];
            cand[threadIdx.x] = make_float4(q.x, q.y, point3D[3 * (size_t)j + 2], has3D[j] ? 1.0f : 0.0f);
        }
        __syncthreads();
        if (live && !has) {
            const int cnt = min(IGD_TILE, N - base);
            for (int k = 0; k < cnt; ++k) {
                const float4 c = cand[k];
                const float du = __fsub_rn(px.x, c.x), dv = __fsub_rn(px.y, c.y);
                const float dist = __fmaf_rn(du, du, __fmul_rn(dv, dv));  // the association the reference compiles to
                const bool ok = c.w != 0.0f && (base + k) != idx && !(dist > max_pixel_dist) && !(dist >= min_dist);
                min_dist = ok ? dist : min_dist;
                depth = ok ? c.z : depth;
            }
        }
    }
    if (!live) return;
    float3 p = make_float3(0.f, 0.f, -1.0f), c = make_float3(0.f, 0.f, 0.f);
    const bool with_point = has || depth > 0.0f;
    if (with_point) {
        if (has) {
            p = make_float3(point3D[3 * (size_t)idx], point3D[3 * (size_t)idx + 1], point3D[3 * (size_t)idx + 2]);
        } else {
            const int ui = (int)px.x, vi = (int)px.y;
            p.x = __fdiv_rn(__fmul_rn(__fsub_rn((float)ui, cx), depth), fx);
            p.y = __fdiv_rn(__fmul_rn(__fsub_rn((float)vi, cy), depth), fy);
            p.z = depth;
        }
        const long long off = (long long)(int)__fmaf_rn(px.y, (float)width, px.x);
        if (off >= 0 && off + 2 < colors_len) c = make_float3(colors[off], colors[off + 1], colors[off + 2]);
    }
    res_p[3 * (size_t)idx] = p.x; res_p[3 * (size_t)idx + 1] = p.y; res_p[3 * (size_t)idx + 2] = p.z;
    res_c[3 * (size_t)idx] = c.x; res_c[3 * (size_t)idx + 1] = c.y; res_c[3 * (size_t)idx + 2] = c.z;
    keep[idx] = p.z > 0.0f ? 1 : 0;  // stereo_vision.cu:203-206
}

// order-preserving compaction by ONE CTA (keypoints per frame are thousands, not millions): chunks of 1024 rows, ballot +
// warp prefix inside the chunk, a running base across chunks
__global__ void __launch_bounds__(1024)
inactive_geo_compact_kernel(int N, const float* __restrict__ res_p, const float* __restrict__ res_c,
                            const unsigned char* __restrict__ keep, float* __restrict__ out_p, float* __restrict__ out_c,
                            int* __restrict__ count) {
    __shared__ int warp_cnt[32];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int start = 0; start < N; start += 1024) {
        const int i = start + threadIdx.x;
        const bool k = i < N && keep[i] != 0;
        const unsigned b = __ballot_sync(0xffffffffu, k);
        if (lane == 0) warp_cnt[warp] = __popc(b);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w) {
            const int c = warp_cnt[w];
            before += w < warp ? c : 0;
            total += c;
        }
        const int base = s_base;
        if (k) {
            const size_t o = (size_t)(base + before + __popc(b & ((1u << lane) - 1u)));
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                out_p[3 * o + r] = res_p[3 * (size_t)i + r];
                out_c[3 * o + r] = res_c[3 * (size_t)i + r];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base = base + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = s_base;
}

}  // namespace lgs

extern "C" size_t lgs_inactive_geo_scratch_bytes(int N) {
    if (N <= 0) return 0;
    return (size_t)N * 25 + 256;  // [N,3] points, [N,3] colours, [N] flags
}

extern "C" int lgs_inactive_geo_densify(int N, int width, float fx, float fy, float cx, float cy, float max_pixel_dist,
                                        const float* kps_pixel, const unsigned char* kps_has3D, const float* kps_point_local,
                                        const float* colors, long long colors_len, float* out_points, float* out_colors,
                                        int* out_count, char* scratch, void* stream) {
    if (N < 0 || width <= 0) return LGS_ERR_INVALID_ARG;
    if (!out_count) return LGS_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (N == 0) {
        LGS_CUDA_TRY(cudaMemsetAsync(out_count, 0, sizeof(int), s));
        return LGS_OK;
    }
    if (!kps_pixel || !kps_has3D || !kps_point_local || !colors || colors_len < 0 || !out_points || !out_colors || !scratch)
        return LGS_ERR_INVALID_ARG;
    if (reinterpret_cast<uintptr_t>(kps_pixel) & 7u) return LGS_ERR_INVALID_ARG;  // rows are read as float2
    char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(scratch) + 15) & ~(uintptr_t)15);
    float* res_p = reinterpret_cast<float*>(p); p += (size_t)N * 12;
    float* res_c = reinterpret_cast<float*>(p); p += (size_t)N * 12;
    unsigned char* keep = reinterpret_cast<unsigned char*>(p);
    inactive_geo_search_kernel<<<(N + IGD_TILE - 1) / IGD_TILE, IGD_TILE, 0, s>>>(
        N, width, fx, fy, cx, cy, max_pixel_dist, reinterpret_cast<const float2*>(kps_pixel), kps_has3D, kps_point_local, colors,
        colors_len, res_p, res_c, keep);
    LGS_LAUNCH_CHECK();
    inactive_geo_compact_kernel<<<1, 1024, 0, s>>>(N, res_p, res_c, keep, out_points, out_colors, out_count);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}
