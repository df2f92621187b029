// densify.cu -- adaptive density control of the Gaussian set with its optimizer-state surgery, fused
// (SURVEY.md section 8f row 1).  Replaces, for one call of GaussianModel::densifyAndPrune (reference
// src/gaussian_model.cpp:806-824), the ~150 ATen kernels of
//     densifyAndClone :775-804 -> densificationPostfix :653-727      (cat of 7 parameters + 14 Adam moments)
//     densifyAndSplit :729-773 -> densificationPostfix, prunePoints  (index + repeat + cat + index of all 21)
//     prunePoints :597-651                                           (index of all 21 + 4 statistics vectors)
// each of which is a full pass over the model, by
//     densify_classify_kernel   one pass over the per-Gaussian statistics -> what happens to every Gaussian
//     3 + 1 exclusive scans     (CUB, CUDA toolkit library) -> where every survivor / clone / child lands
//     densify_map_kernel        destination -> (parent, kind)
//     densify_gather_kernel     ONE gather of every tensor into its final place
// and GaussianModel::addDensificationStats :834-847 + the max_radii2D update (gaussian_mapper.cpp:739-742) by
//     densify_stats_kernel.
//
// The sequence clone -> split -> prune is a pure function of each ORIGINAL Gaussian (clones and split children
// copy their parent's opacity, clones its scale, children scale / 1.6; the gradient of appended points is
// padded with 0 so they are never split; max_radii2D is zeroed by densificationPostfix before the size test,
// SURVEY.md appendix A), so the final layout is known after one classification pass:
//     [ originals that are neither split nor pruned | surviving clones | surviving children, copy 1 | copy 2 ]
// in the reference's order.  Children are placed at R(q/|q|) * (z * exp(scaling)) + xyz with z the caller's
// standard-normal samples (at::normal(0, stds) = z * stds with the same generator state), row k * n_split + m for
// copy k of the m-th selected Gaussian.
#include <cub/cub.cuh>
#include "common.cuh"

namespace lgs {

constexpr uint8_t DF_KEEP = 1;    // the original survives (not split, not pruned)
constexpr uint8_t DF_CLONE = 2;   // a surviving clone
constexpr uint8_t DF_CHILD = 4;   // two surviving split children
constexpr uint8_t DF_SPLIT = 8;   // selected for splitting (consumes two rows of the samples, pruned or not)

__global__ void __launch_bounds__(256)
densify_stats_kernel(int P, const int* __restrict__ radii, const float* __restrict__ dL_dmeans2D,
                     float* __restrict__ accum, float* __restrict__ denom, float* __restrict__ max_radii) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int r = radii[i];
    if (r <= 0) return;  // visibility_filter = radii > 0 (gaussian_renderer.cpp)
    const float gx = dL_dmeans2D[3 * (size_t)i], gy = dL_dmeans2D[3 * (size_t)i + 1];
    accum[i] += sqrtf(gx * gx + gy * gy);
    denom[i] += 1.0f;
    max_radii[i] = fmaxf(max_radii[i], (float)r);
}

__global__ void __launch_bounds__(256)
densify_classify_kernel(int P, const float* __restrict__ accum, const float* __restrict__ denom,
                        const float* __restrict__ scaling, const float* __restrict__ opacity, float max_grad,
                        float min_opacity, float extent, float percent_dense, int use_size_test,
                        uint8_t* __restrict__ flags, uint32_t* __restrict__ cA, uint32_t* __restrict__ cB,
                        uint32_t* __restrict__ cC, uint32_t* __restrict__ cS) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float g = accum[i] / denom[i];  // :810-811
    if (g != g) g = 0.0f;
    const float s0 = expf(scaling[3 * (size_t)i]), s1 = expf(scaling[3 * (size_t)i + 1]), s2 = expf(scaling[3 * (size_t)i + 2]);
    const float smax = fmaxf(s0, fmaxf(s1, s2));
    const float dense = percent_dense * extent;
    const bool clone = (fabsf(g) >= max_grad) && (smax <= dense);  // :780-784
    const bool split = (g >= max_grad) && (smax > dense);          // :738-742 (appended points have g = 0)
    const float op = 1.0f / (1.0f + expf(-opacity[i]));            // sigmoid
    const bool low = op < min_opacity;                             // :815
    const float big = 0.1f * extent;
    const bool prune_self = low || (use_size_test && smax > big);  // :816-820, max_radii2D_ == 0 here
    // children: new_scaling = log(scale / (0.8 * 2)) :753-754, tested through exp again :818
    const float c0 = expf(logf(s0 / 1.6f)), c1 = expf(logf(s1 / 1.6f)), c2 = expf(logf(s2 / 1.6f));
    const bool prune_child = low || (use_size_test && fmaxf(c0, fmaxf(c1, c2)) > big);
    uint8_t f = 0;
    if (!split && !prune_self) f |= DF_KEEP;
    if (clone && !prune_self) f |= DF_CLONE;
    if (split) f |= DF_SPLIT;
    if (split && !prune_child) f |= DF_CHILD;
    flags[i] = f;
    cA[i] = (f & DF_KEEP) ? 1u : 0u;
    cB[i] = (f & DF_CLONE) ? 1u : 0u;
    cC[i] = (f & DF_CHILD) ? 1u : 0u;
    cS[i] = (f & DF_SPLIT) ? 1u : 0u;
}

// totals[0..3] = nA, nB, nC, nS from the exclusive scans (+ the last flag)
__global__ void densify_totals_kernel(int P, const uint8_t* __restrict__ flags, const uint32_t* __restrict__ oA,
                                      const uint32_t* __restrict__ oB, const uint32_t* __restrict__ oC,
                                      const uint32_t* __restrict__ oS, uint32_t* __restrict__ totals) {
    const uint8_t f = flags[P - 1];
    totals[0] = oA[P - 1] + ((f & DF_KEEP) ? 1u : 0u);
    totals[1] = oB[P - 1] + ((f & DF_CLONE) ? 1u : 0u);
    totals[2] = oC[P - 1] + ((f & DF_CHILD) ? 1u : 0u);
    totals[3] = oS[P - 1] + ((f & DF_SPLIT) ? 1u : 0u);
}

// destination row -> parent index | kind << 30 (kind 0 = the original, 1 = clone, 2 = child copy 1, 3 = child copy 2),
// and for children the row of their normal sample
__global__ void __launch_bounds__(256)
densify_map_kernel(int P, const uint8_t* __restrict__ flags, const uint32_t* __restrict__ oA,
                   const uint32_t* __restrict__ oB, const uint32_t* __restrict__ oC, const uint32_t* __restrict__ oS,
                   uint32_t nA, uint32_t nB, uint32_t nC, uint32_t nS, uint32_t* __restrict__ src,
                   uint32_t* __restrict__ sample_row) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint8_t f = flags[i];
    if (f & DF_KEEP) src[oA[i]] = (uint32_t)i;
    if (f & DF_CLONE) src[nA + oB[i]] = (uint32_t)i | (1u << 30);
    if (f & DF_CHILD) {
        const uint32_t d1 = nA + nB + oC[i], d2 = d1 + nC;
        src[d1] = (uint32_t)i | (2u << 30);
        src[d2] = (uint32_t)i | (3u << 30);
        sample_row[d1] = oS[i];
        sample_row[d2] = nS + oS[i];
    }
}

constexpr int DF_MAX_TENSORS = 24;
enum DensifyMode { DM_COPY = 0, DM_ZERO_NEW = 1, DM_XYZ = 2, DM_SCALING = 3 };
struct DensifyTable {
    const float* src[DF_MAX_TENSORS];
    float* dst[DF_MAX_TENSORS];
    int row[DF_MAX_TENSORS];
    int mode[DF_MAX_TENSORS];
    int n;
};

// One warp per destination row: every tensor's row is copied (or zeroed / transformed) by the warp's lanes.
__global__ void __launch_bounds__(256)
densify_gather_kernel(int n_out, const uint32_t* __restrict__ src_map, const uint32_t* __restrict__ sample_row,
                      const float* __restrict__ scaling, const float* __restrict__ rotation,
                      const float* __restrict__ samples, DensifyTable tab) {
    const int o = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (o >= n_out) return;
    const uint32_t sm = src_map[o];
    const uint32_t p = sm & 0x3fffffffu, kind = sm >> 30;
    for (int t = 0; t < tab.n; ++t) {
        const int row = tab.row[t], mode = tab.mode[t];
        const float* s = tab.src[t] + (size_t)p * row;
        float* d = tab.dst[t] + (size_t)o * row;
        if (mode == DM_ZERO_NEW && kind != 0) {  // Adam moments of appended points start at zero (:690-694)
            for (int k = lane; k < row; k += 32) d[k] = 0.0f;
        } else if (mode == DM_XYZ && kind >= 2) {
            if (lane < 3) {
                // samples = z * stds; new_xyz = R(q / |q|) * samples + xyz   (:744-751, general_utils.h:29-53)
                const size_t sr = sample_row[o];
                const float sx = samples[3 * sr] * expf(scaling[3 * (size_t)p]);
                const float sy = samples[3 * sr + 1] * expf(scaling[3 * (size_t)p + 1]);
                const float sz = samples[3 * sr + 2] * expf(scaling[3 * (size_t)p + 2]);
                const float4 q = reinterpret_cast<const float4*>(rotation)[p];
                const float nrm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
                const float r = q.x / nrm, x = q.y / nrm, y = q.z / nrm, z = q.w / nrm;
                float v;
                if (lane == 0) v = (1.f - 2.f * (y * y + z * z)) * sx + (2.f * (x * y - r * z)) * sy + (2.f * (x * z + r * y)) * sz;
                else if (lane == 1) v = (2.f * (x * y + r * z)) * sx + (1.f - 2.f * (x * x + z * z)) * sy + (2.f * (y * z - r * x)) * sz;
                else v = (2.f * (x * z - r * y)) * sx + (2.f * (y * z + r * x)) * sy + (1.f - 2.f * (x * x + y * y)) * sz;
                d[lane] = v + s[lane];
            }
        } else if (mode == DM_SCALING && kind >= 2) {
            if (lane < 3) d[lane] = logf(expf(s[lane]) / 1.6f);  // scaling_inverse_activation(scale / (0.8 N)), N = 2
        } else {
            for (int k = lane; k < row; k += 32) d[k] = s[k];
        }
    }
}

}  // namespace lgs

using namespace lgs;

extern "C" int lgs_densify_stats(int P, const int* radii, const float* dL_dmeans2D, float* xyz_gradient_accum, float* denom,
                                 float* max_radii2D, void* stream) {
    if (P < 0) return LGS_ERR_INVALID_ARG;
    if (P == 0) return LGS_OK;
    if (!radii || !dL_dmeans2D || !xyz_gradient_accum || !denom || !max_radii2D) return LGS_ERR_INVALID_ARG;
    densify_stats_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, radii, dL_dmeans2D, xyz_gradient_accum, denom,
                                                                           max_radii2D);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

static size_t densify_scan_bytes(int P) {
    size_t n = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, n, (uint32_t*)nullptr, (uint32_t*)nullptr, P);
    return (n + 255) & ~(size_t)255;
}
// plan scratch: flags [P] | 4 count arrays [P] | 4 offset arrays [P] | totals [4] | CUB temp
extern "C" size_t lgs_densify_plan_bytes(int P) {
    if (P <= 0) return 0;
    const size_t n = ((size_t)P + 255) & ~(size_t)255;
    return n + 8 * n * sizeof(uint32_t) + 256 + densify_scan_bytes(P) + 256;
}
namespace {
struct PlanView {
    uint8_t* flags;
    uint32_t *c[4], *o[4], *totals;
    char* temp;
    size_t temp_bytes;
};
PlanView plan_view(char* scratch, int P) {
    PlanView v;
    const size_t n = ((size_t)P + 255) & ~(size_t)255;
    char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~(uintptr_t)255);
    v.flags = reinterpret_cast<uint8_t*>(p);
    p += n;
    for (int k = 0; k < 4; ++k) { v.c[k] = reinterpret_cast<uint32_t*>(p); p += n * sizeof(uint32_t); }
    for (int k = 0; k < 4; ++k) { v.o[k] = reinterpret_cast<uint32_t*>(p); p += n * sizeof(uint32_t); }
    v.totals = reinterpret_cast<uint32_t*>(p);
    p += 256;
    v.temp = p;
    v.temp_bytes = densify_scan_bytes(P);
    return v;
}
}  // namespace

extern "C" int lgs_densify_plan(int P, const float* xyz_gradient_accum, const float* denom, const float* scaling,
                                const float* opacity, float max_grad, float min_opacity, float extent, float percent_dense,
                                int max_screen_size, char* plan_scratch, int* totals_host, void* stream) {
    if (P <= 0 || !xyz_gradient_accum || !denom || !scaling || !opacity || !plan_scratch || !totals_host)
        return LGS_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    // 256 bytes of slack in lgs_densify_plan_bytes cover the base alignment
    PlanView v = plan_view(plan_scratch, P);
    densify_classify_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, xyz_gradient_accum, denom, scaling, opacity, max_grad, min_opacity,
                                                            extent, percent_dense, max_screen_size != 0, v.flags, v.c[0], v.c[1],
                                                            v.c[2], v.c[3]);
    LGS_LAUNCH_CHECK();
    for (int k = 0; k < 4; ++k) {
        size_t n = v.temp_bytes;
        LGS_CUDA_TRY(cub::DeviceScan::ExclusiveSum(v.temp, n, v.c[k], v.o[k], P, s));
    }
    densify_totals_kernel<<<1, 1, 0, s>>>(P, v.flags, v.o[0], v.o[1], v.o[2], v.o[3], v.totals);
    LGS_LAUNCH_CHECK();
    uint32_t h[4] = {0, 0, 0, 0};
    LGS_CUDA_TRY(cudaMemcpyAsync(h, v.totals, sizeof(h), cudaMemcpyDeviceToHost, s));
    LGS_CUDA_TRY(cudaStreamSynchronize(s));  // the caller sizes the new tensors (the reference syncs via .item() :767)
    for (int k = 0; k < 4; ++k) totals_host[k] = (int)h[k];
    return LGS_OK;
}

extern "C" int lgs_densify_apply(int P, const char* plan_scratch, const int* totals, int n_tensors, const float* const* src,
                                 float* const* dst, const int* row_floats, const int* modes, const float* scaling,
                                 const float* rotation, const float* samples, uint32_t* map_scratch, void* stream) {
    if (P <= 0 || !plan_scratch || !totals || n_tensors <= 0 || n_tensors > DF_MAX_TENSORS || !src || !dst || !row_floats ||
        !modes || !scaling || !rotation || !map_scratch)
        return LGS_ERR_INVALID_ARG;
    const uint32_t nA = (uint32_t)totals[0], nB = (uint32_t)totals[1], nC = (uint32_t)totals[2], nS = (uint32_t)totals[3];
    const size_t n_out = (size_t)nA + nB + 2 * (size_t)nC;
    if (n_out == 0) return LGS_OK;
    if (nC > 0 && !samples) return LGS_ERR_INVALID_ARG;
    if (reinterpret_cast<uintptr_t>(rotation) & 15u) return LGS_ERR_ALIGNMENT;
    cudaStream_t s = (cudaStream_t)stream;
    PlanView v = plan_view(const_cast<char*>(plan_scratch), P);
    uint32_t* src_map = map_scratch;
    uint32_t* sample_row = map_scratch + n_out;
    densify_map_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, v.flags, v.o[0], v.o[1], v.o[2], v.o[3], nA, nB, nC, nS, src_map,
                                                       sample_row);
    LGS_LAUNCH_CHECK();
    DensifyTable tab;
    tab.n = n_tensors;
    for (int t = 0; t < n_tensors; ++t) {
        if (!src[t] || !dst[t] || row_floats[t] <= 0 || modes[t] < 0 || modes[t] > DM_SCALING) return LGS_ERR_INVALID_ARG;
        tab.src[t] = src[t];
        tab.dst[t] = dst[t];
        tab.row[t] = row_floats[t];
        tab.mode[t] = modes[t];
    }
    const size_t threads = n_out * 32;
    densify_gather_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>((int)n_out, src_map, sample_row, scaling, rotation,
                                                                            samples, tab);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}
