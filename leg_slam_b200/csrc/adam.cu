// adam.cu -- fused multi-tensor Adam for sm_100a: ONE launch updates all Gaussian parameter
// tensors (xyz, f_dc, f_rest, language feature, opacity, scaling, rotation).
//
// Replaces torch::optim::Adam::step over the reference's 7 single-tensor parameter groups
// (reference src/gaussian_model.cpp:483-518, step at src/gaussian_mapper.cpp:793-796):
// libtorch 2.0.1 runs, per tensor, mul_ / add_ / mul_ / addcmul_ / sqrt / div / add_ / addcdiv_
// (8 elementwise kernels, ~10 passes over memory).  Here every element is read once
// (p, g, m, v) and written once (p, m, v): 28 B/element, the algorithmic minimum.
//
// Arithmetic follows libtorch's op sequence (and its FMA contractions) so the update agrees
// with the reference to the last bits:
//   m  = fma(1-b1, g, b1*m)                       exp_avg.mul_(b1).add_(g, 1-b1)
//   v  = fma((1-b2)*g, g, b2*v)                   exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
//   dn = fma(sqrt(v), 1/sqrt(1-b2^t), eps)        (sqrt / scalar -> * reciprocal).add_(eps)
//   p  = fma(-(lr/(1-b1^t)), m / dn, p)           addcdiv_(m, dn, -step_size)
// bias corrections are computed on the host in double like libtorch, then cast to float.
//
// Memory-bound (HBM roofline): grid-stride over 16-byte vectors, streaming loads/stores.
#include <cmath>
#include "common.cuh"
#include "adam_math.cuh"

namespace lgs {

constexpr int ADAM_MAX_TENSORS = 16;
constexpr int ADAM_CHUNK = 4096;  // elements per CTA-chunk (256 threads x 4 float4)

struct AdamTable {
    float* p[ADAM_MAX_TENSORS];
    const float* g[ADAM_MAX_TENSORS];
    float* m[ADAM_MAX_TENSORS];
    float* v[ADAM_MAX_TENSORS];
    long long n[ADAM_MAX_TENSORS];
    float neg_step[ADAM_MAX_TENSORS];  // -(lr / bias_correction1)
    int chunk_start[ADAM_MAX_TENSORS + 1];
    int n_tensors;
};

__device__ __forceinline__ float4 ldcs4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }

__global__ void __launch_bounds__(256)
adam_multi_kernel(const __grid_constant__ AdamTable tab, int total_chunks, float b1, float omb1, float b2,
                  float omb2, float inv_bc2_sqrt, float eps) {
    for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
        int t = 0;
#pragma unroll 1
        while (t + 1 < tab.n_tensors && chunk >= tab.chunk_start[t + 1]) ++t;
        const long long base = (long long)(chunk - tab.chunk_start[t]) * ADAM_CHUNK;
        const long long n = tab.n[t];
        float* __restrict__ P = tab.p[t];
        const float* __restrict__ G = tab.g[t];
        float* __restrict__ M = tab.m[t];
        float* __restrict__ V = tab.v[t];
        const float neg_step = tab.neg_step[t];
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(G) |
                              reinterpret_cast<uintptr_t>(M) | reinterpret_cast<uintptr_t>(V)) & 15u) == 0;
        if (vec_ok && base + ADAM_CHUNK <= n) {
            float4 p4[4], g4[4], m4[4], v4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // all loads first: 16 x 16 B in flight per thread
                const long long i = base + 4 * (threadIdx.x + 256 * k);
                p4[k] = ldcs4(P + i); g4[k] = ldcs4(G + i); m4[k] = ldcs4(M + i); v4[k] = ldcs4(V + i);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                adam_elem(p4[k].x, g4[k].x, m4[k].x, v4[k].x, b1, omb1, b2, omb2, inv_bc2_sqrt, eps, neg_step);
                adam_elem(p4[k].y, g4[k].y, m4[k].y, v4[k].y, b1, omb1, b2, omb2, inv_bc2_sqrt, eps, neg_step);
                adam_elem(p4[k].z, g4[k].z, m4[k].z, v4[k].z, b1, omb1, b2, omb2, inv_bc2_sqrt, eps, neg_step);
                adam_elem(p4[k].w, g4[k].w, m4[k].w, v4[k].w, b1, omb1, b2, omb2, inv_bc2_sqrt, eps, neg_step);
                const long long i = base + 4 * (threadIdx.x + 256 * k);
                *reinterpret_cast<float4*>(P + i) = p4[k];
                __stcs(reinterpret_cast<float4*>(M + i), m4[k]);
                __stcs(reinterpret_cast<float4*>(V + i), v4[k]);
            }
        } else {
            const long long end = (base + ADAM_CHUNK < n) ? base + ADAM_CHUNK : n;
            for (long long i = base + threadIdx.x; i < end; i += 256) {
                float p = P[i], m = M[i], v = V[i];
                adam_elem(p, G[i], m, v, b1, omb1, b2, omb2, inv_bc2_sqrt, eps, neg_step);
                P[i] = p; M[i] = m; V[i] = v;
            }
        }
    }
}

}  // namespace lgs

using namespace lgs;

extern "C" int lgs_adam_multi(int n_tensors, float* const* params, const float* const* grads,
                              float* const* exp_avg, float* const* exp_avg_sq, const int64_t* numel,
                              const double* lr, double beta1, double beta2, double eps, int step, void* stream) {
    if (n_tensors < 0 || n_tensors > ADAM_MAX_TENSORS || step < 1) return LGS_ERR_INVALID_ARG;
    if (n_tensors == 0) return LGS_OK;
    if (!params || !grads || !exp_avg || !exp_avg_sq || !numel || !lr) return LGS_ERR_INVALID_ARG;
    AdamTable tab;
    // libtorch: bias_correction1 = 1 - pow(beta1, step) etc. in double (Adam.cpp of libtorch 2.0.1)
    const double bc1 = 1.0 - std::pow(beta1, (double)step);
    const double bc2 = 1.0 - std::pow(beta2, (double)step);
    const double bc2_sqrt = std::sqrt(bc2);
    long long chunks = 0;
    int nt = 0;
    for (int t = 0; t < n_tensors; ++t) {
        if (numel[t] < 0) return LGS_ERR_INVALID_ARG;
        if (numel[t] == 0) continue;
        if (!params[t] || !grads[t] || !exp_avg[t] || !exp_avg_sq[t]) return LGS_ERR_INVALID_ARG;
        tab.p[nt] = params[t]; tab.g[nt] = grads[t]; tab.m[nt] = exp_avg[t]; tab.v[nt] = exp_avg_sq[t];
        tab.n[nt] = numel[t];
        tab.neg_step[nt] = (float)(-(lr[t] / bc1));
        tab.chunk_start[nt] = (int)chunks;
        chunks += (numel[t] + ADAM_CHUNK - 1) / ADAM_CHUNK;
        if (chunks > 0x7fffffffLL) return LGS_ERR_INVALID_ARG;
        ++nt;
    }
    tab.chunk_start[nt] = (int)chunks;
    tab.n_tensors = nt;
    if (nt == 0) return LGS_OK;
    // scalars reach the ATen kernels as double and are cast to float there
    const float b1 = (float)beta1, b2 = (float)beta2;
    const float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
    const float inv_bc2_sqrt = 1.0f / (float)bc2_sqrt;
    const int grid = (int)(chunks < 148LL * 32 ? chunks : 148LL * 32);
    adam_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tab, (int)chunks, b1, omb1, b2, omb2, inv_bc2_sqrt,
                                                              (float)eps);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}
