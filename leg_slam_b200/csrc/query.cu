// query.cu -- semantic query for sm_100a: cosine similarity of every Gaussian's 64-D language
// feature against a batch of text embeddings, and the reference's min-max inversion.
//
// Replaces the torch sequence of reference eval/find_objects_gaussians.py:160-175
//   F.normalize(lf, dim=1) ; F.normalize(text) ; matmul ; 1 - (s-min)/(max-min)
// (3-5 passes over [P,64] plus the [P,Q] output) by ONE pass: each CTA normalises a tile of
// 64 feature rows into shared memory, every thread owns one query column (its normalised text
// vector lives in 64 registers) and streams the tile's rows from shared memory as broadcast
// float4 loads; the [P,Q] output is written once, coalesced along Q.
//
// Bound: the [P,Q] fp32 output write (HBM) for Q >= ~64; FP32-FMA otherwise.
#include <cfloat>
#include "common.cuh"

namespace lgs {

constexpr int QROWS = 64;    // feature rows per CTA tile
constexpr int QTHREADS = 256;

__global__ void __launch_bounds__(QTHREADS)
cosine_query_kernel(int P, int Q, const float* __restrict__ feats, const float* __restrict__ text,
                    float* __restrict__ out) {
    __shared__ __align__(16) float sF[QROWS][LF];
    const int tid = threadIdx.x;
    for (int qbase = 0; qbase < Q; qbase += QTHREADS) {
        const int q = qbase + tid;
        // this thread's query, normalised like F.normalize (x / max(||x||, 1e-12))
        float t[LF];
        if (q < Q) {
            float ss = 0.f;
#pragma unroll
            for (int k = 0; k < LF / 4; ++k) {
                const float4 v = reinterpret_cast<const float4*>(text + (size_t)q * LF)[k];
                t[4 * k] = v.x; t[4 * k + 1] = v.y; t[4 * k + 2] = v.z; t[4 * k + 3] = v.w;
                ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
            const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
            for (int k = 0; k < LF; ++k) t[k] *= inv;
        } else {
#pragma unroll
            for (int k = 0; k < LF; ++k) t[k] = 0.f;
        }
        for (long long row0 = (long long)blockIdx.x * QROWS; row0 < P; row0 += (long long)gridDim.x * QROWS) {
            __syncthreads();  // previous tile fully consumed
            // stage + normalise 64 rows: 4 threads per row, 16 floats each
            {
                const int r = tid >> 2, part = tid & 3;
                const long long row = row0 + r;
                float4 v[4];
                float ss = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    v[k] = row < P ? __ldcs(reinterpret_cast<const float4*>(feats + (size_t)row * LF) + part * 4 + k)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                    ss += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
                }
                ss += __shfl_xor_sync(0xffffffffu, ss, 1);
                ss += __shfl_xor_sync(0xffffffffu, ss, 2);
                const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    v[k].x *= inv; v[k].y *= inv; v[k].z *= inv; v[k].w *= inv;
                    reinterpret_cast<float4*>(&sF[r][0])[part * 4 + k] = v[k];
                }
            }
            __syncthreads();
            if (q < Q) {
                const int rows = (int)((P - row0) < QROWS ? (P - row0) : QROWS);
#pragma unroll 2
                for (int r = 0; r < rows; ++r) {
                    const float4* f4 = reinterpret_cast<const float4*>(&sF[r][0]);
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                    for (int k = 0; k < LF / 4; ++k) {
                        const float4 f = f4[k];
                        a0 = fmaf(f.x, t[4 * k + 0], a0);
                        a1 = fmaf(f.y, t[4 * k + 1], a1);
                        a2 = fmaf(f.z, t[4 * k + 2], a2);
                        a3 = fmaf(f.w, t[4 * k + 3], a3);
                    }
                    __stcs(out + (size_t)(row0 + r) * Q + q, (a0 + a1) + (a2 + a3));
                }
            }
        }
        __syncthreads();
    }
}

// ---- 1 - (s-min)/(max-min) over one score vector ---------------------------------------------
__device__ __forceinline__ unsigned f2ord(float f) {  // order-preserving float -> uint
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void minmax_init_kernel(unsigned* scratch) {
    scratch[0] = 0xffffffffu;  // min
    scratch[1] = 0u;           // max
}

__global__ void __launch_bounds__(256)
minmax_reduce_kernel(long long n, const float* __restrict__ s, unsigned* __restrict__ scratch) {
    float lo = FLT_MAX, hi = -FLT_MAX;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = s[i];
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&scratch[0], f2ord(lo));
        atomicMax(&scratch[1], f2ord(hi));
    }
}

__global__ void __launch_bounds__(256)
minmax_apply_kernel(long long n, float* __restrict__ s, const unsigned* __restrict__ scratch) {
    const float lo = ord2f(scratch[0]), hi = ord2f(scratch[1]);
    const float range = hi - lo;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s[i] = 1.0f - (s[i] - lo) / range;
}

// ---- per-pixel query over a rendered feature image (reference eval/find_objects_gaussians.py:323) -------------------
// dist[q][px] = F.cosine_similarity(rendered_lf [64,H,W], text_q [64,1,1], dim=0): dot / (max(|lf_px|, 1e-8) * max(|t_q|, 1e-8)).
// Planar input, one pixel per thread, the 64 channel loads of a warp are 64 coalesced 128-byte rows; up to CI_Q text vectors
// share one pass over the image (78.6 MB at 640x480).  HBM-bound: 256 B read + 4 Q B written per pixel.
constexpr int CI_Q = 8;
__global__ void __launch_bounds__(256)
cosine_image_kernel(long long HW, int Q, const float* __restrict__ image, const float* __restrict__ text, float* __restrict__ out) {
    __shared__ float s_t[CI_Q][LF];
    __shared__ float s_inv[CI_Q];
    for (int i = threadIdx.x; i < Q * LF; i += blockDim.x) s_t[i / LF][i % LF] = text[i];
    __syncthreads();
    if (threadIdx.x < Q) {
        float n2 = 0.f;
        for (int c = 0; c < LF; ++c) n2 = fmaf(s_t[threadIdx.x][c], s_t[threadIdx.x][c], n2);
        s_inv[threadIdx.x] = 1.0f / fmaxf(sqrtf(n2), 1e-8f);
    }
    __syncthreads();
    for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < HW; px += (long long)gridDim.x * blockDim.x) {
        float dot[CI_Q];
#pragma unroll
        for (int q = 0; q < CI_Q; ++q) dot[q] = 0.f;
        float n2 = 0.f;
#pragma unroll 8
        for (int c = 0; c < LF; ++c) {
            const float v = __ldcs(image + (size_t)c * HW + px);
            n2 = fmaf(v, v, n2);
#pragma unroll
            for (int q = 0; q < CI_Q; ++q)
                if (q < Q) dot[q] = fmaf(v, s_t[q][c], dot[q]);
        }
        const float inv = 1.0f / fmaxf(sqrtf(n2), 1e-8f);
#pragma unroll
        for (int q = 0; q < CI_Q; ++q)
            if (q < Q) out[(size_t)q * HW + px] = dot[q] * inv * s_inv[q];
    }
}

// ---- heat colours for the "query, then heat-map render" path (BASELINE.json configs[4]) --------------------------------
// colours[p] = ramp(scores[p * stride]): blue (0) -> red (1), the colors_precomp input of a forward without SH / features.
__global__ void __launch_bounds__(256)
heat_colors_kernel(int P, const float* __restrict__ scores, int stride, float* __restrict__ colors) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float s = fminf(fmaxf(scores[(size_t)i * stride], 0.f), 1.f);
    colors[3 * i + 0] = s;
    colors[3 * i + 1] = 1.0f - fabsf(2.0f * s - 1.0f);
    colors[3 * i + 2] = 1.0f - s;
}

}  // namespace lgs

using namespace lgs;

static int query_args_ok(int P, int Q, const float* feats, const float* text, float* out) {
    if (P < 0 || Q < 0) return LGS_ERR_INVALID_ARG;
    if (P == 0 || Q == 0) return -1;  // nothing to do
    if (!feats || !text || !out) return LGS_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(feats) & 15u) || (reinterpret_cast<uintptr_t>(text) & 15u)) return LGS_ERR_ALIGNMENT;
    return LGS_OK;
}

namespace lgs {
int launch_cosine_tc(int P, int Q, const float* feats, const float* text, float* out, cudaStream_t s);
}

// tensor-core path (query_tc.cu: tcgen05 3xTF32, TMEM accumulators)
extern "C" int lgs_cosine_query(int P, int Q, const float* feats, const float* text, float* out, void* stream) {
    const int st = query_args_ok(P, Q, feats, text, out);
    if (st != LGS_OK) return st < 0 ? LGS_OK : st;
    if (reinterpret_cast<uintptr_t>(out) & 15u) return LGS_ERR_ALIGNMENT;
    return launch_cosine_tc(P, Q, feats, text, out, (cudaStream_t)stream);
}

// SIMT fp32 path (round-1 kernel; kept as the cross-check of the tensor-core one)
extern "C" int lgs_cosine_query_simt(int P, int Q, const float* feats, const float* text, float* out, void* stream) {
    const int st = query_args_ok(P, Q, feats, text, out);
    if (st != LGS_OK) return st < 0 ? LGS_OK : st;
    const long long tiles = ((long long)P + QROWS - 1) / QROWS;
    const int grid = (int)(tiles < 148LL * 8 ? tiles : 148LL * 8);
    cosine_query_kernel<<<grid, QTHREADS, 0, (cudaStream_t)stream>>>(P, Q, feats, text, out);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

extern "C" int lgs_minmax_invert(int64_t n, float* scores, float* scratch2, void* stream) {
    if (n < 0) return LGS_ERR_INVALID_ARG;
    if (n == 0) return LGS_OK;
    if (!scores || !scratch2) return LGS_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned* sc = reinterpret_cast<unsigned*>(scratch2);
    const int grid = (int)((n + 255) / 256 < 148LL * 8 ? (n + 255) / 256 : 148LL * 8);
    minmax_init_kernel<<<1, 1, 0, s>>>(sc);
    minmax_reduce_kernel<<<grid, 256, 0, s>>>(n, scores, sc);
    minmax_apply_kernel<<<grid, 256, 0, s>>>(n, scores, sc);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

// per-pixel cosine similarity of a planar [64,H,W] feature image against Q text vectors -> [Q,H,W]
extern "C" int lgs_cosine_image(int64_t HW, int Q, const float* image, const float* text, float* out, void* stream) {
    if (HW < 0 || Q < 0) return LGS_ERR_INVALID_ARG;
    if (HW == 0 || Q == 0) return LGS_OK;
    if (!image || !text || !out) return LGS_ERR_INVALID_ARG;
    const long long blocks = (HW + 255) / 256;
    const int grid = (int)(blocks < 148LL * 8 ? blocks : 148LL * 8);
    for (int q0 = 0; q0 < Q; q0 += CI_Q) {
        const int qn = Q - q0 < CI_Q ? Q - q0 : CI_Q;
        cosine_image_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(HW, qn, image, text + (size_t)q0 * LF, out + (size_t)q0 * HW);
        LGS_LAUNCH_CHECK();
    }
    return LGS_OK;
}

extern "C" int lgs_heat_colors(int P, const float* scores, int stride, float* colors, void* stream) {
    if (P < 0 || stride < 1) return LGS_ERR_INVALID_ARG;
    if (P == 0) return LGS_OK;
    if (!scores || !colors) return LGS_ERR_INVALID_ARG;
    heat_colors_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, scores, stride, colors);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}
