// query_tc.cu -- semantic query on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// sim[p, q] = <f_p/|f_p|, t_q/|t_q|> for P Gaussians x Q text embeddings, K = 64 (reference
// eval/find_objects_gaussians.py:160-172).  This is the one GEMM-shaped piece of the path (north_star:
// "the only place tensor cores are used"); at cfgE (2M x 256) it is bound by the 2 GB fp32 output, not by
// math, so the design goal is to keep the HBM write stream busy:
//
//   * persistent CTAs (one per SM), 128-row tiles.  The normalised text matrix (N <= 256 columns) is split
//     once into TF32 hi + lo parts and stays resident in shared memory in the tensor core's canonical
//     K-major layout (8x16-byte core matrices, no swizzle).
//   * per tile, warps 0-3 normalise 128 feature rows, split them into TF32 hi/lo and store them in the same
//     canonical layout; one elected thread issues 24 tcgen05.mma (kind::tf32, M=128, N<=256, K=8):
//     hi*hi + hi*lo + lo*hi ("3xTF32") so the result keeps fp32-level accuracy (~1e-6), accumulating in TMEM.
//   * accumulators are double-buffered in TMEM (2 x N columns): warps 4-7 drain tile i with tcgen05.ld and
//     stream it to HBM while warps 0-3 prepare and the tensor core computes tile i+1.  Hand-offs are
//     mbarriers: tcgen05.commit -> mma_done[b] (frees the A tile, wakes the epilogue), epilogue -> acc_free[b].
//
// The SIMT kernel in query.cu remains for shapes this one does not take (Q chunk not a multiple of 4).
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"
#include "tc.cuh"

namespace lgs {

constexpr int QT_M = 128;
constexpr int QT_K = 64;
constexpr int QT_THREADS = 256;
constexpr int EPI_PITCH = 36;  // floats per staged row: 144 B keeps 16-byte alignment and spreads banks
constexpr int EPI_BYTES = 4 * 32 * EPI_PITCH * 4;

// one 64-float row -> registers (streaming loads; zeros for rows past the end)
__device__ __forceinline__ void load_row(const float* __restrict__ src, bool valid, float4 (&v)[16]) {
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = valid ? __ldcs(reinterpret_cast<const float4*>(src) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
}
// normalise the row (F.normalize, eps 1e-12), split into TF32 hi / lo, store both in canonical layout
__device__ __forceinline__ void store_row(const float4 (&v)[16], int r, int R, uint8_t* hi, uint8_t* lo) {
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) ss += v[c].x * v[c].x + v[c].y * v[c].y + v[c].z * v[c].z + v[c].w * v[c].w;
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float x0 = v[c].x * inv, x1 = v[c].y * inv, x2 = v[c].z * inv, x3 = v[c].w * inv;
        uint4 h, l;
        h.x = to_tf32(x0); h.y = to_tf32(x1); h.z = to_tf32(x2); h.w = to_tf32(x3);
        l.x = to_tf32(x0 - __uint_as_float(h.x)); l.y = to_tf32(x1 - __uint_as_float(h.y));
        l.z = to_tf32(x2 - __uint_as_float(h.z)); l.w = to_tf32(x3 - __uint_as_float(h.w));
        const uint32_t off = canon_off(r, c, R);
        *reinterpret_cast<uint4*>(hi + off) = h;
        *reinterpret_cast<uint4*>(lo + off) = l;
    }
}
__device__ __forceinline__ void stage_row(const float* __restrict__ src, bool valid, int r, int R, uint8_t* hi, uint8_t* lo) {
    float4 v[16];
    load_row(src, valid, v);
    store_row(v, r, R, hi, lo);
}

__global__ void __launch_bounds__(QT_THREADS, 1)
cosine_tc_kernel(int P, int Q, int q0, int Qn, int N, int tmem_cols, int terms, const float* __restrict__ feats,
                 const float* __restrict__ text, float* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* A_hi = smem;
    uint8_t* A_lo = smem + QT_M * QT_K * 4;
    uint8_t* B_hi = smem + 2 * QT_M * QT_K * 4;
    uint8_t* B_lo = B_hi + (size_t)N * QT_K * 4;
    float* epi_stage = reinterpret_cast<float*>(B_lo + (size_t)N * QT_K * 4);  // 4 warps x [32][EPI_PITCH] floats
    __shared__ __align__(8) uint64_t mma_done[2], acc_free[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&mma_done[0], 1);
        mbar_init(&mma_done[1], 1);
        mbar_init(&acc_free[0], 128);
        mbar_init(&acc_free[1], 128);
        mbar_fence_init();
    }
    if (warp == 0) {  // one warp allocates tensor memory for the whole CTA
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // resident B operand: normalised text rows q0 .. q0+N (zero rows beyond Qn)
    for (int q = tid; q < N; q += QT_THREADS) stage_row(text + (size_t)(q0 + q) * QT_K, q < Qn, q, N, B_hi, B_lo);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 [4,6), A = B = TF32 [7,10) [10,13), both K-major,
    // N>>3 [17,23), M>>4 [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(QT_M >> 4) << 24);
    const int ntiles = (P + QT_M - 1) / QT_M;

    if (warp < 4) {
        // ---------------- feature-tile staging (128 threads = 128 rows) + MMA issue (thread 0) ----------------
        int it = 0;
        float4 rowv[16];  // the NEXT tile's row, fetched while the tensor core and the epilogue work on this one
        {
            const long long row = (long long)blockIdx.x * QT_M + tid;
            load_row(feats + (size_t)row * QT_K, blockIdx.x < ntiles && row < P, rowv);
        }
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            if (it > 0) mbar_wait(&mma_done[(it - 1) & 1], (uint32_t)(((it - 1) >> 1) & 1));  // tensor core done reading A
            store_row(rowv, tid, QT_M, A_hi, A_lo);
            {
                const long long nrow = (long long)(tile + gridDim.x) * QT_M + tid;
                load_row(feats + (size_t)nrow * QT_K, tile + (int)gridDim.x < ntiles && nrow < P, rowv);
            }
            fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tid == 0) {
                if (it >= 2) mbar_wait(&acc_free[it & 1], (uint32_t)(((it >> 1) - 1) & 1));  // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem + (uint32_t)((it & 1) * N);
                const uint32_t a_hi = smem_u32(A_hi), a_lo = smem_u32(A_lo), b_hi = smem_u32(B_hi), b_lo = smem_u32(B_lo);
#pragma unroll
                for (int ks = 0; ks < QT_K / 8; ++ks) {  // K = 8 per instruction = two 16-byte chunks
                    const uint32_t ao = (uint32_t)(2 * ks * QT_M * 16), bo = (uint32_t)(2 * ks * N * 16);
                    const uint64_t dah = make_desc(a_hi + ao, QT_M * 16, 128), dal = make_desc(a_lo + ao, QT_M * 16, 128);
                    const uint64_t dbh = make_desc(b_hi + bo, (uint32_t)N * 16, 128), dbl = make_desc(b_lo + bo, (uint32_t)N * 16, 128);
                    umma_tf32(d, dah, dbh, idesc, ks > 0 ? 1u : 0u);
                    if (terms >= 3) {
                        umma_tf32(d, dah, dbl, idesc, 1u);
                        umma_tf32(d, dal, dbh, idesc, 1u);
                    }
                }
                umma_commit(&mma_done[it & 1]);  // arrives when the 24 MMAs have completed
            }
        }
    } else {
        // ---------------- epilogue: TMEM -> registers -> HBM (warp w reads TMEM lanes 32*(w%4) .. +31) ----------------
        const int q4 = warp & 3;
        const bool vec_ok = ((Q & 3) == 0) && ((q0 & 3) == 0);
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            mbar_wait(&mma_done[it & 1], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const long long row0 = (long long)tile * QT_M + q4 * 32;  // first row of this warp's 32
            float* const stg = epi_stage + (size_t)q4 * 32 * EPI_PITCH;     // this warp's [32 rows][32 cols] transpose buffer
            // chunk c (32 columns): its TMEM load was issued while chunk c-1 was being written out
            auto drain = [&](int c0, const uint32_t (&v)[32]) {
                // TMEM hands each lane one ROW (32 consecutive columns).  Writing that straight out would make every
                // store instruction touch 32 different rows (32 half-filled sectors); transpose through shared memory
                // so that 8 consecutive lanes write one row's 128 contiguous bytes.
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    *reinterpret_cast<uint4*>(stg + lane * EPI_PITCH + 4 * k) = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                __syncwarp();
                const int rs = lane >> 3, c4 = (lane & 7) * 4;
                float4 x[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) x[k] = *reinterpret_cast<const float4*>(stg + (4 * k + rs) * EPI_PITCH + c4);
                float* o = out + (size_t)(row0 + rs) * Q + q0 + c0 + c4;
                const bool full = vec_ok && c0 + c4 + 4 <= Qn;
#pragma unroll
                for (int k = 0; k < 8; ++k, o += (size_t)4 * Q) {
                    if (row0 + 4 * k + rs < P) {
                        if (full) {
                            __stcs(reinterpret_cast<float4*>(o), x[k]);
                        } else {
                            if (c0 + c4 + 0 < Qn) o[0] = x[k].x;
                            if (c0 + c4 + 1 < Qn) o[1] = x[k].y;
                            if (c0 + c4 + 2 < Qn) o[2] = x[k].z;
                            if (c0 + c4 + 3 < Qn) o[3] = x[k].w;
                        }
                    }
                }
            };
            const uint32_t tbase = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)((it & 1) * N);
            uint32_t va[32], vb[32];
            tmem_ld32(tbase, va);
            for (int c0 = 0; c0 < N; c0 += 64) {
                tmem_ld_wait();
                if (c0 + 32 < N) tmem_ld32(tbase + (uint32_t)(c0 + 32), vb);
                drain(c0, va);
                if (c0 + 32 < N) {
                    tmem_ld_wait();
                    if (c0 + 64 < N) tmem_ld32(tbase + (uint32_t)(c0 + 64), va);
                    drain(c0 + 32, vb);
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_free[it & 1]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
    }
}

int launch_cosine_tc(int P, int Q, const float* feats, const float* text, float* out, cudaStream_t s) {
    // per-device facts, looked up on every call (function attributes and the SM count belong to the current device)
    int dev = 0, n_sm = 0;
    LGS_CUDA_TRY(cudaGetDevice(&dev));
    LGS_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    LGS_CUDA_TRY(cudaFuncSetAttribute(cosine_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      2 * QT_M * QT_K * 4 + 2 * 256 * QT_K * 4 + EPI_BYTES));
    const int ntiles = (P + QT_M - 1) / QT_M;
    const int terms = 3;  // hi*hi + hi*lo + lo*hi: fp32-level accuracy (a one-term plain-TF32 mode existed as a timing
                          // experiment behind an environment variable; removed -- the product has one precision)
    for (int q0 = 0; q0 < Q; q0 += 256) {
        const int Qn = Q - q0 < 256 ? Q - q0 : 256;
        const int N = (Qn + 15) & ~15;  // M = 128 needs N % 16 == 0
        int cols = 32;
        while (cols < 2 * N + 32 && cols < 512) cols <<= 1;  // power of two >= 32; the last 32-column read may overhang N
        const size_t smem = 2 * QT_M * QT_K * 4 + 2 * (size_t)N * QT_K * 4 + EPI_BYTES;
        cosine_tc_kernel<<<ntiles < n_sm ? ntiles : n_sm, QT_THREADS, smem, s>>>(P, Q, q0, Qn, N, cols, terms, feats, text, out);
        LGS_LAUNCH_CHECK();
    }
    return LGS_OK;
}

}  // namespace lgs
