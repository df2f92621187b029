// dp_adam.cu -- data-parallel gradient exchange fused with the optimizer, over NVLink peer memory.
//
// New relative to the reference (single GPU).  After every rank has back-propagated its views into its
// flat gradient buffer [123*P], the baseline is  ncclAllReduce(grads) ; Adam on every replica  (SURVEY.md
// 8e).  Here ONE kernel per rank does reduce-scatter + Adam + all-gather on its shard of the flat index
// space, with no NCCL call and no intermediate buffer:
//     g  = sum over ranks of grads_r[i]          multimem.ld_reduce on the NVSwitch multicast address (the
//                                                switch adds), or G peer loads in fixed rank order
//     p, m, v <- Adam(p, g, m, v)                m, v exist only for the owned shard (optimizer state / G)
//     params_r[i] = p  for every rank r          multimem.st (switch broadcast), or G peer stores
// Every element is reduced and updated by exactly one rank and broadcast, so replicas stay bit-identical
// by construction.  Per rank it moves (G-1)/G of the gradient bytes in and of the parameter bytes out --
// the same wire traffic as reduce-scatter + all-gather -- while Adam's HBM traffic drops by G.
// The buffers must be symmetric allocations (torch.distributed._symmetric_memory); the caller brackets the
// launch with cross-rank barriers (gradients complete before, parameters landed after).
#include <cmath>
#include "common.cuh"
#include "adam_math.cuh"

namespace lgs {

constexpr int DP_MAX_SEG = 16;
constexpr int DP_MAX_WORLD = 16;

struct DpTable {
    long long seg_start[DP_MAX_SEG + 1];  // flat offsets of the parameter tensors (multiples of 4)
    float neg_step[DP_MAX_SEG];           // -(lr / bias_correction1) per tensor
    int row_len[DP_MAX_SEG];              // floats per Gaussian of each tensor (sparse exchange only)
    const float* grads[DP_MAX_WORLD];     // every rank's flat gradient buffer (peer pointers)
    float* params[DP_MAX_WORLD];          // every rank's flat parameter buffer (peer pointers)
    int n_seg, world;
    long long n_rows;                     // Gaussians (sparse exchange only)
};

__device__ __forceinline__ float4 mc_ld_reduce_add(const float* mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
    return r;
}
__device__ __forceinline__ void mc_st(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// Four 16-byte vectors per thread and iteration, every load issued before the first use: the peer / multicast loads
// have NVLink latency (microseconds), so the number of requests in flight per SM decides the achieved link bandwidth.
// WORLD > 0 unrolls the peer loop (2, 4, 8 ranks); WORLD = 0 is the generic loop.
constexpr int DP_UNROLL = 4;

// SPARSE (P2P path only): row_mask[gaussian] has bit r set when rank r's gradient row of that Gaussian may be non-zero --
// a Gaussian culled in every view a rank rendered has an exactly zero (+0.0) row there (60 % of the rows with one view per
// rank at cfgB).  A rank's copy of a 16-byte vector is loaded only if one of the (at most four) Gaussians the vector touches
// has its bit set; adding the +0.0 it would have loaded changes nothing, so results are bit-identical to the dense exchange.
template <bool MC, int WORLD, bool SPARSE>
__global__ void __launch_bounds__(256)
dp_adam_kernel(const __grid_constant__ DpTable tab, const float* __restrict__ grads_mc, float* __restrict__ params_mc,
               long long begin, long long end, float* __restrict__ p_local, float* __restrict__ m, float* __restrict__ v,
               float b1, float omb1, float b2, float omb2, float inv_bc2_sqrt, float eps, const uint16_t* __restrict__ row_mask) {
    const long long n4 = (end - begin) >> 2;
    const int world = WORLD > 0 ? WORLD : tab.world;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; k0 < n4; k0 += stride * DP_UNROLL) {
        float4 g[DP_UNROLL], p[DP_UNROLL], mm[DP_UNROLL], vv[DP_UNROLL];
        uint32_t need[DP_UNROLL];
#pragma unroll
        for (int u = 0; u < DP_UNROLL; ++u) {
            need[u] = 0xffffffffu;
            const long long k = k0 + u * stride;
            if (k >= n4) continue;
            const long long i = begin + 4 * k;
            if (SPARSE) {
                int t = 0;
                while (t + 1 < tab.n_seg && i >= tab.seg_start[t + 1]) ++t;
                const long long local = i - tab.seg_start[t];
                const int len = tab.row_len[t];
                const long long g0 = local / len, g1 = (local + 3) / len;
                uint32_t mk = 0;
                for (long long gi = g0; gi <= g1 && gi < tab.n_rows; ++gi) mk |= row_mask[gi];  // past the last row: padding, all zero
                need[u] = mk;
            }
            if (MC) {
                g[u] = mc_ld_reduce_add(grads_mc + i);  // in-switch reduction over all ranks
            } else {
                g[u] = (need[u] & 1u) ? *reinterpret_cast<const float4*>(tab.grads[0] + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            p[u] = *reinterpret_cast<const float4*>(p_local + i);
            mm[u] = *reinterpret_cast<const float4*>(m + 4 * k);
            vv[u] = *reinterpret_cast<const float4*>(v + 4 * k);
        }
        if (!MC) {  // peer gradients, fixed rank order (replicas stay bit-identical)
#pragma unroll
            for (int r = 1; r < (WORLD > 0 ? WORLD : DP_MAX_WORLD); ++r) {
                if (r >= world) break;
                float4 x[DP_UNROLL];
#pragma unroll
                for (int u = 0; u < DP_UNROLL; ++u) {
                    const long long k = k0 + u * stride;
                    x[u] = (k < n4 && ((need[u] >> r) & 1u)) ? *reinterpret_cast<const float4*>(tab.grads[r] + begin + 4 * k)
                                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < DP_UNROLL; ++u) { g[u].x += x[u].x; g[u].y += x[u].y; g[u].z += x[u].z; g[u].w += x[u].w; }
            }
        }
#pragma unroll
        for (int u = 0; u < DP_UNROLL; ++u) {
            const long long k = k0 + u * stride;
            if (k >= n4) continue;
            const long long i = begin + 4 * k;
            int t = 0;
            while (t + 1 < tab.n_seg && i >= tab.seg_start[t + 1]) ++t;
            const float ns = tab.neg_step[t];
            adam_elem(p[u].x, g[u].x, mm[u].x, vv[u].x, b1, omb1, b2, omb2, inv_bc2_sqrt, eps, ns);
            adam_elem(p[u].y, g[u].y, mm[u].y, vv[u].y, b1, omb1, b2, omb2, inv_bc2_sqrt, eps, ns);
            adam_elem(p[u].z, g[u].z, mm[u].z, vv[u].z, b1, omb1, b2, omb2, inv_bc2_sqrt, eps, ns);
            adam_elem(p[u].w, g[u].w, mm[u].w, vv[u].w, b1, omb1, b2, omb2, inv_bc2_sqrt, eps, ns);
            *reinterpret_cast<float4*>(m + 4 * k) = mm[u];
            *reinterpret_cast<float4*>(v + 4 * k) = vv[u];
            if (MC) {
                mc_st(params_mc + i, p[u]);  // switch broadcast to every rank, this one included
            } else {
#pragma unroll
                for (int r = 0; r < (WORLD > 0 ? WORLD : DP_MAX_WORLD); ++r) {
                    if (r >= world) break;
                    *reinterpret_cast<float4*>(tab.params[r] + i) = p[u];
                }
            }
        }
    }
}

}  // namespace lgs

using namespace lgs;

static int dp_adam_shard_impl(int n_seg, const int64_t* seg_start, const double* lr, int world, int rank,
                              const float* const* grads_peers, float* const* params_peers, const float* grads_mc, float* params_mc,
                              int64_t shard_begin, int64_t shard_end, float* exp_avg_shard, float* exp_avg_sq_shard, double beta1,
                              double beta2, double eps, int step, int max_ctas, const int* row_len, const uint16_t* row_mask,
                              int64_t n_rows, void* stream) {
    if (n_seg < 1 || n_seg > DP_MAX_SEG || world < 1 || world > DP_MAX_WORLD || rank < 0 || rank >= world || step < 1)
        return LGS_ERR_INVALID_ARG;
    if (!seg_start || !lr || !grads_peers || !params_peers || !exp_avg_shard || !exp_avg_sq_shard) return LGS_ERR_INVALID_ARG;
    if (shard_begin < 0 || shard_end < shard_begin || ((shard_begin | shard_end) & 3)) return LGS_ERR_INVALID_ARG;
    if (shard_end == shard_begin) return LGS_OK;
    DpTable tab;
    const double bc1 = 1.0 - std::pow(beta1, (double)step), bc2 = 1.0 - std::pow(beta2, (double)step);
    for (int t = 0; t <= n_seg; ++t) {
        if (seg_start[t] & 3) return LGS_ERR_ALIGNMENT;
        tab.seg_start[t] = seg_start[t];
    }
    const bool sparse = row_mask != nullptr;
    for (int t = 0; t < n_seg; ++t) {
        tab.neg_step[t] = (float)(-(lr[t] / bc1));
        tab.row_len[t] = 1;
        if (sparse) {
            // every tensor must be [n_rows, row_len] (plus at most 3 floats of padding up to the next tensor)
            if (!row_len || row_len[t] < 1 || n_rows < 1) return LGS_ERR_INVALID_ARG;
            const int64_t room = seg_start[t + 1] - seg_start[t], used = n_rows * row_len[t];
            if (used > room || room - used > 3) return LGS_ERR_INVALID_ARG;
            tab.row_len[t] = row_len[t];
        }
    }
    for (int r = 0; r < world; ++r) {
        if (!grads_peers[r] || !params_peers[r]) return LGS_ERR_INVALID_ARG;
        tab.grads[r] = grads_peers[r];
        tab.params[r] = params_peers[r];
    }
    tab.n_seg = n_seg;
    tab.world = world;
    tab.n_rows = n_rows;
    const long long n4 = (shard_end - shard_begin) >> 2;
    const long long want = (n4 + 256LL * DP_UNROLL - 1) / (256LL * DP_UNROLL);
    // max_ctas > 0 bounds the grid: a launch that runs underneath other kernels on a side stream must leave them SM slots
    // (the default fills every thread slot of the GPU for the whole exchange)
    const long long cap = max_ctas > 0 ? (long long)max_ctas : 148LL * 8;
    const int grid = (int)(want < cap ? (want > 0 ? want : 1) : cap);
    const bool mc = grads_mc != nullptr && params_mc != nullptr;
    if (mc && sparse) return LGS_ERR_INVALID_ARG;  // the switch reduces every rank's copy: nothing to skip per rank
#define LGS_DP_LAUNCH(MCV, WV, SP)                                                                                              \
    dp_adam_kernel<MCV, WV, SP><<<grid, 256, 0, (cudaStream_t)stream>>>(tab, grads_mc, params_mc, shard_begin, shard_end,       \
                                                                        params_peers[rank], exp_avg_shard, exp_avg_sq_shard,   \
                                                                        (float)beta1, (float)(1.0 - beta1), (float)beta2,      \
                                                                        (float)(1.0 - beta2), 1.0f / (float)std::sqrt(bc2),    \
                                                                        (float)eps, row_mask)
    if (mc) LGS_DP_LAUNCH(true, 0, false);
    else if (sparse) {
        if (world == 2) LGS_DP_LAUNCH(false, 2, true);
        else if (world == 4) LGS_DP_LAUNCH(false, 4, true);
        else if (world == 8) LGS_DP_LAUNCH(false, 8, true);
        else LGS_DP_LAUNCH(false, 0, true);
    } else if (world == 2) LGS_DP_LAUNCH(false, 2, false);
    else if (world == 4) LGS_DP_LAUNCH(false, 4, false);
    else if (world == 8) LGS_DP_LAUNCH(false, 8, false);
    else LGS_DP_LAUNCH(false, 0, false);
#undef LGS_DP_LAUNCH
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

extern "C" int lgs_dp_adam_shard(int n_seg, const int64_t* seg_start, const double* lr, int world, int rank,
                                 const float* const* grads_peers, float* const* params_peers, const float* grads_mc,
                                 float* params_mc, int64_t shard_begin, int64_t shard_end, float* exp_avg_shard,
                                 float* exp_avg_sq_shard, double beta1, double beta2, double eps, int step, int max_ctas,
                                 void* stream) {
    return dp_adam_shard_impl(n_seg, seg_start, lr, world, rank, grads_peers, params_peers, grads_mc, params_mc, shard_begin,
                              shard_end, exp_avg_shard, exp_avg_sq_shard, beta1, beta2, eps, step, max_ctas, nullptr, nullptr, 0,
                              stream);
}

extern "C" int lgs_dp_adam_shard_sparse(int n_seg, const int64_t* seg_start, const double* lr, const int* row_len, int64_t n_rows,
                                        const uint16_t* row_mask, int world, int rank, const float* const* grads_peers,
                                        float* const* params_peers, int64_t shard_begin, int64_t shard_end, float* exp_avg_shard,
                                        float* exp_avg_sq_shard, double beta1, double beta2, double eps, int step, int max_ctas,
                                        void* stream) {
    if (!row_mask || !row_len) return LGS_ERR_INVALID_ARG;
    return dp_adam_shard_impl(n_seg, seg_start, lr, world, rank, grads_peers, params_peers, nullptr, nullptr, shard_begin, shard_end,
                              exp_avg_shard, exp_avg_sq_shard, beta1, beta2, eps, step, max_ctas, row_len, row_mask, n_rows, stream);
}

// ---- which gradient rows can be non-zero: per-rank visibility, published to every rank, combined into one bit mask ----------
namespace lgs {
__global__ void __launch_bounds__(256)
dp_rows_mark_kernel(int P, const int* __restrict__ radii, uint8_t* __restrict__ vis, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint8_t now = radii[i] > 0 ? 1 : 0;
    vis[i] = accumulate ? (uint8_t)(vis[i] | now) : now;
}
struct DpPeerTables {
    uint8_t* t[DP_MAX_WORLD];
};
__global__ void __launch_bounds__(256)
dp_rows_publish_kernel(int P, int world, int rank, const uint8_t* __restrict__ vis, const __grid_constant__ DpPeerTables peers) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 4 Gaussians
    if (4 * i >= P) return;
    uint32_t w = 0;
    if (4 * i + 3 < P) w = *reinterpret_cast<const uint32_t*>(vis + 4 * i);
    else for (int k = 0; 4 * i + k < P; ++k) w |= (uint32_t)vis[4 * i + k] << (8 * k);
    const size_t pitch = ((size_t)P + 3) & ~(size_t)3;
    for (int r = 0; r < world; ++r) *reinterpret_cast<uint32_t*>(peers.t[r] + (size_t)rank * pitch + 4 * i) = w;  // peer stores
}
__global__ void __launch_bounds__(256)
dp_rows_combine_kernel(int P, int world, const uint8_t* __restrict__ table, uint16_t* __restrict__ row_mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const size_t pitch = ((size_t)P + 3) & ~(size_t)3;
    uint32_t mk = 0;
    for (int r = 0; r < world; ++r) mk |= table[(size_t)r * pitch + i] ? (1u << r) : 0u;
    row_mask[i] = (uint16_t)mk;
}
}  // namespace lgs

extern "C" int lgs_dp_rows_mark(int P, const int* radii, unsigned char* vis, int accumulate, void* stream) {
    if (P < 0) return LGS_ERR_INVALID_ARG;
    if (P == 0) return LGS_OK;
    if (!radii || !vis) return LGS_ERR_INVALID_ARG;
    dp_rows_mark_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, radii, vis, accumulate);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}
extern "C" size_t lgs_dp_rows_table_bytes(int P, int world) {
    return (P < 0 || world < 1) ? 0 : (size_t)world * (((size_t)P + 3) & ~(size_t)3);
}
extern "C" int lgs_dp_rows_publish(int P, int world, int rank, const unsigned char* vis, unsigned char* const* tables_peers, void* stream) {
    if (P < 0 || world < 1 || world > DP_MAX_WORLD || rank < 0 || rank >= world) return LGS_ERR_INVALID_ARG;
    if (P == 0) return LGS_OK;
    if (!vis || !tables_peers || (reinterpret_cast<uintptr_t>(vis) & 3u)) return LGS_ERR_INVALID_ARG;
    DpPeerTables pt;
    for (int r = 0; r < world; ++r) {
        if (!tables_peers[r] || (reinterpret_cast<uintptr_t>(tables_peers[r]) & 3u)) return LGS_ERR_INVALID_ARG;
        pt.t[r] = tables_peers[r];
    }
    const int n4 = (P + 3) / 4;
    dp_rows_publish_kernel<<<(n4 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, world, rank, vis, pt);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}
extern "C" int lgs_dp_rows_combine(int P, int world, const unsigned char* table, unsigned short* row_mask, void* stream) {
    if (P < 0 || world < 1 || world > DP_MAX_WORLD) return LGS_ERR_INVALID_ARG;
    if (P == 0) return LGS_OK;
    if (!table || !row_mask) return LGS_ERR_INVALID_ARG;
    dp_rows_combine_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, world, table, row_mask);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}
