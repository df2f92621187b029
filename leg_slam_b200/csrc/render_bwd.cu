// render_bwd.cu -- backward of the per-tile alpha compositing (RGB + depth + 64-D language
// feature) for sm_100a.  Replaces BACKWARD::render / renderCUDA<3,64> (reference
// backward.cu:399-612, launch :703).
//
// The reference keeps, per pixel thread, three 68-float arrays (running "colour behind"
// accumulator, incoming gradient, last colour: 254 registers) and issues 74 global float
// atomics per blended fragment per pixel (backward.cu:557-609).  This file restructures the same
// arithmetic (DESIGN.md "render backward"):
//
//  * dL/dalpha only ever consumes the running accumulator through its dot product with the
//    pixel's incoming gradient g (backward.cu:549-576).  The recurrence is linear, so it is
//    carried as ONE scalar:  A <- last_alpha * d_last + (1-last_alpha) * A  with
//    d_j = <feature_j, g>, and dL/dalpha_j = (d_j - A) * T_j - T_final/(1-alpha_j) * <bg, g_rgb>.
//    Per-pixel state drops from 3x68 floats to 68 (g) + a handful of scalars.
//  * per-Gaussian gradients are sums over the tile's 64 pixels of per-pixel scalars times
//    fixed per-pixel vectors:
//        dL/dfeature[j][ch] = sum_pix  w[j][pix] * g[pix][ch]          w = alpha*T
//        dL/d{mean2D, conic, opacity}[j] = closed forms in the six moments
//        sum_pix t[j][pix] * {1, u, v, u^2, uv, v^2}                     t = G * dL/dalpha
//    (u, v = pixel offset from the tile centre), i.e. a [instances x 64] by [64 x 74] product per
//    tile.  Threads that each own one COLUMN of the tile's [64 x 74] matrix in 64 registers reduce
//    it inside the CTA, so the only global traffic is ONE red.global.add per (tile, Gaussian,
//    column) -- 74 per (tile, Gaussian) instead of 74 per (pixel, Gaussian).
//  * the two roles want different register files (68-float g ROW per pixel thread vs 64-float g
//    COLUMN per channel thread), different thread counts (64 vs 74) and run at different,
//    bursty rates, so they are two kernels: the pixel kernel streams compact w/t records
//    (256 B per instance-half with an active pixel, plus a ballot) to a scratch buffer, the
//    channel kernel streams them back through a TMA ring.  An earlier single-kernel,
//    warp-specialised version (named-barrier hand-off through shared memory) spent 46 % of its
//    warp samples in barrier stalls (profiles/r01_render_bwd_v1.md); the scratch round trip is
//    ~0.6 GB of mostly L2-resident traffic per iteration at cfgB.
//  * Gaussian records / feature rows are staged by TMA bulk copies (cp.async.bulk + mbarrier)
//    two batches ahead, like the forward.
//  * the tile's list is cut at tile_last = max over the tile's pixels of n_contrib (written
//    by the forward): instances behind it contribute to no pixel.
//
// Summation order differs from the reference's atomics (and from run to run), which is why
// the gradient parity gate is 1e-3 relative.
#include "common.cuh"
#include "ptx.cuh"

namespace lgs {

constexpr int BB = 32;       // Gaussians per TMA batch of the pixel kernel
constexpr int BSTAGES = 3;   // its staging ring depth
constexpr int CB = 16;       // instances per TMA batch of the channel kernel
constexpr int CSTAGES = 4;   // its staging ring depth
constexpr int NCOL = 74;     // 64 feature + 3 colour + 1 depth + 6 moment columns
constexpr int PAIR_FLOATS = 128;  // per instance: [half0: W[32] | t[32]] [half1: W[32] | t[32]]

template <bool WITH_LF>
struct BwdStage {
    GaussRec rec[BB];
    float lf[WITH_LF ? BB * LF : 4];
};

struct ChanStage {
    float pairs[CB][PAIR_FLOATS];  // TMA destination: only the active halves are copied
    GaussRec rec[CB];              // TMA destination
    uint2 mask[CB];                // ballots of the two pixel warps
    uint32_t id[CB];               // Gaussian index
};

// ================================ pixel kernel =====================================================
// One CTA (2 warps) per tile, one pixel per thread.  Walks the tile's list back to front, replays
// alpha, evaluates d_j = <feature_j, g_pix>, runs the scalar recurrences and writes, per
// (instance, pixel warp) with at least one active pixel, w = alpha*T and t = G*dL/dalpha for its
// 32 pixels (one coalesced 256-byte record) plus the warp's ballot, to the pair buffer.
template <bool WITH_LF>
__global__ void __launch_bounds__(TILE_PIX)
render_bwd_pix_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
                      const float* __restrict__ bg, const GaussRec* __restrict__ rec,
                      const float* __restrict__ lang_feat, const float* __restrict__ final_T,
                      const uint32_t* __restrict__ n_contrib, const uint32_t* __restrict__ tile_last,
                      const float* __restrict__ dL_dpix, const float* __restrict__ dL_dpix_lf,
                      const float* __restrict__ dL_dpix_depth, float* __restrict__ pair_buf,
                      uint32_t* __restrict__ pair_mask) {
    using Stage = BwdStage<WITH_LF>;
    __shared__ __align__(128) Stage stages[BSTAGES];
    __shared__ __align__(8) uint64_t full_bar[BSTAGES];

    const int tid = threadIdx.x;
    const int lane = tid & 31, wrp = tid >> 5;
    const int tile_id = blockIdx.y * gridDim.x + blockIdx.x;
    const uint2 range = ranges[tile_id];
    const int n = min((int)(range.y - range.x), (int)tile_last[tile_id]);  // entries behind tile_last touch no pixel
    if (n <= 0) return;
    const int nb = (n + BB - 1) / BB;
    const size_t HW = (size_t)H * W;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < BSTAGES; ++s) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const uint32_t pxi = blockIdx.x * TILE + (tid & 7);
    const uint32_t pyi = blockIdx.y * TILE + (tid >> 3);
    const bool inside = pxi < (uint32_t)W && pyi < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * pyi + pxi;
    const float pxf = (float)pxi, pyf = (float)pyi;

    const float T_final = inside ? final_T[pix_id] : 0.f;
    const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;
    float T = T_final;

    float g_lf[WITH_LF ? LF : 1];
    float g_r = 0.f, g_g = 0.f, g_b = 0.f, g_d = 0.f;
#pragma unroll
    for (int k = 0; k < (WITH_LF ? LF : 1); ++k) g_lf[k] = 0.f;
    if (inside) {
        g_r = dL_dpix[0 * HW + pix_id];
        g_g = dL_dpix[1 * HW + pix_id];
        g_b = dL_dpix[2 * HW + pix_id];
        g_d = dL_dpix_depth[pix_id];
        if (WITH_LF) {
#pragma unroll
            for (int k = 0; k < LF; ++k) g_lf[k] = dL_dpix_lf[(size_t)k * HW + pix_id];
        }
    }
    const float bgdot = bg[0] * g_r + bg[1] * g_g + bg[2] * g_b;  // backward.cu:585-588
    float Acc = 0.f, last_alpha = 0.f, last_d = 0.f;

    // producer state (warp 0): Gaussian ids of the next batch to issue (back to front)
    uint32_t pf_id = 0;
    if (tid < BB && tid < n) pf_id = point_list[range.x + (n - 1 - tid)];
    auto issue = [&](int b) {
        if (tid < BB) {
            const int cnt = min(BB, n - b * BB);
            Stage& S = stages[b % BSTAGES];
            uint64_t* bar = &full_bar[b % BSTAGES];
            if (tid == 0)
                mbar_arrive_expect_tx(bar, (uint32_t)cnt * (uint32_t)(sizeof(GaussRec) + (WITH_LF ? LF * 4 : 0)));
            __syncwarp();
            if (tid < cnt) {
                tma_bulk_g2s(&S.rec[tid], rec + pf_id, sizeof(GaussRec), bar);
                if (WITH_LF) tma_bulk_g2s(&S.lf[tid * LF], lang_feat + (size_t)pf_id * LF, LF * 4, bar);
            }
            const int nxt = (b + 1) * BB + tid;
            if (nxt < n) pf_id = point_list[range.x + (n - 1 - nxt)];
        }
    };
    for (int b = 0; b < BSTAGES - 1 && b < nb; ++b) issue(b);

    for (int b = 0; b < nb; ++b) {
        const int cnt = min(BB, n - b * BB);
        const int hi = n - 1 - b * BB;  // list position of slot 0
        __syncthreads();                // both warps are done with stage (b-1) % BSTAGES -> refill it
        if (b + BSTAGES - 1 < nb) issue(b + BSTAGES - 1);
        mbar_wait(&full_bar[b % BSTAGES], (uint32_t)((b / BSTAGES) & 1));
        const Stage& S = stages[b % BSTAGES];

#pragma unroll 1
        for (int j = 0; j < cnt; ++j) {
            const int p = hi - j;
            const float4 q0 = S.rec[j].q0;
            const float4 q1 = S.rec[j].q1;
            float dx, dy;
            const float power = eval_power(q0.x, q0.y, pxf, pyf, q1.x, q1.y, q1.z, dx, dy);
            const float G = expf(power);
            const float alpha = fminf(0.99f, __fmul_rn(q1.w, G));
            // backward.cu:513-530: behind the pixel's last contributor, outside the falloff, or below
            // the alpha threshold -> no contribution
            const bool act = (p < last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
            const uint32_t m = __ballot_sync(0xffffffffu, act);
            const size_t inst = (size_t)range.x + (size_t)p;
            if (lane == 0) pair_mask[2 * inst + wrp] = m;
            if (m == 0) continue;

            const float4 q2 = S.rec[j].q2;
            float d0 = q2.x * g_r, d1 = q2.y * g_g, d2 = q2.z * g_b, d3 = q0.z * g_d;
            if (WITH_LF) {
                const float4* f4 = reinterpret_cast<const float4*>(&S.lf[j * LF]);
#pragma unroll
                for (int k = 0; k < LF / 4; ++k) {
                    const float4 f = f4[k];
                    d0 = fmaf(f.x, g_lf[4 * k + 0], d0);
                    d1 = fmaf(f.y, g_lf[4 * k + 1], d1);
                    d2 = fmaf(f.z, g_lf[4 * k + 2], d2);
                    d3 = fmaf(f.w, g_lf[4 * k + 3], d3);
                }
            }
            const float d = (d0 + d1) + (d2 + d3);
            float Wv = 0.f, Tv = 0.f;
            if (act) {
                const float inv = __frcp_rn(1.0f - alpha);
                T = T * inv;                                                // :536
                Acc = fmaf(last_alpha, last_d, (1.0f - last_alpha) * Acc);  // :549,563,573 (dotted with g)
                last_d = d;
                last_alpha = alpha;
                const float dL_dalpha = fmaf(d - Acc, T, -T_final * inv * bgdot);  // :553-589
                Wv = alpha * T;
                Tv = G * dL_dalpha;
            }
            float* dst = pair_buf + inst * PAIR_FLOATS + wrp * 64 + lane;
            __stcg(dst, Wv);
            __stcg(dst + 32, Tv);
        }
    }
}

// ================================ channel kernel ===================================================
// One CTA per tile; every thread owns one of the 74 columns (64 feature channels, 3 colour, depth,
// 6 moment bases) of the tile's [64 pixel x 74] matrix in registers and reduces the pixel kernel's
// w / t records against it: ONE red.global.add per (tile, Gaussian, column).
template <bool WITH_LF>
__global__ void __launch_bounds__(WITH_LF ? 96 : 32)
render_bwd_chan_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
                       const GaussRec* __restrict__ rec, const uint32_t* __restrict__ tile_last,
                       const float* __restrict__ dL_dpix, const float* __restrict__ dL_dpix_lf,
                       const float* __restrict__ dL_dpix_depth, const float* __restrict__ pair_buf,
                       const uint32_t* __restrict__ pair_mask, float* __restrict__ dL_dmean2D,
                       float* __restrict__ dL_dconic, float* __restrict__ dL_dopacity,
                       float* __restrict__ dL_dcolor, float* __restrict__ dL_dlang_feat,
                       float* __restrict__ dL_ddepth) {
    __shared__ __align__(128) ChanStage stages[CSTAGES];
    __shared__ __align__(8) uint64_t full_bar[CSTAGES];

    const int tid = threadIdx.x;
    const int tile_id = blockIdx.y * gridDim.x + blockIdx.x;
    const uint2 range = ranges[tile_id];
    const int n = min((int)(range.y - range.x), (int)tile_last[tile_id]);
    if (n <= 0) return;
    const int nb = (n + CB - 1) / CB;
    const size_t HW = (size_t)H * W;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < CSTAGES; ++s) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }

    const int col = tid + (WITH_LF ? 0 : LF);  // 0..63 lf, 64..66 rgb, 67 depth, 68..73 moments
    float colv[TILE_PIX];
#pragma unroll
    for (int i = 0; i < TILE_PIX; ++i) {
        const uint32_t pxi = blockIdx.x * TILE + (i & 7);
        const uint32_t pyi = blockIdx.y * TILE + (i >> 3);
        const bool inside = pxi < (uint32_t)W && pyi < (uint32_t)H;
        const size_t pix_id = (size_t)W * pyi + pxi;
        const float u = (float)(i & 7) - 3.5f, v = (float)(i >> 3) - 3.5f;
        float x = 0.f;
        if (col < LF) {
            if (WITH_LF && inside) x = dL_dpix_lf[(size_t)col * HW + pix_id];
        } else if (col < LF + 3) {
            if (inside) x = dL_dpix[(size_t)(col - LF) * HW + pix_id];
        } else if (col == LF + 3) {
            if (inside) x = dL_dpix_depth[pix_id];
        } else {
            const int k = col - (LF + 4);
            x = k == 0 ? 1.f : k == 1 ? u : k == 2 ? v : k == 3 ? u * u : k == 4 ? u * v : v * v;
        }
        colv[i] = x;
    }
    const int toff = (col >= LF + 4) ? 32 : 0;  // moment columns reduce t, the others reduce w
    const bool live = col < NCOL;
    const float ddelx_dx = 0.5f * (float)W, ddely_dy = 0.5f * (float)H;  // backward.cu:478-479
    const float cxf = (float)(blockIdx.x * TILE) + 3.5f, cyf = (float)(blockIdx.y * TILE) + 3.5f;
    __syncthreads();

    // producer (warp 0, lanes 0..CB-1): one instance per lane.  Fills stage b % CSTAGES: mask/id by
    // plain stores, record + active halves by TMA bulk copies.
    auto issue = [&](int b) {
        if (tid < 32) {
            const int cnt = min(CB, n - b * CB);
            ChanStage& S = stages[b % CSTAGES];
            uint64_t* bar = &full_bar[b % CSTAGES];
            uint2 m = make_uint2(0u, 0u);
            uint32_t id = 0;
            const size_t inst = (size_t)range.x + (size_t)(b * CB + tid);
            if (tid < cnt) {
                m = *reinterpret_cast<const uint2*>(pair_mask + 2 * inst);
                id = point_list[inst];
                S.mask[tid] = m;
                S.id[tid] = id;
            }
            const bool any = (m.x | m.y) != 0;
            uint32_t bytes = any ? (uint32_t)sizeof(GaussRec) + (m.x ? 256u : 0u) + (m.y ? 256u : 0u) : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
            if (tid == 0) mbar_arrive_expect_tx(bar, bytes);
            __syncwarp();
            if (any) {
                tma_bulk_g2s(&S.rec[tid], rec + id, sizeof(GaussRec), bar);
                if (m.x) tma_bulk_g2s(&S.pairs[tid][0], pair_buf + inst * PAIR_FLOATS, 256, bar);
                if (m.y) tma_bulk_g2s(&S.pairs[tid][64], pair_buf + inst * PAIR_FLOATS + 64, 256, bar);
            }
        }
    };
    for (int b = 0; b < CSTAGES - 1 && b < nb; ++b) issue(b);

    for (int b = 0; b < nb; ++b) {
        const int cnt = min(CB, n - b * CB);
        __syncthreads();  // every warp is done with stage (b-1) % CSTAGES; its mask/id stores are ordered too
        if (b + CSTAGES - 1 < nb) issue(b + CSTAGES - 1);
        mbar_wait(&full_bar[b % CSTAGES], (uint32_t)((b / CSTAGES) & 1));
        const ChanStage& S = stages[b % CSTAGES];
#pragma unroll 1
        for (int j = 0; j < cnt; ++j) {
            const uint2 m = S.mask[j];
            if ((m.x | m.y) == 0) continue;
            const float* src = &S.pairs[j][toff];
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            if (m.x != 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 x = reinterpret_cast<const float4*>(src)[k];
                    a0 = fmaf(x.x, colv[4 * k + 0], a0);
                    a1 = fmaf(x.y, colv[4 * k + 1], a1);
                    a2 = fmaf(x.z, colv[4 * k + 2], a2);
                    a3 = fmaf(x.w, colv[4 * k + 3], a3);
                }
            }
            if (m.y != 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 x = reinterpret_cast<const float4*>(src + 64)[k];
                    a0 = fmaf(x.x, colv[32 + 4 * k + 0], a0);
                    a1 = fmaf(x.y, colv[32 + 4 * k + 1], a1);
                    a2 = fmaf(x.z, colv[32 + 4 * k + 2], a2);
                    a3 = fmaf(x.w, colv[32 + 4 * k + 3], a3);
                }
            }
            const float acc = (a0 + a1) + (a2 + a3);
            const uint32_t id = S.id[j];
            if (col < LF) {
                red_add_f32(dL_dlang_feat + (size_t)id * LF + col, acc);
            } else {
                // extras warp: lanes 0-2 colour, 3 depth, 4-9 moments
                const float S0 = __shfl_sync(0xffffffffu, acc, 4);
                const float Su = __shfl_sync(0xffffffffu, acc, 5);
                const float Sv = __shfl_sync(0xffffffffu, acc, 6);
                const float Suu = __shfl_sync(0xffffffffu, acc, 7);
                const float Suv = __shfl_sync(0xffffffffu, acc, 8);
                const float Svv = __shfl_sync(0xffffffffu, acc, 9);
                if (live) {
                    const float4 q0 = S.rec[j].q0, q1 = S.rec[j].q1;
                    const float gx = q0.x - cxf, gy = q0.y - cyf, ca = q1.x, cb = q1.y, cc = q1.z, op = q1.w;
                    // sums over the tile's pixels of t*dx, t*dy (dx = gx - u, dy = gy - v)
                    const float sdx = gx * S0 - Su, sdy = gy * S0 - Sv;
                    switch (col - LF) {
                        case 0: case 1: case 2:
                            red_add_f32(dL_dcolor + (size_t)id * 3 + (col - LF), acc);
                            break;
                        case 3:
                            if (dL_ddepth != nullptr) red_add_f32(dL_ddepth + id, acc);
                            break;
                        case 4:  // dL_dmean2D.x  (backward.cu:592-601)
                            red_add_f32(dL_dmean2D + (size_t)id * 3 + 0, -op * ddelx_dx * (ca * sdx + cb * sdy));
                            break;
                        case 5:  // dL_dmean2D.y
                            red_add_f32(dL_dmean2D + (size_t)id * 3 + 1, -op * ddely_dy * (cc * sdy + cb * sdx));
                            break;
                        case 6:  // dL_dconic.x   (:604)
                            red_add_f32(dL_dconic + (size_t)id * 4 + 0, -0.5f * op * (gx * gx * S0 - 2.f * gx * Su + Suu));
                            break;
                        case 7:  // dL_dconic.y   (:605)
                            red_add_f32(dL_dconic + (size_t)id * 4 + 1,
                                        -0.5f * op * (gx * gy * S0 - gx * Sv - gy * Su + Suv));
                            break;
                        case 8:  // dL_dconic.w   (:606)
                            red_add_f32(dL_dconic + (size_t)id * 4 + 3, -0.5f * op * (gy * gy * S0 - 2.f * gy * Sv + Svv));
                            break;
                        case 9:  // dL_dopacity   (:609)
                            red_add_f32(dL_dopacity + id, S0);
                            break;
                        default: break;
                    }
                }
            }
        }
    }
}

// one fused memset of the accumulated-into gradient arrays (the reference's caller does
// torch::zeros per tensor, rasterize_points.cu:157-167)
__global__ void __launch_bounds__(256)
zero_grads_kernel(int P, float* __restrict__ mean2D, float* __restrict__ conic, float* __restrict__ opacity,
                  float* __restrict__ color, float4* __restrict__ lf, float* __restrict__ depth) {
    const int stride = gridDim.x * blockDim.x;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (lf != nullptr) {
        const int n4 = P * (LF / 4);
        for (int i = t0; i < n4; i += stride) lf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = t0; i < P; i += stride) {
        mean2D[3 * i + 0] = 0.f; mean2D[3 * i + 1] = 0.f; mean2D[3 * i + 2] = 0.f;
        reinterpret_cast<float4*>(conic)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        opacity[i] = 0.f;
        color[3 * i + 0] = 0.f; color[3 * i + 1] = 0.f; color[3 * i + 2] = 0.f;
        if (depth != nullptr) depth[i] = 0.f;
    }
}

int launch_zero_grads(int P, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                      float* dL_dlang_feat, float* dL_ddepth, bool include_lf, cudaStream_t s) {
    const int blocks = min((P * (LF / 4) + 255) / 256, 148 * 16);
    zero_grads_kernel<<<max(blocks, 1), 256, 0, s>>>(P, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor,
                                                     include_lf ? reinterpret_cast<float4*>(dL_dlang_feat) : nullptr,
                                                     dL_ddepth);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

size_t render_bwd_scratch_bytes(int R) {
    const size_t n = (size_t)(R > 0 ? R : 1);
    return n * PAIR_FLOATS * sizeof(float) + n * 2 * sizeof(uint32_t) + 512;
}

int launch_render_bwd(int P, int W, int H, int R, const GeomState& g, const BinningState& b,
                      const ImageState& im, const float* background, const float* lang_feat,
                      const float* dL_dpix, const float* dL_dpix_lf, const float* dL_dpix_depth,
                      float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                      float* dL_dlang_feat, float* dL_ddepth, bool include_lf, char* scratch, cudaStream_t s) {
    (void)P;
    const dim3 grid((W + TILE - 1) / TILE, (H + TILE - 1) / TILE, 1);
    // scratch: [R][128] floats of w/t records (256-byte aligned) followed by [R][2] ballots
    uintptr_t base = (reinterpret_cast<uintptr_t>(scratch) + 255) & ~(uintptr_t)255;
    float* pair_buf = reinterpret_cast<float*>(base);
    uint32_t* pair_mask = reinterpret_cast<uint32_t*>(pair_buf + (size_t)(R > 0 ? R : 1) * PAIR_FLOATS);
    if (include_lf) {
        render_bwd_pix_kernel<true><<<grid, TILE_PIX, 0, s>>>(im.ranges, b.point_list, W, H, background, g.rec, lang_feat,
                                                              im.final_T, im.n_contrib, im.tile_last, dL_dpix, dL_dpix_lf,
                                                              dL_dpix_depth, pair_buf, pair_mask);
        LGS_LAUNCH_CHECK();
        prof_mark(PM_RENDER_BWD_PIX, s);
        render_bwd_chan_kernel<true><<<grid, 96, 0, s>>>(im.ranges, b.point_list, W, H, g.rec, im.tile_last, dL_dpix,
                                                         dL_dpix_lf, dL_dpix_depth, pair_buf, pair_mask, dL_dmean2D,
                                                         dL_dconic, dL_dopacity, dL_dcolor, dL_dlang_feat, dL_ddepth);
    } else {
        render_bwd_pix_kernel<false><<<grid, TILE_PIX, 0, s>>>(im.ranges, b.point_list, W, H, background, g.rec, lang_feat,
                                                               im.final_T, im.n_contrib, im.tile_last, dL_dpix, dL_dpix_lf,
                                                               dL_dpix_depth, pair_buf, pair_mask);
        LGS_LAUNCH_CHECK();
        prof_mark(PM_RENDER_BWD_PIX, s);
        render_bwd_chan_kernel<false><<<grid, 32, 0, s>>>(im.ranges, b.point_list, W, H, g.rec, im.tile_last, dL_dpix,
                                                          dL_dpix_lf, dL_dpix_depth, pair_buf, pair_mask, dL_dmean2D,
                                                          dL_dconic, dL_dopacity, dL_dcolor, dL_dlang_feat, dL_ddepth);
    }
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

}  // namespace lgs
