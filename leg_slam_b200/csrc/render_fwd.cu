// render_fwd.cu -- per-tile front-to-back alpha compositing of RGB + depth + 64-D language
// feature for sm_100a.  Replaces FORWARD::render / renderCUDA<3,64> (reference
// forward.cu:261-392, launch :411).
//
// Design (see DESIGN.md "render forward"):
//  * one CTA (64 threads = 2 warps) per 8x8 tile, one pixel per thread, the 68 per-pixel
//    accumulators (3 colour + 64 feature + depth) live in registers;
//  * the tile's sorted instance list is consumed in batches of 32.  For each batch the
//    48-byte render record and the 256-byte feature row of every Gaussian are staged in
//    shared memory by TMA bulk copies (cp.async.bulk, one pair per Gaussian, issued by
//    the lanes of warp 0) into a 3-deep ring of stages; completion is tracked with one
//    mbarrier per stage, so the gather of batch b+2 overlaps the blend of batch b.  The
//    reference instead re-reads 68 operands per fragment per pixel from global memory
//    (forward.cu:360-368);
//  * each fragment's alpha test is evaluated per pixel with the reference's exact float
//    sequence (so alpha, T, the termination decision and n_contrib are bit-identical);
//    a warp ballot skips the 68 FMAs + 17 shared loads when none of the warp's 32 pixels
//    blends the Gaussian;
//  * blending uses w = alpha*T once per fragment and acc += w*v (one FFMA per channel,
//    feature operands broadcast from shared memory as float4).
//
// LGS_NO_TMA=1 (environment, read at launch) selects a plain LDG->STS staging path with the
// same math: a differential-debugging aid, not a fallback for other hardware.
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"

namespace lgs {

constexpr int FB = 32;      // instances per batch
constexpr int FSTAGES = 3;  // ring depth

template <bool WITH_LF>
struct FwdStage {
    GaussRec rec[FB];
    float lf[WITH_LF ? FB * LF : 4];
};

template <bool WITH_LF, bool USE_TMA>
__global__ void __launch_bounds__(TILE_PIX)
render_fwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
                  const GaussRec* __restrict__ rec, const float* __restrict__ lang_feat,
                  const float* __restrict__ bg, float* __restrict__ final_T,
                  uint32_t* __restrict__ n_contrib, uint32_t* __restrict__ tile_last,
                  float* __restrict__ out_color, float* __restrict__ out_lf, float* __restrict__ out_depth) {
    using Stage = FwdStage<WITH_LF>;
    __shared__ __align__(128) Stage stages[USE_TMA ? FSTAGES : 1];
    __shared__ __align__(8) uint64_t full_bar[FSTAGES];
    __shared__ uint32_t s_tile_last[2];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int tile_id = blockIdx.y * gridDim.x + blockIdx.x;
    const uint32_t pxi = blockIdx.x * TILE + (tid & 7);
    const uint32_t pyi = blockIdx.y * TILE + (tid >> 3);
    const bool inside = pxi < (uint32_t)W && pyi < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * pyi + pxi;
    const float pxf = (float)pxi, pyf = (float)pyi;

    const uint2 range = ranges[tile_id];
    const int n = (int)(range.y - range.x);
    const int nb = (n + FB - 1) / FB;

    if (USE_TMA) {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < FSTAGES; ++s) mbar_init(&full_bar[s], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }

    bool done = !inside;
    float T = 1.0f;
    uint32_t last_contributor = 0;
    const float wx0 = (float)(blockIdx.x * TILE), wy0 = (float)(blockIdx.y * TILE + 4 * (tid >> 5));  // this warp's 8x4 pixels
    float C0 = 0.f, C1 = 0.f, C2 = 0.f, Dacc = 0.f;
    float LFacc[WITH_LF ? LF : 1];
#pragma unroll
    for (int k = 0; k < (WITH_LF ? LF : 1); ++k) LFacc[k] = 0.f;

    // ---- producer state (warp 0 only): instance id of the next batch to issue, prefetched
    int issued = 0, consumed = 0;
    uint32_t pf_id = 0;
    if (USE_TMA && tid < FB && tid < n) pf_id = point_list[range.x + tid];

    auto issue = [&](int b) {  // called by every thread, acts on warp 0
        if (tid < 32) {  // the whole of warp 0 (FB <= 32 lanes carry an instance each)
            const int cnt = min(FB, n - b * FB);
            Stage& S = stages[b % FSTAGES];
            uint64_t* bar = &full_bar[b % FSTAGES];
            if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)cnt * (uint32_t)(sizeof(GaussRec) + (WITH_LF ? LF * 4 : 0)));
            __syncwarp();
            if (tid < cnt) {
                tma_bulk_g2s(&S.rec[tid], rec + pf_id, sizeof(GaussRec), bar);
                if (WITH_LF) tma_bulk_g2s(&S.lf[tid * LF], lang_feat + (size_t)pf_id * LF, LF * 4, bar);
            }
            const int nxt = (b + 1) * FB + tid;
            if (tid < FB && nxt < n) pf_id = point_list[range.x + nxt];
        }
        ++issued;
    };

    if (USE_TMA) {
        for (int b = 0; b < FSTAGES - 1 && b < nb; ++b) issue(b);
    }

    for (int b = 0; b < nb; ++b) {
        const int cnt = min(FB, n - b * FB);
        Stage* Sp;
        if (USE_TMA) {
            if (b + FSTAGES - 1 < nb) issue(b + FSTAGES - 1);
            mbar_wait(&full_bar[b % FSTAGES], (uint32_t)((b / FSTAGES) & 1));
            ++consumed;
            Sp = &stages[b % FSTAGES];
        } else {
            Sp = &stages[0];
            // plain staging: 32 records (96 float4) + 32 feature rows (512 float4) by 64 threads
            const uint32_t* ids = point_list + range.x + b * FB;
            for (int i = tid; i < cnt * 3; i += TILE_PIX) {
                const int g = i / 3, q = i - 3 * g;
                reinterpret_cast<float4*>(&Sp->rec[g])[q] = reinterpret_cast<const float4*>(rec + ids[g])[q];
            }
            if (WITH_LF) {
                for (int i = tid; i < cnt * (LF / 4); i += TILE_PIX) {
                    const int g = i / (LF / 4), q = i - (LF / 4) * g;
                    reinterpret_cast<float4*>(&Sp->lf[g * LF])[q] =
                        reinterpret_cast<const float4*>(lang_feat + (size_t)ids[g] * LF)[q];
                }
            }
            __syncthreads();
        }
        const Stage& S = *Sp;

        // which of this batch's Gaussians can touch this warp's 8x4 pixels at all?  lane j tests Gaussian j
        // (opacity-aware bounding box, common.cuh); the ballot lets the warp skip the others in 3 instructions
        // instead of evaluating 32 alphas
        bool touch = false;
        if (lane < cnt) {
            const float4 t0 = S.rec[lane].q0, t1 = S.rec[lane].q1;
            touch = footprint_touches(t0.x, t0.y, t1.x, t1.y, t1.z, t1.w, wx0, wx0 + 7.0f, wy0, wy0 + 3.0f);
        }
        const uint32_t vis = __ballot_sync(0xffffffffu, touch);

        // Visible Gaussians are taken four at a time: the four alpha evaluations (load, falloff, expf) are
        // independent chains that overlap (ILP), the state-dependent part (T, early termination, blend) then
        // runs in list order.
        uint32_t v = vis;
#pragma unroll 1
        while (v != 0) {
            int jj[4];
            float al[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool valid = v != 0;
                jj[u] = valid ? (__ffs(v) - 1) : 0;
                v &= v - 1;  // 0 stays 0
                const float4 q0 = S.rec[jj[u]].q0;  // x, y, depth
                const float4 q1 = S.rec[jj[u]].q1;  // conic a,b,c, opacity
                float dx, dy;
                const float power = eval_power(q0.x, q0.y, pxf, pyf, q1.x, q1.y, q1.z, dx, dy);
                // forward.cu:342-357 (same comparisons, same float ops)
                al[u] = fminf(0.99f, __fmul_rn(q1.w, expf(power)));
                ok[u] = valid && !(power > 0.0f) && !(al[u] < 1.0f / 255.0f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = jj[u];
                const float alpha = al[u];
                const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                bool act = ok[u] && !done;
                if (act && test_T < 0.0001f) {
                    done = true;
                    act = false;
                }
                if (__any_sync(0xffffffffu, act)) {
                    const float w = act ? __fmul_rn(alpha, T) : 0.0f;
                    const float4 q2 = S.rec[j].q2;
                    C0 = fmaf(w, q2.x, C0);
                    C1 = fmaf(w, q2.y, C1);
                    C2 = fmaf(w, q2.z, C2);
                    Dacc = fmaf(w, S.rec[j].q0.z, Dacc);
                    if (WITH_LF) {
                        const float4* f4 = reinterpret_cast<const float4*>(&S.lf[j * LF]);
#pragma unroll
                        for (int k = 0; k < LF / 4; ++k) {
                            const float4 f = f4[k];
                            LFacc[4 * k + 0] = fmaf(w, f.x, LFacc[4 * k + 0]);
                            LFacc[4 * k + 1] = fmaf(w, f.y, LFacc[4 * k + 1]);
                            LFacc[4 * k + 2] = fmaf(w, f.z, LFacc[4 * k + 2]);
                            LFacc[4 * k + 3] = fmaf(w, f.w, LFacc[4 * k + 3]);
                        }
                    }
                    if (act) {
                        T = test_T;
                        last_contributor = (uint32_t)(b * FB + j + 1);  // 1-based position in the tile's list
                    }
                }
            }
        }
        // everyone is finished with this stage; also the tile-wide termination vote
        // (forward.cu:315 __syncthreads_count)
        if (__syncthreads_and(done)) break;
    }

    if (USE_TMA) {
        // never leave the CTA with bulk copies still in flight into its shared memory
        for (int b = consumed; b < issued; ++b) mbar_wait(&full_bar[b % FSTAGES], (uint32_t)((b / FSTAGES) & 1));
    }

    // ---- epilogue (forward.cu:377-391)
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, last_contributor);
    if (lane == 0) s_tile_last[tid >> 5] = wmax;
    if (inside) {
        const size_t HW = (size_t)H * W;
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
        out_color[0 * HW + pix_id] = fmaf(T, bg[0], C0);
        out_color[1 * HW + pix_id] = fmaf(T, bg[1], C1);
        out_color[2 * HW + pix_id] = fmaf(T, bg[2], C2);
        if (WITH_LF) {
#pragma unroll
            for (int k = 0; k < LF; ++k) out_lf[(size_t)k * HW + pix_id] = LFacc[k];
        }
        out_depth[pix_id] = Dacc;
    }
    __syncthreads();
    if (tid == 0) tile_last[tile_id] = max(s_tile_last[0], s_tile_last[1]);
}

static bool env_no_tma() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("LGS_NO_TMA");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

// LGS_FWD=simt (environment, read at the first launch) keeps this file's SIMT kernel for the 64-D feature path: a
// differential-debugging aid for the tensor-core kernel (render_fwd_tc.cu), not a fallback for other hardware.
static bool env_simt_fwd() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("LGS_FWD");
        v = (e && e[0] == 's') ? 1 : 0;
    }
    return v == 1;
}

int launch_render_fwd(int W, int H, int R, const GeomState& g, const BinningState& b, ImageState& im,
                      const float* background, const float* lang_feat, float* out_color,
                      float* out_lang_feat, float* out_depth, bool include_lf, cudaStream_t s) {
    if (include_lf && !env_simt_fwd())
        return launch_render_fwd_tc(W, H, g, b, im, background, lang_feat, out_color, out_lang_feat, out_depth, s);
    (void)R;
    const dim3 grid((W + TILE - 1) / TILE, (H + TILE - 1) / TILE, 1);
    const bool tma = !env_no_tma();
#define LGS_FWD_LAUNCH(LFV, TMAV)                                                                        \
    render_fwd_kernel<LFV, TMAV><<<grid, TILE_PIX, 0, s>>>(im.ranges, b.point_list, W, H, g.rec, lang_feat, \
                                                            background, im.final_T, im.n_contrib,           \
                                                            im.tile_last, out_color, out_lang_feat, out_depth)
    if (include_lf) {
        if (tma) LGS_FWD_LAUNCH(true, true); else LGS_FWD_LAUNCH(true, false);
    } else {
        if (tma) LGS_FWD_LAUNCH(false, true); else LGS_FWD_LAUNCH(false, false);
    }
#undef LGS_FWD_LAUNCH
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

}  // namespace lgs
