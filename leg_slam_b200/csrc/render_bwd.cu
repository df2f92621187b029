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
//    bursty rates, so they are two kernels: the pixel kernel appends compact 272-byte
//    half-records (id, centre, w[32], t[32]; only for instance-halves with an active pixel) to
//    per-(tile, warp) streams in a scratch buffer, the channel kernel streams them back through
//    a ring of 4 KB TMA bulk copies.  An earlier single-kernel,
//    warp-specialised version (named-barrier hand-off through shared memory) spent 46 % of its
//    warp samples in barrier stalls (profiles/r01_render_bwd_v1.md); the scratch round trip is
//    ~0.6 GB of mostly L2-resident traffic per iteration at cfgB.
//  * Gaussian records / feature rows are gathered into shared memory two batches ahead by 16-byte asynchronous
//    copies (cp.async, LDGSTS) issued by both warps; the channel kernels stream the contiguous half-records back
//    with TMA bulk copies (cp.async.bulk + mbarrier).
//  * the tile's list is cut at tile_last = max over the tile's pixels of n_contrib (written
//    by the forward): instances behind it contribute to no pixel.
//
// Summation order differs from the reference's atomics (and from run to run), which is why
// the gradient parity gate is 1e-3 relative.
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"

namespace lgs {

constexpr int BB = 32;       // Gaussians per TMA batch of the pixel kernel
constexpr int BSTAGES = 2;   // its staging ring depth (2 x 9.7 KB: 8 CTAs per SM; 3 stages would allow only 7)
constexpr int CB = 8;        // half-records per TMA batch of the channel kernel
constexpr int CSTAGES = 4;   // its staging ring depth

// gradient arrays for the pixel kernel to clear on behalf of the channel kernel (launch_render_bwd fills it)
constexpr int ZT_MAX = 6;
struct ZeroTargets {
    float* ptr[ZT_MAX];        // 16-byte aligned array starts
    size_t len16[ZT_MAX];      // whole 16-byte vectors in each
    float* tail_ptr[ZT_MAX];   // the floats behind the last whole vector of each array ...
    int tail_len[ZT_MAX];      // ... 0 to 3 of them
    int n, n_tail;             // arrays; tail slots (4 per array)
    size_t n16;                // total vectors
};

template <bool WITH_LF>
struct BwdStage {
    GaussRec rec[BB];
    float lf[WITH_LF ? BB * LF : 4];
};

// ================================ pixel kernel =====================================================
// One CTA (2 warps) per tile, one pixel per thread.  Walks the tile's list back to front, replays
// alpha, evaluates d_j = <feature_j, g_pix>, runs the scalar recurrences and, for every
// (instance, pixel warp) with at least one active pixel, APPENDS one 272-byte half-record
// {Gaussian id, centre relative to the tile, w = alpha*T and t = G*dL/dalpha of the warp's 32
// pixels} to that (tile, warp)'s contiguous stream in the scratch buffer.  Streams need no atomics:
// each warp owns its stream and counts in a register; the count is stored once at the end.
template <bool WITH_LF>
__global__ void __launch_bounds__(TILE_PIX)
render_bwd_pix_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
                      const float* __restrict__ bg, const GaussRec* __restrict__ rec,
                      const float* __restrict__ lang_feat, const float* __restrict__ final_T,
                      const uint32_t* __restrict__ n_contrib, const uint32_t* __restrict__ tile_last,
                      const float* __restrict__ dL_dpix, const float* __restrict__ dL_dpix_lf,
                      const float* __restrict__ dL_dpix_depth, float* __restrict__ hrec_buf,
                      uint32_t* __restrict__ hrec_count, uint32_t* __restrict__ work_counter, const ZeroTargets zt) {
    using Stage = BwdStage<WITH_LF>;
    __shared__ __align__(128) Stage stages[BSTAGES];
    __shared__ uint32_t s_ids[BSTAGES][BB];  // Gaussian ids of the batches in flight (batch b in slot b % BSTAGES)

    const int tid = threadIdx.x;
    const int lane = tid & 31, wrp = tid >> 5;
    const int tile_id = blockIdx.y * gridDim.x + blockIdx.x;
    if (tile_id == 0 && tid == 0) *work_counter = 0;  // the tensor-core channel kernel's tile queue (next launch in stream order)
    // Side job: clear the gradient arrays the CHANNEL kernel (the next launch) accumulates into.  This kernel never touches
    // them, its stores are fire-and-forget underneath its own latency-bound work, and it saves the separate 152 MB memset
    // launch (0.028 ms at cfgB) that the reference's caller does with torch::zeros (rasterize_points.cu:157-167).
    if (zt.n16 > 0) {
        const size_t gt = (size_t)tile_id * TILE_PIX + tid, T = (size_t)gridDim.x * gridDim.y * TILE_PIX;
        // the first array (the feature gradients, 90 % of the bytes) with a plain strided loop, the small ones with a lookup
        float4* const p0 = reinterpret_cast<float4*>(zt.ptr[0]);
        const size_t n0 = zt.len16[0];
#pragma unroll 4
        for (size_t i = gt; i < n0; i += T) p0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (size_t i = gt; i < zt.n16 - n0; i += T) {
            size_t k = i;
            int a = 1;
            while (a + 1 < zt.n && k >= zt.len16[a]) { k -= zt.len16[a]; ++a; }
            reinterpret_cast<float4*>(zt.ptr[a])[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (size_t i = gt; i < (size_t)zt.n_tail; i += T)
            if ((int)(i & 3) < zt.tail_len[i >> 2]) zt.tail_ptr[i >> 2][i & 3] = 0.f;
    }
    const uint2 range = ranges[tile_id];
    const int n_all = (int)(range.y - range.x);
    const int n = min(n_all, (int)tile_last[tile_id]);  // entries behind tile_last touch no pixel
    if (n <= 0) {
        if (tid < 2) hrec_count[2 * tile_id + tid] = 0;
        return;
    }
    const int nb = (n + BB - 1) / BB;
    const size_t HW = (size_t)H * W;

    const uint32_t pxi = blockIdx.x * TILE + (tid & 7);
    const uint32_t pyi = blockIdx.y * TILE + (tid >> 3);
    const bool inside = pxi < (uint32_t)W && pyi < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * pyi + pxi;
    const float pxf = (float)pxi, pyf = (float)pyi;
    const float cxf = (float)(blockIdx.x * TILE) + 3.5f, cyf = (float)(blockIdx.y * TILE) + 3.5f;
    const float wx0 = (float)(blockIdx.x * TILE), wy0 = (float)(blockIdx.y * TILE + 4 * wrp);  // this warp's 8x4 pixels

    const float T_final = inside ? final_T[pix_id] : 0.f;
    const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;
    float T = T_final;

    // the pixel's 64 feature gradients as register pairs: the dot product below runs on packed FFMA2 (sm_100)
    float2 g_lf2[WITH_LF ? LF / 2 : 1];
    float g_r = 0.f, g_g = 0.f, g_b = 0.f, g_d = 0.f;
#pragma unroll
    for (int k = 0; k < (WITH_LF ? LF / 2 : 1); ++k) g_lf2[k] = make_float2(0.f, 0.f);
    if (inside) {
        g_r = dL_dpix[0 * HW + pix_id];
        g_g = dL_dpix[1 * HW + pix_id];
        g_b = dL_dpix[2 * HW + pix_id];
        g_d = dL_dpix_depth[pix_id];
        if (WITH_LF) {
#pragma unroll
            for (int k = 0; k < LF / 2; ++k)
                g_lf2[k] = make_float2(dL_dpix_lf[(size_t)(2 * k) * HW + pix_id], dL_dpix_lf[(size_t)(2 * k + 1) * HW + pix_id]);
        }
    }
    const float bgdot = bg[0] * g_r + bg[1] * g_g + bg[2] * g_b;  // backward.cu:585-588
    float Acc = 0.f, last_alpha = 0.f, last_d = 0.f;

    // this warp's half-record stream: capacity n_all records, after the other warp's
    float* out = hrec_buf + ((size_t)2 * range.x + (size_t)wrp * n_all) * HREC_FLOATS;
    uint32_t nrec = 0;

    // ---- staging: batch b = list positions n-1-b*BB ... (back to front); its records and feature rows are gathered into
    // stages[b % BSTAGES] by 16-byte asynchronous copies spread over BOTH warps (19 chunks per Gaussian), two batches
    // ahead.  (One TMA bulk copy per row, issued by warp 0, made that warp the critical path: 19 % of all warp samples
    // were warp 1 waiting at the batch barrier, profiles/r01_bwd_pix_v6_ncu.txt.)
    constexpr int CH = 3 + (WITH_LF ? LF / 4 : 0);  // 16-byte chunks per Gaussian: record, feature row
    auto put_ids = [&](int b) {  // threads 0..BB-1: ids of batch b -> s_ids[b % BSTAGES]
        const int pos = b * BB + tid;
        if (tid < BB && pos < n) s_ids[b % BSTAGES][tid] = point_list[range.x + (n - 1 - pos)];
    };
    // chunk k = tid, tid + 64, ... of a batch is chunk q = k % CH of Gaussian j = k / CH; stepping k by 64 steps (j, q) by
    // (64 / CH, 64 % CH) with one carry -- no division in the copy loop
    const int j_first = tid / CH, q_first = tid - j_first * CH;
    const char* const rec_g = reinterpret_cast<const char*>(rec);
    const char* const lf_g = reinterpret_cast<const char*>(lang_feat);
    auto issue = [&](int b) {  // all threads; one commit group per call, empty past the end
        if (b < nb) {
            const int cnt = min(BB, n - b * BB);
            Stage& S = stages[b % BSTAGES];
            const uint32_t rec_s = smem_u32(&S.rec[0]), lf_s = smem_u32(&S.lf[0]);
            const uint32_t* ids = s_ids[b % BSTAGES];
            int j = j_first, q = q_first;
            while (j < cnt) {
                const uint32_t id = ids[j];
                if (q < 3) cp_async16(rec_s + j * (uint32_t)sizeof(GaussRec) + q * 16, rec_g + (size_t)id * sizeof(GaussRec) + q * 16);
                else cp_async16(lf_s + j * (LF * 4) + (q - 3) * 16, lf_g + (size_t)id * (LF * 4) + (q - 3) * 16);
                j += TILE_PIX / CH;
                q += TILE_PIX % CH;
                if (q >= CH) { q -= CH; ++j; }
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int b = 0; b < BSTAGES; ++b) put_ids(b);
    __syncthreads();
#pragma unroll
    for (int b = 0; b < BSTAGES - 1; ++b) issue(b);

    for (int b = 0; b < nb; ++b) {
        const int cnt = min(BB, n - b * BB);
        const int hi = n - 1 - b * BB;  // list position of slot 0
        cp_async_wait<BSTAGES - 2>();   // this thread's copies of batch b have landed
        __syncthreads();                // ... everyone's; and both warps are done with stage (b-1) % BSTAGES -> refill it
        issue(b + BSTAGES - 1);
        put_ids(b + BSTAGES);           // slot b % BSTAGES: read by issue(b) two barriers ago
        const Stage& S = stages[b % BSTAGES];

        // lane j tests Gaussian j's alpha >= 1/255 ellipse against this warp's 8x4 pixels (common.cuh)
        bool touch = false;
        if (lane < cnt) {
            const float4 t0 = S.rec[lane].q0, t1 = S.rec[lane].q1;
            touch = footprint_touches(t0.x, t0.y, t1.x, t1.y, t1.z, t1.w, wx0, wx0 + 7.0f, wy0, wy0 + 3.0f);
        }
        const uint32_t vis = __ballot_sync(0xffffffffu, touch);

        // Visible Gaussians four at a time: the alpha replays (load, falloff, expf) are stateless, independent
        // chains that overlap; the scalar recurrences then run in list order.
        uint32_t v = vis;
#pragma unroll 1
        while (v != 0) {
            int jj[4];
            float al[4], Gs[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool valid = v != 0;
                jj[u] = valid ? (__ffs(v) - 1) : 0;
                v &= v - 1;  // 0 stays 0
                const float4 q0 = S.rec[jj[u]].q0;  // x, y, depth, id bits
                const float4 q1 = S.rec[jj[u]].q1;
                float dx, dy;
                const float power = eval_power(q0.x, q0.y, pxf, pyf, q1.x, q1.y, q1.z, dx, dy);
                Gs[u] = expf(power);
                al[u] = fminf(0.99f, __fmul_rn(q1.w, Gs[u]));
                // backward.cu:513-530: behind the pixel's last contributor, outside the falloff, or below
                // the alpha threshold -> no contribution
                ok[u] = valid && ((hi - jj[u]) < last_contributor) && !(power > 0.0f) && !(al[u] < 1.0f / 255.0f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool act = ok[u];
                if (!__any_sync(0xffffffffu, act)) continue;
                const int j = jj[u];
                const float alpha = al[u], G = Gs[u];
                const float4 q0 = S.rec[j].q0;
                const float4 q2 = S.rec[j].q2;
                // four interleaved partial sums d0..d3 (component-wise identical to scalar fmaf chains), two per FFMA2
                float2 d01 = make_float2(q2.x * g_r, q2.y * g_g), d23 = make_float2(q2.z * g_b, q0.z * g_d);
                if (WITH_LF) {
                    const float4* f4 = reinterpret_cast<const float4*>(&S.lf[j * LF]);
#pragma unroll
                    for (int k = 0; k < LF / 4; ++k) {
                        const float4 f = f4[k];
                        d01 = __ffma2_rn(make_float2(f.x, f.y), g_lf2[2 * k], d01);
                        d23 = __ffma2_rn(make_float2(f.z, f.w), g_lf2[2 * k + 1], d23);
                    }
                }
                const float d = (d01.x + d01.y) + (d23.x + d23.y);
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
                float* dst = out + (size_t)nrec * HREC_FLOATS;
                if (lane == 0) __stcg(reinterpret_cast<float4*>(dst), make_float4(q0.x - cxf, q0.y - cyf, 0.f, q0.w));
                __stcg(dst + 4 + lane, Wv);
                __stcg(dst + 36 + lane, Tv);
                ++nrec;
            }
        }
    }
    if (lane == 0) hrec_count[2 * tile_id + wrp] = nrec;
}

// ================================ channel kernel ===================================================
// One single-warp CTA per (tile, column group): group 0/1 = feature channels 0-31 / 32-63, group 2 =
// 3 colour + depth + 6 moment columns.  Every lane owns one column of the tile's [64 pixel x 74]
// matrix in 64 registers and reduces the pixel kernel's half-records against it: ONE
// red.global.add per (half-record, column).  Records arrive through the warp's own ring of TMA
// bulk copies (one 2 KB copy per batch of 8 records); there is no CTA-wide barrier anywhere, so a
// slow group never stalls the others (the first version with one 3-warp CTA per tile and a
// __syncthreads per batch spent 47 % of its warp samples in barrier stalls).
//
// Moment lanes accumulate sums of t, t*dx, t*dy, t*dx^2, t*dx*dy, t*dy^2 (dx, dy = Gaussian centre
// minus pixel) into dL_dopacity / dL_dmean2D.xy / dL_dconic.{x,y,w}; the per-Gaussian factors
// (opacity, conic, viewport scale) are applied once per Gaussian by the preprocess backward.
template <bool WITH_LF>
__global__ void __launch_bounds__(32)
render_bwd_chan_kernel(const uint2* __restrict__ ranges, int W, int H, int tiles_x,
                       const float* __restrict__ dL_dpix, const float* __restrict__ dL_dpix_lf,
                       const float* __restrict__ dL_dpix_depth, const float* __restrict__ hrec_buf,
                       const uint32_t* __restrict__ hrec_count, float* __restrict__ dL_dmean2D,
                       float* __restrict__ dL_dconic, float* __restrict__ dL_dopacity,
                       float* __restrict__ dL_dcolor, float* __restrict__ dL_dlang_feat,
                       float* __restrict__ dL_ddepth) {
    __shared__ __align__(128) float stages[CSTAGES][CB * HREC_FLOATS];
    __shared__ __align__(8) uint64_t full_bar[CSTAGES];

    const int lane = threadIdx.x;
    const int tile_id = WITH_LF ? (int)(blockIdx.x / 3) : (int)blockIdx.x;
    const int grp = WITH_LF ? (int)(blockIdx.x % 3) : 2;
    const uint32_t cnt0 = hrec_count[2 * tile_id], cnt1 = hrec_count[2 * tile_id + 1];
    if (cnt0 + cnt1 == 0) return;
    const uint2 range = ranges[tile_id];
    const int n_all = (int)(range.y - range.x);
    const size_t HW = (size_t)H * W;
    const uint32_t tx0 = (uint32_t)(tile_id % tiles_x) * TILE, ty0 = (uint32_t)(tile_id / tiles_x) * TILE;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < CSTAGES; ++s) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }

    const int col = grp * 32 + lane;  // 0..63 lf, 64..66 rgb, 67 depth, 68..73 moments, 74.. idle
    float colv[TILE_PIX];
#pragma unroll
    for (int i = 0; i < TILE_PIX; ++i) {
        const uint32_t pxi = tx0 + (i & 7);
        const uint32_t pyi = ty0 + (i >> 3);
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
            x = k == 0 ? 1.f : k == 1 ? u : k == 2 ? v : k == 3 ? u * u : k == 4 ? u * v : k == 5 ? v * v : 0.f;
        }
        colv[i] = x;
    }
    const int toff = (col >= LF + 4) ? 36 : 4;  // moment columns reduce t, the others reduce w
    // per-lane output slot: out_base[id * out_stride]
    float* out_base = nullptr;
    uint32_t out_stride = 0;
    if (col < LF) { out_base = dL_dlang_feat + col; out_stride = LF; }
    else if (col < LF + 3) { out_base = dL_dcolor + (col - LF); out_stride = 3; }
    else if (col == LF + 3) { out_base = dL_ddepth; out_stride = 1; }          // may be NULL
    else if (col == LF + 4) { out_base = dL_dopacity; out_stride = 1; }         // sum t
    else if (col == LF + 5) { out_base = dL_dmean2D; out_stride = 3; }          // sum t*dx
    else if (col == LF + 6) { out_base = dL_dmean2D + 1; out_stride = 3; }      // sum t*dy
    else if (col == LF + 7) { out_base = dL_dconic; out_stride = 4; }           // sum t*dx*dx
    else if (col == LF + 8) { out_base = dL_dconic + 1; out_stride = 4; }       // sum t*dx*dy
    else if (col == LF + 9) { out_base = dL_dconic + 3; out_stride = 4; }       // sum t*dy*dy
    // Moment lanes hold S0, Su, Sv, Suu, Suv, Svv with u, v = pixel - tile centre.  With dx = gx - u,
    // dy = gy - v:  sum t*dx = gx*S0 - Su,  sum t*dx^2 = gx^2*S0 - 2*gx*Su + Suu,  ... written branch-free as
    //   out = k_own*own + S0*(kxx*gx*gx + kxy*gx*gy + kyy*gy*gy + kx*gx + ky*gy) + Su*(cux*gx + cuy*gy) + Sv*(cvx*gx + cvy*gy)
    const int mk = col - (LF + 4);
    const float k_own = (mk == 1 || mk == 2) ? -1.f : 1.f;
    const float kx = mk == 1 ? 1.f : 0.f, ky = mk == 2 ? 1.f : 0.f;
    const float kxx = mk == 3 ? 1.f : 0.f, kxy = mk == 4 ? 1.f : 0.f, kyy = mk == 5 ? 1.f : 0.f;
    const float cux = mk == 3 ? -2.f : 0.f, cuy = mk == 4 ? -1.f : 0.f;
    const float cvx = mk == 4 ? -1.f : 0.f, cvy = mk == 5 ? -2.f : 0.f;
    __syncwarp();

    int gb = 0;  // global batch counter: the ring's stage / parity sequence runs across both halves
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        const int n = (int)(half == 0 ? cnt0 : cnt1);
        if (n == 0) continue;
        const int nb = (n + CB - 1) / CB;
        const float* src_base = hrec_buf + ((size_t)2 * range.x + (size_t)half * n_all) * HREC_FLOATS;
        auto issue = [&](int b) {  // b = batch within this half
            if (lane == 0) {
                const int cnt = min(CB, n - b * CB);
                const int s = (gb + b) % CSTAGES;
                const uint32_t bytes = (uint32_t)cnt * HREC_FLOATS * 4u;
                mbar_arrive_expect_tx(&full_bar[s], bytes);
                tma_bulk_g2s(&stages[s][0], src_base + (size_t)b * CB * HREC_FLOATS, bytes, &full_bar[s]);
            }
        };
        for (int b = 0; b < CSTAGES - 1 && b < nb; ++b) issue(b);
        for (int b = 0; b < nb; ++b) {
            const int cnt = min(CB, n - b * CB);
            __syncwarp();  // all lanes are done with the stage about to be refilled (batch b-1's)
            if (b + CSTAGES - 1 < nb) issue(b + CSTAGES - 1);
            const int s = (gb + b) % CSTAGES;
            mbar_wait(&full_bar[s], (uint32_t)(((gb + b) / CSTAGES) & 1));
            const float* S = &stages[s][0];
            // two records per iteration: their load -> FMA chains are independent and overlap
#pragma unroll 1
            for (int j = 0; j < cnt; j += 2) {
                const bool two = j + 1 < cnt;
                const float* r0 = S + j * HREC_FLOATS;
                const float* r1 = S + (two ? j + 1 : j) * HREC_FLOATS;
                const float4 hd0 = *reinterpret_cast<const float4*>(r0);  // gx, gy, -, id
                const float4 hd1 = *reinterpret_cast<const float4*>(r1);
                const float4* s0 = reinterpret_cast<const float4*>(r0 + toff);
                const float4* s1 = reinterpret_cast<const float4*>(r1 + toff);
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
                if (half == 0) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float4 x = s0[k], y = s1[k];
                        a0 = fmaf(x.x, colv[4 * k + 0], a0); b0 = fmaf(y.x, colv[4 * k + 0], b0);
                        a1 = fmaf(x.y, colv[4 * k + 1], a1); b1 = fmaf(y.y, colv[4 * k + 1], b1);
                        a2 = fmaf(x.z, colv[4 * k + 2], a2); b2 = fmaf(y.z, colv[4 * k + 2], b2);
                        a3 = fmaf(x.w, colv[4 * k + 3], a3); b3 = fmaf(y.w, colv[4 * k + 3], b3);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float4 x = s0[k], y = s1[k];
                        a0 = fmaf(x.x, colv[32 + 4 * k + 0], a0); b0 = fmaf(y.x, colv[32 + 4 * k + 0], b0);
                        a1 = fmaf(x.y, colv[32 + 4 * k + 1], a1); b1 = fmaf(y.y, colv[32 + 4 * k + 1], b1);
                        a2 = fmaf(x.z, colv[32 + 4 * k + 2], a2); b2 = fmaf(y.z, colv[32 + 4 * k + 2], b2);
                        a3 = fmaf(x.w, colv[32 + 4 * k + 3], a3); b3 = fmaf(y.w, colv[32 + 4 * k + 3], b3);
                    }
                }
                float acc0 = (a0 + a1) + (a2 + a3), acc1 = (b0 + b1) + (b2 + b3);
                if (grp == 2) {
                    const float S0a = __shfl_sync(0xffffffffu, acc0, 4), S0b = __shfl_sync(0xffffffffu, acc1, 4);
                    const float Sua = __shfl_sync(0xffffffffu, acc0, 5), Sub = __shfl_sync(0xffffffffu, acc1, 5);
                    const float Sva = __shfl_sync(0xffffffffu, acc0, 6), Svb = __shfl_sync(0xffffffffu, acc1, 6);
                    {
                        const float gx = hd0.x, gy = hd0.y;
                        const float c0 = gx * (kxx * gx + kxy * gy + kx) + gy * (kyy * gy + ky);
                        acc0 = k_own * acc0 + S0a * c0 + Sua * (cux * gx + cuy * gy) + Sva * (cvx * gx + cvy * gy);
                    }
                    {
                        const float gx = hd1.x, gy = hd1.y;
                        const float c0 = gx * (kxx * gx + kxy * gy + kx) + gy * (kyy * gy + ky);
                        acc1 = k_own * acc1 + S0b * c0 + Sub * (cux * gx + cuy * gy) + Svb * (cvx * gx + cvy * gy);
                    }
                }
                if (out_base != nullptr) {
                    red_add_f32(out_base + (size_t)__float_as_uint(hd0.w) * out_stride, acc0);
                    if (two) red_add_f32(out_base + (size_t)__float_as_uint(hd1.w) * out_stride, acc1);
                }
            }
        }
        gb += nb;
        __syncwarp();
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

// LGS_BWD_CHAN=simt (environment, read at the first launch) keeps the SIMT channel kernel for the 64-D feature path: a
// differential-debugging aid for the tensor-core kernel (render_bwd_tc.cu), not a fallback for other hardware.
static bool env_simt_chan() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("LGS_BWD_CHAN");
        v = (e && e[0] == 's') ? 1 : 0;
    }
    return v == 1;
}

size_t render_bwd_scratch_bytes(int R, int W, int H) {
    const size_t n = (size_t)(R > 0 ? R : 1);
    const size_t tiles = (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    return 2 * n * HREC_FLOATS * sizeof(float) + tiles * 2 * sizeof(uint32_t) + 1024;
}

int launch_render_bwd(int P, int W, int H, int R, const GeomState& g, const BinningState& b,
                      const ImageState& im, const float* background, const float* lang_feat,
                      const float* dL_dpix, const float* dL_dpix_lf, const float* dL_dpix_depth,
                      float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                      float* dL_dlang_feat, float* dL_ddepth, bool include_lf, char* scratch, bool zero_outputs, cudaStream_t s) {
    const dim3 grid((W + TILE - 1) / TILE, (H + TILE - 1) / TILE, 1);
    // the accumulated-into arrays are cleared by the pixel kernel itself (see there) when the caller asked for it
    ZeroTargets zt;
    zt.n = zt.n_tail = 0;
    zt.n16 = 0;
    if (zero_outputs) {
        struct { float* p; size_t n; } arr[ZT_MAX] = {{include_lf ? dL_dlang_feat : nullptr, (size_t)P * LF}, {dL_dconic, (size_t)P * 4},
                                                      {dL_dmean2D, (size_t)P * 3}, {dL_dcolor, (size_t)P * 3},
                                                      {dL_dopacity, (size_t)P}, {dL_ddepth, (size_t)P}};
        bool ok = true;
        for (int a = 0; a < ZT_MAX; ++a)
            if (arr[a].p && (reinterpret_cast<uintptr_t>(arr[a].p) & 15u)) ok = false;
        if (!ok) {  // an unaligned array: the plain memset kernel
            const int st = launch_zero_grads(P, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dlang_feat, dL_ddepth, include_lf, s);
            if (st != LGS_OK) return st;
        } else {
            for (int a = 0; a < ZT_MAX; ++a) {
                if (!arr[a].p) continue;
                zt.ptr[zt.n] = arr[a].p;
                zt.len16[zt.n] = arr[a].n / 4;
                zt.n16 += arr[a].n / 4;
                zt.tail_ptr[zt.n] = arr[a].p + (arr[a].n & ~(size_t)3);
                zt.tail_len[zt.n] = (int)(arr[a].n & 3);
                ++zt.n;
            }
            zt.n_tail = 4 * zt.n;
        }
    }
    const unsigned tiles = grid.x * grid.y;
    // scratch: [2R] half-records of 272 B (256-byte aligned base) followed by [tiles][2] record counts
    uintptr_t base = (reinterpret_cast<uintptr_t>(scratch) + 255) & ~(uintptr_t)255;
    float* hrec = reinterpret_cast<float*>(base);
    uint32_t* hcount = reinterpret_cast<uint32_t*>(hrec + (size_t)2 * (R > 0 ? R : 1) * HREC_FLOATS);
    uint32_t* work_counter = hcount + 2 * (size_t)tiles;
    if (include_lf) {
        render_bwd_pix_kernel<true><<<grid, TILE_PIX, 0, s>>>(im.ranges, b.point_list, W, H, background, g.rec, lang_feat, im.final_T,
                                                              im.n_contrib, im.tile_last, dL_dpix, dL_dpix_lf, dL_dpix_depth, hrec,
                                                              hcount, work_counter, zt);
        LGS_LAUNCH_CHECK();
        prof_mark(PM_RENDER_BWD_PIX, s);
        if (!env_simt_chan())
            return launch_render_bwd_chan_tc(W, H, im, dL_dpix, dL_dpix_lf, dL_dpix_depth, hrec, hcount, work_counter, dL_dmean2D,
                                             dL_dconic, dL_dopacity, dL_dcolor, dL_dlang_feat, dL_ddepth, s);
        render_bwd_chan_kernel<true><<<3 * tiles, 32, 0, s>>>(im.ranges, W, H, (int)grid.x, dL_dpix, dL_dpix_lf, dL_dpix_depth,
                                                              hrec, hcount, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor,
                                                              dL_dlang_feat, dL_ddepth);
    } else {
        render_bwd_pix_kernel<false><<<grid, TILE_PIX, 0, s>>>(im.ranges, b.point_list, W, H, background, g.rec, lang_feat,
                                                               im.final_T, im.n_contrib, im.tile_last, dL_dpix, dL_dpix_lf,
                                                               dL_dpix_depth, hrec, hcount, work_counter, zt);
        LGS_LAUNCH_CHECK();
        prof_mark(PM_RENDER_BWD_PIX, s);
        render_bwd_chan_kernel<false><<<tiles, 32, 0, s>>>(im.ranges, W, H, (int)grid.x, dL_dpix, dL_dpix_lf, dL_dpix_depth,
                                                           hrec, hcount, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor,
                                                           dL_dlang_feat, dL_ddepth);
    }
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

}  // namespace lgs
