// loss.cu -- the mapper's loss and its gradient w.r.t. the rendered images, fused.
//
// Replaces, for one keyframe, the ~60 stock libtorch kernels of the reference's loss forward +
// backward (src/gaussian_mapper.cpp:707-724 with include/loss_utils.h:27-131):
//     gt_lf  = interpolate(keyframe feature map -> HxW)            (nearest)
//     image, lf, depth *= mask                                      (undistortion mask, optional)
//     loss   = (1-l)*L1(image, gt) + l*(1 - SSIM(image, gt)) +/- mean_px cos(lf, gt_lf) + L1(depth, gt_depth)
// and autograd's walk back through them, by four launches that read each image once:
//   loss_pix_kernel    per pixel: L1 terms, the 64-channel cosine (feature vector kept in registers, 16 channels per
//                      thread, low-res ground truth gathered), and dL/dlf, dL/ddepth, the L1 part of dL/dimage
//   ssim_fwd_kernel    per 16x16 tile and channel: separable 11-tap Gaussian statistics in shared memory,
//                      SSIM map sum, and the three partial-derivative maps the backward needs
//   ssim_bwd_kernel    per tile: separable filter of those maps, adds the SSIM part of dL/dimage
//   loss_finalize      assembles the scalar(s)
// The [64,H,W] feature image is read once and its gradient written once (157 MB at 640x480) instead of
// ~15 passes; the up-sampled ground-truth feature image (78.6 MB) is never materialised.
//
// Semantics follow torch: |x| has derivative sign(x) (0 at 0); cosine_similarity divides each vector by
// max(norm, 1e-8); SSIM windows are zero padded (conv2d padding = 5), C1 = 0.01^2, C2 = 0.03^2, sigma 1.5.
#include "common.cuh"

namespace lgs {

constexpr int SSIM_R = 5;          // window radius (11 taps)
constexpr int SSIM_T = 16;         // tile edge
constexpr int SSIM_E = SSIM_T + 2 * SSIM_R;  // tile + halo = 26

// the reference's 11-tap window (loss_utils.h:42-57: exp(-x^2 / (2 sigma^2)) in float, sigma 1.5, divided by its float sum),
// evaluated once with the same float expressions and kept as literals: no per-process / per-device initialisation state
__constant__ float c_gauss[2 * SSIM_R + 1] = {0.00102838036f, 0.00759875868f, 0.0360007733f, 0.109360695f, 0.213005543f, 0.266011745f,
                                              0.213005543f, 0.109360695f, 0.0360007733f, 0.00759875868f, 0.00102838036f};

__device__ __forceinline__ float sgn(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

__device__ __forceinline__ float block_sum(float v, float* red, int tid) {  // 256 threads, tid = linear thread id
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = tid >> 5, l = tid & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float r = (tid < 8) ? red[tid] : 0.f;
    if (w == 0) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;  // valid in thread tid == 0
}

// 256 threads = 64 pixels x 4 channel groups of 16: thread (grp, px) keeps 16 feature values in registers, the three
// dot products are completed across the groups through shared memory.  (One thread per pixel with all 64 channels
// in registers ran at 2 CTAs/SM and a third of the HBM rate: too few loads in flight per SM.)
constexpr int LP_PIX = 64;
constexpr int LP_GRP = 4;
constexpr int LP_CH = LF / LP_GRP;  // 16

__device__ __forceinline__ void
loss_pix_block(unsigned block, int W, int H, int lw, int lh, const float* __restrict__ image, const float* __restrict__ lf,
               const float* __restrict__ depth, const float* __restrict__ gt_image, const float* __restrict__ gt_lf,
               const float* __restrict__ gt_depth, const float* __restrict__ mask, float w_l1, float w_cos, float w_depth,
               float* __restrict__ dL_dimage, float* __restrict__ dL_dlf, float* __restrict__ dL_ddepth,
               float* __restrict__ acc) {
    __shared__ float red[8];
    __shared__ float part[3][LP_GRP][LP_PIX];
    const size_t HW = (size_t)H * W;
    const int px = threadIdx.x & (LP_PIX - 1), grp = threadIdx.x / LP_PIX;
    const size_t p = (size_t)block * LP_PIX + px;
    const bool live = p < HW;
    float s_l1 = 0.f, s_cos = 0.f, s_d = 0.f;
    const float m0 = (live && mask) ? mask[p] : 1.f;
    if (live && grp == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float mc = mask ? mask[c * HW + p] : 1.f;
            const float d = image[c * HW + p] * mc - gt_image[c * HW + p];
            s_l1 += fabsf(d);
            dL_dimage[c * HW + p] = w_l1 * sgn(d) * mc;
        }
        const float d = depth[p] * m0 - gt_depth[p];
        s_d = fabsf(d);
        dL_ddepth[p] = w_depth * sgn(d) * m0;
    }
    float a[LP_CH], bv[LP_CH];
    float w12 = 0.f, w1 = 0.f, w2 = 0.f;
    if (live) {
        // nearest-neighbour source pixel of the low-resolution ground-truth feature map (torch `nearest`)
        const int py = (int)(p / W), pxx = (int)(p - (size_t)py * W);
        const int sy = min((int)floorf(py * ((float)lh / (float)H)), lh - 1);
        const int sx = min((int)floorf(pxx * ((float)lw / (float)W)), lw - 1);
        const size_t bstride = (size_t)lh * lw;
        const float* b = gt_lf + (size_t)(grp * LP_CH) * bstride + (size_t)sy * lw + sx;
        const float* ap = lf + (size_t)(grp * LP_CH) * HW + p;
#pragma unroll
        for (int k = 0; k < LP_CH; ++k) {
            a[k] = __ldcs(ap + k * HW) * m0;
            bv[k] = __ldg(b + k * bstride);
        }
#pragma unroll
        for (int k = 0; k < LP_CH; ++k) {
            w12 = fmaf(a[k], bv[k], w12);
            w1 = fmaf(a[k], a[k], w1);
            w2 = fmaf(bv[k], bv[k], w2);
        }
    }
    part[0][grp][px] = w12;
    part[1][grp][px] = w1;
    part[2][grp][px] = w2;
    __syncthreads();
    if (live) {
        // fixed summation order over the groups: every thread of a pixel gets the same totals
        w12 = (part[0][0][px] + part[0][1][px]) + (part[0][2][px] + part[0][3][px]);
        w1 = (part[1][0][px] + part[1][1][px]) + (part[1][2][px] + part[1][3][px]);
        w2 = (part[2][0][px] + part[2][1][px]) + (part[2][2][px] + part[2][3][px]);
        const float na_true = sqrtf(w1);
        const float na = fmaxf(na_true, 1e-8f), nb = fmaxf(sqrtf(w2), 1e-8f);
        const float inv_na = 1.f / na, inv_nb = 1.f / nb;
        const float c = w12 * inv_na * inv_nb;
        if (grp == 0) s_cos = c;
        // d cos / d a = (b/|b| - cos * a/|a|) / |a|   (the second term vanishes where the norm is clamped)
        const float k_b = w_cos * m0 * inv_na * inv_nb;
        const float k_a = (na_true > 1e-8f) ? -w_cos * m0 * c * inv_na * inv_na : 0.f;
        float* gp = dL_dlf + (size_t)(grp * LP_CH) * HW + p;
#pragma unroll
        for (int k = 0; k < LP_CH; ++k) __stcs(gp + k * HW, fmaf(k_b, bv[k], k_a * a[k]));
    }
    const float t_l1 = block_sum(s_l1, red, threadIdx.x);
    const float t_cos = block_sum(s_cos, red, threadIdx.x);
    const float t_d = block_sum(s_d, red, threadIdx.x);
    if (threadIdx.x == 0) {
        atomicAdd(acc + 0, t_l1);
        atomicAdd(acc + 2, t_cos);
        atomicAdd(acc + 3, t_d);
    }
}

// Forward SSIM statistics for one 16x16 tile (bx, by) of one channel; x = image*mask, y = gt.  256 threads, tid = 16 * ty + tx.
__device__ __forceinline__ void
ssim_fwd_block(int bx, int by, int ch, int W, int H, const float* __restrict__ image, const float* __restrict__ mask,
               const float* __restrict__ gt, float* __restrict__ maps, float* __restrict__ acc) {
    __shared__ float sx[SSIM_E][SSIM_E + 1], sy[SSIM_E][SSIM_E + 1];
    __shared__ float h[5][SSIM_E][SSIM_T + 1];  // horizontally filtered x, y, xx, yy, xy
    __shared__ float red[8];
    const size_t HW = (size_t)H * W;
    const int x0 = bx * SSIM_T - SSIM_R, y0 = by * SSIM_T - SSIM_R;
    const int tid = threadIdx.x, tx = tid & (SSIM_T - 1), ty = tid / SSIM_T;
    for (int i = tid; i < SSIM_E * SSIM_E; i += SSIM_T * SSIM_T) {
        const int ly = i / SSIM_E, lx = i - ly * SSIM_E;
        const int gx = x0 + lx, gy = y0 + ly;
        float vx = 0.f, vy = 0.f;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            const size_t q = ch * HW + (size_t)gy * W + gx;
            vx = image[q] * (mask ? mask[q] : 1.f);
            vy = gt[q];
        }
        sx[ly][lx] = vx;
        sy[ly][lx] = vy;
    }
    __syncthreads();
    for (int i = tid; i < SSIM_E * SSIM_T; i += SSIM_T * SSIM_T) {
        const int ly = i / SSIM_T, lx = i - ly * SSIM_T;
        float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * SSIM_R; ++k) {
            const float w = c_gauss[k], u = sx[ly][lx + k], v = sy[ly][lx + k];
            a = fmaf(w, u, a); b = fmaf(w, v, b); aa = fmaf(w, u * u, aa); bb = fmaf(w, v * v, bb); ab = fmaf(w, u * v, ab);
        }
        h[0][ly][lx] = a; h[1][ly][lx] = b; h[2][ly][lx] = aa; h[3][ly][lx] = bb; h[4][ly][lx] = ab;
    }
    __syncthreads();
    const int gx = bx * SSIM_T + tx, gy = by * SSIM_T + ty;
    float ssim = 0.f;
    if (gx < W && gy < H) {
        float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * SSIM_R; ++k) {
            const float w = c_gauss[k];
            mu1 = fmaf(w, h[0][ty + k][tx], mu1);
            mu2 = fmaf(w, h[1][ty + k][tx], mu2);
            e11 = fmaf(w, h[2][ty + k][tx], e11);
            e22 = fmaf(w, h[3][ty + k][tx], e22);
            e12 = fmaf(w, h[4][ty + k][tx], e12);
        }
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        const float mu1s = mu1 * mu1, mu2s = mu2 * mu2, mu12 = mu1 * mu2;
        const float s1 = e11 - mu1s, s2 = e22 - mu2s, s12 = e12 - mu12;
        const float A = 2.f * mu12 + C1, B = 2.f * s12 + C2, Cc = mu1s + mu2s + C1, D = s1 + s2 + C2;
        const float inv_cd = 1.f / (Cc * D);
        ssim = A * B * inv_cd;
        // partial derivatives of the map w.r.t. mu1 (total, through sigma1^2 and sigma12), E[x^2], E[xy]
        const float d_s1 = -ssim / D;
        const float d_s12 = 2.f * A * inv_cd;
        const float d_mu1 = 2.f * mu2 * B * inv_cd - 2.f * mu1 * ssim / Cc - 2.f * mu1 * d_s1 - mu2 * d_s12;
        const size_t q = ch * HW + (size_t)gy * W + gx;
        maps[q] = d_mu1;
        maps[3 * HW + q] = d_s1;
        maps[6 * HW + q] = d_s12;
    }
    const float t = block_sum(ssim, red, tid);
    if (tid == 0) atomicAdd(acc + 1, t);
}


// One launch for the two independent halves of the loss forward: the per-pixel terms (HBM-bound: 157 MB of feature image in
// and gradient out) and the SSIM statistics (shared-memory / latency-bound, 11 MB): their CTAs are interleaved in one grid so
// that the SSIM tiles run underneath the feature stream instead of after it (0.080 -> ~0.055 ms at 640x480).
__global__ void __launch_bounds__(LP_PIX * LP_GRP)
loss_fwd_kernel(unsigned n_pix_blocks, unsigned n_ssim_blocks, int ssim_gx, int ssim_gy, int W, int H, int lw, int lh,
                const float* __restrict__ image, const float* __restrict__ lf, const float* __restrict__ depth,
                const float* __restrict__ gt_image, const float* __restrict__ gt_lf, const float* __restrict__ gt_depth,
                const float* __restrict__ mask, float w_l1, float w_cos, float w_depth, float* __restrict__ dL_dimage,
                float* __restrict__ dL_dlf, float* __restrict__ dL_ddepth, float* __restrict__ maps, float* __restrict__ acc) {
    static_assert(LP_PIX * LP_GRP == SSIM_T * SSIM_T, "both halves use 256-thread CTAs");
    const unsigned b = blockIdx.x, m = n_pix_blocks < n_ssim_blocks ? n_pix_blocks : n_ssim_blocks;
    bool ssim;
    unsigned k;
    if (b < 2 * m) { ssim = (b & 1u) == 0; k = b >> 1; }              // alternate while both kinds last
    else { ssim = n_ssim_blocks > n_pix_blocks; k = b - m; }          // then the rest of the larger kind
    if (ssim) {
        const int per = ssim_gx * ssim_gy, ch = (int)k / per, r = (int)k - ch * per;
        ssim_fwd_block(r % ssim_gx, r / ssim_gx, ch, W, H, image, mask, gt_image, maps, acc);
    } else {
        loss_pix_block(k, W, H, lw, lh, image, lf, depth, gt_image, gt_lf, gt_depth, mask, w_l1, w_cos, w_depth, dL_dimage, dL_dlf,
                       dL_ddepth, acc);
    }
}

// dL/dx += w_ssim * mask * ( conv(d_mu1) + 2 x conv(d_s1) + y conv(d_s12) )
__global__ void __launch_bounds__(SSIM_T * SSIM_T)
ssim_bwd_kernel(int W, int H, const float* __restrict__ image, const float* __restrict__ mask,
                const float* __restrict__ gt, const float* __restrict__ maps, float w_ssim,
                float* __restrict__ dL_dimage) {
    __shared__ float s[3][SSIM_E][SSIM_E + 1];
    __shared__ float h[3][SSIM_E][SSIM_T + 1];
    const size_t HW = (size_t)H * W;
    const int ch = blockIdx.z;
    const int x0 = blockIdx.x * SSIM_T - SSIM_R, y0 = blockIdx.y * SSIM_T - SSIM_R;
    const int tid = threadIdx.y * SSIM_T + threadIdx.x;
    for (int i = tid; i < SSIM_E * SSIM_E; i += SSIM_T * SSIM_T) {
        const int ly = i / SSIM_E, lx = i - ly * SSIM_E;
        const int gx = x0 + lx, gy = y0 + ly;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            const size_t q = ch * HW + (size_t)gy * W + gx;
            v0 = maps[q]; v1 = maps[3 * HW + q]; v2 = maps[6 * HW + q];
        }
        s[0][ly][lx] = v0; s[1][ly][lx] = v1; s[2][ly][lx] = v2;
    }
    __syncthreads();
    for (int i = tid; i < SSIM_E * SSIM_T; i += SSIM_T * SSIM_T) {
        const int ly = i / SSIM_T, lx = i - ly * SSIM_T;
        float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * SSIM_R; ++k) {
            const float w = c_gauss[k];
            a = fmaf(w, s[0][ly][lx + k], a); b = fmaf(w, s[1][ly][lx + k], b); c = fmaf(w, s[2][ly][lx + k], c);
        }
        h[0][ly][lx] = a; h[1][ly][lx] = b; h[2][ly][lx] = c;
    }
    __syncthreads();
    const int gx = blockIdx.x * SSIM_T + threadIdx.x, gy = blockIdx.y * SSIM_T + threadIdx.y;
    if (gx < W && gy < H) {
        float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * SSIM_R; ++k) {
            const float w = c_gauss[k];
            a = fmaf(w, h[0][threadIdx.y + k][threadIdx.x], a);
            b = fmaf(w, h[1][threadIdx.y + k][threadIdx.x], b);
            c = fmaf(w, h[2][threadIdx.y + k][threadIdx.x], c);
        }
        const size_t q = ch * HW + (size_t)gy * W + gx;
        const float m = mask ? mask[q] : 1.f;
        const float x = image[q] * m, y = gt[q];
        dL_dimage[q] += w_ssim * m * (a + 2.f * x * b + y * c);
    }
}

__global__ void loss_finalize_kernel(const float* __restrict__ acc, float n_img, float n_pix, float lambda, int cos_sign,
                                     float* __restrict__ out) {
    const float l1 = acc[0] / n_img, ssim = acc[1] / n_img, cs = acc[2] / n_pix, dp = acc[3] / n_pix;
    const float cos_term = cos_sign >= 0 ? cs : 1.f - cs;
    out[0] = (1.f - lambda) * l1 + lambda * (1.f - ssim) + cos_term + dp;
    out[1] = l1; out[2] = ssim; out[3] = cs; out[4] = dp;
}

// ---- activations (reference src/gaussian_model.cpp:46-68) -------------------------------------------
// forward: scales = exp(scaling), rotations = normalize(rotation) (F.normalize, eps 1e-12), opacities =
// sigmoid(opacity), shs = cat(features_dc [P,1,3], features_rest [P,15,3]).  One thread per Gaussian.
__global__ void __launch_bounds__(256)
activations_fwd_kernel(int P, const float* __restrict__ scaling, const float* __restrict__ rotation,
                       const float* __restrict__ opacity, float* __restrict__ scales, float* __restrict__ rots,
                       float* __restrict__ opac) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) scales[3 * (size_t)i + k] = expf(scaling[3 * (size_t)i + k]);
    const float4 q = reinterpret_cast<const float4*>(rotation)[i];
    const float inv = 1.f / fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
    reinterpret_cast<float4*>(rots)[i] = make_float4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
    opac[i] = 1.f / (1.f + expf(-opacity[i]));
}

// shs [P, 1+n_rest, 3] = cat(f_dc [P,1,3], f_rest [P,n_rest,3]) and its inverse, one ELEMENT per thread so
// that both sides are coalesced (a thread-per-Gaussian copy strides 192 B between lanes).
template <bool SCATTER>
__global__ void __launch_bounds__(256)
sh_cat_kernel(long long n, int row, const float* __restrict__ a_dc, const float* __restrict__ a_rest,
              float* __restrict__ shs, const float* __restrict__ g_shs, float* __restrict__ g_dc,
              float* __restrict__ g_rest, int accumulate) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long g = i / row;
        const int c = (int)(i - g * row);
        if (!SCATTER) {
            shs[i] = c < 3 ? a_dc[g * 3 + c] : a_rest[g * (row - 3) + (c - 3)];
        } else {
            float* dst = c < 3 ? g_dc + g * 3 + c : g_rest + g * (row - 3) + (c - 3);
            const float v = __ldcs(g_shs + i);
            *dst = accumulate ? *dst + v : v;
        }
    }
}

// backward: raw-parameter gradients from the gradients w.r.t. the activated tensors; `accumulate` adds
// to the outputs (several views per iteration) instead of overwriting them.
__global__ void __launch_bounds__(256)
activations_bwd_kernel(int P, int accumulate, const float* __restrict__ rotation,
                       const float* __restrict__ scales, const float* __restrict__ opac,
                       const float* __restrict__ g_scales, const float* __restrict__ g_rots,
                       const float* __restrict__ g_opac, float* __restrict__ g_scaling,
                       float* __restrict__ g_rotation, float* __restrict__ g_opacity) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float keep = accumulate ? 1.f : 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {  // d exp(x) = exp(x)
        const size_t j = 3 * (size_t)i + k;
        g_scaling[j] = fmaf(keep, accumulate ? g_scaling[j] : 0.f, g_scales[j] * scales[j]);
    }
    {  // F.normalize backward: (g - n (n.g)) / |q|
        const float4 q = reinterpret_cast<const float4*>(rotation)[i];
        const float4 g = reinterpret_cast<const float4*>(g_rots)[i];
        const float nrm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
        const float d = fmaxf(nrm, 1e-12f), inv = 1.f / d;
        const float nx = q.x * inv, ny = q.y * inv, nz = q.z * inv, nw = q.w * inv;
        const float dot = (nrm > 1e-12f) ? (nx * g.x + ny * g.y + nz * g.z + nw * g.w) : 0.f;
        float4 r = make_float4((g.x - nx * dot) * inv, (g.y - ny * dot) * inv, (g.z - nz * dot) * inv, (g.w - nw * dot) * inv);
        if (accumulate) {
            const float4 o = reinterpret_cast<const float4*>(g_rotation)[i];
            r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
        }
        reinterpret_cast<float4*>(g_rotation)[i] = r;
    }
    {
        const float s = opac[i];
        const float v = g_opac[i] * s * (1.f - s);
        g_opacity[i] = accumulate ? g_opacity[i] + v : v;
    }
}

}  // namespace lgs

using namespace lgs;

extern "C" size_t lgs_mapping_loss_scratch_bytes(int W, int H) {
    if (W <= 0 || H <= 0) return 0;
    return 256 + (size_t)9 * W * H * sizeof(float) + 256;
}

extern "C" int lgs_mapping_loss(int W, int H, int lf_w, int lf_h, const float* image, const float* lf, const float* depth,
                                const float* gt_image, const float* gt_lf, const float* gt_depth, const float* mask,
                                float lambda_dssim, int cos_sign, float* dL_dimage, float* dL_dlf, float* dL_ddepth,
                                float* loss_out, char* scratch, void* stream) {
    if (W <= 0 || H <= 0 || lf_w <= 0 || lf_h <= 0) return LGS_ERR_INVALID_ARG;
    if (!image || !lf || !depth || !gt_image || !gt_lf || !gt_depth || !dL_dimage || !dL_dlf || !dL_ddepth || !loss_out ||
        !scratch)
        return LGS_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    float* acc = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~(uintptr_t)255);
    float* maps = acc + 64;
    const size_t HW = (size_t)W * H;
    const float n_img = 3.f * (float)HW, n_pix = (float)HW;
    LGS_CUDA_TRY(cudaMemsetAsync(acc, 0, 4 * sizeof(float), s));
    const float w_cos = (cos_sign >= 0 ? 1.f : -1.f) / n_pix;
    const dim3 grid((W + SSIM_T - 1) / SSIM_T, (H + SSIM_T - 1) / SSIM_T, 3), block(SSIM_T, SSIM_T, 1);
    const unsigned n_pix_blocks = (unsigned)((HW + LP_PIX - 1) / LP_PIX), n_ssim_blocks = grid.x * grid.y * 3;
    loss_fwd_kernel<<<n_pix_blocks + n_ssim_blocks, LP_PIX * LP_GRP, 0, s>>>(n_pix_blocks, n_ssim_blocks, (int)grid.x, (int)grid.y, W, H, lf_w,
                                                                           lf_h, image, lf, depth, gt_image, gt_lf, gt_depth, mask,
                                                                           (1.f - lambda_dssim) / n_img, w_cos, 1.f / n_pix, dL_dimage,
                                                                           dL_dlf, dL_ddepth, maps, acc);
    LGS_LAUNCH_CHECK();
    ssim_bwd_kernel<<<grid, block, 0, s>>>(W, H, image, mask, gt_image, maps, -lambda_dssim / n_img, dL_dimage);
    LGS_LAUNCH_CHECK();
    loss_finalize_kernel<<<1, 1, 0, s>>>(acc, n_img, n_pix, lambda_dssim, cos_sign, loss_out);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

extern "C" int lgs_activations_fwd(int P, int n_rest, const float* scaling, const float* rotation, const float* opacity,
                                   const float* features_dc, const float* features_rest, float* scales, float* rotations,
                                   float* opacities, float* shs, void* stream) {
    if (P < 0 || n_rest < 0) return LGS_ERR_INVALID_ARG;
    if (P == 0) return LGS_OK;
    if (!scaling || !rotation || !opacity || !scales || !rotations || !opacities) return LGS_ERR_INVALID_ARG;
    if (shs && (!features_dc || (n_rest > 0 && !features_rest))) return LGS_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(rotation) | reinterpret_cast<uintptr_t>(rotations)) & 15u) return LGS_ERR_ALIGNMENT;
    activations_fwd_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, scaling, rotation, opacity, scales, rotations,
                                                                              opacities);
    LGS_LAUNCH_CHECK();
    if (!shs) return LGS_OK;
    const int row = 3 * (n_rest + 1);
    const long long n = (long long)P * row;
    const int grid = (int)((n + 255) / 256 < 148LL * 32 ? (n + 255) / 256 : 148LL * 32);
    sh_cat_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(n, row, features_dc, features_rest, shs, nullptr, nullptr,
                                                                 nullptr, 0);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

extern "C" int lgs_activations_bwd(int P, int n_rest, int accumulate, const float* rotation, const float* scales,
                                   const float* opacities, const float* dL_dscales, const float* dL_drotations,
                                   const float* dL_dopacities, const float* dL_dshs, float* dL_dscaling, float* dL_drotation,
                                   float* dL_dopacity, float* dL_dfeatures_dc, float* dL_dfeatures_rest, void* stream) {
    if (P < 0 || n_rest < 0) return LGS_ERR_INVALID_ARG;
    if (P == 0) return LGS_OK;
    if (!rotation || !scales || !opacities || !dL_dscales || !dL_drotations || !dL_dopacities || !dL_dscaling ||
        !dL_drotation || !dL_dopacity)
        return LGS_ERR_INVALID_ARG;
    if (dL_dshs && (!dL_dfeatures_dc || (n_rest > 0 && !dL_dfeatures_rest))) return LGS_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(rotation) | reinterpret_cast<uintptr_t>(dL_drotations) |
         reinterpret_cast<uintptr_t>(dL_drotation)) & 15u)
        return LGS_ERR_ALIGNMENT;
    activations_bwd_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        P, accumulate, rotation, scales, opacities, dL_dscales, dL_drotations, dL_dopacities, dL_dscaling, dL_drotation,
        dL_dopacity);
    LGS_LAUNCH_CHECK();
    if (!dL_dshs) return LGS_OK;
    const int row = 3 * (n_rest + 1);
    const long long n = (long long)P * row;
    const int grid = (int)((n + 255) / 256 < 148LL * 32 ? (n + 255) / 256 : 148LL * 32);
    sh_cat_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(n, row, nullptr, nullptr, nullptr, dL_dshs, dL_dfeatures_dc,
                                                                dL_dfeatures_rest, accumulate);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}
