// preprocess.cu -- per-Gaussian forward preprocessing for sm_100a.
//
// Replaces FORWARD::preprocess / preprocessCUDA<3> (reference forward.cu:155-256 with
// in_frustum auxiliary.h:139-164, computeCov3D forward.cu:118-152, computeCov2D
// forward.cu:74-113, getRect auxiliary.h:46-56, ndc2Pix auxiliary.h:41-44,
// computeColorFromSH forward.cu:20-71) and checkFrustum (rasterizer_impl.cu:54-66).
//
// Parity contract: depth bits, radii and tile rectangles feed the
// sort keys and must be BIT-EXACT with the reference kernel as nvcc 12.9 compiles
// it for sm_100.  Floating-point contraction is therefore pinned explicitly: every
// operation on the key-affecting chain (view/proj transform, Sigma = (SR)^T(SR),
// EWA projection, determinant, eigenvalue radius, pixel centre) is written with
// __fmaf_rn/__fmul_rn/__fadd_rn/... in exactly the association the reference
// compiles to (documented in DESIGN.md section "bit-exact preprocess"); the compiler may
// not re-fuse intrinsics.  oracle/lgs_oracle.c restates the same sequence with fmaf().
//
// Layout: one thread per Gaussian, 256 threads per CTA.  Outputs go to the packed
// 48-byte render record (common.cuh) written as three float4 stores; SH coefficients
// are read as 12 float4 (a Gaussian's 16x3 coefficients are one 192-byte row).
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include "common.cuh"

namespace lgs {

__device__ __forceinline__ float dot3p(float a0, float a1, float a2, float b0, float b1, float b2) {
    // association the reference compiles to for every 3-term product sum:
    // fma(a2,b2, fma(a0,b0, a1*b1))
    return __fmaf_rn(a2, b2, __fmaf_rn(a0, b0, __fmul_rn(a1, b1)));
}

// row r of a column-major 4x4 times (x,y,z,1):  m[12+r] + fma(z,m[8+r], fma(x,m[r], y*m[4+r]))
__device__ __forceinline__ float xform_row(const float* __restrict__ m, int r, float x, float y, float z) {
    return __fadd_rn(m[12 + r], __fmaf_rn(z, m[8 + r], __fmaf_rn(x, m[r], __fmul_rn(y, m[4 + r]))));
}

__constant__ float kSH_C1 = 0.4886025119029199f;
__constant__ float kSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                -1.0925484305920792f, 0.5462742152960396f};
__constant__ float kSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                -0.5900435899266435f};

// Real SH basis up to degree 3 evaluated at unit direction (x,y,z); b[0] is the DC term.
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float* b) {
    b[0] = 0.28209479177387814f;
    if (deg > 0) {
        b[1] = -kSH_C1 * y;
        b[2] = kSH_C1 * z;
        b[3] = -kSH_C1 * x;
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = kSH_C2[0] * xy;
            b[5] = kSH_C2[1] * yz;
            b[6] = kSH_C2[2] * (2.0f * zz - xx - yy);
            b[7] = kSH_C2[3] * xz;
            b[8] = kSH_C2[4] * (xx - yy);
            if (deg > 2) {
                b[9] = kSH_C3[0] * y * (3.0f * xx - yy);
                b[10] = kSH_C3[1] * xy * z;
                b[11] = kSH_C3[2] * y * (4.0f * zz - xx - yy);
                b[12] = kSH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
                b[13] = kSH_C3[4] * x * (4.0f * zz - xx - yy);
                b[14] = kSH_C3[5] * z * (xx - yy);
                b[15] = kSH_C3[6] * x * (xx - 3.0f * yy);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
preprocess_kernel(int P, int D, int M,
                  const float* __restrict__ means, const float* __restrict__ scales, float mod,
                  const float* __restrict__ rots, const float* __restrict__ opac,
                  const float* __restrict__ shs, const float* __restrict__ shs_rest,
                  const float* __restrict__ cov3D_pre,
                  const float* __restrict__ colors_pre, const float* __restrict__ view,
                  const float* __restrict__ proj, const float* __restrict__ campos,
                  int W, int H, float tan_fovx, float tan_fovy, float focal_x, float focal_y,
                  int tiles_x, int tiles_y,
                  int* __restrict__ radii_user, int* __restrict__ radii, GaussRec* __restrict__ rec,
                  float* __restrict__ cov3Ds,
                  uint8_t* __restrict__ clamped, uint32_t* __restrict__ tiles_touched,
                  uint32_t* __restrict__ total_touched) {
    __shared__ float sV[16], sP[16];
    __shared__ uint32_t s_touched, s_dmax;  // this block's share of R and of the largest depth bit pattern
    if (threadIdx.x < 16) sV[threadIdx.x] = view[threadIdx.x];
    else if (threadIdx.x < 32) sP[threadIdx.x - 16] = proj[threadIdx.x - 16];
    else if (threadIdx.x == 32) { s_touched = 0; s_dmax = 0; }
    __syncthreads();

    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;

    // forward.cu:188-189: default to "not rendered"
    int out_radius = 0;
    uint32_t out_tiles = 0;

    const float px = means[3 * idx + 0], py = means[3 * idx + 1], pz = means[3 * idx + 2];

    // in_frustum (auxiliary.h:147-154): only the near test survives in the reference.
    const float depth = xform_row(sV, 2, px, py, pz);
    bool live = !(depth <= 0.2f);  // same NaN behaviour as the reference's `<=` test

    float cx = 0.f, cy = 0.f, cz = 0.f, pix_x = 0.f, pix_y = 0.f;
    float my_radius = 0.f;
    if (live) {
        // forward.cu:198-201  p_hom, p_w, p_proj (z never used)
        const float hx = xform_row(sP, 0, px, py, pz);
        const float hy = xform_row(sP, 1, px, py, pz);
        const float hw = xform_row(sP, 3, px, py, pz);
        const float p_w = __frcp_rn(__fadd_rn(hw, 0.0000001f));
        const float projx = __fmul_rn(hx, p_w);
        const float projy = __fmul_rn(hy, p_w);

        // forward.cu:205-214  Sigma (6 unique entries)
        float c0, c1, c2, c3, c4, c5;
        if (cov3D_pre != nullptr) {
            const float* c = cov3D_pre + 6 * (size_t)idx;
            c0 = c[0]; c1 = c[1]; c2 = c[2]; c3 = c[3]; c4 = c[4]; c5 = c[5];
        } else {
            // computeCov3D (forward.cu:118-152): M = S*R with the quaternion used as given
            // (not normalised, :127), Sigma = M^T M.
            const float sx = __fmul_rn(mod, scales[3 * idx + 0]);
            const float sy = __fmul_rn(mod, scales[3 * idx + 1]);
            const float sz = __fmul_rn(mod, scales[3 * idx + 2]);
            const float4 q = reinterpret_cast<const float4*>(rots)[idx];
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            // association read off the reference's sm_100 SASS (DESIGN.md "bit-exact preprocess"):
            // shared products xz, rx, rz, yy, zz are rounded once, the other product of each
            // pair is fused into the add.
            const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
            const float xz = __fmul_rn(x, z), rx = __fmul_rn(r, x), rz = __fmul_rn(r, z);
            const float xx_zz = __fmaf_rn(x, x, zz);
            const float xx_yy = __fmaf_rn(x, x, yy);
            const float yy_zz = __fadd_rn(yy, zz);
            // R[c][r] (glm column-major), factor 2 applied as t+t like the compiled reference
            const float t01 = __fmaf_rn(x, y, -rz), t02 = __fmaf_rn(r, y, xz), t10 = __fmaf_rn(x, y, rz);
            const float t12 = __fmaf_rn(y, z, -rx), t20 = __fmaf_rn(-r, y, xz), t21 = __fmaf_rn(y, z, rx);
            const float R00 = __fsub_rn(1.0f, __fadd_rn(yy_zz, yy_zz));
            const float R01 = __fadd_rn(t01, t01);
            const float R02 = __fadd_rn(t02, t02);
            const float R10 = __fadd_rn(t10, t10);
            const float R11 = __fsub_rn(1.0f, __fadd_rn(xx_zz, xx_zz));
            const float R12 = __fadd_rn(t12, t12);
            const float R20 = __fadd_rn(t20, t20);
            const float R21 = __fadd_rn(t21, t21);
            const float R22 = __fsub_rn(1.0f, __fadd_rn(xx_yy, xx_yy));
            const float M00 = __fmul_rn(sx, R00), M01 = __fmul_rn(sy, R01), M02 = __fmul_rn(sz, R02);
            const float M10 = __fmul_rn(sx, R10), M11 = __fmul_rn(sy, R11), M12 = __fmul_rn(sz, R12);
            const float M20 = __fmul_rn(sx, R20), M21 = __fmul_rn(sy, R21), M22 = __fmul_rn(sz, R22);
            c0 = dot3p(M00, M01, M02, M00, M01, M02);
            c1 = dot3p(M10, M11, M12, M00, M01, M02);
            c2 = dot3p(M20, M21, M22, M00, M01, M02);
            c3 = dot3p(M10, M11, M12, M10, M11, M12);
            c4 = dot3p(M20, M21, M22, M10, M11, M12);
            c5 = dot3p(M20, M21, M22, M20, M21, M22);
            float2* cs = reinterpret_cast<float2*>(cov3Ds + 6 * (size_t)idx);
            cs[0] = make_float2(c0, c1);
            cs[1] = make_float2(c2, c3);
            cs[2] = make_float2(c4, c5);
        }

        // computeCov2D (forward.cu:74-113)
        const float tx = xform_row(sV, 0, px, py, pz);
        const float ty = xform_row(sV, 1, px, py, pz);
        const float tz = depth;
        const float limx = __fmul_rn(tan_fovx, 1.3f), limy = __fmul_rn(tan_fovy, 1.3f);
        const float txtz = __fdiv_rn(tx, tz), tytz = __fdiv_rn(ty, tz);
        const float clx = fminf(limx, fmaxf(-limx, txtz));
        const float cly = fminf(limy, fmaxf(-limy, tytz));
        const float ntz = -tz;
        const float tz2 = __fmul_rn(tz, tz);
        const float J00 = __fdiv_rn(focal_x, tz);
        const float J02 = __fdiv_rn(__fmul_rn(focal_x, __fmul_rn(clx, ntz)), tz2);
        const float J11 = __fdiv_rn(focal_y, tz);
        const float J12 = __fdiv_rn(__fmul_rn(focal_y, __fmul_rn(cly, ntz)), tz2);
        // T = W*J, rows 0/1 only (third column of J is zero)
        const float T00 = __fmaf_rn(sV[2], J02, __fmul_rn(sV[0], J00));
        const float T01 = __fmaf_rn(sV[6], J02, __fmul_rn(sV[4], J00));
        const float T02 = __fmaf_rn(J02, sV[10], __fmul_rn(sV[8], J00));
        const float T10 = __fmaf_rn(sV[2], J12, __fmul_rn(J11, sV[1]));
        const float T11 = __fmaf_rn(sV[6], J12, __fmul_rn(J11, sV[5]));
        const float T12 = __fmaf_rn(J12, sV[10], __fmul_rn(J11, sV[9]));
        // A = T^T Vrk ;  cov = A T
        const float A00 = dot3p(T00, T01, T02, c0, c1, c2);
        const float A01 = dot3p(T10, T11, T12, c0, c1, c2);
        const float A10 = dot3p(T00, T01, T02, c1, c3, c4);
        const float A11 = dot3p(T10, T11, T12, c1, c3, c4);
        const float A20 = dot3p(T00, T01, T02, c2, c4, c5);
        const float A21 = dot3p(T10, T11, T12, c2, c4, c5);
        const float cov00 = dot3p(T00, T01, T02, A00, A10, A20);
        const float cov01 = dot3p(T00, T01, T02, A01, A11, A21);
        const float cov11 = dot3p(T10, T11, T12, A01, A11, A21);
        const float a = __fadd_rn(cov00, 0.3f);
        const float c = __fadd_rn(cov11, 0.3f);
        const float b = cov01;

        // forward.cu:219-223 conic
        const float det = __fmaf_rn(a, c, -__fmul_rn(b, b));
        if (det == 0.0f) {
            live = false;
        } else {
            const float det_inv = __frcp_rn(det);
            cx = __fmul_rn(c, det_inv);
            cy = __fmul_rn(det_inv, -b);
            cz = __fmul_rn(a, det_inv);
            // forward.cu:229-232 radius from the larger eigenvalue
            const float mid = __fmul_rn(__fadd_rn(a, c), 0.5f);
            const float sq = __fsqrt_rn(fmaxf(__fmaf_rn(mid, mid, -det), 0.1f));
            const float lam = fmaxf(__fadd_rn(mid, sq), __fsub_rn(mid, sq));
            my_radius = ceilf(__fmul_rn(__fsqrt_rn(lam), 3.0f));
            // ndc2Pix in double (auxiliary.h:41-44)
            pix_x = (float)__dmul_rn(__fma_rn(__dadd_rn((double)projx, 1.0), (double)W, -1.0), 0.5);
            pix_y = (float)__dmul_rn(__fma_rn(__dadd_rn((double)projy, 1.0), (double)H, -1.0), 0.5);
            // getRect (auxiliary.h:46-56): radius truncated to int first
            const int ri = (int)my_radius;
            const float rf = (float)ri;
            const int x0 = min(tiles_x, max(0, (int)__fmul_rn(__fsub_rn(pix_x, rf), 0.125f)));
            const int y0 = min(tiles_y, max(0, (int)__fmul_rn(__fsub_rn(pix_y, rf), 0.125f)));
            const int x1 = min(tiles_x, max(0, (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(pix_x, rf), 8.0f), -1.0f), 0.125f)));
            const int y1 = min(tiles_y, max(0, (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(pix_y, rf), 8.0f), -1.0f), 0.125f)));
            const uint32_t nt = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0);
            if (nt == 0) {
                live = false;
            } else {
                out_tiles = nt;
                out_radius = (int)my_radius;
            }
        }
    }

    radii[idx] = out_radius;  // internal copy: key emission and the backward read this one
    if (radii_user != nullptr) radii_user[idx] = out_radius;
    tiles_touched[idx] = out_tiles;
    {   // R = sum of tiles_touched (and the depth bound of the sort keys) without a scan pass: warp-aggregated into shared
        // memory, one pair of global atomics per block.  Threads past P have exited; the barrier covers the rest.
        namespace cg = cooperative_groups;
        auto grp = cg::coalesced_threads();
        const uint32_t sum = cg::reduce(grp, out_tiles, cg::plus<uint32_t>());
        const uint32_t dmax = cg::reduce(grp, out_tiles ? __float_as_uint(depth) : 0u, cg::greater<uint32_t>());
        if (grp.thread_rank() == 0 && sum != 0) {
            atomicAdd(&s_touched, sum);
            atomicMax(&s_dmax, dmax);
        }
        __syncthreads();
        if (threadIdx.x == 0 && s_touched != 0) {
            atomicAdd(total_touched + HDR_R, s_touched);
            atomicMax(total_touched + HDR_MAX_DEPTH, s_dmax);
        }
    }
    if (!live) return;

    // colour: precomputed, or SH -> RGB (+0.5, clamp at 0, remember which channel clamped)
    float cr, cg, cb;
    if (colors_pre != nullptr) {
        cr = colors_pre[3 * idx + 0];
        cg = colors_pre[3 * idx + 1];
        cb = colors_pre[3 * idx + 2];
    } else {
        float dx = px - campos[0], dy = py - campos[1], dz = pz - campos[2];
        const float inv_len = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
        dx *= inv_len; dy *= inv_len; dz *= inv_len;
        float basis[16];
        sh_basis(D, dx, dy, dz, basis);
        const int ncoef = (D + 1) * (D + 1);
        float acc[3] = {0.f, 0.f, 0.f};
        if (shs_rest != nullptr) {
            // split layout: coefficient 0 in features_dc [P,1,3], coefficients 1.. in features_rest [P,M-1,3]
            // (the reference's two parameter tensors, gaussian_model.cpp:58-62, without the per-iteration cat)
            const float* d0 = shs + (size_t)idx * 3;
            const float* rs = shs_rest + (size_t)idx * (M - 1) * 3;
            acc[0] = basis[0] * __ldg(d0 + 0);
            acc[1] = basis[0] * __ldg(d0 + 1);
            acc[2] = basis[0] * __ldg(d0 + 2);
#pragma unroll 5
            for (int k = 1; k < ncoef; ++k) {
                acc[0] += basis[k] * __ldg(rs + 3 * (k - 1) + 0);
                acc[1] += basis[k] * __ldg(rs + 3 * (k - 1) + 1);
                acc[2] += basis[k] * __ldg(rs + 3 * (k - 1) + 2);
            }
        } else if (M == 16 && ncoef == 16) {
            const float* sh = shs + (size_t)idx * M * 3;
            const float4* sh4 = reinterpret_cast<const float4*>(sh);  // 192-byte row
            float v[48];
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const float4 t = __ldg(sh4 + k);
                v[4 * k + 0] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                acc[0] += basis[k] * v[3 * k + 0];
                acc[1] += basis[k] * v[3 * k + 1];
                acc[2] += basis[k] * v[3 * k + 2];
            }
        } else {
            const float* sh = shs + (size_t)idx * M * 3;
            for (int k = 0; k < ncoef; ++k) {
                acc[0] += basis[k] * sh[3 * k + 0];
                acc[1] += basis[k] * sh[3 * k + 1];
                acc[2] += basis[k] * sh[3 * k + 2];
            }
        }
        acc[0] += 0.5f; acc[1] += 0.5f; acc[2] += 0.5f;
        uint8_t cl = 0;
        if (acc[0] < 0.f) cl |= 1;
        if (acc[1] < 0.f) cl |= 2;
        if (acc[2] < 0.f) cl |= 4;
        clamped[idx] = cl;
        cr = fmaxf(acc[0], 0.f); cg = fmaxf(acc[1], 0.f); cb = fmaxf(acc[2], 0.f);
    }

    GaussRec r;
    r.q0 = make_float4(pix_x, pix_y, depth, __int_as_float(idx));
    r.q1 = make_float4(cx, cy, cz, opac[idx]);
    r.q2 = make_float4(cr, cg, cb, 0.f);
    rec[idx] = r;
}

__global__ void __launch_bounds__(256)
mark_visible_kernel(int P, const float* __restrict__ means, const float* __restrict__ view,
                    unsigned char* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float x = means[3 * idx], y = means[3 * idx + 1], z = means[3 * idx + 2];
    const float depth = xform_row(view, 2, x, y, z);
    present[idx] = !(depth <= 0.2f) ? 1 : 0;
}

int launch_preprocess(int P, int D, int M, const float* means3D, const float* shs, const float* shs_rest,
                      const float* colors_precomp, const float* opacities, const float* scales,
                      float scale_modifier, const float* rotations, const float* cov3D_precomp,
                      const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                      int W, int H, float tan_fovx, float tan_fovy, int prefiltered,
                      GeomState& g, int* radii, cudaStream_t s) {
    (void)prefiltered;
    const float focal_y = H / (2.0f * tan_fovy);  // rasterizer_impl.cu:224-225
    const float focal_x = W / (2.0f * tan_fovx);
    const int tiles_x = (W + TILE - 1) / TILE, tiles_y = (H + TILE - 1) / TILE;
    // the frame header and the sort's digit histograms start from zero (one memset: they are contiguous)
    LGS_CUDA_TRY(cudaMemsetAsync(g.hdr, 0, ((size_t)HDR_WORDS + 3 * RS_BINS) * sizeof(uint32_t), s));
    preprocess_kernel<<<(P + 255) / 256, 256, 0, s>>>(
        P, D, M, means3D, scales, scale_modifier, rotations, opacities, shs, shs_rest, cov3D_precomp,
        colors_precomp, viewmatrix, projmatrix, cam_pos, W, H, tan_fovx, tan_fovy, focal_x, focal_y,
        tiles_x, tiles_y, radii == g.internal_radii ? nullptr : radii, g.internal_radii, g.rec, g.cov3D,
        g.clamped, g.tiles_touched, g.hdr);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

int launch_mark_visible(int P, const float* means3D, const float* viewmatrix,
                        unsigned char* present, cudaStream_t s) {
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, viewmatrix, present);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

}  // namespace lgs
