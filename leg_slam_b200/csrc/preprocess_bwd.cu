// preprocess_bwd.cu -- per-Gaussian backward of the preprocessing for sm_100a, ONE kernel.
//
// Replaces BACKWARD::preprocess = computeCov2DCUDA + preprocessCUDA<3> (reference
// backward.cu:144-274 and :346-396, with the SH backward :20-139 and the covariance
// backward :278-341; launches :642,:659).  The reference runs two kernels that each re-read
// the mean / radii and pass dL_dcov3D and dL_dmean3D through global memory; here one thread
// per Gaussian keeps them in registers: dL/dconic -> dL/dcov2D -> (dL/dcov3D, dL/dmean via T),
// dL/dmean2D -> dL/dmean3D through the projection, dL/dcolour -> dL/dsh (+ view-direction
// term of dL/dmean), dL/dcov3D -> (dL/dscale, dL/drot).
//
// Gradient parity gate: 1e-3 relative, so the arithmetic follows the reference's formulas but
// not its rounding order.  Quirks kept on purpose: the quaternion is used un-normalised and no
// normalisation Jacobian is applied (backward.cu:281,340); dL_dcov3D is also written for the
// cov3D_precomp path (the caller returns it).
#include "common.cuh"

namespace lgs {

__constant__ float bSH_C1 = 0.4886025119029199f;
__constant__ float bSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                -1.0925484305920792f, 0.5462742152960396f};
__constant__ float bSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                -0.5900435899266435f};

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 ld3(const float* p) { return V3{p[0], p[1], p[2]}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// dL/dsh rows and d(colour)/d(direction), reference backward.cu:20-139
// sh1 points at coefficient 1 of the Gaussian's row (coefficient 0 has no direction derivative)
__device__ __forceinline__ V3 sh_backward(int deg, int M, const float* __restrict__ sh1, V3 dir_orig, V3 dL_dRGB,
                                          float* __restrict__ dL_dsh) {
    const float inv_len = 1.0f / sqrtf(dot(dir_orig, dir_orig));
    const float x = dir_orig.x * inv_len, y = dir_orig.y * inv_len, z = dir_orig.z * inv_len;
    V3 dx{0.f, 0.f, 0.f}, dy{0.f, 0.f, 0.f}, dz{0.f, 0.f, 0.f};
    auto st = [&](int k, float w) {
        dL_dsh[3 * k + 0] = w * dL_dRGB.x;
        dL_dsh[3 * k + 1] = w * dL_dRGB.y;
        dL_dsh[3 * k + 2] = w * dL_dRGB.z;
    };
    auto axpy = [&](V3& acc, float w, int k) {
        acc.x = fmaf(w, sh1[3 * (k - 1) + 0], acc.x);
        acc.y = fmaf(w, sh1[3 * (k - 1) + 1], acc.y);
        acc.z = fmaf(w, sh1[3 * (k - 1) + 2], acc.z);
    };
    st(0, 0.28209479177387814f);
    if (deg > 0) {
        st(1, -bSH_C1 * y);
        st(2, bSH_C1 * z);
        st(3, -bSH_C1 * x);
        axpy(dx, -bSH_C1, 3);
        axpy(dy, -bSH_C1, 1);
        axpy(dz, bSH_C1, 2);
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            st(4, bSH_C2[0] * xy);
            st(5, bSH_C2[1] * yz);
            st(6, bSH_C2[2] * (2.f * zz - xx - yy));
            st(7, bSH_C2[3] * xz);
            st(8, bSH_C2[4] * (xx - yy));
            axpy(dx, bSH_C2[0] * y, 4); axpy(dx, bSH_C2[2] * 2.f * -x, 6); axpy(dx, bSH_C2[3] * z, 7); axpy(dx, bSH_C2[4] * 2.f * x, 8);
            axpy(dy, bSH_C2[0] * x, 4); axpy(dy, bSH_C2[1] * z, 5); axpy(dy, bSH_C2[2] * 2.f * -y, 6); axpy(dy, bSH_C2[4] * 2.f * -y, 8);
            axpy(dz, bSH_C2[1] * y, 5); axpy(dz, bSH_C2[2] * 4.f * z, 6); axpy(dz, bSH_C2[3] * x, 7);
            if (deg > 2) {
                st(9, bSH_C3[0] * y * (3.f * xx - yy));
                st(10, bSH_C3[1] * xy * z);
                st(11, bSH_C3[2] * y * (4.f * zz - xx - yy));
                st(12, bSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
                st(13, bSH_C3[4] * x * (4.f * zz - xx - yy));
                st(14, bSH_C3[5] * z * (xx - yy));
                st(15, bSH_C3[6] * x * (xx - 3.f * yy));
                axpy(dx, bSH_C3[0] * 6.f * xy, 9); axpy(dx, bSH_C3[1] * yz, 10); axpy(dx, bSH_C3[2] * -2.f * xy, 11);
                axpy(dx, bSH_C3[3] * -6.f * xz, 12); axpy(dx, bSH_C3[4] * (-3.f * xx + 4.f * zz - yy), 13);
                axpy(dx, bSH_C3[5] * 2.f * xz, 14); axpy(dx, bSH_C3[6] * 3.f * (xx - yy), 15);
                axpy(dy, bSH_C3[0] * 3.f * (xx - yy), 9); axpy(dy, bSH_C3[1] * xz, 10);
                axpy(dy, bSH_C3[2] * (-3.f * yy + 4.f * zz - xx), 11); axpy(dy, bSH_C3[3] * -6.f * yz, 12);
                axpy(dy, bSH_C3[4] * -2.f * xy, 13); axpy(dy, bSH_C3[5] * -2.f * yz, 14); axpy(dy, bSH_C3[6] * -6.f * xy, 15);
                axpy(dz, bSH_C3[1] * xy, 10); axpy(dz, bSH_C3[2] * 8.f * yz, 11);
                axpy(dz, bSH_C3[3] * 3.f * (2.f * zz - xx - yy), 12); axpy(dz, bSH_C3[4] * 8.f * xz, 13);
                axpy(dz, bSH_C3[5] * (xx - yy), 14);
            }
        }
    }
    // rows above the active degree stay zero (the reference leaves its pre-zeroed buffer untouched)
    for (int k = (deg + 1) * (deg + 1); k < M; ++k) st(k, 0.f);
    // direction gradient, then through the normalisation dir = v/|v| (dnormvdv, auxiliary.h:106-116)
    const V3 dL_ddir{dot(dx, dL_dRGB), dot(dy, dL_dRGB), dot(dz, dL_dRGB)};
    const V3 v = dir_orig;
    const float sum2 = dot(v, v);
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    V3 r;
    r.x = ((sum2 - v.x * v.x) * dL_ddir.x - v.y * v.x * dL_ddir.y - v.z * v.x * dL_ddir.z) * invsum32;
    r.y = (-v.x * v.y * dL_ddir.x + (sum2 - v.y * v.y) * dL_ddir.y - v.z * v.y * dL_ddir.z) * invsum32;
    r.z = (-v.x * v.z * dL_ddir.x - v.y * v.z * dL_ddir.y + (sum2 - v.z * v.z) * dL_ddir.z) * invsum32;
    return r;
}

constexpr int PB_THREADS = 128;
constexpr int PB_ROW = 48;  // 16 SH coefficients x 3 channels

// The warp's 32 staged rows (columns COL0 .. COL0 + ROWF - 1 of s_sh; rows without data are zeros) as ONE contiguous block of
// 32 * ROWF floats in global memory, written (or accumulated into) with 16-byte accesses.  `dst` is 16-byte aligned because
// the warp's first Gaussian index is a multiple of 32 and the tensors are.
template <int ROWF, int COL0>
__device__ __forceinline__ void flush_rows(float* __restrict__ dst, const float (*s_sh)[PB_ROW + 1], int w0, uint32_t dmask, int lane,
                                           bool accumulate) {
    static_assert((32 * ROWF) % 4 == 0, "block must be a whole number of 16-byte vectors");
    constexpr int N4 = 32 * ROWF / 4;
    for (int e4 = lane; e4 < N4; e4 += 32) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = 4 * e4 + j, r = e / ROWF, c = e - r * ROWF;  // compile-time divisor
            v[j] = ((dmask >> r) & 1u) ? s_sh[w0 + r][COL0 + c] : 0.f;
        }
        float4* p = reinterpret_cast<float4*>(dst) + e4;
        float4 o = make_float4(v[0], v[1], v[2], v[3]);
        if (accumulate) {
            const float4 old = *p;
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *p = o;
    }
}

__global__ void __launch_bounds__(PB_THREADS)
preprocess_bwd_kernel(int P, int D, int M, const float* __restrict__ means, const int* __restrict__ radii,
                      const float* __restrict__ shs, const float* __restrict__ shs_rest,
                      const uint8_t* __restrict__ clamped,
                      const float* __restrict__ scales, const float* __restrict__ rots, float mod,
                      const float* __restrict__ cov3Ds, const float* __restrict__ view,
                      const float* __restrict__ proj, const float* __restrict__ campos, float h_x, float h_y,
                      float tan_fovx, float tan_fovy, const GaussRec* __restrict__ rec, float ddelx_dx, float ddely_dy,
                      float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconics,
                      const float* __restrict__ dL_dcolor,
                      float* __restrict__ dL_dmeans, float* __restrict__ dL_dcov, float* __restrict__ dL_dsh,
                      float* __restrict__ dL_dsh_rest, bool accumulate_sh,
                      float* __restrict__ dL_dscale, float* __restrict__ dL_drot, bool write_zeros) {
    __shared__ float sV[16], sP[16];
    // dL_dsh rows (up to 48 floats per Gaussian) are staged here and flushed by the whole warp, so the
    // 192-byte rows leave as full 128-byte transactions instead of 32 scattered 4-byte stores per instruction
    __shared__ float s_sh[PB_THREADS][PB_ROW + 1];
    if (threadIdx.x < 16) sV[threadIdx.x] = view[threadIdx.x];
    else if (threadIdx.x < 32) sP[threadIdx.x - 16] = proj[threadIdx.x - 16];
    __syncthreads();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = 3 * M;
    const bool stage_sh = shs != nullptr && row <= PB_ROW;
    const bool visible = idx < P && radii[idx] > 0;
    bool sh_written = false;  // does this thread's staged row hold data to flush?
    bool sh_zero = false;     // ... or is its row to be zero-filled (culled Gaussian, write_zeros)?
    if (idx < P && !visible && write_zeros) {
        dL_dmeans[3 * idx + 0] = 0.f; dL_dmeans[3 * idx + 1] = 0.f; dL_dmeans[3 * idx + 2] = 0.f;
#pragma unroll
        for (int i = 0; i < 6; ++i) dL_dcov[6 * (size_t)idx + i] = 0.f;
        if (shs != nullptr && !accumulate_sh) {
            if (stage_sh) {
                sh_zero = true;  // the flush below writes the zero row without staging it
            } else {
                for (int i = 0; i < row; ++i) dL_dsh[(size_t)idx * row + i] = 0.f;
            }
        }
        if (scales != nullptr) {
            dL_dscale[3 * idx + 0] = 0.f; dL_dscale[3 * idx + 1] = 0.f; dL_dscale[3 * idx + 2] = 0.f;
            reinterpret_cast<float4*>(dL_drot)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    if (visible) {

    const V3 mean = ld3(means + 3 * (size_t)idx);
    const float* cov3D = cov3Ds + 6 * (size_t)idx;
    const float c0 = cov3D[0], c1 = cov3D[1], c2 = cov3D[2], c3 = cov3D[3], c4 = cov3D[4], c5 = cov3D[5];

    // ------------------------------------------------------------------ cov2D backward (:144-274)
    // The render backward left raw pixel sums here (render_bwd.cu, channel kernel): sum t*dx^2,
    // sum t*dx*dy, -, sum t*dy^2 in the conic slots and sum t*dx, sum t*dy in mean2D.xy.  Apply the
    // per-Gaussian factors of backward.cu:592-606 once, and publish the finished gradients.
    float dLx, dLy, dLz, gm2x, gm2y;
    {
        const float4 q1 = rec[idx].q1;  // conic a, b, c, opacity
        const float4 sc = reinterpret_cast<const float4*>(dL_dconics)[idx];
        const float sdx = dL_dmean2D[3 * (size_t)idx + 0], sdy = dL_dmean2D[3 * (size_t)idx + 1];
        const float h = -0.5f * q1.w;
        dLx = h * sc.x; dLy = h * sc.y; dLz = h * sc.w;
        gm2x = -q1.w * ddelx_dx * (q1.x * sdx + q1.y * sdy);
        gm2y = -q1.w * ddely_dy * (q1.z * sdy + q1.y * sdx);
        reinterpret_cast<float4*>(dL_dconics)[idx] = make_float4(dLx, dLy, 0.f, dLz);
        dL_dmean2D[3 * (size_t)idx + 0] = gm2x;
        dL_dmean2D[3 * (size_t)idx + 1] = gm2y;
    }
    float tx = sV[0] * mean.x + sV[4] * mean.y + sV[8] * mean.z + sV[12];
    float ty = sV[1] * mean.x + sV[5] * mean.y + sV[9] * mean.z + sV[13];
    const float tz = sV[2] * mean.x + sV[6] * mean.y + sV[10] * mean.z + sV[14];
    const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
    const float txtz = tx / tz, tytz = ty / tz;
    tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
    ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
    const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;

    const float J00 = h_x / tz, J02 = -(h_x * tx) / (tz * tz), J11 = h_y / tz, J12 = -(h_y * ty) / (tz * tz);
    // T[c][r] = W*J (glm column-major), third column zero
    const float T00 = sV[0] * J00 + sV[2] * J02, T01 = sV[4] * J00 + sV[6] * J02, T02 = sV[8] * J00 + sV[10] * J02;
    const float T10 = sV[1] * J11 + sV[2] * J12, T11 = sV[5] * J11 + sV[6] * J12, T12 = sV[9] * J11 + sV[10] * J12;
    // VT_r = Vrk * T[r][:]
    const float v00 = T00 * c0 + T01 * c1 + T02 * c2, v01 = T00 * c1 + T01 * c3 + T02 * c4, v02 = T00 * c2 + T01 * c4 + T02 * c5;
    const float v10 = T10 * c0 + T11 * c1 + T12 * c2, v11 = T10 * c1 + T11 * c3 + T12 * c4, v12 = T10 * c2 + T11 * c4 + T12 * c5;
    const float a = T00 * v00 + T01 * v01 + T02 * v02 + 0.3f;
    const float b = T00 * v10 + T01 * v11 + T02 * v12;
    const float c = T10 * v10 + T11 * v11 + T12 * v12 + 0.3f;

    const float denom = a * c - b * b;
    float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    float dcv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (denom2inv != 0.f) {
        dL_da = denom2inv * (-c * c * dLx + 2.f * b * c * dLy + (denom - a * c) * dLz);
        dL_dc = denom2inv * (-a * a * dLz + 2.f * a * b * dLy + (denom - a * c) * dLx);
        dL_db = denom2inv * 2.f * (b * c * dLx - (denom + 2.f * b * b) * dLy + a * b * dLz);
        dcv[0] = T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc;
        dcv[3] = T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc;
        dcv[5] = T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc;
        dcv[1] = 2.f * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2.f * T10 * T11 * dL_dc;
        dcv[2] = 2.f * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2.f * T10 * T12 * dL_dc;
        dcv[4] = 2.f * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2.f * T11 * T12 * dL_dc;
    }
    {
        float2* o = reinterpret_cast<float2*>(dL_dcov + 6 * (size_t)idx);
        o[0] = make_float2(dcv[0], dcv[1]);
        o[1] = make_float2(dcv[2], dcv[3]);
        o[2] = make_float2(dcv[4], dcv[5]);
    }
    const float dL_dT00 = 2.f * v00 * dL_da + v10 * dL_db, dL_dT01 = 2.f * v01 * dL_da + v11 * dL_db;
    const float dL_dT02 = 2.f * v02 * dL_da + v12 * dL_db;
    const float dL_dT10 = 2.f * v10 * dL_dc + v00 * dL_db, dL_dT11 = 2.f * v11 * dL_dc + v01 * dL_db;
    const float dL_dT12 = 2.f * v12 * dL_dc + v02 * dL_db;
    const float dL_dJ00 = sV[0] * dL_dT00 + sV[4] * dL_dT01 + sV[8] * dL_dT02;
    const float dL_dJ02 = sV[2] * dL_dT00 + sV[6] * dL_dT01 + sV[10] * dL_dT02;
    const float dL_dJ11 = sV[1] * dL_dT10 + sV[5] * dL_dT11 + sV[9] * dL_dT12;
    const float dL_dJ12 = sV[2] * dL_dT10 + sV[6] * dL_dT11 + sV[10] * dL_dT12;
    const float itz = 1.f / tz, itz2 = itz * itz, itz3 = itz2 * itz;
    const float dL_dtx = x_grad_mul * -h_x * itz2 * dL_dJ02;
    const float dL_dty = y_grad_mul * -h_y * itz2 * dL_dJ12;
    const float dL_dtz = -h_x * itz2 * dL_dJ00 - h_y * itz2 * dL_dJ11 + (2.f * h_x * tx) * itz3 * dL_dJ02 +
                         (2.f * h_y * ty) * itz3 * dL_dJ12;
    V3 dmean;
    dmean.x = sV[0] * dL_dtx + sV[1] * dL_dty + sV[2] * dL_dtz;
    dmean.y = sV[4] * dL_dtx + sV[5] * dL_dty + sV[6] * dL_dtz;
    dmean.z = sV[8] * dL_dtx + sV[9] * dL_dty + sV[10] * dL_dtz;

    // ------------------------------------------------------- mean2D -> mean3D through proj (:366-385)
    {
        const float hx = sP[0] * mean.x + sP[4] * mean.y + sP[8] * mean.z + sP[12];
        const float hy = sP[1] * mean.x + sP[5] * mean.y + sP[9] * mean.z + sP[13];
        const float hw = sP[3] * mean.x + sP[7] * mean.y + sP[11] * mean.z + sP[15];
        const float m_w = 1.0f / (hw + 0.0000001f);
        const float mul1 = hx * m_w * m_w, mul2 = hy * m_w * m_w;
        const float gx = gm2x, gy = gm2y;
        dmean.x += (sP[0] * m_w - sP[3] * mul1) * gx + (sP[1] * m_w - sP[3] * mul2) * gy;
        dmean.y += (sP[4] * m_w - sP[7] * mul1) * gx + (sP[5] * m_w - sP[7] * mul2) * gy;
        dmean.z += (sP[8] * m_w - sP[11] * mul1) * gx + (sP[9] * m_w - sP[11] * mul2) * gy;
    }

    // ------------------------------------------------------------------ SH backward (:20-139)
    if (shs != nullptr) {
        const uint8_t cl = clamped[idx];
        V3 dRGB = ld3(dL_dcolor + 3 * (size_t)idx);
        if (cl & 1) dRGB.x = 0.f;
        if (cl & 2) dRGB.y = 0.f;
        if (cl & 4) dRGB.z = 0.f;
        const V3 dir{mean.x - campos[0], mean.y - campos[1], mean.z - campos[2]};
        const float* sh1 = shs_rest != nullptr ? shs_rest + (size_t)idx * 3 * (M - 1) : shs + (size_t)idx * 3 * M + 3;
        const V3 dm = sh_backward(D, M, sh1, dir, dRGB,
                                  stage_sh ? &s_sh[threadIdx.x][0] : dL_dsh + (size_t)idx * 3 * M);
        sh_written = stage_sh;
        dmean.x += dm.x; dmean.y += dm.y; dmean.z += dm.z;
    }
    dL_dmeans[3 * (size_t)idx + 0] = dmean.x;
    dL_dmeans[3 * (size_t)idx + 1] = dmean.y;
    dL_dmeans[3 * (size_t)idx + 2] = dmean.z;

    // --------------------------------------------------- cov3D -> scale, rotation (:278-341)
    if (scales != nullptr) {
        const float4 q = reinterpret_cast<const float4*>(rots)[idx];
        const float r = q.x, x = q.y, y = q.z, z = q.w;
        // R[c][r] column-major as in forward
        const float R00 = 1.f - 2.f * (y * y + z * z), R01 = 2.f * (x * y - r * z), R02 = 2.f * (x * z + r * y);
        const float R10 = 2.f * (x * y + r * z), R11 = 1.f - 2.f * (x * x + z * z), R12 = 2.f * (y * z - r * x);
        const float R20 = 2.f * (x * z - r * y), R21 = 2.f * (y * z + r * x), R22 = 1.f - 2.f * (x * x + y * y);
        const float sx = mod * scales[3 * (size_t)idx + 0], sy = mod * scales[3 * (size_t)idx + 1];
        const float sz = mod * scales[3 * (size_t)idx + 2];
        // M[c][r] = s_r * R[c][r]
        const float M00 = sx * R00, M01 = sy * R01, M02 = sz * R02;
        const float M10 = sx * R10, M11 = sy * R11, M12 = sz * R12;
        const float M20 = sx * R20, M21 = sy * R21, M22 = sz * R22;
        // dL_dSigma (symmetric; off-diagonals halved), indexed [c][r]
        const float S00 = dcv[0], S01 = 0.5f * dcv[1], S02 = 0.5f * dcv[2];
        const float S11 = dcv[3], S12 = 0.5f * dcv[4], S22 = dcv[5];
        // dL_dM = 2 * M * dL_dSigma  (glm: (A*B)[c][r] = sum_k A[k][r] * B[c][k])
        const float G00 = 2.f * (M00 * S00 + M10 * S01 + M20 * S02), G01 = 2.f * (M01 * S00 + M11 * S01 + M21 * S02);
        const float G02 = 2.f * (M02 * S00 + M12 * S01 + M22 * S02);
        const float G10 = 2.f * (M00 * S01 + M10 * S11 + M20 * S12), G11 = 2.f * (M01 * S01 + M11 * S11 + M21 * S12);
        const float G12 = 2.f * (M02 * S01 + M12 * S11 + M22 * S12);
        const float G20 = 2.f * (M00 * S02 + M10 * S12 + M20 * S22), G21 = 2.f * (M01 * S02 + M11 * S12 + M21 * S22);
        const float G22 = 2.f * (M02 * S02 + M12 * S12 + M22 * S22);
        // Rt[k] = row k of R as a vector = (R[0][k], R[1][k], R[2][k]); dL_dMt[k] = (G[0][k], G[1][k], G[2][k])
        dL_dscale[3 * (size_t)idx + 0] = R00 * G00 + R10 * G10 + R20 * G20;
        dL_dscale[3 * (size_t)idx + 1] = R01 * G01 + R11 * G11 + R21 * G21;
        dL_dscale[3 * (size_t)idx + 2] = R02 * G02 + R12 * G12 + R22 * G22;
        // dL_dMt[k] *= s_k ; element dL_dMt[k][c] = G[c][k] * s_k
        const float t00 = G00 * sx, t01 = G10 * sx, t02 = G20 * sx;  // dL_dMt[0][0..2]
        const float t10 = G01 * sy, t11 = G11 * sy, t12 = G21 * sy;  // dL_dMt[1][0..2]
        const float t20 = G02 * sz, t21 = G12 * sz, t22 = G22 * sz;  // dL_dMt[2][0..2]
        float4 dq;
        dq.x = 2.f * z * (t01 - t10) + 2.f * y * (t20 - t02) + 2.f * x * (t12 - t21);
        dq.y = 2.f * y * (t10 + t01) + 2.f * z * (t20 + t02) + 2.f * r * (t12 - t21) - 4.f * x * (t22 + t11);
        dq.z = 2.f * x * (t10 + t01) + 2.f * r * (t20 - t02) + 2.f * z * (t12 + t21) - 4.f * y * (t22 + t00);
        dq.w = 2.f * r * (t01 - t10) + 2.f * x * (t20 + t02) + 2.f * y * (t12 + t21) - 4.f * z * (t11 + t00);
        reinterpret_cast<float4*>(dL_drot)[idx] = dq;
    }
    }  // visible

    if (stage_sh) {
        // warp-cooperative flush of the staged rows: lane-consecutive addresses in global memory
        __syncwarp();
        const int lane = threadIdx.x & 31, w0 = threadIdx.x & ~31;
        const uint32_t dmask = __ballot_sync(0xffffffffu, sh_written), zmask = __ballot_sync(0xffffffffu, sh_zero);
        const uint32_t wmask = dmask | zmask;
        const long long g0 = (long long)blockIdx.x * blockDim.x + w0;  // first Gaussian of this warp
        const bool al16 = ((reinterpret_cast<uintptr_t>(dL_dsh) | reinterpret_cast<uintptr_t>(dL_dsh_rest)) & 15u) == 0;
        if (wmask == 0xffffffffu && row == PB_ROW && al16) {
            // every row of the warp's 32 x 48 block is written (the mapping configuration: SH degree 3, culled rows zero-filled):
            // 16-byte stores over the contiguous block(s) -- 12 (concatenated) or 1 + 12 (split layout) iterations instead of 48
            if (dL_dsh_rest == nullptr) flush_rows<PB_ROW, 0>(dL_dsh + g0 * PB_ROW, s_sh, w0, dmask, lane, accumulate_sh);
            else {
                flush_rows<3, 0>(dL_dsh + g0 * 3, s_sh, w0, dmask, lane, accumulate_sh);
                flush_rows<PB_ROW - 3, 3>(dL_dsh_rest + g0 * (PB_ROW - 3), s_sh, w0, dmask, lane, accumulate_sh);
            }
            return;
        }
        // element i = r * row + c of the warp's 32 x row block; (r, c) advance with i (no division in the loop)
        int r = lane / row, c = lane - r * row;
        for (int i = lane; i < 32 * row; i += 32) {
            if ((wmask >> r) & 1u) {
                // split layout (dL_dsh_rest != NULL): coefficient 0 -> dL_dsh [P,1,3], the rest -> dL_dsh_rest [P,M-1,3]
                float* dst = dL_dsh + g0 * row + i;
                if (dL_dsh_rest != nullptr) dst = c < 3 ? dL_dsh + (g0 + r) * 3 + c : dL_dsh_rest + (g0 + r) * (row - 3) + (c - 3);
                const float v = ((dmask >> r) & 1u) ? s_sh[w0 + r][c] : 0.f;
                *dst = accumulate_sh ? *dst + v : v;
            }
            c += 32;
            while (c >= row) { c -= row; ++r; }
        }
    }
}

int launch_preprocess_bwd(int P, int D, int M, const float* means3D, const int* radii, const float* shs,
                          const float* shs_rest, const float* scales, const float* rotations, float scale_modifier, const float* cov3D,
                          const float* viewmatrix, const float* projmatrix, const float* cam_pos, int W, int H,
                          float tan_fovx, float tan_fovy, const GeomState& g, float* dL_dmean2D,
                          float* dL_dconic, float* dL_dmean3D, const float* dL_dcolor, float* dL_dcov3D,
                          float* dL_dsh, float* dL_dsh_rest, bool accumulate_sh, float* dL_dscale, float* dL_drot,
                          bool write_zeros, cudaStream_t s) {
    const float focal_y = H / (2.0f * tan_fovy);  // rasterizer_impl.cu:392-393
    const float focal_x = W / (2.0f * tan_fovx);
    preprocess_bwd_kernel<<<(P + PB_THREADS - 1) / PB_THREADS, PB_THREADS, 0, s>>>(
        P, D, M, means3D, radii, shs, shs_rest, g.clamped, scales, rotations, scale_modifier, cov3D, viewmatrix, projmatrix,
        cam_pos, focal_x, focal_y, tan_fovx, tan_fovy, g.rec, 0.5f * (float)W, 0.5f * (float)H, dL_dmean2D, dL_dconic,
        dL_dcolor, dL_dmean3D, dL_dcov3D,
        dL_dsh, dL_dsh_rest, accumulate_sh, dL_dscale, dL_drot, write_zeros);
    LGS_LAUNCH_CHECK();
    return LGS_OK;
}

}  // namespace lgs
