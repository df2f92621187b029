/*
 * lgs_oracle.c -- CPU restatement of LEG-SLAM's mapping hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may build, load or call
 * this file.  Nothing under leg_slam_b200/ links or imports it; the product path is CUDA only.
 *
 * Each function restates, in plain scalar C, what the reference computes (file:line cited per
 * function, paths relative to /root/reference; the Python-extension copy under
 * eval/submodules/diff-gaussian-rasterization-legs-slam/ is byte-identical).  It is a
 * restatement, not a copy: per-pixel loops instead of CUDA blocks, explicit fmaf() where the
 * reference's sm_100 build contracts, so that everything feeding the sort keys (depth bits,
 * radii, tile rectangles) is BIT-EXACT with the compiled reference.
 *
 * Pinning: the reference ships no golden vectors.  This oracle is pinned
 * against outputs of the UNMODIFIED reference kernels compiled for sm_100 (oracle/build_ref.py)
 * and run on a B200: tests/golden/ holds those outputs with the script that produced them
 * (tests/golden/make_golden.py); tests/test_oracle_golden.py checks this file against them.
 *
 * Threading: OpenMP over Gaussians / tiles (the cpu_baseline leg reports the thread count).
 * Gradient accumulation across tiles uses `omp atomic`, so its summation order varies run to
 * run exactly like the reference's atomicAdd.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC lgs_oracle.c -o _build/liblgs_oracle.so -lm
 *        (-ffp-contract=off: only the fmaf() written below may fuse)
 */
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TILE 8
#define LF 64
#define NCH 3

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

/* a0*b0 + a1*b1 + a2*b2 as the reference's sm_100 build evaluates every glm 3-term product:
 * fma(a2,b2, fma(a0,b0, a1*b1)) */
static inline float dot3p(float a0, float a1, float a2, float b0, float b1, float b2) {
    return fmaf(a2, b2, fmaf(a0, b0, a1 * b1));
}
/* transformPoint4x3/4x4 row r (auxiliary.h:58-77): m[12+r] + fma(z,m[8+r], fma(x,m[r], y*m[4+r])) */
static inline float xform_row(const float* m, int r, float x, float y, float z) {
    return m[12 + r] + fmaf(z, m[8 + r], fmaf(x, m[r], y * m[4 + r]));
}
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* getRect (auxiliary.h:46-56) with BLOCK_X = BLOCK_Y = 8 */
static inline void get_rect(float px, float py, int radius, int tiles_x, int tiles_y, int* x0, int* y0, int* x1,
                            int* y1) {
    const float rf = (float)radius;
    *x0 = imin(tiles_x, imax(0, (int)((px - rf) / 8.0f)));
    *y0 = imin(tiles_y, imax(0, (int)((py - rf) / 8.0f)));
    *x1 = imin(tiles_x, imax(0, (int)((((px + rf) + 8.0f) - 1.0f) / 8.0f)));
    *y1 = imin(tiles_y, imax(0, (int)((((py + rf) + 8.0f) - 1.0f) / 8.0f)));
}

/* ------------------------------------------------------------------------------------------
 * preprocessCUDA<3> forward (cuda_rasterizer/forward.cu:155-256) with in_frustum
 * (auxiliary.h:139-164), computeCov3D (forward.cu:118-152), computeCov2D (:74-113),
 * computeColorFromSH (:20-71), ndc2Pix (auxiliary.h:41-44).
 * Outputs use the reference's GeometryState arrays (rasterizer_impl.h:33-48).
 * Arrays of culled Gaussians are left untouched except radii/tiles_touched = 0.
 * ---------------------------------------------------------------------------------------- */
void oracle_preprocess(int P, int D, int M, const float* means, const float* scales, float mod, const float* rots,
                       const float* opac, const float* shs, const float* cov3D_pre, const float* colors_pre,
                       const float* view, const float* proj, const float* campos, int W, int H, float tan_fovx,
                       float tan_fovy, int* radii, float* means2D, float* depths, float* cov3Ds, float* conic_opacity,
                       float* rgb, uint8_t* clamped, uint32_t* tiles_touched) {
    const float focal_y = H / (2.0f * tan_fovy); /* rasterizer_impl.cu:224-225 */
    const float focal_x = W / (2.0f * tan_fovx);
    const int tiles_x = (W + TILE - 1) / TILE, tiles_y = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; ++idx) {
        radii[idx] = 0;
        tiles_touched[idx] = 0;
        const float px = means[3 * idx], py = means[3 * idx + 1], pz = means[3 * idx + 2];
        const float depth = xform_row(view, 2, px, py, pz);
        if (depth <= 0.2f) continue; /* auxiliary.h:154 */
        const float hx = xform_row(proj, 0, px, py, pz), hy = xform_row(proj, 1, px, py, pz);
        const float hw = xform_row(proj, 3, px, py, pz);
        const float p_w = 1.0f / (hw + 0.0000001f);
        const float projx = hx * p_w, projy = hy * p_w;

        float c0, c1, c2, c3, c4, c5;
        if (cov3D_pre) {
            const float* c = cov3D_pre + 6 * (size_t)idx;
            c0 = c[0]; c1 = c[1]; c2 = c[2]; c3 = c[3]; c4 = c[4]; c5 = c[5];
        } else {
            /* M = S*R (quaternion NOT normalised, forward.cu:127), Sigma = M^T M */
            const float sx = mod * scales[3 * idx], sy = mod * scales[3 * idx + 1], sz = mod * scales[3 * idx + 2];
            const float r = rots[4 * idx], x = rots[4 * idx + 1], y = rots[4 * idx + 2], z = rots[4 * idx + 3];
            const float yy = y * y, zz = z * z, xz = x * z, rx = r * x, rz = r * z;
            const float xx_zz = fmaf(x, x, zz), xx_yy = fmaf(x, x, yy), yy_zz = yy + zz;
            const float t01 = fmaf(x, y, -rz), t02 = fmaf(r, y, xz), t10 = fmaf(x, y, rz);
            const float t12 = fmaf(y, z, -rx), t20 = fmaf(-r, y, xz), t21 = fmaf(y, z, rx);
            const float R00 = 1.0f - (yy_zz + yy_zz), R01 = t01 + t01, R02 = t02 + t02;
            const float R10 = t10 + t10, R11 = 1.0f - (xx_zz + xx_zz), R12 = t12 + t12;
            const float R20 = t20 + t20, R21 = t21 + t21, R22 = 1.0f - (xx_yy + xx_yy);
            const float M00 = sx * R00, M01 = sy * R01, M02 = sz * R02;
            const float M10 = sx * R10, M11 = sy * R11, M12 = sz * R12;
            const float M20 = sx * R20, M21 = sy * R21, M22 = sz * R22;
            c0 = dot3p(M00, M01, M02, M00, M01, M02);
            c1 = dot3p(M10, M11, M12, M00, M01, M02);
            c2 = dot3p(M20, M21, M22, M00, M01, M02);
            c3 = dot3p(M10, M11, M12, M10, M11, M12);
            c4 = dot3p(M20, M21, M22, M10, M11, M12);
            c5 = dot3p(M20, M21, M22, M20, M21, M22);
            float* cs = cov3Ds + 6 * (size_t)idx;
            cs[0] = c0; cs[1] = c1; cs[2] = c2; cs[3] = c3; cs[4] = c4; cs[5] = c5;
        }
        /* EWA projection (forward.cu:74-113) */
        const float tx = xform_row(view, 0, px, py, pz), ty = xform_row(view, 1, px, py, pz), tz = depth;
        const float limx = tan_fovx * 1.3f, limy = tan_fovy * 1.3f;
        const float txtz = tx / tz, tytz = ty / tz;
        const float clx = fminf(limx, fmaxf(-limx, txtz)), cly = fminf(limy, fmaxf(-limy, tytz));
        const float tz2 = tz * tz;
        const float J00 = focal_x / tz, J02 = (focal_x * (clx * -tz)) / tz2;
        const float J11 = focal_y / tz, J12 = (focal_y * (cly * -tz)) / tz2;
        const float T00 = fmaf(view[2], J02, view[0] * J00), T01 = fmaf(view[6], J02, view[4] * J00);
        const float T02 = fmaf(J02, view[10], view[8] * J00);
        const float T10 = fmaf(view[2], J12, J11 * view[1]), T11 = fmaf(view[6], J12, J11 * view[5]);
        const float T12 = fmaf(J12, view[10], J11 * view[9]);
        const float A00 = dot3p(T00, T01, T02, c0, c1, c2), A01 = dot3p(T10, T11, T12, c0, c1, c2);
        const float A10 = dot3p(T00, T01, T02, c1, c3, c4), A11 = dot3p(T10, T11, T12, c1, c3, c4);
        const float A20 = dot3p(T00, T01, T02, c2, c4, c5), A21 = dot3p(T10, T11, T12, c2, c4, c5);
        const float a = dot3p(T00, T01, T02, A00, A10, A20) + 0.3f;
        const float b = dot3p(T00, T01, T02, A01, A11, A21);
        const float c = dot3p(T10, T11, T12, A01, A11, A21) + 0.3f;
        const float det = fmaf(a, c, -(b * b)); /* forward.cu:219 */
        if (det == 0.0f) continue;
        const float det_inv = 1.0f / det;
        const float mid = (a + c) * 0.5f;
        const float sq = sqrtf(fmaxf(fmaf(mid, mid, -det), 0.1f));
        const float lam = fmaxf(mid + sq, mid - sq);
        const float my_radius = ceilf(sqrtf(lam) * 3.0f); /* :229-232 */
        const float pix_x = (float)(fma((double)projx + 1.0, (double)W, -1.0) * 0.5);
        const float pix_y = (float)(fma((double)projy + 1.0, (double)H, -1.0) * 0.5);
        int x0, y0, x1, y1;
        get_rect(pix_x, pix_y, (int)my_radius, tiles_x, tiles_y, &x0, &y0, &x1, &y1);
        if ((x1 - x0) * (y1 - y0) == 0) continue;

        if (!colors_pre) { /* computeColorFromSH, forward.cu:20-71 */
            float dx = px - campos[0], dy = py - campos[1], dz = pz - campos[2];
            const float len = sqrtf(dx * dx + dy * dy + dz * dz);
            dx /= len; dy /= len; dz /= len;
            float basis[16];
            basis[0] = SH_C0;
            if (D > 0) {
                basis[1] = -SH_C1 * dy; basis[2] = SH_C1 * dz; basis[3] = -SH_C1 * dx;
                if (D > 1) {
                    const float xx = dx * dx, yy = dy * dy, zz = dz * dz, xy = dx * dy, yz = dy * dz, xz = dx * dz;
                    basis[4] = SH_C2[0] * xy; basis[5] = SH_C2[1] * yz; basis[6] = SH_C2[2] * (2.0f * zz - xx - yy);
                    basis[7] = SH_C2[3] * xz; basis[8] = SH_C2[4] * (xx - yy);
                    if (D > 2) {
                        basis[9] = SH_C3[0] * dy * (3.0f * xx - yy);
                        basis[10] = SH_C3[1] * xy * dz;
                        basis[11] = SH_C3[2] * dy * (4.0f * zz - xx - yy);
                        basis[12] = SH_C3[3] * dz * (2.0f * zz - 3.0f * xx - 3.0f * yy);
                        basis[13] = SH_C3[4] * dx * (4.0f * zz - xx - yy);
                        basis[14] = SH_C3[5] * dz * (xx - yy);
                        basis[15] = SH_C3[6] * dx * (xx - 3.0f * yy);
                    }
                }
            }
            const float* sh = shs + (size_t)idx * M * 3;
            for (int ch = 0; ch < 3; ++ch) {
                float acc = 0.f;
                for (int k = 0; k < (D + 1) * (D + 1); ++k) acc += basis[k] * sh[3 * k + ch];
                acc += 0.5f;
                clamped[3 * idx + ch] = acc < 0.f;
                rgb[3 * idx + ch] = fmaxf(acc, 0.f);
            }
        }
        depths[idx] = depth;
        radii[idx] = (int)my_radius;
        means2D[2 * idx] = pix_x;
        means2D[2 * idx + 1] = pix_y;
        conic_opacity[4 * idx + 0] = c * det_inv;
        conic_opacity[4 * idx + 1] = -b * det_inv;
        conic_opacity[4 * idx + 2] = a * det_inv;
        conic_opacity[4 * idx + 3] = opac[idx];
        tiles_touched[idx] = (uint32_t)((y1 - y0) * (x1 - x0));
    }
}

/* checkFrustum / markVisible (rasterizer_impl.cu:54-66,141-153) */
void oracle_mark_visible(int P, const float* means, const float* view, uint8_t* present) {
    for (int i = 0; i < P; ++i)
        present[i] = !(xform_row(view, 2, means[3 * i], means[3 * i + 1], means[3 * i + 2]) <= 0.2f);
}

/* InclusiveSum total (rasterizer_impl.cu:277-282) */
int64_t oracle_num_rendered(int P, const uint32_t* tiles_touched) {
    int64_t r = 0;
    for (int i = 0; i < P; ++i) r += tiles_touched[i];
    return r;
}

/* getHigherMsb (rasterizer_impl.cu:35-50): bits needed to write n */
static int key_bits(uint32_t n) {
    int bits = 1;
    while ((n >> bits) != 0 && bits < 32) ++bits;
    return bits;
}

/* duplicateWithKeys + SortPairs + identifyTileRanges (rasterizer_impl.cu:70-138,290-320).
 * Emission order: Gaussian-major, then tile y, then x.  Sort: stable LSD radix sort, 8 bits per
 * pass over key bits [0, 32+bit) -- a stable sort has one answer, so the digit size is free.
 * ranges is [tiles][2], zeroed first (cudaMemset :311). */
void oracle_binning(int P, int W, int H, const float* means2D, const float* depths, const int* radii,
                    const uint32_t* tiles_touched, int64_t R, uint64_t* keys_unsorted, uint32_t* vals_unsorted,
                    uint64_t* keys_sorted, uint32_t* point_list, uint32_t* ranges) {
    const int tiles_x = (W + TILE - 1) / TILE, tiles_y = (H + TILE - 1) / TILE;
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)tiles_x * tiles_y);
    int64_t off = 0;
    for (int idx = 0; idx < P; ++idx) {
        if (!(radii[idx] > 0)) continue;
        int x0, y0, x1, y1;
        get_rect(means2D[2 * idx], means2D[2 * idx + 1], radii[idx], tiles_x, tiles_y, &x0, &y0, &x1, &y1);
        uint32_t dbits;
        memcpy(&dbits, &depths[idx], 4);
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) {
                keys_unsorted[off] = ((uint64_t)(uint32_t)(y * tiles_x + x) << 32) | dbits;
                vals_unsorted[off] = (uint32_t)idx;
                ++off;
            }
        (void)tiles_touched;
    }
    if (R <= 0) return;
    const int end_bit = 32 + key_bits((uint32_t)(tiles_x * tiles_y));
    uint64_t* ka = (uint64_t*)malloc(sizeof(uint64_t) * R);
    uint64_t* kb = (uint64_t*)malloc(sizeof(uint64_t) * R);
    uint32_t* va = (uint32_t*)malloc(sizeof(uint32_t) * R);
    uint32_t* vb = (uint32_t*)malloc(sizeof(uint32_t) * R);
    memcpy(ka, keys_unsorted, sizeof(uint64_t) * R);
    memcpy(va, vals_unsorted, sizeof(uint32_t) * R);
    for (int shift = 0; shift < end_bit; shift += 8) {
        const int nb = end_bit - shift < 8 ? end_bit - shift : 8;
        const uint64_t mask = (1u << nb) - 1;
        int64_t count[257];
        memset(count, 0, sizeof(count));
        for (int64_t i = 0; i < R; ++i) count[((ka[i] >> shift) & mask) + 1]++;
        for (int d = 0; d < 256; ++d) count[d + 1] += count[d];
        for (int64_t i = 0; i < R; ++i) {
            const int64_t dst = count[(ka[i] >> shift) & mask]++;
            kb[dst] = ka[i];
            vb[dst] = va[i];
        }
        uint64_t* tk = ka; ka = kb; kb = tk;
        uint32_t* tv = va; va = vb; vb = tv;
    }
    memcpy(keys_sorted, ka, sizeof(uint64_t) * R);
    memcpy(point_list, va, sizeof(uint32_t) * R);
    free(ka); free(kb); free(va); free(vb);
    for (int64_t i = 0; i < R; ++i) { /* identifyTileRanges :116-138 */
        const uint32_t cur = (uint32_t)(keys_sorted[i] >> 32);
        if (i == 0) ranges[2 * cur] = 0;
        else {
            const uint32_t prev = (uint32_t)(keys_sorted[i - 1] >> 32);
            if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)i; ranges[2 * cur] = (uint32_t)i; }
        }
        if (i == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
    }
}

/* power = -0.5*(a dx^2 + c dy^2) - b dx dy as the reference's renderCUDA evaluates it on sm_100
 * (forward.cu:337-341): s = fma(dx, dx*a, dy*(dy*c)); power = fma(s, -0.5, -(dy*(dx*b))) */
static inline float eval_power(float dx, float dy, float a, float b, float c) {
    const float s = fmaf(dx, dx * a, dy * (dy * c));
    return fmaf(s, -0.5f, -(dy * (dx * b)));
}

/* renderCUDA<3,64> forward (forward.cu:261-392), one pixel at a time.  colors is [P,3]
 * (geomState.rgb or colors_precomp, rasterizer_impl.cu:323).  Blending follows the compiled
 * reference: C = fma(T, alpha*v, C).  Returns the number of blended fragments. */
int64_t oracle_render_fwd(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* means2D,
                          const float* colors, const float* lang_feat, const float* depths,
                          const float* conic_opacity, const float* bg, int include_lf, float* final_T,
                          uint32_t* n_contrib, float* out_color, float* out_lf, float* out_depth) {
    const int tiles_x = (W + TILE - 1) / TILE;
    const size_t HW = (size_t)H * W;
    int64_t blended = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : blended)
    for (int py = 0; py < H; ++py) {
        for (int px = 0; px < W; ++px) {
            const uint32_t tile = (uint32_t)((py / TILE) * tiles_x + px / TILE);
            const uint32_t beg = ranges[2 * tile], end = ranges[2 * tile + 1];
            float T = 1.0f, C[NCH] = {0.f, 0.f, 0.f}, L[LF], Dp = 0.f;
            for (int k = 0; k < LF; ++k) L[k] = 0.f;
            uint32_t contributor = 0, last = 0;
            for (uint32_t i = beg; i < end; ++i) {
                ++contributor;
                const uint32_t id = point_list[i];
                const float dx = means2D[2 * id] - (float)px, dy = means2D[2 * id + 1] - (float)py;
                const float* co = conic_opacity + 4 * (size_t)id;
                const float power = eval_power(dx, dy, co[0], co[1], co[2]);
                if (power > 0.0f) continue;
                const float alpha = fminf(0.99f, co[3] * expf(power));
                if (alpha < 1.0f / 255.0f) continue;
                const float test_T = T * (1.0f - alpha);
                if (test_T < 0.0001f) break; /* done = true */
                for (int ch = 0; ch < NCH; ++ch) C[ch] = fmaf(T, alpha * colors[3 * (size_t)id + ch], C[ch]);
                if (include_lf)
                    for (int ch = 0; ch < LF; ++ch) L[ch] = fmaf(T, alpha * lang_feat[LF * (size_t)id + ch], L[ch]);
                Dp = fmaf(T, alpha * depths[id], Dp);
                T = test_T;
                last = contributor;
                ++blended;
            }
            const size_t pix = (size_t)py * W + px;
            final_T[pix] = T;
            n_contrib[pix] = last;
            for (int ch = 0; ch < NCH; ++ch) out_color[ch * HW + pix] = fmaf(T, bg[ch], C[ch]);
            if (include_lf)
                for (int ch = 0; ch < LF; ++ch) out_lf[ch * HW + pix] = L[ch];
            out_depth[pix] = Dp;
        }
    }
    return blended;
}

static inline void atomic_addf(float* p, float v) {
#pragma omp atomic
    *p += v;
}

/* renderCUDA<3,64> backward (backward.cu:399-612), one pixel at a time, back to front, with the
 * reference's per-channel recurrences (accum_rec / last_color, :546-577).  Outputs are
 * accumulated into (caller zeroes them): dL_dmean2D [P,3], dL_dconic [P,4] (slots 0,1,3),
 * dL_dopacity [P], dL_dcolors [P,3], dL_dlang [P,64], dL_ddepths [P]. */
void oracle_render_bwd(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* bg,
                       const float* means2D, const float* conic_opacity, const float* colors, const float* lang_feat,
                       const float* depths, const float* final_T, const uint32_t* n_contrib, const float* dL_dpix,
                       const float* dL_dpix_lf, const float* dL_dpix_depth, int include_lf, float* dL_dmean2D,
                       float* dL_dconic, float* dL_dopacity, float* dL_dcolors, float* dL_dlang, float* dL_ddepths) {
    const int tiles_x = (W + TILE - 1) / TILE;
    const size_t HW = (size_t)H * W;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
#pragma omp parallel for schedule(dynamic, 8)
    for (int py = 0; py < H; ++py) {
        for (int px = 0; px < W; ++px) {
            const size_t pix = (size_t)py * W + px;
            const uint32_t tile = (uint32_t)((py / TILE) * tiles_x + px / TILE);
            const uint32_t beg = ranges[2 * tile];
            const uint32_t last = n_contrib[pix];
            const float T_final = final_T[pix];
            float T = T_final;
            float g[NCH], gl[LF], acc_c[NCH] = {0, 0, 0}, last_c[NCH] = {0, 0, 0}, acc_l[LF], last_l[LF];
            float acc_d = 0.f, last_d = 0.f, last_alpha = 0.f;
            for (int ch = 0; ch < NCH; ++ch) g[ch] = dL_dpix[ch * HW + pix];
            for (int ch = 0; ch < LF; ++ch) {
                gl[ch] = include_lf ? dL_dpix_lf[ch * HW + pix] : 0.f;
                acc_l[ch] = 0.f;
                last_l[ch] = 0.f;
            }
            const float gd = dL_dpix_depth[pix];
            float bg_dot = 0.f;
            for (int ch = 0; ch < NCH; ++ch) bg_dot += bg[ch] * g[ch];
            for (uint32_t k = last; k-- > 0;) { /* positions last-1 .. 0 */
                const uint32_t id = point_list[beg + k];
                const float dx = means2D[2 * id] - (float)px, dy = means2D[2 * id + 1] - (float)py;
                const float* co = conic_opacity + 4 * (size_t)id;
                const float power = eval_power(dx, dy, co[0], co[1], co[2]);
                if (power > 0.0f) continue;
                const float G = expf(power);
                const float alpha = fminf(0.99f, co[3] * G);
                if (alpha < 1.0f / 255.0f) continue;
                T = T / (1.f - alpha);
                const float w = alpha * T;
                float dL_dalpha = 0.f;
                for (int ch = 0; ch < NCH; ++ch) {
                    const float c = colors[3 * (size_t)id + ch];
                    acc_c[ch] = last_alpha * last_c[ch] + (1.f - last_alpha) * acc_c[ch];
                    last_c[ch] = c;
                    dL_dalpha += (c - acc_c[ch]) * g[ch];
                    atomic_addf(&dL_dcolors[3 * (size_t)id + ch], w * g[ch]);
                }
                if (include_lf)
                    for (int ch = 0; ch < LF; ++ch) {
                        const float c = lang_feat[LF * (size_t)id + ch];
                        acc_l[ch] = last_alpha * last_l[ch] + (1.f - last_alpha) * acc_l[ch];
                        last_l[ch] = c;
                        dL_dalpha += (c - acc_l[ch]) * gl[ch];
                        atomic_addf(&dL_dlang[LF * (size_t)id + ch], w * gl[ch]);
                    }
                const float dth = depths[id];
                acc_d = last_alpha * last_d + (1.f - last_alpha) * acc_d;
                last_d = dth;
                dL_dalpha += (dth - acc_d) * gd;
                atomic_addf(&dL_ddepths[id], w * gd);
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot; /* :585-589 */
                const float dL_dG = co[3] * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                const float dG_ddelx = -gdx * co[0] - gdy * co[1];
                const float dG_ddely = -gdy * co[2] - gdx * co[1];
                atomic_addf(&dL_dmean2D[3 * (size_t)id + 0], dL_dG * dG_ddelx * ddelx_dx);
                atomic_addf(&dL_dmean2D[3 * (size_t)id + 1], dL_dG * dG_ddely * ddely_dy);
                atomic_addf(&dL_dconic[4 * (size_t)id + 0], -0.5f * gdx * dx * dL_dG);
                atomic_addf(&dL_dconic[4 * (size_t)id + 1], -0.5f * gdx * dy * dL_dG);
                atomic_addf(&dL_dconic[4 * (size_t)id + 3], -0.5f * gdy * dy * dL_dG);
                atomic_addf(&dL_dopacity[id], G * dL_dalpha);
            }
        }
    }
}

/* computeCov2DCUDA + preprocessCUDA<3> backward (backward.cu:144-274, 346-396) with the SH
 * backward (:20-139) and computeCov3D backward (:278-341).  Writes only visible Gaussians
 * (radii > 0); the caller zeroes the outputs. */
void oracle_preprocess_bwd(int P, int D, int M, const float* means, const int* radii, const float* shs,
                           const uint8_t* clamped, const float* scales, const float* rots, float mod,
                           const float* cov3Ds, const float* view, const float* proj, const float* campos, int W,
                           int H, float tan_fovx, float tan_fovy, const float* dL_dmean2D, const float* dL_dconic,
                           const float* dL_dcolor, float* dL_dmeans, float* dL_dcov, float* dL_dsh, float* dL_dscale,
                           float* dL_drot) {
    const float h_y = H / (2.0f * tan_fovy), h_x = W / (2.0f * tan_fovx);
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; ++idx) {
        if (!(radii[idx] > 0)) continue;
        const float mx = means[3 * idx], my = means[3 * idx + 1], mz = means[3 * idx + 2];
        const float* cv = cov3Ds + 6 * (size_t)idx;
        /* ---- cov2D backward */
        const float dLx = dL_dconic[4 * idx], dLy = dL_dconic[4 * idx + 1], dLz = dL_dconic[4 * idx + 3];
        float tx = view[0] * mx + view[4] * my + view[8] * mz + view[12];
        float ty = view[1] * mx + view[5] * my + view[9] * mz + view[13];
        const float tz = view[2] * mx + view[6] * my + view[10] * mz + view[14];
        const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
        const float txtz = tx / tz, tytz = ty / tz;
        tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
        ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
        const float xg = (txtz < -limx || txtz > limx) ? 0.f : 1.f, yg = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        /* J, W, T = W*J, Vrk as 3x3 arrays indexed [col][row] like glm */
        float J[3][3] = {{h_x / tz, 0.f, -(h_x * tx) / (tz * tz)}, {0.f, h_y / tz, -(h_y * ty) / (tz * tz)}, {0, 0, 0}};
        float Wm[3][3] = {{view[0], view[4], view[8]}, {view[1], view[5], view[9]}, {view[2], view[6], view[10]}};
        float V[3][3] = {{cv[0], cv[1], cv[2]}, {cv[1], cv[3], cv[4]}, {cv[2], cv[4], cv[5]}};
        float Tm[3][3], A[3][3], C2[3][3];
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) Tm[c][r] = Wm[0][r] * J[c][0] + Wm[1][r] * J[c][1] + Wm[2][r] * J[c][2];
        for (int c = 0; c < 3; ++c) /* A = T^T * V^T : A[c][r] = sum_k T[r][k] * V[k][c] */
            for (int r = 0; r < 3; ++r) A[c][r] = Tm[r][0] * V[0][c] + Tm[r][1] * V[1][c] + Tm[r][2] * V[2][c];
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) C2[c][r] = A[0][r] * Tm[c][0] + A[1][r] * Tm[c][1] + A[2][r] * Tm[c][2];
        const float a = C2[0][0] + 0.3f, b = C2[0][1], c = C2[1][1] + 0.3f;
        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float* dc = dL_dcov + 6 * (size_t)idx;
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dLx + 2 * b * c * dLy + (denom - a * c) * dLz);
            dL_dc = denom2inv * (-a * a * dLz + 2 * a * b * dLy + (denom - a * c) * dLx);
            dL_db = denom2inv * 2 * (b * c * dLx - (denom + 2 * b * b) * dLy + a * b * dLz);
            dc[0] = Tm[0][0] * Tm[0][0] * dL_da + Tm[0][0] * Tm[1][0] * dL_db + Tm[1][0] * Tm[1][0] * dL_dc;
            dc[3] = Tm[0][1] * Tm[0][1] * dL_da + Tm[0][1] * Tm[1][1] * dL_db + Tm[1][1] * Tm[1][1] * dL_dc;
            dc[5] = Tm[0][2] * Tm[0][2] * dL_da + Tm[0][2] * Tm[1][2] * dL_db + Tm[1][2] * Tm[1][2] * dL_dc;
            dc[1] = 2 * Tm[0][0] * Tm[0][1] * dL_da + (Tm[0][0] * Tm[1][1] + Tm[0][1] * Tm[1][0]) * dL_db + 2 * Tm[1][0] * Tm[1][1] * dL_dc;
            dc[2] = 2 * Tm[0][0] * Tm[0][2] * dL_da + (Tm[0][0] * Tm[1][2] + Tm[0][2] * Tm[1][0]) * dL_db + 2 * Tm[1][0] * Tm[1][2] * dL_dc;
            dc[4] = 2 * Tm[0][2] * Tm[0][1] * dL_da + (Tm[0][1] * Tm[1][2] + Tm[0][2] * Tm[1][1]) * dL_db + 2 * Tm[1][1] * Tm[1][2] * dL_dc;
        } else {
            for (int i = 0; i < 6; ++i) dc[i] = 0;
        }
        float dT0[3], dT1[3];
        for (int k = 0; k < 3; ++k) {
            const float t0v = Tm[0][0] * V[k][0] + Tm[0][1] * V[k][1] + Tm[0][2] * V[k][2];
            const float t1v = Tm[1][0] * V[k][0] + Tm[1][1] * V[k][1] + Tm[1][2] * V[k][2];
            dT0[k] = 2 * t0v * dL_da + t1v * dL_db;
            dT1[k] = 2 * t1v * dL_dc + t0v * dL_db;
        }
        const float dJ00 = Wm[0][0] * dT0[0] + Wm[0][1] * dT0[1] + Wm[0][2] * dT0[2];
        const float dJ02 = Wm[2][0] * dT0[0] + Wm[2][1] * dT0[1] + Wm[2][2] * dT0[2];
        const float dJ11 = Wm[1][0] * dT1[0] + Wm[1][1] * dT1[1] + Wm[1][2] * dT1[2];
        const float dJ12 = Wm[2][0] * dT1[0] + Wm[2][1] * dT1[1] + Wm[2][2] * dT1[2];
        const float itz = 1.f / tz, itz2 = itz * itz, itz3 = itz2 * itz;
        const float dtx = xg * -h_x * itz2 * dJ02, dty = yg * -h_y * itz2 * dJ12;
        const float dtz = -h_x * itz2 * dJ00 - h_y * itz2 * dJ11 + (2 * h_x * tx) * itz3 * dJ02 + (2 * h_y * ty) * itz3 * dJ12;
        float dm[3] = {view[0] * dtx + view[1] * dty + view[2] * dtz, view[4] * dtx + view[5] * dty + view[6] * dtz,
                       view[8] * dtx + view[9] * dty + view[10] * dtz};
        /* ---- mean2D -> mean3D through the projection (:366-385) */
        {
            const float hx = proj[0] * mx + proj[4] * my + proj[8] * mz + proj[12];
            const float hy = proj[1] * mx + proj[5] * my + proj[9] * mz + proj[13];
            const float hw = proj[3] * mx + proj[7] * my + proj[11] * mz + proj[15];
            const float m_w = 1.0f / (hw + 0.0000001f);
            const float mul1 = hx * m_w * m_w, mul2 = hy * m_w * m_w;
            const float gx = dL_dmean2D[3 * idx], gy = dL_dmean2D[3 * idx + 1];
            dm[0] += (proj[0] * m_w - proj[3] * mul1) * gx + (proj[1] * m_w - proj[3] * mul2) * gy;
            dm[1] += (proj[4] * m_w - proj[7] * mul1) * gx + (proj[5] * m_w - proj[7] * mul2) * gy;
            dm[2] += (proj[8] * m_w - proj[11] * mul1) * gx + (proj[9] * m_w - proj[11] * mul2) * gy;
        }
        /* ---- SH backward (:20-139) */
        if (shs) {
            const float* sh = shs + (size_t)idx * M * 3;
            float* dsh = dL_dsh + (size_t)idx * M * 3;
            float dRGB[3];
            for (int ch = 0; ch < 3; ++ch) dRGB[ch] = clamped[3 * idx + ch] ? 0.f : dL_dcolor[3 * idx + ch];
            const float ox = mx - campos[0], oy = my - campos[1], oz = mz - campos[2];
            const float len = sqrtf(ox * ox + oy * oy + oz * oz);
            const float x = ox / len, y = oy / len, z = oz / len;
            float w[16], ddx[16], ddy[16], ddz[16]; /* basis and its derivatives w.r.t. dir */
            for (int k = 0; k < 16; ++k) w[k] = ddx[k] = ddy[k] = ddz[k] = 0.f;
            w[0] = SH_C0;
            if (D > 0) {
                w[1] = -SH_C1 * y; w[2] = SH_C1 * z; w[3] = -SH_C1 * x;
                ddx[3] = -SH_C1; ddy[1] = -SH_C1; ddz[2] = SH_C1;
                if (D > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    w[4] = SH_C2[0] * xy; w[5] = SH_C2[1] * yz; w[6] = SH_C2[2] * (2.f * zz - xx - yy);
                    w[7] = SH_C2[3] * xz; w[8] = SH_C2[4] * (xx - yy);
                    ddx[4] = SH_C2[0] * y; ddx[6] = SH_C2[2] * 2.f * -x; ddx[7] = SH_C2[3] * z; ddx[8] = SH_C2[4] * 2.f * x;
                    ddy[4] = SH_C2[0] * x; ddy[5] = SH_C2[1] * z; ddy[6] = SH_C2[2] * 2.f * -y; ddy[8] = SH_C2[4] * 2.f * -y;
                    ddz[5] = SH_C2[1] * y; ddz[6] = SH_C2[2] * 2.f * 2.f * z; ddz[7] = SH_C2[3] * x;
                    if (D > 2) {
                        w[9] = SH_C3[0] * y * (3.f * xx - yy); w[10] = SH_C3[1] * xy * z;
                        w[11] = SH_C3[2] * y * (4.f * zz - xx - yy);
                        w[12] = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                        w[13] = SH_C3[4] * x * (4.f * zz - xx - yy); w[14] = SH_C3[5] * z * (xx - yy);
                        w[15] = SH_C3[6] * x * (xx - 3.f * yy);
                        ddx[9] = SH_C3[0] * 3.f * 2.f * xy; ddx[10] = SH_C3[1] * yz; ddx[11] = SH_C3[2] * -2.f * xy;
                        ddx[12] = SH_C3[3] * -3.f * 2.f * xz; ddx[13] = SH_C3[4] * (-3.f * xx + 4.f * zz - yy);
                        ddx[14] = SH_C3[5] * 2.f * xz; ddx[15] = SH_C3[6] * 3.f * (xx - yy);
                        ddy[9] = SH_C3[0] * 3.f * (xx - yy); ddy[10] = SH_C3[1] * xz;
                        ddy[11] = SH_C3[2] * (-3.f * yy + 4.f * zz - xx); ddy[12] = SH_C3[3] * -3.f * 2.f * yz;
                        ddy[13] = SH_C3[4] * -2.f * xy; ddy[14] = SH_C3[5] * -2.f * yz; ddy[15] = SH_C3[6] * -3.f * 2.f * xy;
                        ddz[10] = SH_C3[1] * xy; ddz[11] = SH_C3[2] * 4.f * 2.f * yz;
                        ddz[12] = SH_C3[3] * 3.f * (2.f * zz - xx - yy); ddz[13] = SH_C3[4] * 4.f * 2.f * xz;
                        ddz[14] = SH_C3[5] * (xx - yy);
                    }
                }
            }
            float dir_g[3] = {0, 0, 0};
            const int ncoef = (D + 1) * (D + 1);
            for (int k = 0; k < ncoef; ++k)
                for (int ch = 0; ch < 3; ++ch) {
                    dsh[3 * k + ch] = w[k] * dRGB[ch];
                    dir_g[0] += ddx[k] * sh[3 * k + ch] * dRGB[ch];
                    dir_g[1] += ddy[k] * sh[3 * k + ch] * dRGB[ch];
                    dir_g[2] += ddz[k] * sh[3 * k + ch] * dRGB[ch];
                }
            /* dnormvdv (auxiliary.h:106-116) */
            const float sum2 = ox * ox + oy * oy + oz * oz;
            const float inv32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
            dm[0] += ((sum2 - ox * ox) * dir_g[0] - oy * ox * dir_g[1] - oz * ox * dir_g[2]) * inv32;
            dm[1] += (-ox * oy * dir_g[0] + (sum2 - oy * oy) * dir_g[1] - oz * oy * dir_g[2]) * inv32;
            dm[2] += (-ox * oz * dir_g[0] - oy * oz * dir_g[1] + (sum2 - oz * oz) * dir_g[2]) * inv32;
        }
        dL_dmeans[3 * idx] = dm[0]; dL_dmeans[3 * idx + 1] = dm[1]; dL_dmeans[3 * idx + 2] = dm[2];
        /* ---- cov3D -> scale / rotation (:278-341), no quaternion-normalisation Jacobian (:340) */
        if (scales) {
            const float r = rots[4 * idx], x = rots[4 * idx + 1], y = rots[4 * idx + 2], z = rots[4 * idx + 3];
            float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                             {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                             {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
            const float s[3] = {mod * scales[3 * idx], mod * scales[3 * idx + 1], mod * scales[3 * idx + 2]};
            float Mm[3][3], Sg[3][3], dM[3][3];
            for (int c2 = 0; c2 < 3; ++c2)
                for (int r2 = 0; r2 < 3; ++r2) Mm[c2][r2] = s[r2] * R[c2][r2]; /* M = S*R */
            Sg[0][0] = dc[0]; Sg[1][1] = dc[3]; Sg[2][2] = dc[5];
            Sg[0][1] = Sg[1][0] = 0.5f * dc[1]; Sg[0][2] = Sg[2][0] = 0.5f * dc[2]; Sg[1][2] = Sg[2][1] = 0.5f * dc[4];
            for (int c2 = 0; c2 < 3; ++c2) /* dL_dM = 2 * M * dL_dSigma */
                for (int r2 = 0; r2 < 3; ++r2)
                    dM[c2][r2] = 2.0f * (Mm[0][r2] * Sg[c2][0] + Mm[1][r2] * Sg[c2][1] + Mm[2][r2] * Sg[c2][2]);
            float dMt[3][3]; /* dMt[k][c] = dM[c][k] */
            for (int k = 0; k < 3; ++k) {
                float acc = 0.f;
                for (int c2 = 0; c2 < 3; ++c2) { dMt[k][c2] = dM[c2][k]; acc += R[c2][k] * dM[c2][k]; }
                dL_dscale[3 * idx + k] = acc;
                for (int c2 = 0; c2 < 3; ++c2) dMt[k][c2] *= s[k];
            }
            float* dq = dL_drot + 4 * (size_t)idx;
            dq[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
            dq[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
            dq[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
            dq[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
        }
    }
}

/* torch::optim::Adam::step of libtorch 2.0.1 for one tensor (reference src/gaussian_model.cpp:
 * 483-518; step at src/gaussian_mapper.cpp:793-796): beta/eps/lr are doubles cast to float at
 * each ATen op; division by the scalar sqrt(bias_correction2) is a multiply by its float
 * reciprocal; a + alpha*b contracts to fma on the GPU. */
void oracle_adam(int64_t n, float* p, const float* g, float* m, float* v, double lr, double beta1, double beta2,
                 double eps, int step) {
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    const float b1 = (float)beta1, b2 = (float)beta2, omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
    const float inv_bc2_sqrt = 1.0f / (float)sqrt(bc2), neg_step = (float)(-(lr / bc1)), epsf = (float)eps;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        m[i] = fmaf(omb1, g[i], m[i] * b1);
        v[i] = fmaf(omb2 * g[i], g[i], v[i] * b2);
        const float dn = sqrtf(v[i]) * inv_bc2_sqrt + epsf;
        p[i] = fmaf(neg_step, m[i] / dn, p[i]);
    }
}

/* F.normalize(feats, dim=1) @ F.normalize(text, dim=1).T  (eval/find_objects_gaussians.py:
 * 160-172), accumulated in double so it is the more accurate side of the comparison. */
void oracle_cosine(int P, int Q, const float* feats, const float* text, float* out) {
#pragma omp parallel for schedule(static)
    for (int p = 0; p < P; ++p) {
        double fn = 0;
        for (int k = 0; k < LF; ++k) fn += (double)feats[(size_t)p * LF + k] * feats[(size_t)p * LF + k];
        fn = fmax(sqrt(fn), 1e-12);
        for (int q = 0; q < Q; ++q) {
            double tn = 0, d = 0;
            for (int k = 0; k < LF; ++k) {
                tn += (double)text[(size_t)q * LF + k] * text[(size_t)q * LF + k];
                d += (double)feats[(size_t)p * LF + k] * text[(size_t)q * LF + k];
            }
            tn = fmax(sqrt(tn), 1e-12);
            out[(size_t)p * Q + q] = (float)(d / (fn * tn));
        }
    }
}

/* GaussianModel::exponLrFunc (reference src/gaussian_model.cpp:1143-1157): the xyz learning rate of an iteration, in float
 * with libm's float functions exactly as the compiled reference evaluates it (std::log / std::exp / std::sin on floats,
 * M_PI_2f32, std::clamp). */
float oracle_expon_lr(int step, float lr_init, float lr_final, float lr_delay_mult, int lr_delay_steps, int max_steps) {
    if (step < 0 || (lr_init == 0.0f && lr_final == 0.0f)) return 0.0f;
    float delay_rate = 1.0f;
    if (lr_delay_steps > 0) {
        float x = (float)step / lr_delay_steps;
        x = x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
        delay_rate = lr_delay_mult + (1.0f - lr_delay_mult) * sinf(1.57079632679489661923f * x);
    }
    float t = (float)step / max_steps;
    t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
    const float log_lerp = expf(logf(lr_init) * (1 - t) + logf(lr_final) * t);
    return delay_rate * log_lerp;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
