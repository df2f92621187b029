// l0_driver.cpp -- exercises the drop-in L0 C++ interface exactly as the reference's L1 does
// (src/rasterize_points.cu:29-35,85-118,169-205): std::function allocators that grow device
// buffers, CudaRasterizer::Rasterizer::forward, pre-zeroed gradient buffers, ::backward.
// Reads one test case from a flat binary file, writes every output to another.
//   l0_driver in.bin out.bin
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

#include "cuda_rasterizer/rasterizer.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(2); } } while (0)

struct DevBuf {
    char* p = nullptr;
    size_t n = 0;
    std::function<char*(size_t)> resizer() {
        return [this](size_t N) {
            if (p) cudaFree(p);
            CK(cudaMalloc(&p, N ? N : 1));
            n = N;
            return p;
        };
    }
};

static std::vector<float> rd(FILE* f, size_t n) {
    std::vector<float> v(n);
    if (n && fread(v.data(), 4, n, f) != n) { fprintf(stderr, "short read\n"); exit(3); }
    return v;
}
static float* up(const std::vector<float>& v) {
    if (v.empty()) return nullptr;
    float* d;
    CK(cudaMalloc(&d, v.size() * 4));
    CK(cudaMemcpy(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
    return d;
}
static float* zeros(size_t n) {
    float* d;
    CK(cudaMalloc(&d, (n ? n : 1) * 4));
    CK(cudaMemset(d, 0, (n ? n : 1) * 4));
    return d;
}
static void wr(FILE* f, const float* d, size_t n) {
    std::vector<float> v(n);
    if (n) CK(cudaMemcpy(v.data(), d, n * 4, cudaMemcpyDeviceToHost));
    fwrite(v.data(), 4, n, f);
}

int main(int argc, char** argv) {
    if (argc < 3) return 1;
    FILE* f = fopen(argv[1], "rb");
    int hdr[6];  // P, W, H, D, M, include_lf
    if (!f || fread(hdr, 4, 6, f) != 6) return 1;
    const int P = hdr[0], W = hdr[1], H = hdr[2], D = hdr[3], M = hdr[4], lf_on = hdr[5];
    float sc[3];  // tanfovx, tanfovy, scale_modifier
    if (fread(sc, 4, 3, f) != 3) return 1;
    auto bg = rd(f, 3), means = rd(f, (size_t)P * 3), shs = rd(f, (size_t)P * M * 3), lf = rd(f, (size_t)P * 64),
         opac = rd(f, P), scales = rd(f, (size_t)P * 3), rots = rd(f, (size_t)P * 4), view = rd(f, 16), proj = rd(f, 16),
         cam = rd(f, 3), gcol = rd(f, (size_t)3 * H * W), glf = rd(f, (size_t)64 * H * W), gdep = rd(f, (size_t)H * W);
    fclose(f);
    float *d_bg = up(bg), *d_means = up(means), *d_shs = up(shs), *d_lf = up(lf), *d_opac = up(opac), *d_scales = up(scales),
          *d_rots = up(rots), *d_view = up(view), *d_proj = up(proj), *d_cam = up(cam), *d_gcol = up(gcol), *d_glf = up(glf),
          *d_gdep = up(gdep);
    const size_t HW = (size_t)H * W;
    float *out_color = zeros(3 * HW), *out_lf = zeros(64 * HW), *out_depth = zeros(HW);
    int* radii;
    CK(cudaMalloc(&radii, (size_t)(P ? P : 1) * 4));
    CK(cudaMemset(radii, 0, (size_t)(P ? P : 1) * 4));
    DevBuf geom, binning, img;
    const int R = CudaRasterizer::Rasterizer::forward(geom.resizer(), binning.resizer(), img.resizer(), P, D, M, d_bg, W, H,
                                                      d_means, d_shs, nullptr, d_lf, d_opac, d_scales, sc[2], d_rots, nullptr,
                                                      d_view, d_proj, d_cam, sc[0], sc[1], false, out_color, out_lf, out_depth,
                                                      radii, lf_on != 0);
    float *g_m2d = zeros((size_t)P * 3), *g_conic = zeros((size_t)P * 4), *g_op = zeros(P), *g_col = zeros((size_t)P * 3),
          *g_lf = zeros((size_t)P * 64), *g_dep = zeros(P), *g_m3d = zeros((size_t)P * 3), *g_cov = zeros((size_t)P * 6),
          *g_sh = zeros((size_t)P * M * 3), *g_sc = zeros((size_t)P * 3), *g_rot = zeros((size_t)P * 4);
    CudaRasterizer::Rasterizer::backward(P, D, M, R, d_bg, W, H, d_means, d_shs, nullptr, d_lf, d_scales, sc[2], d_rots, nullptr,
                                         d_view, d_proj, d_cam, sc[0], sc[1], radii, geom.p, binning.p, img.p, d_gcol, d_glf,
                                         d_gdep, g_m2d, g_conic, g_op, g_col, g_lf, g_dep, g_m3d, g_cov, g_sh, g_sc, g_rot,
                                         lf_on != 0);
    bool* present;
    CK(cudaMalloc(&present, P ? P : 1));
    CudaRasterizer::Rasterizer::markVisible(P, d_means, d_view, d_proj, present);
    CK(cudaDeviceSynchronize());
    FILE* o = fopen(argv[2], "wb");
    fwrite(&R, 4, 1, o);
    std::vector<int> hr(P);
    CK(cudaMemcpy(hr.data(), radii, (size_t)P * 4, cudaMemcpyDeviceToHost));
    fwrite(hr.data(), 4, P, o);
    std::vector<unsigned char> hp(P);
    CK(cudaMemcpy(hp.data(), present, P, cudaMemcpyDeviceToHost));
    fwrite(hp.data(), 1, P, o);
    wr(o, out_color, 3 * HW); wr(o, out_lf, 64 * HW); wr(o, out_depth, HW);
    wr(o, g_m2d, (size_t)P * 3); wr(o, g_col, (size_t)P * 3); wr(o, g_lf, (size_t)P * 64); wr(o, g_op, P);
    wr(o, g_m3d, (size_t)P * 3); wr(o, g_cov, (size_t)P * 6); wr(o, g_sh, (size_t)P * M * 3); wr(o, g_sc, (size_t)P * 3);
    wr(o, g_rot, (size_t)P * 4);
    fclose(o);
    printf("l0_driver ok R=%d\n", R);
    return 0;
}
