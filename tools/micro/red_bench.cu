// Microbenchmark: throughput of red.global.add.f32 / .v2 / .v4 with the render-backward access pattern
// (random 256-byte rows of a [P][64] fp32 array, one warp instruction covers 128/256/512 contiguous bytes).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
template <int V>
__global__ void red_kernel(float* out, int P, int recs_per_warp) {
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    // V floats per lane: a warp instruction covers 32*V floats = V/2 records of 64 floats
    for (int r = 0; r < recs_per_warp * 2 / V; ++r) {
        if (V == 1) {
            for (int h = 0; h < 2; ++h) {
                const uint32_t id = hash(gw * 4099u + r) % P;
                float* p = out + (size_t)id * 64 + h * 32 + lane;
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.0f) : "memory");
            }
        } else if (V == 2) {
            const uint32_t id = hash(gw * 4099u + r) % P;
            float* p = out + (size_t)id * 64 + lane * 2;
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(1.0f), "f"(1.0f) : "memory");
        } else {
            const uint32_t id = hash(gw * 4099u + 2 * r + (lane >> 4)) % P;
            float* p = out + (size_t)id * 64 + (lane & 15) * 4;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.0f), "f"(1.0f), "f"(1.0f), "f"(1.0f) : "memory");
        }
    }
}
// V = 4, transposed layout: every lane owns ONE record (TMEM lane = record) and walks its 16 groups of 4 channels: a warp
// instruction touches 32 different rows, 16 bytes each
__global__ void red_kernel_lane_per_record(float* out, int P, int recs_per_warp) {
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int r = 0; r < recs_per_warp / 32; ++r) {
        const uint32_t id = hash(gw * 4099u + r * 32 + lane) % P;
        float* p = out + (size_t)id * 64;
#pragma unroll
        for (int c = 0; c < 16; ++c)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p + 4 * c), "f"(1.0f), "f"(1.0f), "f"(1.0f), "f"(1.0f) : "memory");
    }
}
float run_lpr(float* out, int P, int blocks, int threads, int total_recs) {
    const int warps = blocks * threads / 32;
    const int rpw = (total_recs / warps) & ~31;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    red_kernel_lane_per_record<<<blocks, threads>>>(out, P, rpw);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) red_kernel_lane_per_record<<<blocks, threads>>>(out, P, rpw);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("V=4 lane-per-record blocks=%d threads=%d recs=%d: %.3f ms per launch, %.1f G float-adds/s\n", blocks, threads, rpw * warps, ms / 5, (double)rpw * warps * 64 / (ms / 5) / 1e6);
    return ms / 5;
}
template <int V>
float run(float* out, int P, int blocks, int threads, int total_recs) {
    const int warps = blocks * threads / 32;
    const int rpw = (total_recs / warps) & ~1;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    red_kernel<V><<<blocks, threads>>>(out, P, rpw);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) red_kernel<V><<<blocks, threads>>>(out, P, rpw);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("V=%d blocks=%d threads=%d recs=%d: %.3f ms per launch, %.1f G float-adds/s\n", V, blocks, threads, rpw * warps, ms / 5, (double)rpw * warps * 64 / (ms / 5) / 1e6);
    return ms / 5;
}
int main() {
    const int P = 200000; float* out; cudaMalloc(&out, (size_t)P * 64 * 4); cudaMemset(out, 0, (size_t)P * 64 * 4);
    const int total = 1800000;  // records of 64 floats
    for (int cfg = 0; cfg < 3; ++cfg) {
        const int blocks = cfg == 0 ? 296 : cfg == 1 ? 148 * 8 : 148 * 16, threads = cfg == 0 ? 64 : cfg == 1 ? 128 : 128;
        run<1>(out, P, blocks, threads, total); run<2>(out, P, blocks, threads, total); run<4>(out, P, blocks, threads, total);
        run_lpr(out, P, blocks, threads, total);
    }
    cudaError_t e = cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(e));
    return 0;
}
