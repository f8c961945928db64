// Probe: how fast can the dense maps be zero-filled, and what slows the mask-aware variant down?
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

// variant 0: unconditional, 4 maps interleaved per warp chunk (512 voxels), like the shipped fill
// variant 1: mask-aware (loads mask words, conditional stores)
// variant 2: unconditional, one map after another (4 sequential passes inside one kernel)
// variant 3: mask-aware, each thread handles 1 word (4 voxels) per round, grid-stride, words consecutive across lanes
template <int V>
__global__ void __launch_bounds__(256) fill(float* m0, float* m1, float* m2, float* m3, const uint8_t* mask, int64_t n) {
    const float4 z = make_float4(0, 0, 0, 0);
    const int lane = threadIdx.x & 31;
    const int64_t gw = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * 256) >> 5;
    if (V == 2) {
        float* maps[4] = {m0, m1, m2, m3};
        for (int m = 0; m < 4; ++m)
            for (int64_t c = gw; c * 512 < n; c += nw) {
                float4* p = reinterpret_cast<float4*>(maps[m] + c * 512 + lane * 4);
#pragma unroll
                for (int j = 0; j < 4; ++j) p[j * 32] = z;
            }
        return;
    }
    if (V == 3) {
        const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x, nt = (int64_t)gridDim.x * 256;
        for (int64_t wi = t; wi * 4 < n; wi += nt) {
            const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(mask) + wi);
            reinterpret_cast<float4*>(m3)[wi] = z;
            if (w == 0) { reinterpret_cast<float4*>(m0)[wi] = z; reinterpret_cast<float4*>(m1)[wi] = z; reinterpret_cast<float4*>(m2)[wi] = z; }
        }
        return;
    }
    for (int64_t c = gw; c * 512 < n; c += nw) {
        const int64_t off = c * 512 + lane * 4;
        uint32_t w[4] = {0, 0, 0, 0};
        if (V == 1) {
            const uint32_t* pm = reinterpret_cast<const uint32_t*>(mask + c * 512) + lane;
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = __ldg(pm + j * 32);
        }
        float4 *p0 = reinterpret_cast<float4*>(m0 + off), *p1 = reinterpret_cast<float4*>(m1 + off),
               *p2 = reinterpret_cast<float4*>(m2 + off), *p3 = reinterpret_cast<float4*>(m3 + off);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            p3[j * 32] = z;
            if (w[j] == 0) { p0[j * 32] = z; p1[j * 32] = z; p2[j * 32] = z; }
        }
    }
}

int main() {
    const int64_t n = 256LL * 256 * 256;
    float* maps; uint8_t* mask;
    CK(cudaMalloc(&maps, 4 * n * sizeof(float)));
    CK(cudaMalloc(&mask, n));
    std::vector<uint8_t> h(n, 0);
    for (int64_t z = 60; z < 196; ++z) for (int64_t y = 40; y < 216; ++y) for (int64_t x = 58; x < 198; ++x) h[(z * 256 + y) * 256 + x] = 1;
    CK(cudaMemcpy(mask, h.data(), n, cudaMemcpyHostToDevice));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](int v, int grid) {
        for (int rep = 0; rep < 3; ++rep) {
            if (rep == 2) cudaEventRecord(a);
            for (int it = 0; it < (rep == 2 ? 20 : 3); ++it) {
                switch (v) {
                    case 0: fill<0><<<grid, 256>>>(maps, maps + n, maps + 2 * n, maps + 3 * n, mask, n); break;
                    case 1: fill<1><<<grid, 256>>>(maps, maps + n, maps + 2 * n, maps + 3 * n, mask, n); break;
                    case 2: fill<2><<<grid, 256>>>(maps, maps + n, maps + 2 * n, maps + 3 * n, mask, n); break;
                    case 3: fill<3><<<grid, 256>>>(maps, maps + n, maps + 2 * n, maps + 3 * n, mask, n); break;
                }
            }
        }
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("variant %d grid %5d: %7.2f us  (%.2f TB/s of 268 MB)\n", v, grid, ms / 20 * 1e3, 268.4e6 / (ms / 20 * 1e-3) / 1e12);
    };
    for (int v = 0; v < 4; ++v) for (int grid : {148, 296, 592, 1184, 4096, 6328, 16384}) run(v, grid);
    for (int it = 0; it < 3; ++it) cudaMemsetAsync(maps, 0, 4 * n * sizeof(float));
    cudaEventRecord(a);
    for (int it = 0; it < 20; ++it) cudaMemsetAsync(maps, 0, 4 * n * sizeof(float));
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("cudaMemsetAsync: %7.2f us\n", ms / 20 * 1e3);
    return 0;
}
