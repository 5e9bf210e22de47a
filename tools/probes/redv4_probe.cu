// redv4_probe.cu -- correctness of red.global.add.v4.f32 under intra-warp address collisions.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ void red_v4(float *p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
template <int MODE>
__global__ void k(const uint32_t *slot, uint32_t n, float *acc)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = slot[i];
    float *row = acc + (size_t)s * 8;
    if (MODE == 0) { red_v4(row, 1.f, 1.f, 1.f, 1.f); red_v4(row + 4, 1.f, 1.f, 1.f, 1.f); }
    if (MODE == 1) { for (int q = 0; q < 8; ++q) atomicAdd(row + q, 1.0f); }
    if (MODE == 2) { atomicAdd(reinterpret_cast<float4 *>(row), make_float4(1.f, 1.f, 1.f, 1.f)); atomicAdd(reinterpret_cast<float4 *>(row + 4), make_float4(1.f, 1.f, 1.f, 1.f)); }
}
int main()
{
    const uint32_t n = 1 << 21, rows = 1 << 16;
    for (int pattern = 0; pattern < 3; ++pattern) {
        std::vector<uint32_t> h(n);
        std::vector<float> cnt(rows, 0.f);
        uint64_t st = 88172645463325252ull;
        auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
        for (uint32_t i = 0; i < n; ++i) {
            uint32_t s;
            if (pattern == 0) s = rnd() % rows;                       // random
            else if (pattern == 1) s = (i / 7) % rows;                // runs of 7 consecutive lanes share a row
            else s = ((i / 32) * 4 + (i % 4)) % rows;                 // each warp: 4 rows, 8 lanes each, interleaved
            h[i] = s; cnt[s] += 1.f;
        }
        uint32_t *slot; float *acc;
        CK(cudaMalloc(&slot, n * 4)); CK(cudaMalloc(&acc, (size_t)rows * 32));
        CK(cudaMemcpy(slot, h.data(), n * 4, cudaMemcpyHostToDevice));
        for (int mode = 0; mode < 3; ++mode) {
            CK(cudaMemset(acc, 0, (size_t)rows * 32));
            if (mode == 0) k<0><<<n / 256, 256>>>(slot, n, acc);
            if (mode == 1) k<1><<<n / 256, 256>>>(slot, n, acc);
            if (mode == 2) k<2><<<n / 256, 256>>>(slot, n, acc);
            CK(cudaDeviceSynchronize());
            std::vector<float> out((size_t)rows * 8);
            CK(cudaMemcpy(out.data(), acc, (size_t)rows * 32, cudaMemcpyDeviceToHost));
            long bad = 0; double lost = 0;
            for (uint32_t r = 0; r < rows; ++r) for (int q = 0; q < 8; ++q) if (out[(size_t)r * 8 + q] != cnt[r]) { ++bad; lost += cnt[r] - out[(size_t)r * 8 + q]; }
            printf("pattern %d mode %d (%s): %ld wrong elements, %.0f lost increments\n", pattern, mode,
                   mode == 0 ? "red.v4.f32" : mode == 1 ? "scalar atomicAdd" : "atomicAdd(float4)", bad, lost);
        }
        cudaFree(slot); cudaFree(acc);
    }
    return 0;
}
