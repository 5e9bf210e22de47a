// atomics_probe.cu -- measures L2 atomic / reduction throughput on B200 for the access pattern of
// the voxelizer's insert pass: N points, each hitting one of ~600k occupied cells out of 2M.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o atomics_probe atomics_probe.cu
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
__device__ __forceinline__ void red_v2(float *p, float a, float b)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

template <int MODE>
__global__ void k(const uint32_t *__restrict__ slot, uint32_t n, uint32_t *first, uint32_t *cnt, float *acc, uint32_t *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = slot[i];
    if (s == 0xFFFFFFFFu) return;
    if (MODE == 0) atomicMin(first + s, i);                                  // RED.MIN
    if (MODE == 1) out[i] = atomicAdd(cnt + s, 1u);                          // ATOM.ADD with return
    if (MODE == 2) atomicAdd(acc + (size_t)s * 8, 1.0f);                     // RED.ADD.F32
    if (MODE == 3) red_v4(acc + (size_t)s * 8, 1.f, 2.f, 3.f, 4.f);
    if (MODE == 4) { red_v4(acc + (size_t)s * 8, 1.f, 2.f, 3.f, 4.f); red_v4(acc + (size_t)s * 8 + 4, 1.f, 2.f, 3.f, 4.f); }
    if (MODE == 5) { atomicMin(first + s, i); red_v4(acc + (size_t)s * 8, 1.f, 2.f, 3.f, 4.f); red_v4(acc + (size_t)s * 8 + 4, 1.f, 2.f, 3.f, 4.f); }
    if (MODE == 6) { atomicMin(first + s, i); atomicAdd(cnt + s, 1u); }      // two REDs (current K1 without aggregation)
    if (MODE == 7) {                                                          // warp-aggregated min + add (current K1)
        const unsigned peers = __match_any_sync(__activemask(), s);
        if ((int)(threadIdx.x & 31u) == __ffs(peers) - 1) { atomicMin(first + s, i); atomicAdd(cnt + s, (uint32_t)__popc(peers)); }
    }
    if (MODE == 8) { for (int q = 0; q < 8; ++q) atomicAdd(acc + (size_t)s * 8 + q, 1.0f); }   // 8 scalar REDs
    if (MODE == 9) red_v2(acc + (size_t)s * 8, 1.f, 2.f);
    if (MODE == 10) out[i] = s;                                              // plain coalesced store baseline
    if (MODE == 11) first[s] = i;                                            // plain scattered store
    if (MODE == 12) out[i] = first[s];                                       // plain gather
    if (MODE == 13) {                                                        // 64-bit packed: min in high, can't add; ATOM.MIN.64 w/ return
        out[i] = (uint32_t)atomicMin((unsigned long long *)(acc + (size_t)s * 8), (unsigned long long)i);
    }
}

template <int MODE>
static float run(const char *name, const uint32_t *slot, uint32_t n, uint32_t *first, uint32_t *cnt, float *acc, uint32_t *out, int lanes)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) k<MODE><<<(n + 255) / 256, 256>>>(slot, n, first, cnt, acc, out);
    CK(cudaDeviceSynchronize());
    const int it = 20;
    CK(cudaEventRecord(e0));
    for (int w = 0; w < it; ++w) k<MODE><<<(n + 255) / 256, 256>>>(slot, n, first, cnt, acc, out);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const float us = ms * 1000.f / it;
    printf("%-44s %8.2f us   %7.2f G ops/s\n", name, us, (double)lanes * 1e-3 / us);
    return us;
}

// Sector-pairing probes: the same 2 x 16-byte reductions per point as MODE 4, but both halves of a
// point's 32-byte row leave in ONE warp instruction from two different lanes.
//   PAIR 0: lanes 0-15 carry the first halves of 16 points, lanes 16-31 the second halves
//   PAIR 1: adjacent lanes carry the two halves of one point
//   PAIR 2: control -- 16 points per instruction, first halves only (half the reductions)
//   PAIR 3: 32-byte plain store per lane (st.v8 as two float4 halves by lane pairs), control for request cost
template <int PAIR>
__global__ void kp(const uint32_t *__restrict__ slot, uint32_t n, float *acc)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = t & 31u, wbase = (t >> 5) * 16u;
    uint32_t i, half;
    if (PAIR == 1) { i = wbase + (lane >> 1); half = lane & 1u; }
    else { i = wbase + (lane & 15u); half = lane >> 4; }
    if (i >= n) return;
    const uint32_t s = slot[i];
    if (s == 0xFFFFFFFFu) return;
    if (PAIR == 2 && half) return;
    red_v4(acc + (size_t)s * 8 + 4 * half, 1.f, 2.f, 3.f, 4.f);
}
template <int PAIR>
static float runp(const char *name, const uint32_t *slot, uint32_t n, float *acc, int lanes)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const uint32_t threads = 2 * n;
    for (int w = 0; w < 3; ++w) kp<PAIR><<<(threads + 255) / 256, 256>>>(slot, n, acc);
    CK(cudaDeviceSynchronize());
    const int it = 20;
    CK(cudaEventRecord(e0));
    for (int w = 0; w < it; ++w) kp<PAIR><<<(threads + 255) / 256, 256>>>(slot, n, acc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const float us = ms * 1000.f / it;
    printf("%-44s %8.2f us   %7.2f G points/s\n", name, us, (double)lanes * 1e-3 / us);
    return us;
}

int main(int argc, char **argv)
{
    const uint32_t B = 8, cells = 512 * 512, per = 295000, occupied = 76650;
    const uint32_t n = B * per;
    for (int pattern = 0; pattern < 2; ++pattern) {
        std::vector<uint32_t> h(n);
        std::vector<uint32_t> occ(occupied);
        uint64_t st = 88172645463325252ull;
        auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
        int live = 0;
        for (uint32_t b = 0; b < B; ++b) {
            for (uint32_t q = 0; q < occupied; ++q) occ[q] = (uint32_t)(rnd() % cells);
            for (uint32_t j = 0; j < per; ++j) {
                uint32_t s;
                if (rnd() % 100 < 22) s = 0xFFFFFFFFu;   // out of range
                else if (pattern == 0) s = b * cells + occ[rnd() % occupied];                   // random occupied cell
                else s = b * cells + occ[(uint32_t)(((uint64_t)j * occupied) / per + rnd() % 3) % occupied];  // ring-like: neighbours hit nearby list entries
                h[b * per + j] = s;
                live += s != 0xFFFFFFFFu;
            }
        }
        uint32_t *slot, *first, *cnt, *out; float *acc;
        CK(cudaMalloc(&slot, n * 4)); CK(cudaMalloc(&out, n * 4));
        CK(cudaMalloc(&first, (size_t)B * cells * 4)); CK(cudaMalloc(&cnt, (size_t)B * cells * 4));
        CK(cudaMalloc(&acc, (size_t)B * cells * 32));
        CK(cudaMemcpy(slot, h.data(), n * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(first, 0xFF, (size_t)B * cells * 4)); CK(cudaMemset(cnt, 0, (size_t)B * cells * 4));
        CK(cudaMemset(acc, 0, (size_t)B * cells * 32));
        printf("pattern %d: n=%u live=%d\n", pattern, n, live);
        run<10>("coalesced store baseline", slot, n, first, cnt, acc, out, live);
        run<11>("scattered store", slot, n, first, cnt, acc, out, live);
        run<12>("gather 4B", slot, n, first, cnt, acc, out, live);
        run<0>("RED.MIN.u32", slot, n, first, cnt, acc, out, live);
        run<1>("ATOM.ADD.u32 (return)", slot, n, first, cnt, acc, out, live);
        run<2>("RED.ADD.f32", slot, n, first, cnt, acc, out, live);
        run<9>("RED.ADD.v2.f32", slot, n, first, cnt, acc, out, live);
        run<3>("RED.ADD.v4.f32", slot, n, first, cnt, acc, out, live);
        run<4>("2 x RED.ADD.v4.f32 (one 32 B row)", slot, n, first, cnt, acc, out, live);
        run<5>("RED.MIN + 2 x RED.ADD.v4.f32", slot, n, first, cnt, acc, out, live);
        run<6>("RED.MIN + RED.ADD.u32", slot, n, first, cnt, acc, out, live);
        run<7>("match_any aggregated RED.MIN + RED.ADD", slot, n, first, cnt, acc, out, live);
        run<8>("8 x RED.ADD.f32", slot, n, first, cnt, acc, out, live);
        run<13>("ATOM.MIN.u64 (return)", slot, n, first, cnt, acc, out, live);
        runp<0>("row halves from lanes l / l+16, 1 instr", slot, n, acc, live);
        runp<1>("row halves from adjacent lanes, 1 instr", slot, n, acc, live);
        runp<2>("control: first halves only, 16 pts / instr", slot, n, acc, live);
        cudaFree(slot); cudaFree(out); cudaFree(first); cudaFree(cnt); cudaFree(acc);
    }
    return 0;
}
