// tc_gemm.cu -- tcgen05 (5th-gen tensor core) building block for the PFN linear layers (sm_100a).
//
//   D[M, N] = A[M, K] . B[N, K]^T      fp32 in / fp32 out, accumulators in TMEM
//
// fp32 parity on TF32 tensor cores uses the 3xTF32 split: x = hi + lo with hi = x truncated to
// TF32 (the hardware ignores the low 13 mantissa bits), lo = x - hi (exact), and
//   A.B ~= A_hi.B_hi + A_lo.B_hi + A_hi.B_lo      (the lo.lo term is below fp32 resolution),
// three tcgen05.mma.kind::tf32 per 8-wide K step accumulating into the same fp32 TMEM tile.
//
// One CTA (128 threads) per 128-row tile: operands are written to shared memory by the threads
// in the canonical K-major no-swizzle UMMA layout (8 x 16-byte core matrices), one elected thread
// issues the MMAs and commits to an mbarrier, the four warps read their 32 TMEM lanes back with
// tcgen05.ld and store the rows.  This is the linear layer of PFNLayer
// (det3d/models/readers/pillar_encoder.py:41,50) as a plain GEMM; the PFN kernel builds on it.
#include "tc_common.cuh"

// D[M, N] = A[M, K] . B[N, K]^T ; N in {16..256, multiple of 16}, K multiple of 8, K <= 64.
// variant bit 0: swap the LBO / SBO roles in the descriptor (bring-up aid).
__global__ void __launch_bounds__(TC_THREADS) k_tc_gemm(const float *__restrict__ A, const float *__restrict__ B,
                                                        float *__restrict__ D, int M, int N, int K, int variant)
{
    extern __shared__ __align__(128) float smem[];
    float *a_hi = smem;                          // [TC_M x K] canonical
    float *a_lo = a_hi + TC_M * K;
    float *b_hi = a_lo + TC_M * K;               // [N x K] canonical
    float *b_lo = b_hi + N * K;
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row_base = blockIdx.x * TC_M;
    uint32_t ncols = 32;
    while ((int)ncols < N) ncols <<= 1;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc_smem_u32(&s_tmem)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) tc_mbar_init(tc_smem_u32(&s_mbar), 1);

    // operands -> shared memory (split into hi / lo), canonical layout
    for (int e = tid; e < TC_M * K; e += TC_THREADS) {
        const int r = e / K, k = e - r * K;
        const float x = (row_base + r < M) ? A[(size_t)(row_base + r) * K + k] : 0.0f;
        float hi, lo;
        tc_split(x, hi, lo);
        const uint32_t o = tc_canon(r, k, TC_M);
        a_hi[o] = hi; a_lo[o] = lo;
    }
    for (int e = tid; e < N * K; e += TC_THREADS) {
        const int r = e / K, k = e - r * K;
        float hi, lo;
        tc_split(B[(size_t)r * K + k], hi, lo);
        const uint32_t o = tc_canon(r, k, N);
        b_hi[o] = hi; b_lo[o] = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (tid == 0) {
        const uint32_t idesc = tc_idesc_tf32(TC_M, N);
        // strides between core matrices: along K = (rows / 8) * 128 bytes, along M/N = 128 bytes
        const uint32_t a_k = (TC_M / 8) * 128, b_k = (N / 8) * 128, mn = 128;
        for (int ks = 0; ks < K / 8; ++ks) {
            // one MMA consumes K = 8 = two core matrices along K
            const uint32_t a_off = ks * 2 * a_k, b_off = ks * 2 * b_k;
            unsigned long long dah, dal, dbh, dbl;
            if (variant & 1) {
                dah = tc_desc(tc_smem_u32(a_hi) + a_off, mn, a_k); dal = tc_desc(tc_smem_u32(a_lo) + a_off, mn, a_k);
                dbh = tc_desc(tc_smem_u32(b_hi) + b_off, mn, b_k); dbl = tc_desc(tc_smem_u32(b_lo) + b_off, mn, b_k);
            } else {
                dah = tc_desc(tc_smem_u32(a_hi) + a_off, a_k, mn); dal = tc_desc(tc_smem_u32(a_lo) + a_off, a_k, mn);
                dbh = tc_desc(tc_smem_u32(b_hi) + b_off, b_k, mn); dbl = tc_desc(tc_smem_u32(b_lo) + b_off, b_k, mn);
            }
            tc_mma_tf32(tmem, dal, dbh, idesc, ks > 0 ? 1u : 0u);     // small terms first
            tc_mma_tf32(tmem, dah, dbl, idesc, 1u);
            tc_mma_tf32(tmem, dah, dbh, idesc, 1u);
        }
        tc_commit(tc_smem_u32(&s_mbar));
    }
    tc_mbar_wait(tc_smem_u32(&s_mbar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: warp w owns TMEM lanes [32 w, 32 w + 32) = tile rows
    const int row = row_base + warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tc_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        if (row < M) {
#pragma unroll
            for (int k = 0; k < 32; ++k)
                if (c0 + k < N) D[(size_t)row * N + c0 + k] = v[k];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(ncols) : "memory");
}

extern "C" int pv_tc_gemm_tf32x3(const float *a, const float *b, int32_t m, int32_t n, int32_t k, float *d,
                                 int32_t variant, pv_stream_t stream)
{
    if (!a || !b || !d || m <= 0) return PV_ERR_BAD_ARGUMENT;
    if (n < 16 || n > 256 || (n & 15) || k < 8 || k > 64 || (k & 7)) return PV_ERR_UNSUPPORTED;
    const size_t smem = sizeof(float) * 2 * ((size_t)TC_M * k + (size_t)n * k);
    if (cudaFuncSetAttribute(k_tc_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PV_ERR_CUDA;
    k_tc_gemm<<<(m + TC_M - 1) / TC_M, TC_THREADS, smem, (cudaStream_t)stream>>>(a, b, d, m, n, k, variant);
    return pv_last_cuda_error();
}
