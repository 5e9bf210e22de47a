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
#include "pv_common.cuh"

#define TC_M 128
#define TC_THREADS 128

__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 64-bit shared-memory matrix descriptor (sm_100 format, version 1), no swizzle.
// lbo / sbo in bytes: strides between 8x16-byte core matrices along K and along M/N.
__device__ __forceinline__ unsigned long long tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3FFFu);
    d |= (unsigned long long)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (unsigned long long)((sbo >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;                 // descriptor version (sm_100)
    return d;                        // base offset 0, lbo mode 0, layout type 0 = no swizzle
}

// 32-bit instruction descriptor, kind::tf32: D = F32, A = B = TF32, both K-major.
__device__ __forceinline__ uint32_t tc_idesc_tf32(uint32_t m, uint32_t n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, unsigned long long a_desc, unsigned long long b_desc,
                                            uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tc_commit(uint32_t mbar_saddr)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar_saddr) : "memory");
}

__device__ __forceinline__ void tc_mbar_init(uint32_t mbar_saddr, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar_saddr), "r"(count) : "memory");
}

__device__ __forceinline__ void tc_mbar_wait(uint32_t mbar_saddr, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n"
        :: "r"(mbar_saddr), "r"(parity) : "memory");
}

// 32 lanes x 32 columns of fp32 accumulators: thread l of the warp gets row (lane base + l).
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
}

// Canonical K-major no-swizzle placement of element (row, k) of a [rows x K] fp32 operand:
// core matrix = 8 rows x 4 elements (16 bytes per row, 128 bytes), core matrices ordered
// K-group-major: offset(floats) = ((k / 4) * (rows / 8) + row / 8) * 32 + (row % 8) * 4 + (k % 4).
__device__ __forceinline__ uint32_t tc_canon(uint32_t row, uint32_t k, uint32_t rows)
{
    return ((k >> 2) * (rows >> 3) + (row >> 3)) * 32u + (row & 7u) * 4u + (k & 3u);
}

__device__ __forceinline__ void tc_split(float x, float &hi, float &lo)
{
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);   // what the tensor core will read
    lo = __fsub_rn(x, hi);
}

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

// ---------------------------------------------------------------------------------------------
// Second PFN layer on the tensor cores.  PFNLayer.forward_static (pillar_encoder.py:49-61) for the
// last layer of a two-layer PillarFeatureNet:
//     x = Linear([x0 | repeat(x_max0)])      K = 2 * u0 (64), N = units (128)     -> tcgen05, 3xTF32
//     y = ReLU(BatchNorm_eval(x)) ;  out[voxel] = max over the voxel's rows (padded slots included:
//     the layer-0 stage exported one representative padded row per non-full voxel).
// Persistent CTAs of 128 threads walk 128-row tiles.  Per tile: the A operand [x0 | x_max0[voxel]]
// is assembled straight into the canonical UMMA layout (hi / lo split), one thread issues the
// 3 * K/8 MMAs into a 128 x N fp32 TMEM tile, each warp pulls its 32 TMEM lanes, applies BN + ReLU
// and parks the tile transposed in shared memory, then thread c reduces column c over the runs of
// equal voxel id (rows are grouped by voxel): interior runs are plain stores, the first and last
// run of a tile may continue in a neighbouring tile and use atomicMax (values are >= 0, so the
// integer order is the float order; `out` is zeroed beforehand).
// ---------------------------------------------------------------------------------------------
struct PtcArgs {
    const float *x0; const uint32_t *row_vox; const float *vmax0; const uint32_t *total_rows;
    const float *w, *mean, *var, *gamma, *beta;
    float eps;
    int u0, n;
    float *out;
};

#define PTC_STAGE_LD 129     // transposed staging tile: [column][row], odd stride = conflict-free both ways

__global__ void __launch_bounds__(TC_THREADS, 1) k_pfn_tc_layer(const __grid_constant__ PtcArgs a)
{
    extern __shared__ __align__(128) float smem[];
    const int K = 2 * a.u0, N = a.n;
    float *a_hi = smem;                          // [128 x K] canonical
    float *a_lo = a_hi + TC_M * K;
    float *b_hi = a_lo + TC_M * K;               // [N x K] canonical
    float *b_lo = b_hi + N * K;
    float *s_stage = b_lo + N * K;               // [N][PTC_STAGE_LD]
    float *s_bn = s_stage + N * PTC_STAGE_LD;    // mean, invstd, gamma, beta: 4 x N
    uint32_t *s_vox = reinterpret_cast<uint32_t *>(s_bn + 4 * N);   // [128]
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t ncols = 32;
    while ((int)ncols < N) ncols <<= 1;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc_smem_u32(&s_tmem)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) tc_mbar_init(tc_smem_u32(&s_mbar), 1);
    // weights (linear.weight is [N, K], K-major already) and BatchNorm constants, once per CTA
    for (int e = tid; e < N * K; e += TC_THREADS) {
        const int r = e / K, k = e - r * K;
        float hi, lo;
        tc_split(__ldg(a.w + e), hi, lo);
        const uint32_t o = tc_canon(r, k, N);
        b_hi[o] = hi; b_lo[o] = lo;
    }
    for (int o = tid; o < N; o += TC_THREADS) {
        s_bn[o] = a.mean[o];
        s_bn[N + o] = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(a.var[o], a.eps)));
        s_bn[2 * N + o] = a.gamma[o];
        s_bn[3 * N + o] = a.beta[o];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t total = *a.total_rows;
    const uint32_t idesc = tc_idesc_tf32(TC_M, N);
    const uint32_t a_k = (TC_M / 8) * 128, b_k = (N / 8) * 128, mn = 128;
    uint32_t phase = 0;

    for (uint32_t row_base = blockIdx.x * TC_M; row_base < total; row_base += gridDim.x * TC_M) {
        const uint32_t rows_here = min((uint32_t)TC_M, total - row_base);
        // ---- A tile: [x0 row | x_max0 of the row's voxel], split, canonical layout ----
        if (tid < TC_M) s_vox[tid] = (uint32_t)tid < rows_here ? __ldg(a.row_vox + row_base + tid) : 0xFFFFFFFFu;
        __syncthreads();
        // float4 granularity: u0 / 4 quads per row half; all loads of a pass are issued before use
        {
            const int qpr = a.u0 >> 2;                        // quads per row half (8 for u0 = 32)
            const int nq = TC_M * qpr;                        // quads per half tile
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                for (int q0 = tid; q0 < nq; q0 += TC_THREADS * 8) {
                    float4 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int q = q0 + u * TC_THREADS;
                        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (q < nq) {
                            const int r = q / qpr, kq = q - r * qpr;
                            if ((uint32_t)r < rows_here)
                                v[u] = half == 0
                                    ? __ldcs(reinterpret_cast<const float4 *>(a.x0 + (size_t)(row_base + r) * a.u0) + kq)
                                    : __ldg(reinterpret_cast<const float4 *>(a.vmax0 + (size_t)s_vox[r] * a.u0) + kq);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int q = q0 + u * TC_THREADS;
                        if (q < nq) {
                            const int r = q / qpr, kq = q - r * qpr;
                            // 4 consecutive k of one row = one 16-byte row of a core matrix
                            const uint32_t o = tc_canon(r, half * a.u0 + 4 * kq, TC_M);
                            float4 hi, lo;
                            tc_split(v[u].x, hi.x, lo.x); tc_split(v[u].y, hi.y, lo.y);
                            tc_split(v[u].z, hi.z, lo.z); tc_split(v[u].w, hi.w, lo.w);
                            *reinterpret_cast<float4 *>(a_hi + o) = hi;
                            *reinterpret_cast<float4 *>(a_lo + o) = lo;
                        }
                    }
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
            for (int ks = 0; ks < K / 8; ++ks) {
                const uint32_t a_off = ks * 2 * a_k, b_off = ks * 2 * b_k;
                const unsigned long long dah = tc_desc(tc_smem_u32(a_hi) + a_off, a_k, mn), dal = tc_desc(tc_smem_u32(a_lo) + a_off, a_k, mn);
                const unsigned long long dbh = tc_desc(tc_smem_u32(b_hi) + b_off, b_k, mn), dbl = tc_desc(tc_smem_u32(b_lo) + b_off, b_k, mn);
                tc_mma_tf32(tmem, dal, dbh, idesc, ks > 0 ? 1u : 0u);
                tc_mma_tf32(tmem, dah, dbl, idesc, 1u);
                tc_mma_tf32(tmem, dah, dbh, idesc, 1u);
            }
            tc_commit(tc_smem_u32(&s_mbar));
        }
        tc_mbar_wait(tc_smem_u32(&s_mbar), phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue 1: TMEM -> BN (ATen eval order) + ReLU -> transposed staging tile ----
        const int r_local = warp * 32 + lane;
        for (int c0 = 0; c0 < N; c0 += 32) {
            float v[32];
            tc_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const int o = c0 + k;
                const float y = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(v[k], s_bn[o]), s_bn[N + o]), s_bn[2 * N + o]), s_bn[3 * N + o]);
                s_stage[o * PTC_STAGE_LD + r_local] = fmaxf(y, 0.0f);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        // ---- epilogue 2: per-voxel max down each column ----
        for (int c = tid; c < N; c += TC_THREADS) {
            const float *col = s_stage + c * PTC_STAGE_LD;
            uint32_t cur = s_vox[0];
            float run = 0.0f;
            bool first_run = true;
            for (uint32_t r = 0; r < rows_here; ++r) {
                const uint32_t v = s_vox[r];
                if (v != cur) {
                    int *dst = reinterpret_cast<int *>(a.out) + (size_t)cur * N + c;
                    if (first_run) atomicMax(dst, __float_as_int(run)); else *dst = __float_as_int(run);
                    first_run = false;
                    cur = v; run = 0.0f;
                }
                run = fmaxf(run, col[r]);
            }
            atomicMax(reinterpret_cast<int *>(a.out) + (size_t)cur * N + c, __float_as_int(run));
        }
        __syncthreads();      // staging tile, s_vox and the A operand are reused by the next tile
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(ncols) : "memory");
}

int pv_pfn_tc_layer(const float *x0, const uint32_t *row_vox, const float *vmax0, const uint32_t *total_rows,
                    int u0, int n, const float *w, const float *bn_mean, const float *bn_var, const float *bn_gamma,
                    const float *bn_beta, float eps, long long m, float *out, cudaStream_t st)
{
    const int K = 2 * u0;
    if (K < 8 || K > 64 || (K & 7) || n < 16 || n > 128 || (n & 15)) return PV_ERR_UNSUPPORTED;
    PtcArgs a;
    a.x0 = x0; a.row_vox = row_vox; a.vmax0 = vmax0; a.total_rows = total_rows;
    a.w = w; a.mean = bn_mean; a.var = bn_var; a.gamma = bn_gamma; a.beta = bn_beta; a.eps = eps;
    a.u0 = u0; a.n = n; a.out = out;
    const size_t smem = sizeof(float) * (2 * ((size_t)TC_M * K + (size_t)n * K) + (size_t)n * PTC_STAGE_LD + 4 * (size_t)n) +
                        sizeof(uint32_t) * TC_M;
    if (cudaMemsetAsync(out, 0, (size_t)m * n * sizeof(float), st) != cudaSuccess) return PV_ERR_CUDA;
    if (cudaFuncSetAttribute(k_pfn_tc_layer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PV_ERR_CUDA;
    k_pfn_tc_layer<<<148, TC_THREADS, smem, st>>>(a);
    return pv_last_cuda_error();
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
