// tc_common.cuh -- tcgen05 / TMEM / mbarrier helpers shared by the tensor-core kernels (sm_100a).
#pragma once
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

