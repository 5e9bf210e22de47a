// pv_common.cuh -- shared device helpers for the polar front end (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "polar_voxel_b200.h"

#define PV_INF 0xFFFFFFFFu
#define PV_DENSE_MAX_CELLS (1u << 20)  // grids up to this many cells/frame use a direct map

// One voxel-map entry.  The whole map is reset with a single 0xFF memset:
//   key = PV_INF (empty), first = PV_INF (ready for atomicMin), cnt = -1 (count - 1), g = PV_INF.
struct __align__(16) PvEntry {
    uint32_t key;    // linear cell index inside the frame (hash mode; unused in dense mode)
    uint32_t first;  // smallest point index that fell into the cell
    uint32_t cnt;    // number of points in the cell, minus one
    uint32_t g;      // batch-global first-occurrence rank of the cell
};

// Workspace carve-up (host computed, passed by value).
struct PvWs {
    // --- zeroed every call ---
    uint32_t *ctrl;           // [0] scan ticket, [1] status bits
    unsigned long long *frame_scan;  // [B+1] packed exclusive scan at each frame start
    int32_t *base;            // [B+1] first output row of each frame
    unsigned long long *tile_state;  // [num_tiles] decoupled look-back state
    // --- set to 0xFF every call ---
    PvEntry *table;           // [B * capf]
    uint32_t *kept;           // [n_cap] per-voxel sorted point lists, CSR by vox_koff
    // --- no init needed ---
    uint32_t *slot;           // [n_cap] map slot of every point (PV_INF = out of range)
    uint32_t *vox_slot;       // [n_cap] map slot of the voxel with global rank g
    uint32_t *vox_koff;       // [n_cap] start of that voxel's list in kept[]
    uint32_t capf;            // map slots per frame (pow2 in hash mode, cells in dense mode)
    uint32_t dense;           // 1 = direct map
    uint32_t num_tiles;
    size_t zero_bytes, ff_bytes, total_bytes;
    char *zero_begin, *ff_begin;
};

struct PvParams {
    float lo[3], vs[3], gridf[3];
    int32_t grid[3];
    int32_t T, V;
    const float *pts;
    const int32_t *offsets;
    int32_t B;
    uint32_t n;
    int32_t c_in, cart, C;
    uint32_t cells;
    PvWs ws;
    int32_t *coors, *num_points, *voxel_counts, *grid_ind, *density;
    float *voxels, *feats, *canvas;
};

// ---------------------------------------------------------------------------------------------
// Scan payload: rank (first points seen) in the high field, kept-point count in the low field.
// 31 bits each so that a 2-bit look-back flag fits in the same 64-bit word.
// ---------------------------------------------------------------------------------------------
#define PV_FIELD 31
#define PV_FIELD_MASK ((1ull << PV_FIELD) - 1)
__device__ __forceinline__ unsigned long long pv_pack(uint32_t rank, uint32_t ksum)
{
    return ((unsigned long long)rank << PV_FIELD) | ksum;
}
__device__ __forceinline__ uint32_t pv_rank(unsigned long long v) { return (uint32_t)((v >> PV_FIELD) & PV_FIELD_MASK); }
__device__ __forceinline__ uint32_t pv_ksum(unsigned long long v) { return (uint32_t)(v & PV_FIELD_MASK); }

__device__ __forceinline__ uint32_t pv_ld_volatile(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned long long pv_ld_volatile64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void pv_st_volatile64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint4 pv_ld_entry(const PvEntry *e)
{
    // entries are written by earlier kernels only: plain 128-bit load through L2
    return __ldcg(reinterpret_cast<const uint4 *>(e));
}

__device__ __forceinline__ uint32_t pv_hash(uint32_t k)
{
    k ^= k >> 16; k *= 0x85ebca6bu; k ^= k >> 13; k *= 0xc2b2ae35u; k ^= k >> 16;
    return k;
}

// Largest b in [0, B) with offsets[b] <= i  (frames may be empty).
__device__ __forceinline__ int pv_frame_of(const int32_t *__restrict__ off, int B, uint32_t i)
{
    int lo = 0, hi = B;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if ((uint32_t)__ldg(off + mid) <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------------
// phi = atan2(y, x) as a fixed sequence of correctly rounded binary32 operations; the CPU
// oracle evaluates the identical sequence, so phi (and every bin derived from it) is
// reproducible bit for bit.  Max error 1.66 ulp; <= 4 ulp from numpy's float32 arctan2
// (det3d/datasets/pipelines/utils.py:41).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float pv_atan2f(float y, float x)
{
    if (x != x || y != y) return __int_as_float(0x7fc00000);
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = ax > ay ? ax : ay;
    const float mn = ax > ay ? ay : ax;
    float a = __fdiv_rn(mn, mx);
    if (mx == 0.0f) a = 0.0f;
    if (mn == __int_as_float(0x7f800000)) a = 1.0f;
    const float s = __fmul_rn(a, a);
    float p = -0x1.d62f3cp-10f;
    p = __fmaf_rn(p, s, 0x1.65a5f8p-7f);
    p = __fmaf_rn(p, s, -0x1.fed102p-6f);
    p = __fmaf_rn(p, s, 0x1.dac9b4p-5f);
    p = __fmaf_rn(p, s, -0x1.583482p-4f);
    p = __fmaf_rn(p, s, 0x1.c099fap-4f);
    p = __fmaf_rn(p, s, -0x1.2421b4p-3f);
    p = __fmaf_rn(p, s, 0x1.9991fep-3f);
    p = __fmaf_rn(p, s, -0x1.55553ep-2f);
    float r = __fmaf_rn(__fmul_rn(a, s), p, a);
    if (ay > ax) r = __fadd_rn(__fsub_rn(0x1.921fb6p+0f, r), -0x1.777a5cp-25f);
    if (__float_as_int(x) < 0) r = __fadd_rn(__fsub_rn(0x1.921fb6p+1f, r), -0x1.777a5cp-24f);
    return copysignf(r, y);
}

// rho = sqrt(x*x + y*y): four separately rounded ops, as numpy evaluates utils.py:40.
__device__ __forceinline__ float pv_rho(float x, float y)
{
    return __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
}

// Channel k of the polar row (rho, phi, z, x, y, feat3..) built from a Cartesian row.
// `row` points at c_in floats.
__device__ __forceinline__ void pv_polar_row(const float *__restrict__ row, int c_in, int cart,
                                             float *__restrict__ out)
{
    if (cart) {
        const float x = row[0], y = row[1];
        out[0] = pv_rho(x, y);
        out[1] = pv_atan2f(y, x);
        out[2] = row[2];
        out[3] = x;
        out[4] = y;
#pragma unroll 4
        for (int k = 3; k < c_in; ++k) out[k + 2] = row[k];
    } else {
#pragma unroll 4
        for (int k = 0; k < c_in; ++k) out[k] = row[k];
    }
}

// Host-side helpers shared by the translation units.
int pv_check_config(const pv_config *cfg);
int pv_make_layout(const pv_config *cfg, int64_t n_cap, int32_t batch, int64_t frame_capacity,
                   void *base, PvWs *out);
int pv_last_cuda_error();
