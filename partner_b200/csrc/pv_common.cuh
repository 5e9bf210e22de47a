// pv_common.cuh -- shared device helpers for the polar front end (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "polar_voxel_b200.h"

#define PV_INF 0xFFFFFFFFu
#define PV_DENSE_MAX_CELLS (1u << 20)  // grids up to this many cells/frame use a direct map
#define PV_DYN_MAX_CELLS (1u << 26)    // dynamic voxelization: cell-order bitmap up to this many cells/frame

// One voxel-map entry (one per grid cell in dense mode, one per hash slot otherwise).
// CLEAN state = all bits set.  The workspace is self-cleaning: k_cell_flags, the only reader,
// restores every entry it consumes, so no per-call memset is needed (pv_workspace_init once).
struct __align__(8) PvEntry {
    uint32_t first;  // smallest point index that fell into the cell        (K1, atomicMin)
    uint32_t cnt;    // number of points in the cell, minus one             (K1, atomicAdd)
};

// Workspace carve-up (host computed from CAPACITIES, so it is stable across calls).
struct PvWs {
    uint32_t *ctrl;           // [1] status bits
    uint32_t *counts_raw;     // [B] first-occurrence cells per frame (before the V cap)
    int32_t *base;            // [B+1] first output row of each frame
    uint32_t *frame_rank0;    // [B] cells of all earlier frames (global rank at the frame start)
    uint32_t *cum_tiles;      // [B+1] scan tiles of all earlier frames (tiles never straddle frames)
    unsigned long long *tile_agg;    // [max_tiles] per-tile sums       rank << 32 | ksum
    unsigned long long *tile_pre;    // [max_tiles] exclusive prefixes  rank << 32 | ksum
    PvEntry *table;           // [B * capf]                              (clean = all ones)
    uint32_t *keys;           // [B * capf] hash mode: cell index of slot (clean = all ones)
    uint32_t *kept;           // [n_cap] per-voxel ascending point lists (clean = all ones)
    unsigned long long *meta; // [B * capf] per-cell word written by the scan for the cells occupied
                              //   in THIS call (never needs cleaning):
                              //   [63:48] arrival cursor | [47:32] min(count, 65535) | [31:0] kg
    uint32_t *slot;           // [n_cap] map slot of every point (PV_INF = out of range)
    uint32_t *pv;             // [n_cap] per-point scan word: bit 31 = first point of its cell,
                              //         bits 30..0 = points in that cell (first points only)
    uint32_t *pcell;          // [n_cap] hash mode: linear cell index of every point
    uint32_t *vox_cell;       // [B * fcap] linear cell index of the first-occurrence cell (b, r)
    uint32_t *vox_kg;         // [B * fcap] its list offset
    uint32_t *vox_c;          // [B * fcap] its point count
    uint32_t capf;            // map slots per frame (pow2 in hash mode, cells in dense mode)
    uint32_t fcap;            // frame capacity (points)
    uint32_t dense;           // 1 = direct map
    uint32_t max_tiles;
    size_t total_bytes;
};

// Workspace of the list-free path (fused.cu): per-cell accumulator rows instead of point lists.
// Self-cleaning like PvWs: every array marked "clean" is back in that state when a call ends.
struct PvF {
    uint32_t *ctrl;        // [16] [3] scan ticket, [4..5] heavy allocator: cells << 32 | candidates (clean 0)
    int32_t *base;         // [B+1] first output row of each frame
    uint32_t *counts_raw;  // [B] occupied cells per frame before the V cap
    uint32_t *first;       // [B * capf] smallest point index of the cell                      (clean INF)
    uint32_t *keys;        // [B * capf] hash mode: linear cell index of the slot              (clean INF)
    float *acc;            // [B * capf * rowf] per-cell row: C feature sums, then the count   (clean 0)
    uint32_t *sa;          // [n_cap] per point: heavy-bitmap index of its cell (INF = out of range)
    uint32_t *bits;        // [B * wcap] bit (i - frame start) set <=> point i is its cell's first point (clean 0)
    uint2 *wb;             // [B * wcap] static path: {first points before this word in its scan chunk, the word};
                           //            dynamic path: {occupied cells before this word in the frame, the word}
    uint32_t *cagg;        // [max_chunks] static path: first points per scan chunk (PF_CHUNK_WORDS bitmap words)
    uint32_t *cbase;       // [max_chunks + 1] first points of the batch before the chunk
    uint32_t *frank0;      // [B] first points of the batch before the frame's first point
    uint32_t *hbits;       // [B * capf / 32 + 1] bitmap of cells holding more than T points   (clean 0)
    uint4 *hinfo;          // [2 * hmax] heavy cell h: {slot, output row, candidate offset, count}, {cell, frame, cursor, -}
    uint32_t *hlist;       // [n_cap + 32] candidate point indices of the heavy cells, one range per cell
    uint32_t capf;         // slots per frame (dense: cells rounded up to 4; hash: pow2)
    uint32_t wcap;         // bitmap words per frame (multiple of 4)
    uint32_t rowf_cap;     // floats per accumulator row the workspace was sized for
    uint32_t rowf;         // floats per row in THIS call: C + 1 rounded up to 4
    uint32_t dense;
    uint32_t dyn_ok;       // the bitmap covers every cell of a frame: dynamic voxelization available
    uint32_t hmax;
    uint32_t max_chunks;
    size_t total_bytes;
};

struct PvParams {
    float lo[3], vs[3], gridf[3];
    float inv_vs[3];          // fl(1 / vs): fast path of pv_bin
    int32_t grid[3];
    int32_t T, V;
    const float *pts;
    const int32_t *offsets;
    int32_t B;
    uint32_t n;
    int32_t c_in, cart, C;
    uint32_t cells;
    uint32_t num_tiles;
    PvWs ws;
    int32_t *coors, *num_points, *voxel_counts, *grid_ind, *density;
    float *voxels, *feats, *canvas;
    // dynamic voxelization (pv_dynamic_voxelize): coors = unq, num_points = unq_cnt, grid_ind = [N, 4]
    int32_t dyn;
    int32_t any_order;        // list-based pipeline: the consumer does not need the lists sorted (pv_forward_pfn_canvas)
    const int32_t *gi_in;     // caller-provided (b, z, y, x) per point, or NULL = bin the points
    int32_t *unq_inv;         // [N] voxel row of every point, or NULL
};

__device__ __forceinline__ uint32_t pv_ld_volatile(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// L2 residency control (B200: 126 MB L2).  Point rows are read twice per call -- streamed by
// k_bin_insert, gathered again by k_emit -- so their first read asks L2 to keep them
// (evict_last), while write-once outputs are stored evict_first.
__device__ __forceinline__ unsigned long long pv_policy_evict_last()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 pv_ld_keep(const float4 *p, unsigned long long pol)
{
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float pv_ld_keep(const float *p, unsigned long long pol)
{
    float v;
    asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}

__device__ __forceinline__ uint32_t pv_hash(uint32_t k)
{
    k ^= k >> 16; k *= 0x85ebca6bu; k ^= k >> 13; k *= 0xc2b2ae35u; k ^= k >> 16;
    return k;
}

// Home slot of cell (cz, cy, cx) in a hash map of mask + 1 slots (power of two >= 1024).  LOCALITY
// PRESERVING: the 16 azimuth-consecutive cells of one (z, range) share a row of 16 consecutive
// slots -- the group is hashed, the low 4 azimuth bits are kept -- so the keys, first-point words,
// map entries and accumulator rows that consecutive points of a LiDAR ring touch share 128-byte
// lines, exactly as in the direct (phi-fastest) map.  Probing moves a whole row at a time
// (PV_PROBE_STEP), which keeps the in-row position; pv_probe_next walks all rows of that position
// before it moves to the next position, so every slot of the map is reachable.
#ifdef PV_HASH_FLAT
#define PV_PROBE_STEP 1u
__device__ __forceinline__ uint32_t pv_slot_home(uint32_t cx, uint32_t cy, uint32_t cz, uint32_t nx, uint32_t ny, uint32_t mask)
{
    return pv_hash((cz * ny + cy) * nx + cx) & mask;
}
#else
#define PV_PROBE_STEP 16u
__device__ __forceinline__ uint32_t pv_slot_home(uint32_t cx, uint32_t cy, uint32_t cz, uint32_t nx, uint32_t ny, uint32_t mask)
{
    if (ny < 16u) return pv_hash((cz * ny + cy) * nx + cx) & mask;      // no azimuth rows to keep together
    const uint32_t group = (cz * nx + cx) * ((ny + 15u) >> 4) + (cy >> 4);
    return ((pv_hash(group) << 4) | (cy & 15u)) & mask;
}
#endif

__device__ __forceinline__ uint32_t pv_probe_next(uint32_t h, uint32_t probe, uint32_t mask)
{
    h = (h + PV_PROBE_STEP) & mask;
    if (PV_PROBE_STEP > 1u && ((probe + 1u) & (mask / PV_PROBE_STEP)) == 0u) h = (h + 1u) & mask;
    return h;
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

// floor((q - lo) / vs) exactly as the reference evaluates it (point_cloud_ops.py:45: float32
// subtract, IEEE divide, floor) without paying for the division: r = t * fl(1/vs) is within
// 2^-23 relative of the exact quotient, so unless r sits within 1e-6 relative of an integer its
// floor equals the floor of the correctly rounded quotient; the rare near-integer (and NaN / inf /
// huge) cases take the IEEE division.
__device__ __forceinline__ float pv_bin(float q, float lo, float vs, float inv)
{
    const float t = __fsub_rn(q, lo);
    const float r = __fmul_rn(t, inv);
    const float fr = floorf(r);
    const float d = __fsub_rn(r, fr);
    const float thr = __fmul_rn(fmaxf(fabsf(r), 1.0f), 1e-6f);
    if (!(d > thr && d < __fsub_rn(1.0f, thr))) return floorf(__fdiv_rn(t, vs));
    return fr;
}

// The near-integer test of pv_bin on its own (r = t * inv, fr = floorf(r)): true = take the division.
__device__ __forceinline__ bool pv_bin_unsure(float r, float fr)
{
    const float d = __fsub_rn(r, fr);
    const float thr = __fmul_rn(fmaxf(fabsf(r), 1.0f), 1e-6f);
    return !(d > thr && d < __fsub_rn(1.0f, thr));
}

// sum / n for n = 1, 2, 3, ... given inv = 1 / n: one Newton correction of the reciprocal product,
// i.e. the fast path of the IEEE division (correctly rounded except for a vanishing fraction of
// operands, where it is 1 ulp off -- far inside the 1e-5 gate); infinities pass through.
__device__ __forceinline__ float pv_div_count(float s, float n, float inv)
{
    const float q = __fmul_rn(s, inv);
    const float r = __fmaf_rn(-q, n, s);
    const float q1 = __fmaf_rn(r, inv, q);
    return fabsf(q) <= 3.0e38f ? q1 : q;
}

// rho = sqrt(x*x + y*y): four separately rounded ops, as numpy evaluates utils.py:40.
__device__ __forceinline__ float pv_rho(float x, float y)
{
    return __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
}

// Point i's raw row, zero padded.  A gathered load costs one load/store slot per lane whatever its
// width, so rows with an even number of floats (8-byte aligned) are fetched as float2.
template <int N>
__device__ __forceinline__ void pv_load_row(const float *__restrict__ pts, uint32_t i, int c_in, float (&in)[N])
{
    const float *row = pts + (size_t)i * c_in;
    if ((c_in & 1) == 0 && (reinterpret_cast<uintptr_t>(pts) & 7u) == 0) {
#pragma unroll
        for (int k = 0; k < N; k += 2) {
            const float2 v = (k < c_in) ? __ldg(reinterpret_cast<const float2 *>(row + k)) : make_float2(0.0f, 0.0f);
            in[k] = v.x;
            if (k + 1 < N) in[k + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) in[k] = (k < c_in) ? __ldg(row + k) : 0.0f;
    }
}

// Loads point i's row and expands it to the C-channel feature row the reference voxelizes:
// Cartesian input -> (rho, phi, z, x, y, feat3..) (utils.py:42-44); polar input -> as is.
__device__ __forceinline__ void pv_feature_row(const float *__restrict__ pts, uint32_t i, int c_in,
                                               int cart, float (&out)[PV_MAX_CHANNELS])
{
    float in[PV_MAX_CHANNELS];
    pv_load_row(pts, i, c_in, in);
    if (cart) {
        out[0] = pv_rho(in[0], in[1]);
        out[1] = pv_atan2f(in[1], in[0]);
        out[2] = in[2]; out[3] = in[0]; out[4] = in[1];
#pragma unroll
        for (int k = 5; k < PV_MAX_CHANNELS; ++k) out[k] = in[k - 2];
    } else {
#pragma unroll
        for (int k = 0; k < PV_MAX_CHANNELS; ++k) out[k] = in[k];
    }
}

// One feature row [C] at row vid of a packed [M, C] array.  A scattered store costs one load/store
// unit slot per lane whatever its width, so 28-byte rows (C = 7) go out as one 16-, one 8- and one
// 4-byte store chosen by the row's alignment, 32-byte rows (C = 8) as two 16-byte stores.
template <int CT>
__device__ __forceinline__ void pv_store_feats(float *feats, int32_t vid, int C, const float (&mean)[CT])
{
    float *o = feats + (size_t)vid * C;
    const bool al16 = (reinterpret_cast<uintptr_t>(feats) & 15u) == 0;
    if (CT >= 7 && C == 7 && al16) {
        float2 *o2; float4 *o4;
        switch (vid & 3) {
        case 0:
            o4 = reinterpret_cast<float4 *>(o); o2 = reinterpret_cast<float2 *>(o + 4);
            __stcs(o4, make_float4(mean[0], mean[1], mean[2], mean[3])); __stcs(o2, make_float2(mean[4 % CT], mean[5 % CT])); __stcs(o + 6, mean[6 % CT]);
            break;
        case 1:     // row starts 12 bytes past a 16-byte boundary
            o2 = reinterpret_cast<float2 *>(o + 5); o4 = reinterpret_cast<float4 *>(o + 1);
            __stcs(o + 0, mean[0]); __stcs(o4, make_float4(mean[1], mean[2], mean[3], mean[4 % CT])); __stcs(o2, make_float2(mean[5 % CT], mean[6 % CT]));
            break;
        case 2:     // 8 bytes past
            o2 = reinterpret_cast<float2 *>(o); o4 = reinterpret_cast<float4 *>(o + 2);
            __stcs(o2, make_float2(mean[0], mean[1])); __stcs(o4, make_float4(mean[2], mean[3], mean[4 % CT], mean[5 % CT])); __stcs(o + 6, mean[6 % CT]);
            break;
        default:    // 4 bytes past
            o2 = reinterpret_cast<float2 *>(o + 1); o4 = reinterpret_cast<float4 *>(o + 3);
            __stcs(o + 0, mean[0]); __stcs(o2, make_float2(mean[1], mean[2])); __stcs(o4, make_float4(mean[3], mean[4 % CT], mean[5 % CT], mean[6 % CT]));
            break;
        }
    } else if (CT >= 8 && C == 8 && al16) {
        __stcs(reinterpret_cast<float4 *>(o), make_float4(mean[0], mean[1], mean[2], mean[3]));
        __stcs(reinterpret_cast<float4 *>(o) + 1, make_float4(mean[4 % CT], mean[5 % CT], mean[6 % CT], mean[7 % CT]));
    } else {
#pragma unroll
        for (int k = 0; k < CT; ++k)
            if (k < C) __stcs(o + k, mean[k]);
    }
}


// Host-side helpers shared by the translation units.
int pv_check_config(const pv_config *cfg);
int pv_make_layout(const pv_config *cfg, int64_t n_cap, int32_t batch, int64_t frame_capacity,
                   void *base, PvWs *out);
int pv_last_cuda_error();
int pv_sm_count();      // SMs of the current device
int pvf_make_layout(const pv_config *cfg, int64_t n_cap, int32_t batch, int64_t frame_capacity,
                    int32_t max_channels, void *base, PvF *out);
int pvf_init(const PvF &f, int32_t batch, int64_t n_cap, cudaStream_t st);
// list-free front end; ev (optional) = PV_PROFILE_STAGES + 1 events recorded at the stage boundaries
int pvf_run(PvParams &p, PvF &f, cudaStream_t st, cudaEvent_t *ev);
int pvf_run_dynamic(PvParams &p, PvF &f, cudaStream_t st);
int pvf_run_grid_ind(PvParams &p, cudaStream_t st);           // binning only: p.grid_ind [n, 4]
int pvf_insert_lists(PvParams &p, PvF &f, cudaStream_t st);     // binning front end of the list-based pipeline
