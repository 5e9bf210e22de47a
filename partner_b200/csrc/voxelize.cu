// voxelize.cu -- batched hard voxelization on the polar grid (sm_100a).
//
// Replaces the sequential numba loop of det3d/ops/point_cloud/point_cloud_ops.py:7-72
// (called through VoxelGenerator.generate, core/input/voxel_generator.py:19-32) with an
// order-free formulation that reproduces it exactly:
//
//   K1 bin_insert  point -> (rho, phi, z) bin -> cell; per cell atomicMin(first point index) and
//                  atomicAdd(count) in a per-frame map (direct map for small grids, hash else);
//                  the binning kernel is shared with the list-free pipeline (fused.cu).
//   K2 cell_flags  streams over the MAP (coalesced): every occupied cell marks its first point
//                  pv[first] = FIRST | count, and the entry is restored to its clean state.
//   K3 scan        frame-segmented scan over pv[], in point order, of (is_first, min(count, T)):
//                  tile_reduce + scan_apply.  A first point's exclusive prefix is its cell's
//                  first-occurrence rank r in the frame and its list offset kg; it stores itself
//                  as list element 0 and publishes (kg, count) for the cell.
//   K4 place       every other point of a kept voxel inserts its index into the voxel's list,
//                  which converges to the T smallest indices in ascending order (atomicMin chain).
//   K5 emit        one thread per voxel: coors (b, z, y, x), num_points, mean feature, optional
//                  padded voxels tensor / density, scatter into the BEV canvas; restores the lists.
//
// Voxel order = first-occurrence order, kept points = the T smallest indices in index order,
// voxels with rank >= V dropped -- the three order-dependent behaviours of the reference loop.
#include <stdlib.h>

#include <algorithm>

#include "pv_common.cuh"
#include "pfn_fused.cuh"

#define K1_THREADS 256
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)
#define PV_FIRST 0x80000000u

// Direct-map slot of cell (cz, cy, cx): phi (y) is the fastest index, so consecutive azimuth
// samples of one LiDAR ring touch adjacent entries.
__device__ __forceinline__ uint32_t pv_dense_slot(const PvParams &p, int cz, int cy, int cx)
{
    return ((uint32_t)cz * (uint32_t)p.grid[0] + (uint32_t)cx) * (uint32_t)p.grid[1] + (uint32_t)cy;
}
// ... and back to the linear cell index (z * ny + y) * nx + x used by coors / density / canvas.
__device__ __forceinline__ uint32_t pv_dense_cell(const PvParams &p, uint32_t local)
{
    const uint32_t ny = p.grid[1], nx = p.grid[0];
    const uint32_t cy = local % ny, t = local / ny;
    const uint32_t cx = t % nx, cz = t / nx;
    return (cz * ny + cy) * nx + cx;
}

// ---------------------------------------------------------------------------------------------
// K1 -- bin + insert: kf_insert<..., PF_MODE_LISTS> of fused.cu (TMA-staged tiles, 4 points per
// thread, reciprocal binning, runs of equal cells merged): per run atomicMin(first) +
// atomicAdd(count) on the {first, cnt} map, per point slot[] / pv[] (/ pcell[]).  This kernel only
// lays out the scan tiles (tiles never straddle frames).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_scan_layout(const __grid_constant__ PvParams p)
{
    if (threadIdx.x != 0) return;
    uint32_t acc = 0;
    for (int b = 0; b < p.B; ++b) {
        p.ws.cum_tiles[b] = acc;
        acc += ((uint32_t)(p.offsets[b + 1] - p.offsets[b]) + SCAN_TILE - 1) / SCAN_TILE;
    }
    p.ws.cum_tiles[p.B] = acc;
}

// ---------------------------------------------------------------------------------------------
// K1 (hash maps) -- bin + insert, one point per thread: the slot claim is a dependent L2 round
// trip per probe, so hash maps want as many independent threads in flight as possible (the
// 4-points-per-thread kernel of fused.cu measured 182 us vs 121 us here on the Waymo batch; round 2:
// a warp-autonomous TMA-ring version in the shape of kf_insert_lanes -- lane = consecutive point, runs
// found with one ballot, the first probe of four groups issued before any answer is read, half the
// instructions per point -- measured 222 vs 117 us on config 4 and 188 vs 123 us on config 5: with a
// quarter of the threads per SM the dependent CAS round trips are no longer covered)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pv_claim(uint32_t *keys, uint32_t mask, uint32_t cell, uint32_t h,
                                             uint32_t *status)
{
    for (uint32_t probe = 0; probe <= mask; ++probe) {
        // CAS first: one L2 round trip whether the slot is free, already ours, or taken
        const uint32_t old = atomicCAS(keys + h, PV_INF, cell);
        if (old == PV_INF || old == cell) return h;
        h = pv_probe_next(h, probe, mask);
    }
    atomicOr(status, 1u);
    return PV_INF;
}

template <bool DENSE>
__global__ void __launch_bounds__(K1_THREADS) k_bin_insert(const __grid_constant__ PvParams p)
{
    __shared__ __align__(16) float s_pts[K1_THREADS * PV_MAX_CHANNELS];
    __shared__ int s_b0;
    const uint32_t tile_base = blockIdx.x * K1_THREADS;
    const uint32_t tid = threadIdx.x;
    const uint32_t n_tile = tile_base < p.n ? min((uint32_t)K1_THREADS, p.n - tile_base) : 0u;
    const int c_in = p.c_in;
    if (tid == 0) s_b0 = pv_frame_of(p.offsets, p.B, tile_base);

    // ---- stage the tile's rows: coalesced 128-bit loads of the contiguous float range ----
    {
        const size_t f0 = (size_t)tile_base * c_in;
        const uint32_t nf = n_tile * c_in;
        const float *src = p.pts + f0;
        const unsigned long long keep = pv_policy_evict_last();   // k_emit gathers these rows again
        if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
            const uint32_t nv = nf >> 2;
            const float4 *src4 = reinterpret_cast<const float4 *>(src);
            float4 *dst4 = reinterpret_cast<float4 *>(s_pts);
            for (uint32_t k = tid; k < nv; k += K1_THREADS) dst4[k] = pv_ld_keep(src4 + k, keep);
            for (uint32_t k = (nv << 2) + tid; k < nf; k += K1_THREADS) s_pts[k] = pv_ld_keep(src + k, keep);
        } else {
            for (uint32_t k = tid; k < nf; k += K1_THREADS) s_pts[k] = pv_ld_keep(src + k, keep);
        }
    }
    __syncthreads();

    const uint32_t i = tile_base + tid;
    const bool live = tid < n_tile;
    bool ok = live;
    uint32_t cell = 0, local = 0, home = 0;
    int b = s_b0;
    if (live) {
        const float *row = s_pts + tid * c_in;
        float q[3];
        if (p.cart) {
            q[0] = pv_rho(row[0], row[1]);
            q[1] = pv_atan2f(row[1], row[0]);
            q[2] = row[2];
        } else {
            q[0] = row[0]; q[1] = row[1]; q[2] = row[2];
        }
        int ci[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            // point_cloud_ops.py:45 -- float32 subtract, IEEE divide, floor
            const float cf = pv_bin(q[j], p.lo[j], p.vs[j], p.inv_vs[j]);
            int c;
            if (cf != cf) { ok = false; c = 0; }
            else if (cf < 0.0f) { ok = false; c = 0; }
            else if (cf >= p.gridf[j]) { ok = false; c = p.grid[j] - 1; }
            else c = (int)cf;
            ci[j] = c;
        }
        if (p.grid_ind) {  // :46-54 clamped (z, y, x) for every point
            int32_t *gi = p.grid_ind + (size_t)i * 3;
            gi[0] = ci[2]; gi[1] = ci[1]; gi[2] = ci[0];
        }
        cell = ((uint32_t)ci[2] * (uint32_t)p.grid[1] + (uint32_t)ci[1]) * (uint32_t)p.grid[0] + (uint32_t)ci[0];
        if (DENSE) local = pv_dense_slot(p, ci[2], ci[1], ci[0]);
        else home = pv_slot_home((uint32_t)ci[0], (uint32_t)ci[1], (uint32_t)ci[2], (uint32_t)p.grid[0], (uint32_t)p.grid[1], p.ws.capf - 1);
        while (b + 1 < p.B && i >= (uint32_t)__ldg(p.offsets + b + 1)) ++b;   // tile may straddle frames
    }

    // ---- warp-aggregated insert: one lane per distinct (frame, cell) issues the atomics ----
    const unsigned lane = tid & 31u;
    unsigned peers;
    if (DENSE) {
        const uint32_t key = ok ? (uint32_t)b * p.ws.capf + local : (PV_INF - lane);
        peers = __match_any_sync(0xffffffffu, key);
    } else {
        const int b0 = __shfl_sync(0xffffffffu, b, 0);
        if (__all_sync(0xffffffffu, !live || b == b0)) {
            peers = __match_any_sync(0xffffffffu, ok ? cell : (PV_INF - lane));
        } else {
            const unsigned long long key = ok ? (((unsigned long long)b << 32) | cell)
                                              : (0xFFFFFFFF00000000ull | lane);
            peers = __match_any_sync(0xffffffffu, key);
        }
    }
    const int leader = __ffs(peers) - 1;
    uint32_t s = PV_INF;
    if (ok && (int)lane == leader) {
        if (!DENSE) local = pv_claim(p.ws.keys + (size_t)b * p.ws.capf, p.ws.capf - 1, cell, home, p.ws.ctrl + 1);
        if (local != PV_INF) {
            s = (uint32_t)b * p.ws.capf + local;
            atomicMin(&p.ws.table[s].first, i);          // lanes are in index order: leader is the min
            atomicAdd(&p.ws.table[s].cnt, (uint32_t)__popc(peers));
        }
    }
    s = __shfl_sync(0xffffffffu, s, leader);
    if (live) {
        p.ws.slot[i] = ok ? s : PV_INF;
        p.ws.pv[i] = 0u;
        if (!DENSE) p.ws.pcell[i] = cell;
    }
}

// ---------------------------------------------------------------------------------------------
// K2 -- stream over the map: mark first points, restore the entries (and hash keys)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cell_flags(const __grid_constant__ PvParams p)
{
    const size_t total = (size_t)p.B * p.ws.capf;
    const size_t e0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;   // two 8-byte entries per thread
    if (e0 >= total) return;
    uint4 *ptr = reinterpret_cast<uint4 *>(p.ws.table + e0);
    if (e0 + 1 < total) {
        const uint4 v = __ldcs(ptr);
        if (v.x == PV_INF && v.z == PV_INF) return;
        if (v.x != PV_INF) p.ws.pv[v.x] = PV_FIRST | min(v.y + 1u, 0x7FFFFFFFu);
        if (v.z != PV_INF) p.ws.pv[v.z] = PV_FIRST | min(v.w + 1u, 0x7FFFFFFFu);
        *ptr = make_uint4(PV_INF, PV_INF, PV_INF, PV_INF);
        if (!p.ws.dense) {
            if (v.x != PV_INF) p.ws.keys[e0] = PV_INF;
            if (v.z != PV_INF) p.ws.keys[e0 + 1] = PV_INF;
        }
    } else {
        PvEntry *e = p.ws.table + e0;
        const uint32_t f = e->first;
        if (f == PV_INF) return;
        p.ws.pv[f] = PV_FIRST | min(e->cnt + 1u, 0x7FFFFFFFu);
        e->first = PV_INF; e->cnt = PV_INF;
        if (!p.ws.dense) p.ws.keys[e0] = PV_INF;
    }
}

// ---------------------------------------------------------------------------------------------
// K3 -- scan over pv[], in point order, of (is_first, min(count, T)).  Tiles never straddle a
// frame (tile t of frame b covers TILE consecutive points of that frame), so first-occurrence
// ranks restart per frame without any segmented arithmetic:
//   tile_reduce : per-tile sums                          -> tile_agg[t] = rank << 32 | ksum
//   scan_mid    : one block; exclusive prefixes (rank per frame, ksum global) -> tile_pre[t],
//                 per-frame voxel counts (capped at V) and output row bases
//   scan_apply  : per-tile scan; a first point's prefix is its rank r and list offset kg.
// ---------------------------------------------------------------------------------------------
struct TileRange { int b; uint32_t lo, hi; };   // frame, [lo, hi) point range; b < 0 = no such tile

__device__ __forceinline__ TileRange pv_tile_range(const PvParams &p, uint32_t tile)
{
    // cum_tiles[b] = tiles of frames < b (written by k_bin_insert's block 0)
    const uint32_t *cum = p.ws.cum_tiles;
    TileRange t;
    if (tile >= cum[p.B]) { t.b = -1; t.lo = t.hi = 0; return t; }
    int lo = 0, hi = p.B;            // cum[lo] <= tile < cum[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (cum[mid] <= tile) lo = mid; else hi = mid;
    }
    t.b = lo;
    t.lo = (uint32_t)p.offsets[lo] + (tile - cum[lo]) * SCAN_TILE;
    t.hi = min(t.lo + SCAN_TILE, (uint32_t)p.offsets[lo + 1]);
    return t;
}

// Loads the tile's pv words and returns this thread's packed (rank << 32 | ksum) items.
__device__ __forceinline__ unsigned long long pv_load_items(const PvParams &p, const TileRange &t,
                                                            uint32_t (&w)[SCAN_ITEMS],
                                                            unsigned long long (&val)[SCAN_ITEMS])
{
    const uint32_t i0 = t.lo + threadIdx.x * SCAN_ITEMS;
    if (i0 + SCAN_ITEMS <= t.hi && ((i0 & 3u) == 0)) {
        const uint4 v0 = __ldcg(reinterpret_cast<const uint4 *>(p.ws.pv + i0));
        const uint4 v1 = __ldcg(reinterpret_cast<const uint4 *>(p.ws.pv + i0 + 4));
        w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w;
        w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
    } else {
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) w[j] = (i0 + j < t.hi) ? __ldcg(p.ws.pv + i0 + j) : 0u;
    }
    unsigned long long sum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        val[j] = (w[j] & PV_FIRST) ? ((1ull << 32) | min(w[j] & 0x7FFFFFFFu, (uint32_t)p.T)) : 0ull;
        sum += val[j];
    }
    return sum;
}

// Exclusive prefixes over the tile sums (rank and ksum), per-frame voxel counts (capped at V) and
// output row bases.  Runs in the LAST tile_reduce block to finish (no extra launch).
__device__ void pv_scan_mid(const PvParams &p, unsigned long long *s_warp /*[SCAN_THREADS/32 + 1]*/)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t *cum = p.ws.cum_tiles;
    const uint32_t ntiles = cum[p.B];
    unsigned long long &s_carry = s_warp[SCAN_THREADS / 32];
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < ntiles; base += SCAN_THREADS) {
        const uint32_t t = base + tid;
        const unsigned long long v = t < ntiles ? __ldcg(p.ws.tile_agg + t) : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (unsigned)d) incl += o;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned long long off = s_carry;
        for (uint32_t k = 0; k < warp; ++k) off += s_warp[k];
        if (t < ntiles) p.ws.tile_pre[t] = off + incl - v;
        __syncthreads();
        if (tid == SCAN_THREADS - 1) s_carry = off + incl;
        __syncthreads();
    }
    // rank prefix at a frame's first tile = cells of all earlier frames
    const unsigned long long total = s_carry;
    for (int b0 = 0; b0 < p.B; b0 += SCAN_THREADS) {
        const int b = b0 + (int)tid;
        uint32_t m = 0;
        if (b < p.B) {
            const uint32_t t0 = cum[b], t1 = cum[b + 1];
            const unsigned long long g0 = t0 < ntiles ? p.ws.tile_pre[t0] : total;
            const unsigned long long g1 = t1 < ntiles ? p.ws.tile_pre[t1] : total;
            const uint32_t raw = (uint32_t)(g1 >> 32) - (uint32_t)(g0 >> 32);
            p.ws.frame_rank0[b] = (uint32_t)(g0 >> 32);
            p.ws.counts_raw[b] = raw;
            m = min(raw, (uint32_t)p.V);
            p.voxel_counts[b] = (int32_t)m;
        }
        // block-wide exclusive sum of m -> output row bases
        uint32_t incl = m;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (unsigned)d) incl += o;
        }
        __syncthreads();
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t off = (b0 == 0) ? 0u : (uint32_t)p.ws.base[b0];
        for (uint32_t k = 0; k < warp; ++k) off += (uint32_t)s_warp[k];
        if (b < p.B) p.ws.base[b + 1] = (int32_t)(off + incl);
        if (b == 0) p.ws.base[0] = 0;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_tile_reduce(const __grid_constant__ PvParams p)
{
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32 + 1];
    __shared__ uint32_t s_last;
    const TileRange t = pv_tile_range(p, blockIdx.x);
    if (t.b >= 0) {
        uint32_t w[SCAN_ITEMS];
        unsigned long long val[SCAN_ITEMS];
        unsigned long long sum = pv_load_items(p, t, w, val);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        if ((threadIdx.x & 31u) == 0) s_warp[threadIdx.x >> 5] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long tot = 0;
#pragma unroll
            for (int k = 0; k < SCAN_THREADS / 32; ++k) tot += s_warp[k];
            p.ws.tile_agg[blockIdx.x] = tot;
        }
    }
    // last block to arrive scans the tile sums (counter restores itself for the next call)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t done = atomicAdd(p.ws.ctrl + 2, 1u);
        s_last = (done == gridDim.x - 1) ? 1u : 0u;
        if (s_last) p.ws.ctrl[2] = 0u;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        pv_scan_mid(p, s_warp);
    }
}

// meta word of a cell: [63:48] arrival cursor | [47:32] min(count, 65535) | [31:0] kg
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const __grid_constant__ PvParams p)
{
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
    // per-voxel records of the tile, indexed by rank inside the tile: a thread owns SCAN_ITEMS
    // consecutive points, so writing them straight out would put the lanes of a store 32 bytes
    // apart; staged here they leave as coalesced rows
    __shared__ uint32_t s_cell[SCAN_TILE], s_kg[SCAN_TILE], s_c[SCAN_TILE];
    const TileRange t = pv_tile_range(p, blockIdx.x);
    if (t.b < 0) return;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // everything this thread will need is requested up front: one memory round trip, not three
    const uint32_t i0 = t.lo + tid * SCAN_ITEMS;
    const unsigned long long tile_pre = __ldcg(p.ws.tile_pre + blockIdx.x);
    const uint32_t frame_rank0 = __ldcg(p.ws.frame_rank0 + t.b);
    uint32_t sl[SCAN_ITEMS];                               // all slot loads in flight before any use
    uint32_t pc[SCAN_ITEMS];                               // hash maps: the cell index travels per point
    if (i0 + SCAN_ITEMS <= t.hi && (i0 & 3u) == 0) {   // 16-byte loads: 4 lines per warp, not 32 sectors
#pragma unroll
        for (int h = 0; h < SCAN_ITEMS / 4; ++h) {
            const uint4 a = __ldcs(reinterpret_cast<const uint4 *>(p.ws.slot + i0) + h);
            sl[4 * h] = a.x; sl[4 * h + 1] = a.y; sl[4 * h + 2] = a.z; sl[4 * h + 3] = a.w;
            const uint4 c4 = p.ws.dense ? make_uint4(0, 0, 0, 0) : __ldcs(reinterpret_cast<const uint4 *>(p.ws.pcell + i0) + h);
            pc[4 * h] = c4.x; pc[4 * h + 1] = c4.y; pc[4 * h + 2] = c4.z; pc[4 * h + 3] = c4.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) sl[j] = i0 + j < t.hi ? __ldcs(p.ws.slot + i0 + j) : 0u;
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) pc[j] = (i0 + j < t.hi && !p.ws.dense) ? __ldcs(p.ws.pcell + i0 + j) : 0u;
    }
    uint32_t w[SCAN_ITEMS];
    unsigned long long val[SCAN_ITEMS];
    const unsigned long long tsum = pv_load_items(p, t, w, val);
    unsigned long long incl = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (unsigned)d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    const unsigned long long tile0 = tile_pre - ((unsigned long long)frame_rank0 << 32);
    unsigned long long excl = tile0, tile_total = 0;
#pragma unroll
    for (uint32_t k = 0; k < SCAN_THREADS / 32; ++k) {
        if (k < warp) excl += s_warp[k];
        tile_total += s_warp[k];
    }
    excl += incl - tsum;
    const uint32_t r_tile = (uint32_t)(tile0 >> 32);                     // rank of the tile's first voxel
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        if (val[j]) {
            const uint32_t i = i0 + j;
            const uint32_t r = (uint32_t)(excl >> 32);
            const uint32_t kg = (uint32_t)excl;
            const uint32_t c = w[j] & 0x7FFFFFFFu;
            const uint32_t s = sl[j];
            const bool keep = r < (uint32_t)p.V;                          // :60-61 max_voxels
            p.ws.meta[s] = ((unsigned long long)min(c, 65535u) << 32) | (keep ? kg : PV_INF);
            if (keep) p.ws.kept[kg] = i;                                  // list element 0 = first point
            const uint32_t k = r - r_tile;                                // < SCAN_TILE: one voxel per first point
            s_cell[k] = p.ws.dense ? pv_dense_cell(p, s - (uint32_t)t.b * p.ws.capf) : pc[j];
            s_kg[k] = kg;
            s_c[k] = c;
        }
        excl += val[j];
    }
    __syncthreads();
    const uint32_t n_vox = (uint32_t)(tile_total >> 32);
    const uint32_t r_kept = min(r_tile + n_vox, (uint32_t)p.V);          // :60-61 max_voxels
    const uint32_t r_end = min(r_kept, p.ws.fcap);                        // kept voxels that fit
    if (tid == 0 && r_kept > p.ws.fcap && r_kept > r_tile) atomicOr(p.ws.ctrl + 1, 1u);   // frame larger than frame_capacity
    for (uint32_t r = r_tile + tid; r < r_end; r += SCAN_THREADS) {
        const size_t v = (size_t)t.b * p.ws.fcap + r;
        p.ws.vox_cell[v] = s_cell[r - r_tile];
        p.ws.vox_kg[v] = s_kg[r - r_tile];
        p.ws.vox_c[v] = s_c[r - r_tile];
    }
}

// ---------------------------------------------------------------------------------------------
// K4 -- place every other point of a kept voxel in the voxel's list.  One 64-bit atomicAdd on
// the cell's meta word returns the list offset, the count class and an arrival position:
//   count <= PV_SORT_MAX : the point goes to its arrival position (k_emit orders the few indices
//                          in registers), one store;
//   larger cells         : atomicMin chain; the list converges to the min(count, T) smallest
//                          indices in ascending order.
// ---------------------------------------------------------------------------------------------
#define PV_SORT_MAX 8u

__global__ void __launch_bounds__(256) k_place(const __grid_constant__ PvParams p)
{
    // one point per thread: the chains below serialise per warp, so warps must stay plentiful
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const uint32_t s = __ldcs(p.ws.slot + i);
    const uint32_t w = __ldcs(p.ws.pv + i);
    if (s == PV_INF || (w & PV_FIRST)) return;            // out of range, or already placed by the scan
    const unsigned long long m = atomicAdd(p.ws.meta + s, 1ull << 48);
    const uint32_t kg = (uint32_t)m;
    if (kg == PV_INF) return;                             // voxel beyond max_voxels: dropped
    const uint32_t c = (uint32_t)(m >> 32) & 0xFFFFu;
    uint32_t *list = p.ws.kept + kg;
    // cells that keep all their points: arrival order through the cursor -- k_emit sorts up to PV_SORT_MAX of them,
    // a consumer that takes maxima and a mean over the voxel (the PFN front end) does not care about the order at all
    if (c <= (uint32_t)p.T && (c <= PV_SORT_MAX || p.any_order)) { list[1u + (uint32_t)(m >> 48)] = i; return; }
    const uint32_t L = min(c, (uint32_t)p.T);
    if (L < 2) return;
    // slots only ever decrease: a tail already below i can never admit i
    if (c > (uint32_t)p.T && pv_ld_volatile(list + L - 1) < i) return;
    uint32_t x = i;
    for (uint32_t k = 1; k < L; ++k) {
        const uint32_t old = atomicMin(list + k, x);
        if (old == PV_INF) break;       // took a free slot, nothing displaced
        x = max(old, x);                // carry the loser to the next slot
    }
}

// ---------------------------------------------------------------------------------------------
// K5 -- emit: one thread per kept voxel (b, r); rows are fetched four at a time.
// ---------------------------------------------------------------------------------------------
#define PV_CSWAP(a, b) do { const uint32_t lo_ = min(a, b), hi_ = max(a, b); a = lo_; b = hi_; } while (0)

template <int CT>
__global__ void __launch_bounds__(256, 4) k_emit(const __grid_constant__ PvParams p)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (r >= (uint32_t)p.voxel_counts[b]) return;
    const size_t v = (size_t)b * p.ws.fcap + r;
    const uint32_t cell = __ldcs(p.ws.vox_cell + v);
    const uint32_t kg = __ldcs(p.ws.vox_kg + v);
    const uint32_t c = __ldcs(p.ws.vox_c + v);
    const int32_t vid = p.ws.base[b] + (int32_t)r;
    const uint32_t L = min(c, (uint32_t)p.T);
    uint32_t *list = p.ws.kept + kg;
    const int C = p.C, c_in = p.c_in;
    const bool fill = p.voxels != nullptr;                // the padded tensor: k_fill_voxels, which also restores the list

    const uint32_t nx = p.grid[0], ny = p.grid[1];
    // (measured and rejected: a separate path for single-point voxels -- five of six on the 3-D grids -- without sort,
    // sum and division: a warp holds both kinds, so both paths run: emit 86 vs 70 us on config 4)
    // first eight indices into registers; small cells hold them in arrival order -> sort
    uint32_t e[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = (uint32_t)k < L ? __ldcg(list + k) : PV_INF;
    if (c > 2 && c <= PV_SORT_MAX) {                      // Batcher odd-even merge sort, 19 exchanges
        PV_CSWAP(e[0], e[1]); PV_CSWAP(e[2], e[3]); PV_CSWAP(e[4], e[5]); PV_CSWAP(e[6], e[7]);
        PV_CSWAP(e[0], e[2]); PV_CSWAP(e[1], e[3]); PV_CSWAP(e[4], e[6]); PV_CSWAP(e[5], e[7]);
        PV_CSWAP(e[1], e[2]); PV_CSWAP(e[5], e[6]);
        PV_CSWAP(e[0], e[4]); PV_CSWAP(e[1], e[5]); PV_CSWAP(e[2], e[6]); PV_CSWAP(e[3], e[7]);
        PV_CSWAP(e[2], e[4]); PV_CSWAP(e[3], e[5]);
        PV_CSWAP(e[1], e[2]); PV_CSWAP(e[3], e[4]); PV_CSWAP(e[5], e[6]);
        if (fill) {                                       // k_fill_voxels reads the list in index order
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if ((uint32_t)k < L) list[k] = e[k];
        }
    }
    float acc[CT];
#pragma unroll
    for (int k = 0; k < CT; ++k) acc[k] = 0.0f;
    for (uint32_t j0 = 0; j0 < L; j0 += 4) {
        uint32_t id[4];
        if (j0 == 0) { id[0] = e[0]; id[1] = e[1]; id[2] = e[2]; id[3] = e[3]; }
        else if (j0 == 4) { id[0] = e[4]; id[1] = e[5]; id[2] = e[6]; id[3] = e[7]; }
        else {
#pragma unroll
            for (int q = 0; q < 4; ++q) id[q] = j0 + q < L ? __ldcg(list + j0 + q) : PV_INF;
        }
        float raw[4][CT];                                 // up to 4 * c_in independent loads in flight
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (id[q] != PV_INF) pv_load_row(p.pts, id[q], c_in, raw[q]);
            else {
#pragma unroll
                for (int k = 0; k < CT; ++k) raw[q][k] = 0.0f;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (id[q] == PV_INF) break;
            if (!fill) list[j0 + q] = PV_INF;             // restore the list for the next call
            float f[CT];
            if (p.cart) {   // utils.py:42-44: (rho, phi, z, x, y, feat3..)
                f[0] = pv_rho(raw[q][0], raw[q][1]);
                f[1] = pv_atan2f(raw[q][1], raw[q][0]);
                f[2] = raw[q][2]; f[3] = raw[q][0]; f[4] = raw[q][1];
#pragma unroll
                for (int k = 5; k < CT; ++k) f[k] = raw[q][k - 2];
            } else {
#pragma unroll
                for (int k = 0; k < CT; ++k) f[k] = raw[q][k];
            }
#pragma unroll
            for (int k = 0; k < CT; ++k) acc[k] = __fadd_rn(acc[k], f[k]);   // index order, like sum(dim=1)
        }
    }
    const uint32_t x = cell % nx, yz = cell / nx;
    __stcs(reinterpret_cast<int4 *>(p.coors) + vid, make_int4(b, (int)(yz / ny), (int)(yz % ny), (int)x));
    __stcs(p.num_points + vid, (int32_t)L);
    if (p.density) p.density[(size_t)b * p.cells + cell] = (int32_t)c;   // :70-71 un-capped count
    const float nf = (float)L;
    float *cv = p.canvas ? p.canvas + (size_t)b * C * p.cells + cell : nullptr;
    float mean[CT];
#pragma unroll
    for (int k = 0; k < CT; ++k) {
        mean[k] = k < C ? __fdiv_rn(acc[k], nf) : 0.0f;                 // voxel_encoder.py:18-22
        if (k < C && cv) __stcs(cv + (size_t)k * p.cells, mean[k]);     // pillar_encoder.py:211-217
    }
    if (p.feats) pv_store_feats<CT>(p.feats, vid, C, mean);
}

__global__ void __launch_bounds__(256) k_transform(const float *__restrict__ in, long long n,
                                                   int c_in, int cylinder, float *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *row = in + i * c_in;
    float *o = out + i * (c_in + 2);
    const float x = row[0], y = row[1];
    const float rho = pv_rho(x, y), phi = pv_atan2f(y, x);
    if (cylinder) {          // utils.py:42-44
        o[0] = rho; o[1] = phi; o[2] = row[2]; o[3] = x; o[4] = y;
        for (int k = 3; k < c_in; ++k) o[k + 2] = row[k];
    } else {                 // utils.py:45-47
        for (int k = 0; k < c_in; ++k) o[k] = row[k];
        o[c_in] = rho; o[c_in + 1] = phi;
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
int pv_last_cuda_error() { return cudaGetLastError() == cudaSuccess ? PV_OK : PV_ERR_CUDA; }

int pv_check_config(const pv_config *cfg)
{
    if (!cfg) return PV_ERR_BAD_ARGUMENT;
    long long cells = 1;
    for (int j = 0; j < 3; ++j) {
        if (cfg->grid[j] <= 0 || !(cfg->vs[j] > 0.0f)) return PV_ERR_BAD_CONFIG;
        cells *= cfg->grid[j];
        if (cells >= (1ll << 31)) return PV_ERR_BAD_CONFIG;
    }
    if (cfg->max_points <= 0 || cfg->max_points > 32767 || cfg->max_voxels <= 0) return PV_ERR_BAD_CONFIG;
    if (cfg->pipeline < 0 || cfg->pipeline > 2) return PV_ERR_BAD_CONFIG;
    return PV_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int pv_make_layout(const pv_config *cfg, int64_t n_cap, int32_t batch, int64_t frame_capacity,
                   void *base, PvWs *w)
{
    int rc = pv_check_config(cfg);
    if (rc) return rc;
    if (batch <= 0 || n_cap < 0 || frame_capacity < 0) return PV_ERR_BAD_ARGUMENT;
    if (n_cap >= (1ll << 30)) return PV_ERR_BAD_ARGUMENT;
    if (frame_capacity > n_cap) frame_capacity = n_cap;
    const uint64_t cells = (uint64_t)cfg->grid[0] * cfg->grid[1] * cfg->grid[2];
    const bool dense = cells <= PV_DENSE_MAX_CELLS;
    uint64_t capf;
    if (dense) capf = cells;
    else {
        const uint64_t want = (uint64_t)frame_capacity + (uint64_t)frame_capacity / 4 + 1;
        capf = 1024;
        while (capf < want) capf <<= 1;
    }
    if (capf * (uint64_t)batch >= 0xFFFFFF00ull) return PV_ERR_BAD_ARGUMENT;
    const size_t n = (size_t)(n_cap > 0 ? n_cap : 1);
    const size_t fcap = (size_t)(frame_capacity > 0 ? frame_capacity : 1);
    w->capf = (uint32_t)capf;
    w->fcap = (uint32_t)fcap;
    w->dense = dense ? 1u : 0u;
    w->max_tiles = (uint32_t)(n_cap / SCAN_TILE + batch + 1);
    char *p0 = (char *)base;
    size_t o = 0;
    const size_t slots = (size_t)capf * batch;
    w->ctrl = (uint32_t *)(p0 + o);                      o = align_up(o + 16 * sizeof(uint32_t), 256);
    w->counts_raw = (uint32_t *)(p0 + o);                o = align_up(o + (size_t)batch * 4, 256);
    w->base = (int32_t *)(p0 + o);                       o = align_up(o + (size_t)(batch + 1) * 4, 256);
    w->frame_rank0 = (uint32_t *)(p0 + o);               o = align_up(o + (size_t)batch * 4, 256);
    w->cum_tiles = (uint32_t *)(p0 + o);                 o = align_up(o + (size_t)(batch + 1) * 4, 256);
    w->tile_agg = (unsigned long long *)(p0 + o);        o = align_up(o + (size_t)w->max_tiles * 8, 256);
    w->tile_pre = (unsigned long long *)(p0 + o);        o = align_up(o + (size_t)w->max_tiles * 8, 256);
    w->table = (PvEntry *)(p0 + o);                      o = align_up(o + slots * sizeof(PvEntry) + 16, 256);
    w->keys = (uint32_t *)(p0 + o);                      if (!dense) o = align_up(o + slots * 4, 256);
    w->kept = (uint32_t *)(p0 + o);                      o = align_up(o + n * 4, 256);
    w->meta = (unsigned long long *)(p0 + o);            o = align_up(o + slots * 8, 256);
    w->slot = (uint32_t *)(p0 + o);                      o = align_up(o + n * 4 + 32, 256);
    w->pv = (uint32_t *)(p0 + o);                        o = align_up(o + n * 4 + 32, 256);
    w->pcell = (uint32_t *)(p0 + o);                     if (!dense) o = align_up(o + n * 4 + 32, 256);
    w->vox_cell = (uint32_t *)(p0 + o);                  o = align_up(o + fcap * batch * 4, 256);
    w->vox_kg = (uint32_t *)(p0 + o);                    o = align_up(o + fcap * batch * 4, 256);
    w->vox_c = (uint32_t *)(p0 + o);                     o = align_up(o + fcap * batch * 4, 256);
    w->total_bytes = o;
    return PV_OK;
}

// The caller's workspace holds both layouts back to back: [list-based PvWs | list-free PvF].
static int make_layouts(const pv_config *cfg, int64_t n_cap, int32_t batch, int64_t frame_capacity,
                        int32_t max_channels, void *base, PvWs *w, PvF *f, size_t *total)
{
    int rc = pv_make_layout(cfg, n_cap, batch, frame_capacity, base, w);
    if (rc) return rc;
    const size_t off = align_up(w->total_bytes, 256);
    rc = pvf_make_layout(cfg, n_cap, batch, frame_capacity, max_channels, (char *)base + off, f);
    if (rc) return rc;
    *total = off + f->total_bytes;
    return PV_OK;
}

// Pipeline choice = pv_config::pipeline (part of the caller's configuration, no process state):
// 0 auto      list-free on direct-map grids, list-based on hash-map grids (per-cell rows in a hash
//             map cost more random DRAM sectors than point lists when most voxels hold 1-2 points)
// 1 lists     list-based everywhere
// 2 list-free list-free wherever the padded voxels tensor is not requested
static bool use_lists(const pv_config *cfg, const PvF &f, bool want_voxels)
{
    if (want_voxels || cfg->pipeline == 1) return true;
    if (cfg->pipeline == 2) return false;
    return !f.dense;
}

static int fill_params(PvParams *p, PvF *f, const pv_config *cfg, const float *points,
                       const int32_t *frame_offsets, int32_t batch, int64_t n_total, int32_t c_in,
                       int32_t is_cartesian, int64_t max_points_total, int64_t frame_capacity,
                       void *workspace, size_t workspace_bytes)
{
    if (!points && n_total > 0) return PV_ERR_BAD_ARGUMENT;
    if (!frame_offsets || !workspace || n_total < 0 || n_total > max_points_total) return PV_ERR_BAD_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return PV_ERR_BAD_ARGUMENT;
    const int C = is_cartesian ? c_in + 2 : c_in;
    if (c_in < 3 || C > PV_MAX_CHANNELS) return PV_ERR_BAD_ARGUMENT;
    size_t total = 0;
    int rc = make_layouts(cfg, max_points_total, batch, frame_capacity, C, workspace, &p->ws, f, &total);
    if (rc) return rc;
    if (total > workspace_bytes) return PV_ERR_WORKSPACE;
    for (int j = 0; j < 3; ++j) {
        p->lo[j] = cfg->lo[j]; p->vs[j] = cfg->vs[j]; p->grid[j] = cfg->grid[j];
        p->gridf[j] = (float)cfg->grid[j];
        p->inv_vs[j] = 1.0f / cfg->vs[j];
    }
    p->T = cfg->max_points; p->V = cfg->max_voxels;
    p->pts = points; p->offsets = frame_offsets; p->B = batch; p->n = (uint32_t)n_total;
    p->c_in = c_in; p->cart = is_cartesian ? 1 : 0; p->C = C;
    p->cells = (uint32_t)((uint64_t)cfg->grid[0] * cfg->grid[1] * cfg->grid[2]);
    p->num_tiles = (uint32_t)(n_total / SCAN_TILE + batch + 1);
    p->coors = p->num_points = p->voxel_counts = p->grid_ind = p->density = nullptr;
    p->voxels = p->feats = p->canvas = nullptr;
    p->dyn = 0; p->any_order = 0; p->gi_in = nullptr; p->unq_inv = nullptr;
    return PV_OK;
}

// ---------------------------------------------------------------------------------------------
// K6 (drop-in calls that return the padded tensor) -- voxels [M, T, C], point_cloud_ops.py:65-68,187:
// one WARP per voxel.  Lane q gathers the row of the voxel's q-th kept point (index order, as
// k_emit left the list), the warp parks the rows in shared memory and writes the voxel's T * C
// floats as whole 128-byte lines, zero padding included -- a thread-per-voxel writer puts every
// lane of a store into a different voxel (560 bytes apart for T = 20, C = 7).  Restores the list.
// ---------------------------------------------------------------------------------------------
#define FILL_WARPS 8
__global__ void __launch_bounds__(FILL_WARPS * 32) k_fill_voxels(const __grid_constant__ PvParams p)
{
    __shared__ float s_rows[FILL_WARPS][32 * PV_MAX_CHANNELS];
    const uint32_t lane = threadIdx.x & 31u, wq = threadIdx.x >> 5;
    const uint32_t r = blockIdx.x * FILL_WARPS + wq;
    const int b = blockIdx.y;
    if (r >= (uint32_t)p.voxel_counts[b]) return;
    const size_t v = (size_t)b * p.ws.fcap + r;
    const uint32_t kg = __ldcs(p.ws.vox_kg + v);
    const uint32_t L = min(__ldcs(p.ws.vox_c + v), (uint32_t)p.T);
    const int32_t vid = p.ws.base[b] + (int32_t)r;
    const int C = p.C, c_in = p.c_in;
    uint32_t *list = p.ws.kept + kg;
    float *vox = p.voxels + (size_t)vid * p.T * C;
    float *rows = s_rows[wq];
    const uint32_t total = (uint32_t)p.T * (uint32_t)C;
    for (uint32_t q0 = 0; q0 < (uint32_t)p.T; q0 += 32) {             // 32 slots per round
        const uint32_t q = q0 + lane;
        if (q < L) {
            const uint32_t i = __ldcg(list + q);
            float f[PV_MAX_CHANNELS];
            pv_feature_row(p.pts, i, c_in, p.cart, f);
#pragma unroll
            for (int k = 0; k < PV_MAX_CHANNELS; ++k)
                if (k < C) rows[lane * C + k] = f[k];
            list[q] = PV_INF;                                            // restore the list for the next call
        }
        __syncwarp();
        const uint32_t e0 = q0 * C, e1 = min(total, (q0 + 32u) * C), filled = L > q0 ? (L - q0) * C : 0u;
        for (uint32_t e = e0 + lane; e < e1; e += 32)
            __stcs(vox + e, e - e0 < filled ? rows[e - e0] : 0.0f);   // zero padding (:187)
        __syncwarp();
    }
}

template <int CT>
static void launch_emit(const PvParams &p, cudaStream_t st)
{
    const uint32_t vmax = min((uint32_t)p.V, p.ws.fcap);
    dim3 grid((vmax + 255) / 256, (unsigned)p.B);
    k_emit<CT><<<grid, 256, 0, st>>>(p);
    if (p.voxels) k_fill_voxels<<<dim3((vmax + FILL_WARPS - 1) / FILL_WARPS, (unsigned)p.B), FILL_WARPS * 32, 0, st>>>(p);
}

// Stage boundaries (for pv_profile_*): ev[k] is recorded BEFORE stage k, ev[PV_STAGES] after the
// last one.  0 bin_insert (+canvas zero fill), 1 cell_flags, 2 scan (tile_reduce incl. the mid scan
// + scan_apply), 3 place, 4 emit (+canvas scatter).
#define PV_STAGES 5
#define PV_MARK(k) do { if (ev && cudaEventRecord(ev[k], st) != cudaSuccess) return PV_ERR_CUDA; } while (0)

// emit == false: stop after the point lists are in place (K1-K4); the caller's own kernel consumes the
// lists (and restores them) instead of k_emit -- the fused PFN front end.
static int run_voxelize(PvParams &p, PvF &f, cudaStream_t st, cudaEvent_t *ev = nullptr, bool emit = true)
{
    const PvWs &w = p.ws;
    if (p.density &&
        cudaMemsetAsync(p.density, 0, (size_t)p.B * p.cells * sizeof(int32_t), st) != cudaSuccess)
        return PV_ERR_CUDA;
    if (p.canvas &&        // k_emit scatters into a zero-filled canvas
        cudaMemsetAsync(p.canvas, 0, (size_t)p.B * p.C * p.cells * sizeof(float), st) != cudaSuccess)
        return PV_ERR_CUDA;
    PV_MARK(0);
    k_scan_layout<<<1, 32, 0, st>>>(p);
    if (w.dense) {
        const int rc = pvf_insert_lists(p, f, st);
        if (rc) return rc;
    } else if (p.n > 0) k_bin_insert<false><<<(p.n + K1_THREADS - 1) / K1_THREADS, K1_THREADS, 0, st>>>(p);
    PV_MARK(1);
    {
        const size_t pairs = ((size_t)p.B * w.capf + 1) / 2;
        k_cell_flags<<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(p);
    }
    PV_MARK(2);
    k_tile_reduce<<<p.num_tiles, SCAN_THREADS, 0, st>>>(p);      // its last block also runs the mid scan
    k_scan_apply<<<p.num_tiles, SCAN_THREADS, 0, st>>>(p);
    PV_MARK(3);
    if (p.n > 0) k_place<<<(p.n + 255) / 256, 256, 0, st>>>(p);
    PV_MARK(4);
    if (!emit) return pv_last_cuda_error();
    switch (p.C) {
    case 3: case 4: case 5: launch_emit<5>(p, st); break;
    case 6: launch_emit<6>(p, st); break;
    case 7: launch_emit<7>(p, st); break;
    case 8: launch_emit<8>(p, st); break;
    case 9: case 10: launch_emit<10>(p, st); break;
    default: launch_emit<PV_MAX_CHANNELS>(p, st); break;
    }
    PV_MARK(5);
    return pv_last_cuda_error();
}

// readers.cu
int pv_scatter_dev(const float *feats, const int32_t *coors, const int32_t *total_rows, int64_t cap, int32_t c,
                   int32_t batch, int32_t ny, int32_t nx, void *workspace, float *canvas, cudaStream_t st);

extern "C" {

int pv_version(void) { return 300; }

int pv_profile_pipeline(const pv_config *cfg)
{
    if (pv_check_config(cfg)) return PV_ERR_BAD_CONFIG;
    const uint64_t cells = (uint64_t)cfg->grid[0] * cfg->grid[1] * cfg->grid[2];
    PvF f;
    f.dense = cells <= PV_DENSE_MAX_CELLS ? 1u : 0u;
    return use_lists(cfg, f, false) ? 1 : 2;
}

const char *pv_error_string(int code)
{
    switch (code) {
    case PV_OK: return "ok";
    case PV_ERR_BAD_CONFIG: return "bad voxel grid configuration";
    case PV_ERR_BAD_ARGUMENT: return "bad argument";
    case PV_ERR_WORKSPACE: return "workspace too small";
    case PV_ERR_CUDA: return "CUDA runtime error";
    case PV_ERR_TABLE_FULL: return "a frame holds more points than frame_capacity";
    case PV_ERR_UNSUPPORTED: return "unsupported layer shape";
    case PV_ERR_INTERNAL: return "a pipeline wait of the tensor-core PFN kernel starved (watchdog)";
    default: return "unknown error";
    }
}

size_t pv_workspace_bytes(const pv_config *cfg, int64_t max_points_total, int32_t batch,
                          int64_t frame_capacity, int32_t channels)
{
    PvWs w;
    PvF f;
    size_t total = 0;
    if (make_layouts(cfg, max_points_total, batch, frame_capacity, channels, nullptr, &w, &f, &total) != PV_OK) return 0;
    return total;
}

int pv_workspace_init(const pv_config *cfg, int64_t max_points_total, int32_t batch,
                      int64_t frame_capacity, int32_t channels, void *workspace,
                      size_t workspace_bytes, pv_stream_t stream)
{
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return PV_ERR_BAD_ARGUMENT;
    PvWs w;
    PvF f;
    size_t total = 0;
    int rc = make_layouts(cfg, max_points_total, batch, frame_capacity, channels, workspace, &w, &f, &total);
    if (rc) return rc;
    if (total > workspace_bytes) return PV_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    rc = pvf_init(f, batch, max_points_total, st);
    if (rc) return rc;
    const size_t zero_bytes = (size_t)((char *)w.table - (char *)workspace);
    const size_t ff_bytes = (size_t)((char *)w.meta - (char *)w.table);   // table, keys, kept
    if (cudaMemsetAsync(workspace, 0, zero_bytes, st) != cudaSuccess) return PV_ERR_CUDA;
    if (cudaMemsetAsync(w.table, 0xFF, ff_bytes, st) != cudaSuccess) return PV_ERR_CUDA;
    return PV_OK;
}

int pv_transform_points(const float *in, int64_t n, int32_t c_in, int32_t cylinder, float *out,
                        pv_stream_t stream)
{
    if (n < 0 || c_in < 3 || (n > 0 && (!in || !out))) return PV_ERR_BAD_ARGUMENT;
    if (n == 0) return PV_OK;
    k_transform<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, n, c_in, cylinder, out);
    return pv_last_cuda_error();
}

int pv_voxelize(const pv_config *cfg, const float *points, const int32_t *frame_offsets,
                int32_t batch, int64_t n_total, int32_t c_in, int32_t is_cartesian,
                int64_t max_points_total, int64_t frame_capacity, void *workspace,
                size_t workspace_bytes, int32_t *coors, int32_t *num_points, int32_t *voxel_counts,
                float *voxels, float *mean_feats, int32_t *pc_grid_ind, int32_t *density,
                pv_stream_t stream)
{
    PvParams p;
    PvF f;
    int rc = fill_params(&p, &f, cfg, points, frame_offsets, batch, n_total, c_in, is_cartesian,
                         max_points_total, frame_capacity, workspace, workspace_bytes);
    if (rc) return rc;
    if (!coors || !num_points || !voxel_counts) return PV_ERR_BAD_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(coors) & 15u) != 0) return PV_ERR_BAD_ARGUMENT;
    p.coors = coors; p.num_points = num_points; p.voxel_counts = voxel_counts;
    p.voxels = voxels; p.feats = mean_feats; p.grid_ind = pc_grid_ind; p.density = density;
    // the padded [M, T, C] tensor needs per-voxel point lists; everything else runs list-free
    if (use_lists(cfg, f, voxels != nullptr)) return run_voxelize(p, f, (cudaStream_t)stream);
    return pvf_run(p, f, (cudaStream_t)stream, nullptr);
}

static int setup_canvas_call(PvParams &p, const pv_config *cfg, int32_t *coors, int32_t *num_points,
                             int32_t *voxel_counts, float *mean_feats, float *canvas)
{
    if (!coors || !num_points || !voxel_counts || !mean_feats) return PV_ERR_BAD_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(coors) & 15u) != 0) return PV_ERR_BAD_ARGUMENT;
    if (canvas) {
        if (cfg->grid[2] != 1 || (cfg->grid[0] & 3) != 0) return PV_ERR_BAD_CONFIG;
        if ((reinterpret_cast<uintptr_t>(canvas) & 15u) != 0) return PV_ERR_BAD_ARGUMENT;
    }
    p.coors = coors; p.num_points = num_points; p.voxel_counts = voxel_counts;
    p.feats = mean_feats; p.canvas = canvas;
    return PV_OK;
}

int pv_forward_mean_canvas(const pv_config *cfg, const float *points, const int32_t *frame_offsets,
                           int32_t batch, int64_t n_total, int32_t c_in, int32_t is_cartesian,
                           int64_t max_points_total, int64_t frame_capacity, void *workspace,
                           size_t workspace_bytes, int32_t *coors, int32_t *num_points,
                           int32_t *voxel_counts, float *mean_feats, float *canvas,
                           pv_stream_t stream)
{
    PvParams p;
    PvF f;
    int rc = fill_params(&p, &f, cfg, points, frame_offsets, batch, n_total, c_in, is_cartesian,
                         max_points_total, frame_capacity, workspace, workspace_bytes);
    if (rc) return rc;
    if (!canvas) return PV_ERR_BAD_ARGUMENT;
    rc = setup_canvas_call(p, cfg, coors, num_points, voxel_counts, mean_feats, canvas);
    if (rc) return rc;
    if (use_lists(cfg, f, false)) return run_voxelize(p, f, (cudaStream_t)stream);
    return pvf_run(p, f, (cudaStream_t)stream, nullptr);
}

int pv_profile_mean_canvas(const pv_config *cfg, const float *points, const int32_t *frame_offsets,
                           int32_t batch, int64_t n_total, int32_t c_in, int32_t is_cartesian,
                           int64_t max_points_total, int64_t frame_capacity, void *workspace,
                           size_t workspace_bytes, int32_t *coors, int32_t *num_points,
                           int32_t *voxel_counts, float *mean_feats, float *canvas,
                           pv_stream_t stream, int32_t iters, float *stage_ms)
{
    PvParams p;
    PvF f;
    int rc = fill_params(&p, &f, cfg, points, frame_offsets, batch, n_total, c_in, is_cartesian,
                         max_points_total, frame_capacity, workspace, workspace_bytes);
    if (rc) return rc;
    if (!stage_ms || iters <= 0) return PV_ERR_BAD_ARGUMENT;
    rc = setup_canvas_call(p, cfg, coors, num_points, voxel_counts, mean_feats, canvas);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t ev[PV_STAGES + 1];
    for (int k = 0; k <= PV_STAGES; ++k)
        if (cudaEventCreate(&ev[k]) != cudaSuccess) return PV_ERR_CUDA;
    for (int k = 0; k < PV_STAGES; ++k) stage_ms[k] = 0.0f;
    for (int it = 0; it < iters && rc == PV_OK; ++it) {
        rc = use_lists(cfg, f, false) ? run_voxelize(p, f, st, ev) : pvf_run(p, f, st, ev);
        if (rc == PV_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = PV_ERR_CUDA;
        for (int k = 0; k < PV_STAGES && rc == PV_OK; ++k) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, ev[k], ev[k + 1]) != cudaSuccess) rc = PV_ERR_CUDA;
            stage_ms[k] += ms / (float)iters;
        }
    }
    for (int k = 0; k <= PV_STAGES; ++k) cudaEventDestroy(ev[k]);
    return rc;
}

int pv_dynamic_voxelize(const pv_config *cfg, const float *points, const int32_t *frame_offsets,
                        const int32_t *grid_ind_in, int32_t batch, int64_t n_total, int32_t c_in,
                        int32_t is_cartesian, int64_t max_points_total, int64_t frame_capacity,
                        void *workspace, size_t workspace_bytes, int32_t *grid_ind_out, int32_t *unq,
                        int32_t *unq_inv, int32_t *unq_cnt, int32_t *voxel_counts, float *mean_feats,
                        float *canvas, pv_stream_t stream)
{
    PvParams p;
    PvF f;
    if (!frame_offsets && !grid_ind_in) return PV_ERR_BAD_ARGUMENT;
    static const int32_t dummy = 0;       // fill_params only checks for NULL; kernels never read it when gi_in is set
    int rc = fill_params(&p, &f, cfg, points, frame_offsets ? frame_offsets : &dummy, batch, n_total, c_in,
                         is_cartesian, max_points_total, frame_capacity, workspace, workspace_bytes);
    if (rc) return rc;
    if (!frame_offsets) p.offsets = nullptr;
    if (!unq || !voxel_counts) return PV_ERR_BAD_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(unq) & 15u) != 0) return PV_ERR_BAD_ARGUMENT;
    if (grid_ind_in && (reinterpret_cast<uintptr_t>(grid_ind_in) & 15u) != 0) return PV_ERR_BAD_ARGUMENT;
    if (grid_ind_out && (reinterpret_cast<uintptr_t>(grid_ind_out) & 15u) != 0) return PV_ERR_BAD_ARGUMENT;
    if (canvas) {
        if (cfg->grid[2] != 1) return PV_ERR_BAD_CONFIG;
        if ((reinterpret_cast<uintptr_t>(canvas) & 15u) != 0) return PV_ERR_BAD_ARGUMENT;
    }
    p.dyn = 1; p.gi_in = grid_ind_in; p.grid_ind = grid_ind_out;
    p.coors = unq; p.num_points = unq_cnt; p.unq_inv = unq_inv; p.voxel_counts = voxel_counts;
    p.feats = mean_feats; p.canvas = canvas;
    return pvf_run_dynamic(p, f, (cudaStream_t)stream);
}

size_t pv_pfn_canvas_workspace_bytes(int32_t batch, int32_t ny, int32_t nx, int64_t max_points_total, int32_t max_voxels)
{
    const size_t map = pv_scatter_workspace_bytes(batch, ny, nx);
    if (!map || max_points_total < 0 || max_voxels <= 0) return 0;
    // [chunk queues + watchdog words | BEV index map | decorated rows, group descriptors, groups per chunk]
    const long long rows = (long long)max_points_total + std::min<long long>((long long)batch * max_voxels, (long long)max_points_total);
    return 256 + ((map + 255) & ~(size_t)255) + pv_pfn_rows_bytes(rows, max_voxels, batch, nullptr, nullptr);
}

int pv_forward_pfn_canvas(const pv_config *cfg, const float *points, const int32_t *frame_offsets,
                          int32_t batch, int64_t n_total, int32_t c_in, int32_t is_cartesian,
                          int64_t max_points_total, int64_t frame_capacity, void *workspace,
                          size_t workspace_bytes, void *aux_workspace, size_t aux_bytes,
                          const pv_pfn_layer *layers, int32_t n_layers, int32_t with_distance, float vx, float vy,
                          float x_off, float y_off, float eps, int32_t *coors, int32_t *num_points,
                          int32_t *voxel_counts, float *pfn_feats, float *canvas, pv_stream_t stream)
{
    PvParams p;
    PvF f;
    int rc = fill_params(&p, &f, cfg, points, frame_offsets, batch, n_total, c_in, is_cartesian,
                         max_points_total, frame_capacity, workspace, workspace_bytes);
    if (rc) return rc;
    if (!coors || !num_points || !voxel_counts || !pfn_feats || !layers || !aux_workspace) return PV_ERR_BAD_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(coors) & 15u) != 0 || (reinterpret_cast<uintptr_t>(pfn_feats) & 15u) != 0 ||
        (reinterpret_cast<uintptr_t>(aux_workspace) & 255u) != 0)
        return PV_ERR_BAD_ARGUMENT;
    if (canvas && (cfg->grid[2] != 1 || (reinterpret_cast<uintptr_t>(canvas) & 15u) != 0)) return PV_ERR_BAD_CONFIG;
    if (aux_bytes < pv_pfn_canvas_workspace_bytes(batch, cfg->grid[1], cfg->grid[0], max_points_total, cfg->max_voxels)) return PV_ERR_WORKSPACE;
    if (n_layers <= 0 || n_layers > PV_MAX_PFN_LAYERS) return PV_ERR_BAD_ARGUMENT;
    for (int l = 0; l < n_layers; ++l)
        if (!layers[l].weight || !layers[l].bn_mean || !layers[l].bn_var || !layers[l].bn_gamma || !layers[l].bn_beta)
            return PV_ERR_BAD_ARGUMENT;
    if (!pv_pfn_fused_supported(layers, n_layers, cfg->max_points, p.C, with_distance)) return PV_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    p.coors = coors; p.num_points = num_points; p.voxel_counts = voxel_counts;
    p.any_order = 1;                                     // max / mean per voxel: no sorted lists needed
    rc = run_voxelize(p, f, st, nullptr, false);         // point lists only: no [M, T, C] tensor, no k_emit
    if (rc) return rc;
    P2Args a = {};
    a.mode = 1;
    a.pts = points; a.c_in = c_in; a.cart = is_cartesian ? 1 : 0;
    a.vox_cell = p.ws.vox_cell; a.vox_kg = p.ws.vox_kg; a.vox_c = p.ws.vox_c; a.kept = p.ws.kept;
    a.base = p.ws.base; a.voxel_counts = voxel_counts; a.fcap = p.ws.fcap;
    a.nx = cfg->grid[0]; a.ny = cfg->grid[1];
    a.coors_out = coors; a.num_out = num_points;
    a.t = cfg->max_points; a.c = p.C; a.with_distance = with_distance ? 1 : 0; a.c0 = p.C + 5 + a.with_distance;
    a.vx = vx; a.vy = vy; a.x_off = x_off; a.y_off = y_off; a.eps = eps;
    a.counter = reinterpret_cast<unsigned int *>(aux_workspace);
    a.status = p.ws.ctrl + 1;
    a.out = pfn_feats;
    const long long vcap = std::min<long long>(cfg->max_voxels, (long long)p.ws.fcap);
    {
        const size_t map = (pv_scatter_workspace_bytes(batch, cfg->grid[1], cfg->grid[0]) + 255) & ~(size_t)255;
        char *rows0 = reinterpret_cast<char *>(aux_workspace) + 256 + map;
        size_t desc_off = 0, ng_off = 0;
        const long long rows = (long long)max_points_total + std::min<long long>((long long)batch * cfg->max_voxels, (long long)max_points_total);
        pv_pfn_rows_bytes(rows, vcap, batch, &desc_off, &ng_off);
        a.drows_out = reinterpret_cast<float4 *>(rows0);
        a.drow_stride = pv_pfn_rows_stride(rows);
        a.desc_out = reinterpret_cast<uint4 *>(rows0 + desc_off);
        a.ngroups_out = reinterpret_cast<uint32_t *>(rows0 + ng_off);
    }
    rc = pv_pfn_fused_launch(a, layers, batch, vcap, st);
    if (rc || !canvas) return rc;
    const int64_t cap = std::min<int64_t>((int64_t)batch * cfg->max_voxels, n_total);
    return pv_scatter_dev(pfn_feats, coors, p.ws.base + batch, cap, layers[n_layers - 1].units, batch, cfg->grid[1], cfg->grid[0],
                          reinterpret_cast<char *>(aux_workspace) + 256, canvas, st);
}

int pv_dynamic_grid_ind(const pv_config *cfg, const float *points, const int32_t *frame_offsets, int32_t batch,
                        int64_t n_total, int32_t c_in, int32_t is_cartesian, int32_t *grid_ind_out, pv_stream_t stream)
{
    int rc = pv_check_config(cfg);
    if (rc) return rc;
    if (!frame_offsets || !grid_ind_out || batch <= 0 || n_total < 0 || n_total >= (1ll << 31) || (!points && n_total > 0))
        return PV_ERR_BAD_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(grid_ind_out) & 15u) != 0) return PV_ERR_BAD_ARGUMENT;
    const int C = is_cartesian ? c_in + 2 : c_in;
    if (c_in < 3 || C > PV_MAX_CHANNELS) return PV_ERR_BAD_ARGUMENT;
    PvParams p = {};
    for (int j = 0; j < 3; ++j) {
        p.lo[j] = cfg->lo[j]; p.vs[j] = cfg->vs[j]; p.grid[j] = cfg->grid[j];
        p.gridf[j] = (float)cfg->grid[j];
        p.inv_vs[j] = 1.0f / cfg->vs[j];
    }
    p.pts = points; p.offsets = frame_offsets; p.B = batch; p.n = (uint32_t)n_total;
    p.c_in = c_in; p.cart = is_cartesian ? 1 : 0; p.C = C;
    p.grid_ind = grid_ind_out;
    return pvf_run_grid_ind(p, (cudaStream_t)stream);
}

int pv_read_status(const void *workspace, pv_stream_t stream)
{
    if (!workspace) return PV_ERR_BAD_ARGUMENT;
    uint32_t ctrl[2] = {0, 0};
    if (cudaMemcpyAsync(ctrl, workspace, sizeof(ctrl), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess)
        return PV_ERR_CUDA;
    if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return PV_ERR_CUDA;
    if (ctrl[1] & 4u) return PV_ERR_INTERNAL;                 // the tensor-core PFN's barrier watchdog fired
    if (ctrl[1] & 1u) return PV_ERR_TABLE_FULL;
    return (ctrl[1] & 2u) ? PV_ERR_BAD_ARGUMENT : PV_OK;     // bit 1: a caller-provided grid_ind row was outside the grid
}

}  // extern "C"
