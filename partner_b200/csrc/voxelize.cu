// voxelize.cu -- batched hard voxelization on the polar grid (sm_100a).
//
// Replaces the sequential numba loop of det3d/ops/point_cloud/point_cloud_ops.py:7-72
// (called through VoxelGenerator.generate, core/input/voxel_generator.py:19-32) with an
// order-free formulation that reproduces it exactly:
//
//   K1 bin_insert   point -> (rho, phi, z) bin -> cell; per cell atomicMin(first point index) and
//                   atomicAdd(count) in a per-frame map (direct map for small grids, hash else),
//                   warp-aggregated so one lane per distinct cell issues the atomics.
//   K2 rank_scan    a point is its cell's first point iff map.first == i.  One decoupled
//                   look-back scan over all points, in point order, of (is_first, min(count, T))
//                   gives every cell its first-occurrence rank g and its list offset.
//   K3 fill_lists   every point of a kept voxel inserts its index into the voxel's list, which
//                   converges to the T smallest indices in ascending order (atomicMin chain).
//   K4 emit         per voxel: coors (b, z, y, x), num_points, mean feature; optional padded
//                   voxels tensor, density, and (pillar grids) the dense BEV canvas.
//
// Voxel order = first-occurrence order, kept points = the T smallest indices in index order,
// voxels with rank >= V dropped -- the three order-dependent behaviours of the reference loop.
#include "pv_common.cuh"

#define K1_THREADS 256
#define K2_THREADS 256
#define K2_ITEMS 4
#define K2_TILE (K2_THREADS * K2_ITEMS)

// ---------------------------------------------------------------------------------------------
// K1
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pv_claim(PvEntry *tab, uint32_t mask, uint32_t cell,
                                             uint32_t *status)
{
    uint32_t h = pv_hash(cell) & mask;
    for (uint32_t probe = 0; probe <= mask; ++probe) {
        uint32_t k = pv_ld_volatile(&tab[h].key);
        if (k == cell) return h;
        if (k == PV_INF) {
            uint32_t old = atomicCAS(&tab[h].key, PV_INF, cell);
            if (old == PV_INF || old == cell) return h;
        }
        h = (h + 1) & mask;
    }
    atomicOr(status, 1u);
    return PV_INF;
}

template <bool DENSE>
__global__ void __launch_bounds__(K1_THREADS) k_bin_insert(const __grid_constant__ PvParams p)
{
    __shared__ __align__(16) float s_pts[K1_THREADS * PV_MAX_CHANNELS];
    const uint32_t tile_base = blockIdx.x * K1_THREADS;
    const uint32_t tid = threadIdx.x;
    const uint32_t n_tile = min((uint32_t)K1_THREADS, p.n - tile_base);
    const int c_in = p.c_in;

    // ---- stage the tile's rows: coalesced 128-bit loads of the contiguous float range ----
    {
        const size_t f0 = (size_t)tile_base * c_in;
        const uint32_t nf = n_tile * c_in;
        const float *src = p.pts + f0;
        if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
            const uint32_t nv = nf >> 2;
            const float4 *src4 = reinterpret_cast<const float4 *>(src);
            float4 *dst4 = reinterpret_cast<float4 *>(s_pts);
            for (uint32_t k = tid; k < nv; k += K1_THREADS) dst4[k] = __ldcs(src4 + k);
            for (uint32_t k = (nv << 2) + tid; k < nf; k += K1_THREADS) s_pts[k] = __ldcs(src + k);
        } else {
            for (uint32_t k = tid; k < nf; k += K1_THREADS) s_pts[k] = __ldcs(src + k);
        }
    }
    __syncthreads();

    const uint32_t i = tile_base + tid;
    const bool live = tid < n_tile;
    bool ok = live;
    uint32_t cell = 0;
    int b = 0;
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
            const float cf = floorf(__fdiv_rn(__fsub_rn(q[j], p.lo[j]), p.vs[j]));
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
        b = pv_frame_of(p.offsets, p.B, i);
    }

    // ---- warp-aggregated insert: one lane per distinct (frame, cell) issues the atomics ----
    const unsigned lane = tid & 31u;
    const unsigned long long key = ok ? (((unsigned long long)b << 32) | cell)
                                      : (0xFFFFFFFF00000000ull | lane);
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    uint32_t s = PV_INF;
    if (ok && (int)lane == leader) {
        PvEntry *tab = p.ws.table + (size_t)b * p.ws.capf;
        uint32_t local;
        if (DENSE) local = cell;
        else local = pv_claim(tab, p.ws.capf - 1, cell, p.ws.ctrl + 1);
        if (local != PV_INF) {
            atomicMin(&tab[local].first, i);             // lanes are in index order: leader is the min
            atomicAdd(&tab[local].cnt, (uint32_t)__popc(peers));
            s = (uint32_t)b * p.ws.capf + local;
        }
    }
    s = __shfl_sync(0xffffffffu, s, leader);
    if (live) p.ws.slot[i] = ok ? s : PV_INF;
}

// ---------------------------------------------------------------------------------------------
// K2 -- single-pass scan with decoupled look-back
// ---------------------------------------------------------------------------------------------
#define PV_FLAG_AGG (1ull << 62)
#define PV_FLAG_PREFIX (2ull << 62)
#define PV_FLAG_MASK (3ull << 62)

__global__ void __launch_bounds__(K2_THREADS) k_rank_scan(const __grid_constant__ PvParams p)
{
    __shared__ unsigned long long s_warp[K2_THREADS / 32];
    __shared__ unsigned long long s_excl[K2_TILE];
    __shared__ unsigned long long s_tile_excl;
    __shared__ uint32_t s_tile;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) s_tile = atomicAdd(p.ws.ctrl, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t tile_base = tile * K2_TILE;
    const uint32_t i0 = tile_base + tid * K2_ITEMS;

    uint32_t sl[K2_ITEMS];
    if (i0 + K2_ITEMS <= p.n) {
        const uint4 v = *reinterpret_cast<const uint4 *>(p.ws.slot + i0);
        sl[0] = v.x; sl[1] = v.y; sl[2] = v.z; sl[3] = v.w;
    } else {
#pragma unroll
        for (int j = 0; j < K2_ITEMS; ++j) sl[j] = (i0 + j < p.n) ? p.ws.slot[i0 + j] : PV_INF;
    }
    unsigned long long val[K2_ITEMS];
    uint4 ent[K2_ITEMS];
#pragma unroll
    for (int j = 0; j < K2_ITEMS; ++j)
        ent[j] = (sl[j] != PV_INF) ? pv_ld_entry(p.ws.table + sl[j]) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < K2_ITEMS; ++j) {
        const bool first = (sl[j] != PV_INF) && (ent[j].y == i0 + j);
        const uint32_t L = min(ent[j].z + 1u, (uint32_t)p.T);
        val[j] = first ? pv_pack(1u, L) : 0ull;
    }
    unsigned long long tsum = 0;
#pragma unroll
    for (int j = 0; j < K2_ITEMS; ++j) tsum += val[j];

    // block-wide exclusive scan of tsum
    const unsigned lane = tid & 31u, warp = tid >> 5;
    unsigned long long incl = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (unsigned)d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long warp_off = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < K2_THREADS / 32; ++w) {
        const unsigned long long t = s_warp[w];
        if ((unsigned)w < warp) warp_off += t;
        block_total += t;
    }
    unsigned long long excl = warp_off + incl - tsum;

    // decoupled look-back (warp 0)
    if (warp == 0) {
        if (lane == 0)
            pv_st_volatile64(p.ws.tile_state + tile,
                             (tile == 0 ? PV_FLAG_PREFIX : PV_FLAG_AGG) | block_total);
        unsigned long long run = 0;
        if (tile > 0) {
            int pred = (int)tile - 1 - (int)lane;
            while (true) {
                unsigned long long w = PV_FLAG_PREFIX;  // virtual tile -1: prefix 0
                if (pred >= 0) {
                    do { w = pv_ld_volatile64(p.ws.tile_state + pred); } while ((w & PV_FLAG_MASK) == 0);
                }
                const unsigned pm = __ballot_sync(0xffffffffu, (w & PV_FLAG_MASK) == PV_FLAG_PREFIX);
                const int firstp = pm ? (__ffs(pm) - 1) : 32;
                unsigned long long c = ((int)lane <= firstp) ? (w & ~PV_FLAG_MASK) : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
                run += c;
                if (pm) break;
                pred -= 32;
            }
            if (lane == 0)
                pv_st_volatile64(p.ws.tile_state + tile, PV_FLAG_PREFIX | (run + block_total));
        }
        if (lane == 0) s_tile_excl = run;
    }
    __syncthreads();
    excl += s_tile_excl;

#pragma unroll
    for (int j = 0; j < K2_ITEMS; ++j) {
        s_excl[tid * K2_ITEMS + j] = excl;
        if (val[j]) {
            const uint32_t g = pv_rank(excl);
            p.ws.table[sl[j]].g = g;
            p.ws.vox_slot[g] = sl[j];
            p.ws.vox_koff[g] = pv_ksum(excl);
        }
        excl += val[j];
    }
    __syncthreads();
    // exclusive scan value at each frame start that falls inside this tile (offset == n lands in
    // the last tile, whose items past n contribute nothing).
    for (int b = tid; b <= p.B; b += K2_THREADS) {
        const uint32_t off = (uint32_t)p.offsets[b];
        if (off >= tile_base && off < tile_base + K2_TILE) p.ws.frame_scan[b] = s_excl[off - tile_base];
    }
}

// ---------------------------------------------------------------------------------------------
// K3 -- per-voxel sorted lists of the T smallest point indices
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fill_lists(const __grid_constant__ PvParams p)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // per-frame voxel counts and output row bases (tiny, serial)
        int32_t acc = 0;
        for (int b = 0; b < p.B; ++b) {
            const uint32_t tot = pv_rank(p.ws.frame_scan[b + 1]) - pv_rank(p.ws.frame_scan[b]);
            const int32_t m = (int32_t)min(tot, (uint32_t)p.V);
            p.ws.base[b] = acc;
            p.voxel_counts[b] = m;
            acc += m;
        }
        p.ws.base[p.B] = acc;
    }
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const uint32_t s = p.ws.slot[i];
    if (s == PV_INF) return;
    const uint4 e = pv_ld_entry(p.ws.table + s);
    const uint32_t b = s / p.ws.capf;
    const uint32_t r = e.w - pv_rank(p.ws.frame_scan[b]);
    if (r >= (uint32_t)p.V) return;                       // voxel beyond max_voxels: dropped
    const uint32_t c = e.z + 1u;
    const uint32_t L = min(c, (uint32_t)p.T);
    uint32_t *list = p.ws.kept + p.ws.vox_koff[e.w];
    if (i == e.y) { list[0] = i; return; }                // rank 0 is known: the first point
    if (L < 2) return;
    if (c == 2) { list[1] = i; return; }
    if (c > (uint32_t)p.T) {
        // slots only ever decrease: a tail already below i can never admit i
        if (pv_ld_volatile(list + L - 1) < i) return;
    }
    uint32_t x = i;
    for (uint32_t k = 1; k < L; ++k) {
        const uint32_t old = atomicMin(list + k, x);
        if (old == PV_INF) break;       // took a free slot, nothing displaced
        x = max(old, x);                // carry the loser to the next slot
    }
}

// ---------------------------------------------------------------------------------------------
// K4 -- emit per-voxel outputs
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int pv_frame_of_row(const int32_t *base, int B, int32_t row)
{
    int lo = 0, hi = B;  // base[lo] <= row < base[hi]; frames may be empty
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (base[mid] <= row) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_emit(const __grid_constant__ PvParams p)
{
    const int32_t vid = blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t total = p.ws.base[p.B];
    if (vid >= total) return;
    const int b = pv_frame_of_row(p.ws.base, p.B, vid);
    const uint32_t g = pv_rank(p.ws.frame_scan[b]) + (uint32_t)(vid - p.ws.base[b]);
    const uint32_t s = p.ws.vox_slot[g];
    const uint4 e = pv_ld_entry(p.ws.table + s);
    const uint32_t cell = p.ws.dense ? (s - (uint32_t)b * p.ws.capf) : e.x;
    const uint32_t nx = p.grid[0], ny = p.grid[1];
    const uint32_t x = cell % nx, yz = cell / nx;
    const uint32_t y = yz % ny, z = yz / ny;
    reinterpret_cast<int4 *>(p.coors)[vid] = make_int4(b, (int)z, (int)y, (int)x);
    const uint32_t c = e.z + 1u;
    const uint32_t L = min(c, (uint32_t)p.T);
    p.num_points[vid] = (int32_t)L;
    if (p.density) p.density[(size_t)b * p.cells + cell] = (int32_t)c;   // :70-71 un-capped count
    if (p.feats) {
        // VoxelFeatureExtractorV3 (voxel_encoder.py:18-22): sum in slot order / num_points
        const uint32_t *list = p.ws.kept + p.ws.vox_koff[g];
        float acc[PV_MAX_CHANNELS];
#pragma unroll
        for (int k = 0; k < PV_MAX_CHANNELS; ++k) acc[k] = 0.0f;
        for (uint32_t j = 0; j < L; ++j) {
            const float *row = p.pts + (size_t)list[j] * p.c_in;
            float in[PV_MAX_CHANNELS];
#pragma unroll
            for (int k = 0; k < PV_MAX_CHANNELS; ++k) in[k] = (k < p.c_in) ? __ldg(row + k) : 0.0f;
            if (p.cart) {
                acc[0] = __fadd_rn(acc[0], pv_rho(in[0], in[1]));
                acc[1] = __fadd_rn(acc[1], pv_atan2f(in[1], in[0]));
                acc[2] = __fadd_rn(acc[2], in[2]);
                acc[3] = __fadd_rn(acc[3], in[0]);
                acc[4] = __fadd_rn(acc[4], in[1]);
#pragma unroll
                for (int k = 5; k < PV_MAX_CHANNELS; ++k) acc[k] = __fadd_rn(acc[k], in[k - 2]);
            } else {
#pragma unroll
                for (int k = 0; k < PV_MAX_CHANNELS; ++k) acc[k] = __fadd_rn(acc[k], in[k]);
            }
        }
        const float nf = (float)L;
        float *o = p.feats + (size_t)vid * p.C;
#pragma unroll
        for (int k = 0; k < PV_MAX_CHANNELS; ++k)
            if (k < p.C) o[k] = __fdiv_rn(acc[k], nf);
    }
}

// Padded voxels tensor [SM, T, C] (point_cloud_ops.py:187,67): one thread per (voxel, slot) row.
__global__ void __launch_bounds__(256) k_emit_voxels(const __grid_constant__ PvParams p)
{
    const long long rowid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)p.ws.base[p.B] * p.T;
    if (rowid >= total) return;
    const int32_t vid = (int32_t)(rowid / p.T);
    const uint32_t t = (uint32_t)(rowid - (long long)vid * p.T);
    const int b = pv_frame_of_row(p.ws.base, p.B, vid);
    const uint32_t g = pv_rank(p.ws.frame_scan[b]) + (uint32_t)(vid - p.ws.base[b]);
    const uint4 e = pv_ld_entry(p.ws.table + p.ws.vox_slot[g]);
    const uint32_t L = min(e.z + 1u, (uint32_t)p.T);
    float out[PV_MAX_CHANNELS];
#pragma unroll
    for (int k = 0; k < PV_MAX_CHANNELS; ++k) out[k] = 0.0f;
    if (t < L) {
        const uint32_t i = p.ws.kept[p.ws.vox_koff[g] + t];
        const float *row = p.pts + (size_t)i * p.c_in;
        float in[PV_MAX_CHANNELS];
#pragma unroll
        for (int k = 0; k < PV_MAX_CHANNELS; ++k) in[k] = (k < p.c_in) ? __ldg(row + k) : 0.0f;
        if (p.cart) {
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
    float *o = p.voxels + (size_t)rowid * p.C;
#pragma unroll
    for (int k = 0; k < PV_MAX_CHANNELS; ++k)
        if (k < p.C) o[k] = out[k];
}

// ---------------------------------------------------------------------------------------------
// K5 -- dense BEV canvas for pillar grids (PointPillarsScatter, pillar_encoder.py:189-225):
// one thread per 4 consecutive x cells; looks the cell up in the voxel map, writes the voxel's
// feature row or zeros, so the canvas is written exactly once and needs no separate zero fill.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int32_t pv_lookup_row(const PvParams &p, int b, uint32_t cell)
{
    const PvEntry *tab = p.ws.table + (size_t)b * p.ws.capf;
    uint4 e;
    if (p.ws.dense) {
        e = pv_ld_entry(tab + cell);
        if (e.y == PV_INF) return -1;
    } else {
        const uint32_t mask = p.ws.capf - 1;
        uint32_t h = pv_hash(cell) & mask;
        while (true) {
            e = pv_ld_entry(tab + h);
            if (e.x == cell) break;
            if (e.x == PV_INF) return -1;
            h = (h + 1) & mask;
        }
    }
    const uint32_t r = e.w - pv_rank(p.ws.frame_scan[b]);
    if (r >= (uint32_t)p.V) return -1;
    return p.ws.base[b] + (int32_t)r;
}

__global__ void __launch_bounds__(256) k_canvas(const __grid_constant__ PvParams p)
{
    const uint32_t cells = p.cells;               // nz == 1: cells = ny * nx
    const uint32_t quads = cells >> 2;            // host guarantees nx % 4 == 0
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (q >= quads) return;
    const uint32_t cell0 = q << 2;
    int32_t row[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) row[k] = pv_lookup_row(p, b, cell0 + k);
    float *dst = p.canvas + (size_t)b * p.C * cells + cell0;
    for (int ch = 0; ch < p.C; ++ch) {
        float4 v;
        v.x = row[0] >= 0 ? p.feats[(size_t)row[0] * p.C + ch] : 0.0f;
        v.y = row[1] >= 0 ? p.feats[(size_t)row[1] * p.C + ch] : 0.0f;
        v.z = row[2] >= 0 ? p.feats[(size_t)row[2] * p.C + ch] : 0.0f;
        v.w = row[3] >= 0 ? p.feats[(size_t)row[3] * p.C + ch] : 0.0f;
        __stcs(reinterpret_cast<float4 *>(dst + (size_t)ch * cells), v);
    }
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
    }
    if (cells >= (1ll << 31)) return PV_ERR_BAD_CONFIG;
    if (cfg->max_points <= 0 || cfg->max_voxels <= 0) return PV_ERR_BAD_CONFIG;
    return PV_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int pv_make_layout(const pv_config *cfg, int64_t n_cap, int32_t batch, int64_t frame_capacity,
                   void *base, PvWs *w)
{
    int rc = pv_check_config(cfg);
    if (rc) return rc;
    if (batch <= 0 || n_cap < 0 || frame_capacity < 0) return PV_ERR_BAD_ARGUMENT;
    if (n_cap >= (1ll << 31) - K2_TILE) return PV_ERR_BAD_ARGUMENT;
    if (frame_capacity > n_cap) frame_capacity = n_cap;
    const uint64_t cells = (uint64_t)cfg->grid[0] * cfg->grid[1] * cfg->grid[2];
    uint64_t capf;
    const bool dense = cells <= PV_DENSE_MAX_CELLS;
    if (dense) capf = cells;
    else {
        uint64_t want = (uint64_t)frame_capacity + (uint64_t)frame_capacity / 4 + 1;
        capf = 1024;
        while (capf < want) capf <<= 1;
    }
    if (capf * (uint64_t)batch >= 0xFFFFFFFFull) return PV_ERR_BAD_ARGUMENT;
    const size_t n = (size_t)(n_cap > 0 ? n_cap : 1);
    w->capf = (uint32_t)capf;
    w->dense = dense ? 1u : 0u;
    w->num_tiles = (uint32_t)(n_cap / K2_TILE + 1);
    char *p0 = (char *)base;
    size_t o = 0;
    w->zero_begin = p0 + o;
    w->ctrl = (uint32_t *)(p0 + o);                      o = align_up(o + 16 * sizeof(uint32_t), 256);
    w->frame_scan = (unsigned long long *)(p0 + o);      o = align_up(o + (size_t)(batch + 1) * 8, 256);
    w->base = (int32_t *)(p0 + o);                       o = align_up(o + (size_t)(batch + 1) * 4, 256);
    w->tile_state = (unsigned long long *)(p0 + o);      o = align_up(o + (size_t)w->num_tiles * 8, 256);
    w->zero_bytes = o;
    w->ff_begin = p0 + o;
    w->table = (PvEntry *)(p0 + o);                      o = align_up(o + (size_t)capf * batch * sizeof(PvEntry), 256);
    w->kept = (uint32_t *)(p0 + o);                      o = align_up(o + n * 4, 256);
    w->ff_bytes = o - w->zero_bytes;
    w->slot = (uint32_t *)(p0 + o);                      o = align_up(o + n * 4 + 16, 256);
    w->vox_slot = (uint32_t *)(p0 + o);                  o = align_up(o + n * 4, 256);
    w->vox_koff = (uint32_t *)(p0 + o);                  o = align_up(o + n * 4, 256);
    w->total_bytes = o;
    return PV_OK;
}

static int fill_params(PvParams *p, const pv_config *cfg, const float *points,
                       const int32_t *frame_offsets, int32_t batch, int64_t n_total, int32_t c_in,
                       int32_t is_cartesian, int64_t frame_capacity, void *workspace,
                       size_t workspace_bytes)
{
    if (!points && n_total > 0) return PV_ERR_BAD_ARGUMENT;
    if (!frame_offsets || !workspace) return PV_ERR_BAD_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return PV_ERR_BAD_ARGUMENT;
    const int C = is_cartesian ? c_in + 2 : c_in;
    if (c_in < 3 || C > PV_MAX_CHANNELS) return PV_ERR_BAD_ARGUMENT;
    int rc = pv_make_layout(cfg, n_total, batch, frame_capacity, workspace, &p->ws);
    if (rc) return rc;
    if (p->ws.total_bytes > workspace_bytes) return PV_ERR_WORKSPACE;
    for (int j = 0; j < 3; ++j) {
        p->lo[j] = cfg->lo[j]; p->vs[j] = cfg->vs[j]; p->grid[j] = cfg->grid[j];
        p->gridf[j] = (float)cfg->grid[j];
    }
    p->T = cfg->max_points; p->V = cfg->max_voxels;
    p->pts = points; p->offsets = frame_offsets; p->B = batch; p->n = (uint32_t)n_total;
    p->c_in = c_in; p->cart = is_cartesian ? 1 : 0; p->C = C;
    p->cells = (uint32_t)((uint64_t)cfg->grid[0] * cfg->grid[1] * cfg->grid[2]);
    p->coors = p->num_points = p->voxel_counts = p->grid_ind = p->density = nullptr;
    p->voxels = p->feats = p->canvas = nullptr;
    return PV_OK;
}

static int run_voxelize(PvParams &p, cudaStream_t st)
{
    const PvWs &w = p.ws;
    if (cudaMemsetAsync(w.zero_begin, 0, w.zero_bytes, st) != cudaSuccess) return PV_ERR_CUDA;
    if (cudaMemsetAsync(w.ff_begin, 0xFF, w.ff_bytes, st) != cudaSuccess) return PV_ERR_CUDA;
    if (p.density &&
        cudaMemsetAsync(p.density, 0, (size_t)p.B * p.cells * sizeof(int32_t), st) != cudaSuccess)
        return PV_ERR_CUDA;
    if (p.n > 0) {
        const unsigned g1 = (p.n + K1_THREADS - 1) / K1_THREADS;
        if (w.dense) k_bin_insert<true><<<g1, K1_THREADS, 0, st>>>(p);
        else k_bin_insert<false><<<g1, K1_THREADS, 0, st>>>(p);
    }
    k_rank_scan<<<p.n / K2_TILE + 1, K2_THREADS, 0, st>>>(p);
    k_fill_lists<<<(p.n + 255) / 256 + (p.n == 0), 256, 0, st>>>(p);
    const long long rows_cap = min((long long)p.B * p.V, (long long)p.n);
    if (rows_cap > 0) {
        k_emit<<<(unsigned)((rows_cap + 255) / 256), 256, 0, st>>>(p);
        if (p.voxels) {
            const long long tr = rows_cap * p.T;
            k_emit_voxels<<<(unsigned)((tr + 255) / 256), 256, 0, st>>>(p);
        }
    }
    return pv_last_cuda_error();
}

extern "C" {

int pv_version(void) { return 100; }

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
    default: return "unknown error";
    }
}

size_t pv_workspace_bytes(const pv_config *cfg, int64_t max_points_total, int32_t batch,
                          int64_t frame_capacity)
{
    PvWs w;
    if (pv_make_layout(cfg, max_points_total, batch, frame_capacity, nullptr, &w) != PV_OK) return 0;
    return w.total_bytes;
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
                int64_t frame_capacity, void *workspace, size_t workspace_bytes, int32_t *coors,
                int32_t *num_points, int32_t *voxel_counts, float *voxels, float *mean_feats,
                int32_t *pc_grid_ind, int32_t *density, pv_stream_t stream)
{
    PvParams p;
    int rc = fill_params(&p, cfg, points, frame_offsets, batch, n_total, c_in, is_cartesian,
                         frame_capacity, workspace, workspace_bytes);
    if (rc) return rc;
    if (!coors || !num_points || !voxel_counts) return PV_ERR_BAD_ARGUMENT;
    p.coors = coors; p.num_points = num_points; p.voxel_counts = voxel_counts;
    p.voxels = voxels; p.feats = mean_feats; p.grid_ind = pc_grid_ind; p.density = density;
    return run_voxelize(p, (cudaStream_t)stream);
}

int pv_forward_mean_canvas(const pv_config *cfg, const float *points, const int32_t *frame_offsets,
                           int32_t batch, int64_t n_total, int32_t c_in, int32_t is_cartesian,
                           int64_t frame_capacity, void *workspace, size_t workspace_bytes,
                           int32_t *coors, int32_t *num_points, int32_t *voxel_counts,
                           float *mean_feats, float *canvas, pv_stream_t stream)
{
    PvParams p;
    int rc = fill_params(&p, cfg, points, frame_offsets, batch, n_total, c_in, is_cartesian,
                         frame_capacity, workspace, workspace_bytes);
    if (rc) return rc;
    if (!coors || !num_points || !voxel_counts || !mean_feats || !canvas) return PV_ERR_BAD_ARGUMENT;
    if (cfg->grid[2] != 1 || (cfg->grid[0] & 3) != 0) return PV_ERR_BAD_CONFIG;
    if ((reinterpret_cast<uintptr_t>(canvas) & 15u) != 0) return PV_ERR_BAD_ARGUMENT;
    p.coors = coors; p.num_points = num_points; p.voxel_counts = voxel_counts;
    p.feats = mean_feats; p.canvas = canvas;
    rc = run_voxelize(p, (cudaStream_t)stream);
    if (rc) return rc;
    const unsigned quads = p.cells >> 2;
    dim3 grid((quads + 255) / 256, (unsigned)batch);
    k_canvas<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    return pv_last_cuda_error();
}

int pv_read_status(const void *workspace, pv_stream_t stream)
{
    if (!workspace) return PV_ERR_BAD_ARGUMENT;
    uint32_t ctrl[2] = {0, 0};
    if (cudaMemcpyAsync(ctrl, workspace, sizeof(ctrl), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess)
        return PV_ERR_CUDA;
    if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return PV_ERR_CUDA;
    return (ctrl[1] & 1u) ? PV_ERR_TABLE_FULL : PV_OK;
}

}  // extern "C"
