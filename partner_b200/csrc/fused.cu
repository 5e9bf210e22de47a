// fused.cu -- list-free polar front end (sm_100a): voxelize + mean VFE + BEV canvas without ever
// materialising per-voxel point lists.
//
// Same contract as the list-based pipeline of voxelize.cu (point_cloud_ops.py:7-72 through
// VoxelGenerator.generate, voxel_encoder.py:15-22, pillar_encoder.py:189-225): voxel order =
// first-occurrence order, kept points = the T smallest indices of the cell, cells of rank >= V
// dropped.  Integer outputs are bit-exact; the mean is a sum of the same addends in a different
// order (fp32 reductions at L2), inside the 1e-5 tolerance of the north star.
//
// Measured on B200 (tools/probes/atomics_probe.cu): a scattered 4-byte load, store or reduction
// costs one LSU slot per lane (~1.3 cycles per lane per SM, ~225 G lane-ops/s per GPU), and a
// 16-byte vector reduction costs the same as a 4-byte one.  The design therefore minimises
// SCATTERED OPERATIONS PER POINT:
//
//   F1 insert      per point: bin -> cell slot (direct maps are PHI FASTEST: consecutive points of a
//                  LiDAR ring are azimuth neighbours, so the lanes of a warp reduce into rows that
//                  share 128-byte lines); a RETURNING atomicMin(first[slot], i) and one 16-byte
//                  RED.ADD.F32x4 per 4 floats of the row {C feature sums, count}.  The returned
//                  minimum feeds the first-point bitmap (bit i <=> point i is its cell's first
//                  point): own bits leave as one aggregated reduction per 32 points, a displaced
//                  earlier minimum is toggled back; XOR commutes, so the bitmap is exact when the
//                  grid has drained.  Rows staged with one TMA bulk copy per tile; consecutive
//                  points of one thread that share a cell are merged first.
//   F2 scan        popcount scan over the first-point bitmap (N/32 words) -> rank of every first
//                  point in its frame = first-occurrence rank of its cell; voxel counts, row bases.
//   F3 finalize    one block per 32 x 32 patch of cells: occupied cell -> rank lookup, mean = sum /
//                  min(n, T), coors / num_points / features rows (azimuth neighbours have
//                  consecutive ranks, so a warp's rows are neighbours); the BEV canvas (and
//                  density) goes through a shared-memory transpose and is written IN CELL ORDER,
//                  every element exactly once, zeros included -- no zero fill, no scatter; the map
//                  is restored to its clean state on the way.
//   F4 heavy       (rare cells, 2-5 % of the points) points of heavy cells append their index to
//                  the cell's candidate range; one warp per heavy cell selects the T smallest
//                  (REDUX.MIN rounds) and re-sums exactly those rows.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#include "pv_common.cuh"

#define PF_THREADS 256
#define PF_PPT 4                          // consecutive points per thread in the point passes
#define PF_TILE (PF_THREADS * PF_PPT)
#ifndef PF_INSERT_MIN_BLOCKS
#define PF_INSERT_MIN_BLOCKS 4
#endif
#define PF_SCAN_THREADS 1024
#define PF_SCAN_VEC 4                    // 4-word vectors per thread per scan iteration

// ---------------------------------------------------------------------------------------------
// small PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pf_red_add_v4(float *p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// the same with an L2 eviction policy (accumulator rows are read again by the finalize kernel)
__device__ __forceinline__ void pf_red_add_v4(float *p, float a, float b, float c, float d, unsigned long long pol)
{
    asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                 ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "l"(pol) : "memory");
}
__device__ __forceinline__ unsigned long long pf_policy_evict_first()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint32_t pf_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pf_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void pf_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pf_mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// TMA bulk copy global -> shared (1-D, 16-byte granules), completion on an mbarrier
__device__ __forceinline__ void pf_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    // the points are read once: do not let them push the map out of L2
    const unsigned long long pol = pf_policy_evict_first();
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}

// Programmatic dependent launch (PDL): the kernels of one call form a chain; each is launched with
// programmatic stream serialization, lets its successor's blocks be scheduled while its own last
// wave is still running (pf_pdl_trigger) and waits for its predecessor to have completed and
// flushed (pf_pdl_wait) before it touches anything the predecessor wrote.  Hides the launch
// latency and the block ramp at every kernel boundary of a single-stream step.
__device__ __forceinline__ void pf_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pf_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Hash-mode slot claim (same protocol as voxelize.cu): keys[] = cell, clean = INF.
__device__ __forceinline__ uint32_t pf_claim(uint32_t *keys, uint32_t mask, uint32_t cell, uint32_t h, uint32_t *status)
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

// ---------------------------------------------------------------------------------------------
// F1 -- insert.  CIN > 0: compile-time row width and transform flag (rows read from shared memory
// with 128-bit loads); CIN == 0: runtime c_in / cart.  NV = 16-byte reductions per row.
// ---------------------------------------------------------------------------------------------
// DYN: dynamic voxelization (voxelization.py:169-172): bins are CLAMPED into the grid instead of
// range-tested, every point lands in a cell; the cell's bit in the cell-order occupancy bitmap
// replaces the first-point minimum (voxels are ordered by cell, not by first occurrence).
// MODE 2 (lists): the same binning front end for the list-based pipeline of voxelize.cu: per run
// RED.MIN(first) + RED.ADD(count) on its {first, cnt} map, per point the map slot, the cleared scan
// word and (hash maps) the cell index; no accumulator rows.
#define PF_MODE_FREE 0
#define PF_MODE_DYN 1
#define PF_MODE_LISTS 2
template <bool DENSE, int CIN, bool CART, int NV, int MODE = PF_MODE_FREE, bool GI = true>
__global__ void __launch_bounds__(PF_THREADS, PF_INSERT_MIN_BLOCKS) kf_insert(const __grid_constant__ PvParams p,
                                                        const __grid_constant__ PvF f)
{
    pf_pdl_trigger();
    extern __shared__ __align__(128) float s_pts[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ int s_b0;
    __shared__ uint32_t s_next;
    constexpr int CT = NV * 4;
    constexpr bool DYN = MODE == PF_MODE_DYN, LISTS = MODE == PF_MODE_LISTS;
    const uint32_t tid = threadIdx.x;
    const int c_in = CIN ? CIN : p.c_in;
    const bool cart = CIN ? CART : (p.cart != 0);
    constexpr int CS = CIN ? CIN + (CART ? 2 : 0) : 0;      // compile-time channel count (0 = runtime)
    const int C = CS ? CS : p.C;
    const uint32_t tile_base = blockIdx.x * PF_TILE;
    const uint32_t n_tile = min((uint32_t)PF_TILE, p.n - tile_base);
    const uint32_t nf = n_tile * (uint32_t)c_in;
    const float *src = p.pts + (size_t)tile_base * c_in;

    // ---- stage the tile's rows: one TMA bulk copy (16-byte granules) + a scalar tail ----
    const bool bulk = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && nf >= 4u;
    const uint32_t bulk_bytes = bulk ? ((nf * 4u) & ~15u) : 0u;
    const uint32_t bar = pf_smem_addr(&s_bar);
    if (tid == 0 && bulk) {
        pf_mbar_init(bar, 1);
        pf_mbar_expect_tx(bar, bulk_bytes);
        pf_bulk_g2s(pf_smem_addr(s_pts), src, bulk_bytes, bar);
    }
    if (tid >= 32 && tid < 64) {             // frame of the tile's first point, while the copy is in flight
        const uint32_t lane = tid & 31u;
        int cnt = 0;                         // = #{b >= 1 : offsets[b] <= tile_base}
        uint32_t nxt = PV_INF;               // smallest offsets[b] > tile_base
        for (int bb = 1 + (int)lane; bb < p.B && p.offsets; bb += 32) {   // offsets == NULL: frames come from gi_in
            const uint32_t o = (uint32_t)__ldg(p.offsets + bb);
            if (o <= tile_base) ++cnt; else nxt = min(nxt, o);
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        nxt = __reduce_min_sync(0xffffffffu, nxt);
        if (lane == 0) { s_b0 = cnt; s_next = nxt; }
    }
    for (uint32_t k = (bulk_bytes >> 2) + tid; k < nf; k += PF_THREADS) s_pts[k] = __ldg(src + k);
    __syncthreads();                        // barrier initialised + tail visible
    if (bulk) pf_mbar_wait(bar, 0);

    const uint32_t t0 = tid * PF_PPT;       // first point of this thread inside the tile
    float rows[CIN ? PF_PPT * CIN : 1];      // this thread's 4 rows = CIN consecutive 16-byte words
    if constexpr (CIN > 0) {
        const float4 *q4 = reinterpret_cast<const float4 *>(s_pts) + tid * CIN;
#pragma unroll
        for (int k = 0; k < CIN; ++k) {
            const float4 v = q4[k];
            rows[4 * k] = v.x; rows[4 * k + 1] = v.y; rows[4 * k + 2] = v.z; rows[4 * k + 3] = v.w;
        }
    }

    int b = s_b0;
    uint32_t next_off = s_next;              // first point index of the next non-empty frame boundary
    const float lo0 = p.lo[0], lo1 = p.lo[1], lo2 = p.lo[2];
    const float iv0 = p.inv_vs[0], iv1 = p.inv_vs[1], iv2 = p.inv_vs[2];
    const float g0 = p.gridf[0], g1 = p.gridf[1], g2 = p.gridf[2];
    const uint32_t nx = (uint32_t)p.grid[0], ny = (uint32_t)p.grid[1];
    uint32_t cur_s = PV_INF, cur_i = 0;      // current run of consecutive points in one cell
    float cur[CT], cur_n = 0.0f;             // its feature sums and point count
    uint32_t sa_out[PF_PPT];
    // First-point bitmap (bit i <=> point i is its cell's first point), built on the fly: the
    // minimum is taken with a RETURNING atomic; a run that became its cell's minimum (old > i)
    // marks its own bit and, if it displaced an earlier minimum, toggles that one's bit.  Every bit
    // is toggled at most twice -- once by its owner, once by the run that displaced it -- and XOR
    // commutes, so when the grid has drained exactly the final first points are set, whatever the
    // arrival order.  The returned values are only looked at when the thread has issued all its
    // runs (one wait for up to four round trips in flight, not one per run); own bits are
    // collected per thread and leave as ONE reduction per 32 points (8 lanes x 4 points = one
    // bitmap word).
    uint32_t olds[PF_PPT];                   // per flush site: what the minimum was before (0: site unused)
    uint32_t starts = 0;                     // 2 bits per site: the run's first point, relative to t0
#pragma unroll
    for (int j = 0; j < PF_PPT; ++j) olds[j] = 0u;
    auto flush = [&](const int site) {
        if (cur_s != PV_INF) {
            if constexpr (LISTS) {           // list-based map entry {first, count - 1}
                atomicMin(&p.ws.table[cur_s].first, cur_i);
                atomicAdd(&p.ws.table[cur_s].cnt, (uint32_t)cur_n);
                return;
            }
            if constexpr (DYN) atomicOr(f.bits + (cur_i >> 5), 1u << (cur_i & 31u));   // cur_i = the cell's bit address
            else {
                // straight into the site's register: a copy of the result would wait for the round trip
                olds[site] = atomicMin(f.first + cur_s, cur_i);
                starts |= (cur_i - (tile_base + t0)) << (2 * site);
            }
            float o[CT];                     // the count rides in channel C of the row
#pragma unroll
            for (int k = 0; k < CT; ++k) o[k] = k < C ? cur[k] : (k == C ? cur_n : 0.0f);
            float *row = f.acc + (size_t)cur_s * f.rowf;
            // rows are read again by the finalize kernel: ask L2 to keep them (insert 38.2 -> 36.8 us,
            // finalize 33.0 -> 31.7 us together with the evict-first point tiles)
            const unsigned long long keep = pv_policy_evict_last();
#pragma unroll
            for (int q = 0; q < NV; ++q) pf_red_add_v4(row + 4 * q, o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3], keep);
        }
    };
#pragma unroll
    for (int j = 0; j < PF_PPT; ++j) {
        sa_out[j] = PV_INF;
        const uint32_t i = tile_base + t0 + j;
        if (t0 + j >= n_tile) continue;
        float in[CT];                        // the raw row, zero padded
#pragma unroll
        for (int k = 0; k < CT; ++k) {
            if constexpr (CIN > 0) in[k] = k < CIN ? rows[j * CIN + (k < CIN ? k : 0)] : 0.0f;
            else in[k] = k < c_in ? s_pts[(t0 + j) * c_in + k] : 0.0f;
        }
        float v[CT];                         // the feature row the reference voxelizes
        if (cart) {                          // utils.py:42-44: (rho, phi, z, x, y, feat3..)
#pragma unroll
            for (int k = 0; k < CT; ++k) v[k] = k >= 5 ? in[k >= 5 ? k - 2 : 0] : 0.0f;
            v[0] = pv_rho(in[0], in[1]);
            if (CT > 1) v[1] = pv_atan2f(in[1], in[0]);
            if (CT > 2) v[2] = in[2];
            if (CT > 3) v[3] = in[0];
            if (CT > 4) v[CT > 4 ? 4 : 0] = in[1];
        } else {
#pragma unroll
            for (int k = 0; k < CT; ++k) v[k] = in[k];
        }
        // point_cloud_ops.py:45 -- floor of the float32 quotient, via pv_bin's reciprocal fast path;
        // the three near-integer tests share one branch
        const float t0f = __fsub_rn(v[0], lo0), t1f = __fsub_rn(v[CT > 1 ? 1 : 0], lo1), t2f = __fsub_rn(v[CT > 2 ? 2 : 0], lo2);
        const float r0 = __fmul_rn(t0f, iv0), r1 = __fmul_rn(t1f, iv1), r2 = __fmul_rn(t2f, iv2);
        float c0 = floorf(r0), c1 = floorf(r1), c2 = floorf(r2);
        const bool n0 = pv_bin_unsure(r0, c0), n1 = pv_bin_unsure(r1, c1), n2 = pv_bin_unsure(r2, c2);
        if (n0 | n1 | n2) {
            if (n0) c0 = floorf(__fdiv_rn(t0f, p.vs[0]));
            if (n1) c1 = floorf(__fdiv_rn(t1f, p.vs[1]));
            if (n2) c2 = floorf(__fdiv_rn(t2f, p.vs[2]));
        }
        bool ok = c0 >= 0.0f && c0 < g0 && c1 >= 0.0f && c1 < g1 && c2 >= 0.0f && c2 < g2;   // NaN fails
        while (i >= next_off) {              // crossed into the next frame (frames may be empty)
            ++b;
            next_off = b + 1 < p.B ? (uint32_t)__ldg(p.offsets + b + 1) : PV_INF;
        }
        if constexpr (DYN) {                 // voxelization.py:170: clip(q, 0, grid - 1) before the floor (NaN -> 0)
            c0 = fminf(fmaxf(c0, 0.0f), g0 - 1.0f); c1 = fminf(fmaxf(c1, 0.0f), g1 - 1.0f); c2 = fminf(fmaxf(c2, 0.0f), g2 - 1.0f);
            ok = true;
            if (p.gi_in) {                   // drop-in reader: the caller's (b, z, y, x) rows replace the binning
                const int4 g = __ldg(reinterpret_cast<const int4 *>(p.gi_in) + i);
                b = g.x; c2 = (float)g.y; c1 = (float)g.z; c0 = (float)g.w;
                ok = (unsigned)g.x < (unsigned)p.B && (unsigned)g.y < (unsigned)p.grid[2] && (unsigned)g.z < ny && (unsigned)g.w < nx;
                if (!ok) atomicOr(p.ws.ctrl + 1, 2u);
            }
            if (p.grid_ind) reinterpret_cast<int4 *>(p.grid_ind)[i] = make_int4(b, (int)c2, (int)c1, (int)c0);
        } else if (GI && p.grid_ind) {       // :46-54 clamped (z, y, x) for every point (NaN -> 0)
            int32_t *gi = p.grid_ind + (size_t)i * 3;
            gi[0] = (int)fminf(fmaxf(c2, 0.0f), g2 - 1.0f);
            gi[1] = (int)fminf(fmaxf(c1, 0.0f), g1 - 1.0f);
            gi[2] = (int)fminf(fmaxf(c0, 0.0f), g0 - 1.0f);
        }
        if (!ok) continue;
        const uint32_t cx = (uint32_t)(int)c0, cy = (uint32_t)(int)c1, cz = (uint32_t)(int)c2;
        const uint32_t cell = (cz * ny + cy) * nx + cx;
        uint32_t s, sa;
        if (LISTS) {
            if (DENSE) s = (uint32_t)b * p.ws.capf + (cz * nx + cx) * ny + cy;     // voxelize.cu's direct map: phi fastest
            else {
                const uint32_t h = pf_claim(p.ws.keys + (size_t)b * p.ws.capf, p.ws.capf - 1, cell,
                                          pv_slot_home(cx, cy, cz, nx, ny, p.ws.capf - 1), p.ws.ctrl + 1);
                if (h == PV_INF) continue;
                s = (uint32_t)b * p.ws.capf + h;
                p.ws.pcell[i] = cell;
            }
            sa = s;
        } else if (DYN) {
            if (DENSE) s = (uint32_t)b * f.capf + cell;
            else {                                               // 3-D grids: rows live in the hash map, the bitmap stays in cell order
                const uint32_t h = pf_claim(f.keys + (size_t)b * f.capf, f.capf - 1, cell,
                                          pv_slot_home(cx, cy, cz, nx, ny, f.capf - 1), p.ws.ctrl + 1);
                if (h == PV_INF) continue;
                s = (uint32_t)b * f.capf + h;
            }
            sa = (uint32_t)b * (f.wcap * 32u) + cell;            // bit address in the occupancy bitmap
        } else if (DENSE) {
            // direct map, PHI FASTEST: consecutive points of a LiDAR ring are azimuth neighbours, so
            // the rows they reduce into (and the heavy-bitmap words) share 128-byte lines
            s = (uint32_t)b * f.capf + (cz * nx + cx) * ny + cy;
            sa = s;
        } else {
            const uint32_t h = pf_claim(f.keys + (size_t)b * f.capf, f.capf - 1, cell,
                                      pv_slot_home(cx, cy, cz, nx, ny, f.capf - 1), p.ws.ctrl + 1);
            if (h == PV_INF) continue;       // map full: status bit set
            s = (uint32_t)b * f.capf + h;
            sa = s;
        }
        sa_out[j] = sa;
        if (s == cur_s) {
            if constexpr (!LISTS) {
#pragma unroll
                for (int k = 0; k < CT; ++k)
                    if (k < C) cur[k] = __fadd_rn(cur[k], v[k]);
            }
            cur_n += 1.0f;
        } else {
            flush(j);                        // (never a run at j == 0: site 0 is the final flush's)
            cur_s = s; cur_i = DYN ? sa : i; cur_n = 1.0f;
#pragma unroll
            for (int k = 0; k < CT; ++k) cur[k] = v[k];
        }
    }
    flush(0);
    if (DYN && !p.unq_inv) return;           // the per-point map is only needed for the inverse index
    uint32_t *sa_dst = LISTS ? p.ws.slot : f.sa;
    if (t0 + PF_PPT <= n_tile) {
        __stcs(reinterpret_cast<uint4 *>(sa_dst + tile_base + t0), make_uint4(sa_out[0], sa_out[1], sa_out[2], sa_out[3]));
        if (LISTS) *reinterpret_cast<uint4 *>(p.ws.pv + tile_base + t0) = make_uint4(0u, 0u, 0u, 0u);
    } else {
#pragma unroll
        for (int j = 0; j < PF_PPT; ++j)
            if (t0 + j < n_tile) {
                sa_dst[tile_base + t0 + j] = sa_out[j];
                if (LISTS) p.ws.pv[tile_base + t0 + j] = 0u;
            }
    }
    if constexpr (MODE == PF_MODE_FREE) {
        uint32_t mine = 0;                   // bit j: point t0 + j became its cell's minimum
#pragma unroll
        for (int site = 0; site < PF_PPT; ++site) {
            const uint32_t j0 = (starts >> (2 * site)) & 3u, i0 = tile_base + t0 + j0, old = olds[site];
            if (old > i0) {
                mine |= 1u << j0;
                if (old != PV_INF) atomicXor(f.bits + (old >> 5), 1u << (old & 31u));
            }
        }
        uint32_t word = mine << (4u * (tid & 7u));
        word |= __shfl_xor_sync(0xffffffffu, word, 1);
        word |= __shfl_xor_sync(0xffffffffu, word, 2);
        word |= __shfl_xor_sync(0xffffffffu, word, 4);
        if ((tid & 7u) == 0 && word) atomicXor(f.bits + ((tile_base + t0) >> 5), word);
    }
}

// ---------------------------------------------------------------------------------------------
// F1, streaming form: kf_insert_lanes (direct maps, list-free, no pc_grid_ind -- the fused front end's hot path).
// Same contract as kf_insert<true, CIN, CART, NV, PF_MODE_FREE, false>, restructured around what the
// ncu captures of that kernel showed (issue slots 52 % busy, 220 instructions per point, 41 % of
// the stall samples on the returned minima, warps idle while their tile is in flight):
//   * every WARP is its own software pipeline over tiles of 128 consecutive points (32 lanes x 4):
//     a two-stage ring of 128 * CIN * 4-byte shared-memory buffers per warp, each filled by one TMA
//     bulk copy that the warp's lane 0 issues as soon as the previous occupant has been read into
//     registers -- no block-level barrier anywhere, so warps drift apart instead of bunching
//     their reductions; the load of tile k+1 is in flight while tile k is binned;
//   * the minima returned for tile k are consumed just before tile k+1 issues its own atomics
//     (same registers, no copy), i.e. a full tile of arithmetic after they were requested;
//   * the per-point path is branch free: bins through the reciprocal with a CONSTANT guard band
//     (3e-7 * (grid + 2) >= the 1.8e-7 |q| worst-case distance between t * fl(1 / vs) and the
//     correctly rounded quotient; inside the band, and for NaN / huge values, the IEEE division
//     decides -- one rare divergent branch per thread), range test on the converted integers,
//     runs of equal cells merged with selects / shuffles, atomics predicated inside the PTX.
// Persistent grid: 4 warps per block, 4 blocks per SM; warp g takes tiles g, g + G, g + 2G, ...
// ---------------------------------------------------------------------------------------------
#ifndef KI_WARPS
#define KI_WARPS 4
#endif
#ifndef KI_STAGES
#define KI_STAGES 2
#endif
#ifndef KI_BLOCKS_PER_SM
#define KI_BLOCKS_PER_SM 4
#endif
#ifndef KI_EXP
#define KI_EXP 0      // timing experiments (never defined in product builds)
#endif
#ifndef KF_EXP
#define KF_EXP 0
#endif

// predicated returning minimum: old stays what it was (0) when pred == 0
__device__ __forceinline__ void ki_atom_min(uint32_t &old, uint32_t *addr, uint32_t val, uint32_t pred)
{
    const unsigned long long pol = pv_policy_evict_last();    // first[] is part of the map: keep it in L2
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q atom.global.min.L2::cache_hint.u32 %0, [%1], %2, %4;\n\t}"
                 : "+r"(old) : "l"(addr), "r"(val), "r"(pred), "l"(pol) : "memory");
}
__device__ __forceinline__ void ki_red_add_v4(float *p, float a, float b, float c, float d, unsigned long long pol, uint32_t pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %6, 0;\n\t"
                 "@q red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;\n\t}"
                 ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "l"(pol), "r"(pred) : "memory");
}

// ---------------------------------------------------------------------------------------------
// kf_insert_lanes: the streaming insert.  The 128 points of a tile are taken as four groups of 32
// consecutive points, lane l of group j = point 32 j + l.  Runs of equal cells lie ACROSS lanes:
// they are summed with two segmented shuffle steps (lane l ends up with the sum of lanes l .. l+3
// of its run) and the lanes at run positions 0, 4, 8, ... issue the atomics, so a run of up to
// four points costs one set.  The own first-point bits of a group are one ballot -> one RED.XOR
// per 32 points.  (The first version gave every lane four consecutive points and merged runs inside
// the lane; a warp-wide reduction then spans the cells of 128 points, ~16 lines of accumulator
// rows, instead of ~4.  Measured in the same run: 41.6 vs 39.8 us for the stage, 0.0712 vs 0.0709
// ms for the overlapped step -- what an atomic costs is the lane operation, not the line.)
// ---------------------------------------------------------------------------------------------
template <int CIN, bool CART, int NV>
__global__ void __launch_bounds__(KI_WARPS * 32, KI_BLOCKS_PER_SM) kf_insert_lanes(const __grid_constant__ PvParams p,
                                                                                 const __grid_constant__ PvF f,
                                                                                 const uint32_t n_tiles, const uint32_t n_warps)
{
    constexpr int C = CIN + (CART ? 2 : 0);
    constexpr int CT = NV * 4;
    constexpr uint32_t TILE_FLOATS = 128u * CIN, TILE_BYTES = TILE_FLOATS * 4u;
    constexpr uint32_t FULL = 0xffffffffu;
    extern __shared__ __align__(128) float s_ring[];                 // [KI_WARPS][KI_STAGES][TILE_FLOATS]
    __shared__ __align__(8) unsigned long long s_bar[KI_WARPS][KI_STAGES];
    pf_pdl_trigger();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * KI_WARPS + warp;
    float *ring = s_ring + (size_t)warp * KI_STAGES * TILE_FLOATS;
    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < KI_STAGES; ++st) pf_mbar_init(pf_smem_addr(&s_bar[warp][st]), 1);
    }
    __syncwarp();
    auto issue = [&](uint32_t tile, int st) {             // lane 0: one bulk copy of a FULL tile into stage st
        const uint32_t bar = pf_smem_addr(&s_bar[warp][st]);
        pf_mbar_expect_tx(bar, TILE_BYTES);
        pf_bulk_g2s(pf_smem_addr(ring + (size_t)st * TILE_FLOATS), p.pts + (size_t)tile * TILE_FLOATS, TILE_BYTES, bar);
    };
    const uint32_t n_full = p.n >> 7;                      // tiles [0, n_full) are full; tile n_full (if any) is the tail
    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < KI_STAGES; ++st) {
            const uint32_t t = gw + (uint32_t)st * n_warps;
            if (t < n_full) issue(t, st);
        }
    }
    const float lo0 = p.lo[0], lo1 = p.lo[1], lo2 = p.lo[2];
    const float iv0 = p.inv_vs[0], iv1 = p.inv_vs[1], iv2 = p.inv_vs[2];
    const uint32_t nx = (uint32_t)p.grid[0], ny = (uint32_t)p.grid[1], nz = (uint32_t)p.grid[2];
    const float th0 = 3e-7f * (p.gridf[0] + 2.0f), th1 = 3e-7f * (p.gridf[1] + 2.0f), th2 = 3e-7f * (p.gridf[2] + 2.0f);
    const unsigned long long keep = pv_policy_evict_last();
    uint32_t olds[PF_PPT] = {0u, 0u, 0u, 0u};              // minima returned for the previous tile (0: site unused)
    uint32_t prev_base = 0;
    bool prev_valid = false;
    // first-point bits of a finished tile: a site whose returned minimum was larger than its own point
    // became the cell's minimum (bit on) and displaced the previous one (bit toggled back), see kf_insert
    auto settle = [&]() {
#pragma unroll
        for (int site = 0; site < PF_PPT; ++site) {
            const uint32_t i_run = prev_base + 32u * site + lane, old = olds[site];
            const bool won = old > i_run;
            if (won && old != PV_INF) atomicXor(f.bits + (old >> 5), 1u << (old & 31u));
            const uint32_t word = __ballot_sync(FULL, won);
            if (lane == 0 && word) atomicXor(f.bits + (prev_base >> 5) + site, word);
            olds[site] = 0u;
        }
    };

    const uint32_t my_off = (int)lane + 1 < p.B ? (uint32_t)__ldg(p.offsets + lane + 1) : PV_INF;   // interior frame boundaries
    uint32_t it = 0;
    for (uint32_t tile = gw; tile < n_tiles; tile += n_warps, ++it) {
        const int st = (int)(it % KI_STAGES);
        const uint32_t tile_base = tile << 7;
        float rows[PF_PPT * CIN];
        float *stage = ring + (size_t)st * TILE_FLOATS;
        if (tile < n_full) pf_mbar_wait(pf_smem_addr(&s_bar[warp][st]), (it / KI_STAGES) & 1u);
        else {                                             // the batch's last, partial tile: guarded loads into the stage
            const uint32_t nf = (p.n - tile_base) * CIN;
            for (uint32_t e = lane; e < TILE_FLOATS; e += 32u) stage[e] = e < nf ? __ldg(p.pts + (size_t)tile_base * CIN + e) : 0.0f;
            __syncwarp();
        }
        {
#pragma unroll
            for (int j = 0; j < PF_PPT; ++j) {
                const float *in = stage + (32u * j + lane) * CIN;
                if (CIN % 4 == 0) {
#pragma unroll
                    for (int k = 0; k < CIN / 4; ++k) {
                        const float4 v4 = reinterpret_cast<const float4 *>(in)[k];
                        rows[j * CIN + 4 * k] = v4.x; rows[j * CIN + 4 * k + 1] = v4.y; rows[j * CIN + 4 * k + 2] = v4.z; rows[j * CIN + 4 * k + 3] = v4.w;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < CIN; ++k) rows[j * CIN + k] = in[k];
                }
            }
            __syncwarp();                                  // every lane has its rows: the stage may be refilled
            const uint32_t nt = tile + KI_STAGES * n_warps;
            if (lane == 0 && nt < n_full) issue(nt, st);
        }
        // ---- frame of the tile's first point (lane l holds offsets[l + 1]); boundaries inside the tile are rare ----
        int b0 = 0;
        uint32_t next_off = PV_INF;
        if (p.B <= 33) {
            b0 = __popc(__ballot_sync(FULL, my_off <= tile_base));
            next_off = __reduce_min_sync(FULL, my_off > tile_base ? my_off : PV_INF);
        } else {
            for (int bb0 = 1; bb0 < p.B; bb0 += 32) {
                const int bb = bb0 + (int)lane;
                const uint32_t o = bb < p.B ? (uint32_t)__ldg(p.offsets + bb) : PV_INF;
                b0 += __popc(__ballot_sync(FULL, o <= tile_base));
                next_off = min(next_off, __reduce_min_sync(FULL, o > tile_base ? o : PV_INF));
            }
        }
        const bool straddle = next_off < tile_base + 128u;
        // ---- per point: transform, bins, slot ----
        float v[PF_PPT][CT];
        uint32_t slot[PF_PPT];
        float t[PF_PPT][3], c[PF_PPT][3];
        uint32_t unsure = 0;
#pragma unroll
        for (int j = 0; j < PF_PPT; ++j) {
            const float *in = rows + j * CIN;
            if (CART) {                                    // utils.py:42-44: (rho, phi, z, x, y, feat3..)
                v[j][0] = pv_rho(in[0], in[1]);
                v[j][1] = pv_atan2f(in[1], in[0]);
                v[j][2] = in[2]; v[j][3] = in[0]; v[j][4] = in[1];
#pragma unroll
                for (int k = 5; k < CT; ++k) v[j][k] = k < C ? in[k - 2 < CIN ? k - 2 : 0] : 0.0f;
            } else {
#pragma unroll
                for (int k = 0; k < CT; ++k) v[j][k] = k < C ? in[k < CIN ? k : 0] : 0.0f;
            }
            v[j][C] = 1.0f;                                // the count rides in channel C of the row
            t[j][0] = __fsub_rn(v[j][0], lo0); t[j][1] = __fsub_rn(v[j][1], lo1); t[j][2] = __fsub_rn(v[j][2], lo2);
            const float r0 = __fmul_rn(t[j][0], iv0), r1 = __fmul_rn(t[j][1], iv1), r2 = __fmul_rn(t[j][2], iv2);
            c[j][0] = floorf(r0); c[j][1] = floorf(r1); c[j][2] = floorf(r2);
            const float d0 = __fsub_rn(__fsub_rn(r0, c[j][0]), 0.5f), d1 = __fsub_rn(__fsub_rn(r1, c[j][1]), 0.5f),
                        d2 = __fsub_rn(__fsub_rn(r2, c[j][2]), 0.5f);
            if (!(fabsf(d0) < 0.5f - th0)) unsure |= 1u << (3 * j);
            if (!(fabsf(d1) < 0.5f - th1)) unsure |= 2u << (3 * j);
            if (!(fabsf(d2) < 0.5f - th2)) unsure |= 4u << (3 * j);
        }
        if (unsure) {                                      // rare: the IEEE division decides (NaN -> outside)
#pragma unroll
            for (int j = 0; j < PF_PPT; ++j)
#pragma unroll
                for (int d = 0; d < 3; ++d)
                    if ((unsure >> (3 * j + d)) & 1u) {
                        const float q = floorf(__fdiv_rn(t[j][d], p.vs[d]));
                        c[j][d] = q == q ? q : -1.0f;
                    }
        }
        const uint32_t sb0 = (uint32_t)b0 * f.capf;
#pragma unroll
        for (int j = 0; j < PF_PPT; ++j) {
            const uint32_t cx = (uint32_t)__float2int_rz(c[j][0]), cy = (uint32_t)__float2int_rz(c[j][1]), cz = (uint32_t)__float2int_rz(c[j][2]);
            const bool ok = cx < nx && cy < ny && cz < nz && tile_base + 32u * j + lane < p.n;
            slot[j] = ok ? sb0 + (cz * nx + cx) * ny + cy : PV_INF;   // direct map, PHI FASTEST
        }
        if (straddle) {                                    // a frame boundary inside the tile: per-point frame
#pragma unroll
            for (int j = 0; j < PF_PPT; ++j) {
                int b = b0;
                uint32_t nxt = next_off;
                while (tile_base + 32u * j + lane >= nxt) {   // frames may be empty
                    ++b;
                    nxt = b + 1 < p.B ? (uint32_t)__ldg(p.offsets + b + 1) : PV_INF;
                }
                if (slot[j] != PV_INF) slot[j] += (uint32_t)(b - b0) * f.capf;
            }
        }
        // ---- runs of consecutive points in one cell lie across lanes: lane l takes the sum of lanes
        // l .. l+3 of its run, the lanes at run positions 0, 4, 8, .. emit ----
        uint32_t emit = 0;
#pragma unroll
        for (int j = 0; j < PF_PPT; ++j) {
            const uint32_t s = slot[j];
            const uint32_t up = __shfl_up_sync(FULL, s, 1), d1 = __shfl_down_sync(FULL, s, 1), d2 = __shfl_down_sync(FULL, s, 2);
            const uint32_t heads = __ballot_sync(FULL, lane == 0 || up != s);
            const uint32_t pos = lane - (31u - (uint32_t)__clz(heads & (FULL >> (31u - lane))));
            const bool s1 = lane < 31u && d1 == s, s2 = s1 && lane < 30u && d2 == s;   // lane l+2 only through l+1 (cells A, B, A are two runs of A)
#pragma unroll
            for (int k = 0; k <= C; ++k) {
                const float o = __shfl_down_sync(FULL, v[j][k], 1);
                v[j][k] = __fadd_rn(v[j][k], s1 ? o : 0.0f);
            }
#pragma unroll
            for (int k = 0; k <= C; ++k) {
                const float o = __shfl_down_sync(FULL, v[j][k], 2);
                v[j][k] = __fadd_rn(v[j][k], s2 ? o : 0.0f);
            }
            emit |= (s != PV_INF && (pos & 3u) == 0u) ? 1u << j : 0u;
        }
        // A run can only become its cell's minimum if the cell's current minimum is larger: look first (a plain load
        // costs a fraction of a returning atomic, and first[] only ever decreases, so a stale value errs on the side
        // of issuing the atomic).  Tiles are taken in roughly ascending order: sweeps after the first mostly find the
        // cell claimed (-29 % returning atomics)
        uint32_t seen[PF_PPT];
#pragma unroll
        for (int j = 0; j < PF_PPT; ++j) {
            seen[j] = 0u;
            if ((emit >> j) & 1u)
                asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(seen[j]) : "l"(f.first + slot[j]), "l"(keep));
        }
        // ---- the previous tile's minima have had a whole tile of arithmetic to come back ----
        if (prev_valid) settle();
        prev_base = tile_base; prev_valid = true;
#pragma unroll
        for (int j = 0; j < PF_PPT; ++j) {
            const uint32_t pred = (emit >> j) & 1u;
            const uint32_t sj = pred ? slot[j] : 0u;       // keep the address computation in range when predicated off
            const uint32_t i_run = tile_base + 32u * j + lane;
            if (KI_EXP != 1 && KI_EXP != 3) ki_atom_min(olds[j], f.first + sj, i_run, (KI_EXP == 5 || seen[j] > i_run) ? pred : 0u);
            float *row = f.acc + (size_t)sj * f.rowf;
            if (KI_EXP != 1 && KI_EXP != 4) {
#pragma unroll
                for (int q = 0; q < NV; ++q) ki_red_add_v4(row + 4 * q, v[j][4 * q], v[j][4 * q + 1], v[j][4 * q + 2], v[j][4 * q + 3], keep, pred);
            }
        }
#pragma unroll
        for (int j = 0; j < PF_PPT; ++j)
            if (tile_base + 32u * j + lane < p.n) __stcs(f.sa + tile_base + 32u * j + lane, slot[j]);
    }
    if (prev_valid) settle();
}

// ---------------------------------------------------------------------------------------------
// F3 -- popcount scan over the first-point bitmap, one block per frame; consumes (zeroes) bits[]
// and publishes wb[] = {prefix, word}.  The last block to finish computes the output row bases.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PF_SCAN_THREADS) kf_scan(const __grid_constant__ PvParams p, const __grid_constant__ PvF f)
{
    __shared__ uint32_t s_warp[PF_SCAN_THREADS / 32];
    __shared__ uint32_t s_carry, s_last;
    pf_pdl_trigger();
    pf_pdl_wait();
    const int b = blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // static: one bit per point of the frame; dynamic: one bit per cell of the grid
    const uint32_t n_b = p.dyn ? p.cells : (uint32_t)(p.offsets[b + 1] - p.offsets[b]);
    const uint32_t nw = min((n_b + 31u) >> 5, f.wcap);
    uint32_t *bits = f.bits + (size_t)b * f.wcap;
    uint2 *wb = f.wb + (size_t)b * f.wcap;
    if (tid == 0) {
        s_carry = 0;
        if (b == 0) *reinterpret_cast<unsigned long long *>(f.ctrl + 4) = 0ull;   // heavy-cell allocator of this call
    }
    __syncthreads();
    // each thread owns PF_SCAN_VEC consecutive 4-word vectors per iteration: one block scan covers
    // 16k words (512k points), so a LiDAR frame is a single iteration
    for (uint32_t w0 = 0; w0 < nw; w0 += PF_SCAN_THREADS * 4 * PF_SCAN_VEC) {
        const uint32_t wt = w0 + tid * (4u * PF_SCAN_VEC);
        uint4 v[PF_SCAN_VEC];
        uint32_t tsum = 0;
#pragma unroll
        for (int q = 0; q < PF_SCAN_VEC; ++q) {            // wcap % 4 == 0: a vector never straddles the end
            v[q] = wt + 4u * q < nw ? __ldcg(reinterpret_cast<const uint4 *>(bits + wt) + q) : make_uint4(0, 0, 0, 0);
            tsum += __popc(v[q].x) + __popc(v[q].y) + __popc(v[q].z) + __popc(v[q].w);
        }
        uint32_t incl = tsum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (unsigned)d) incl += o;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t wsum = lane < PF_SCAN_THREADS / 32 ? s_warp[lane] : 0u;     // every warp scans the warp totals
        uint32_t wincl = wsum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, wincl, d);
            if (lane >= (unsigned)d) wincl += o;
        }
        const uint32_t warp_excl = __shfl_sync(0xffffffffu, wincl - wsum, warp);
        const uint32_t total = __shfl_sync(0xffffffffu, wincl, 31);
        uint32_t run = s_carry + warp_excl + incl - tsum;
#pragma unroll
        for (int q = 0; q < PF_SCAN_VEC; ++q) {
            if (wt + 4u * q < nw) {
                const uint32_t c0 = __popc(v[q].x), c1 = __popc(v[q].y), c2 = __popc(v[q].z), c3 = __popc(v[q].w);
                uint4 *dst = reinterpret_cast<uint4 *>(wb + wt + 4u * q);
                dst[0] = make_uint4(run, v[q].x, run + c0, v[q].y);
                dst[1] = make_uint4(run + c0 + c1, v[q].z, run + c0 + c1 + c2, v[q].w);
                // restore the bitmap; issued after the loaded value was consumed (a store to an
                // address with a load still in flight stalls the load/store unit for the round trip)
                if (c0 + c1 + c2 + c3) reinterpret_cast<uint4 *>(bits + wt)[q] = make_uint4(0, 0, 0, 0);
                run += c0 + c1 + c2 + c3;
            }
        }
        __syncthreads();
        if (tid == 0) s_carry += total;
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t raw = s_carry;
        f.counts_raw[b] = raw;
        p.voxel_counts[b] = (int32_t)(p.dyn ? raw : min(raw, (uint32_t)p.V));    // no max_voxels cap on the dynamic path
        __threadfence();
        const uint32_t done = atomicAdd(f.ctrl + 3, 1u);
        s_last = done == gridDim.x - 1 ? 1u : 0u;
        if (s_last) f.ctrl[3] = 0u;
    }
    __syncthreads();
    if (s_last && warp == 0) {                           // row bases: exclusive sum of the capped counts
        __threadfence();
        uint32_t carry = 0;
        for (int b0 = 0; b0 < p.B; b0 += 32) {
            const int bb = b0 + (int)lane;
            const uint32_t m = bb < p.B ? (uint32_t)__ldcg(p.voxel_counts + bb) : 0u;
            uint32_t incl = m;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= (unsigned)d) incl += o;
            }
            if (bb < p.B) f.base[bb] = (int32_t)(carry + incl - m);
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) f.base[p.B] = (int32_t)carry;
    }
}

// ---------------------------------------------------------------------------------------------
// F2 (static path) -- popcount scan over the first-point bitmap.  The bitmap is addressed by the
// GLOBAL point index, so the scan ignores frames: one block per chunk of PF_CHUNK_WORDS words
// (16k points) publishes wb[w] = {first points of the chunk before word w, the word} and the
// chunk's total; the LAST block to finish turns the totals into chunk bases and evaluates the
// batch-wide prefix G(i) = first points with index < i at the frame boundaries:
//     frank0[b] = G(offsets[b]),  raw count of frame b = G(offsets[b+1]) - G(offsets[b]),
// so a first point fi of frame b has first-occurrence rank G(fi) - frank0[b] (pf_rank).  One wave
// of independent blocks + one short tail instead of one block per frame (10 -> 3 us for 8 frames).
// The bitmap is cleared later by kf_heavy_points.
// ---------------------------------------------------------------------------------------------
#define PF_CHUNK_WORDS 512
#define PF_CHUNK_SHIFT 9
#define PF_SCAN2_THREADS 128             // 4 words per thread

// G(i): first points of the batch with index < i (valid once kf_scan2's last block has published cbase)
__device__ __forceinline__ uint32_t pf_prefix_at(const PvF &f, uint32_t i)
{
    const uint32_t w = i >> 5;
    const uint2 wv = __ldcg(f.wb + w);
    return __ldcg(f.cbase + (w >> PF_CHUNK_SHIFT)) + wv.x + __popc(wv.y & ((1u << (i & 31u)) - 1u));
}

__global__ void __launch_bounds__(PF_SCAN2_THREADS) kf_scan2(const __grid_constant__ PvParams p, const __grid_constant__ PvF f)
{
    __shared__ uint32_t s_warp[PF_SCAN2_THREADS / 32];
    __shared__ uint32_t s_last, s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t nw = (p.n >> 5) + 1u;                    // words [0, nw) cover bit n (always clear)
    const uint32_t w = blockIdx.x * PF_CHUNK_WORDS + tid * 4u;
    pf_pdl_trigger();
    if (blockIdx.x == 0 && tid == 0) *reinterpret_cast<unsigned long long *>(f.ctrl + 4) = 0ull;   // heavy-cell allocator of this call
    pf_pdl_wait();                           // the bitmap of kf_insert
    // the bitmap is padded with clear words past nw (pvf_make_layout), so a vector never reads outside
    const uint4 x = w < nw ? __ldcg(reinterpret_cast<const uint4 *>(f.bits + w)) : make_uint4(0, 0, 0, 0);
    const uint32_t c0 = __popc(x.x), c1 = __popc(x.y), c2 = __popc(x.z), c3 = __popc(x.w);
    const uint32_t tsum = c0 + c1 + c2 + c3;
    uint32_t incl = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (unsigned)d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t run = incl - tsum, total = 0;
#pragma unroll
    for (int k = 0; k < PF_SCAN2_THREADS / 32; ++k) {
        const uint32_t v = s_warp[k];
        if ((uint32_t)k < warp) run += v;
        total += v;
    }
    if (w < nw) {                            // wb is sized like the bitmap: the padding words may be written too
        uint4 *dst = reinterpret_cast<uint4 *>(f.wb + w);
        dst[0] = make_uint4(run, x.x, run + c0, x.y);
        dst[1] = make_uint4(run + c0 + c1, x.z, run + c0 + c1 + c2, x.w);
    }
    if (tid == 0) {
        f.cagg[blockIdx.x] = total;
        __threadfence();
        const uint32_t done = atomicAdd(f.ctrl + 3, 1u);
        s_last = done == gridDim.x - 1 ? 1u : 0u;
        if (s_last) f.ctrl[3] = 0u;
        s_carry = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- tail (one block): chunk bases, then the per-frame quantities ----
    const uint32_t nchunks = gridDim.x;
    for (uint32_t c0b = 0; c0b < nchunks; c0b += PF_SCAN2_THREADS) {
        const uint32_t c = c0b + tid;
        const uint32_t v = c < nchunks ? __ldcg(f.cagg + c) : 0u;
        uint32_t in2 = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, in2, d);
            if (lane >= (unsigned)d) in2 += o;
        }
        if (lane == 31) s_warp[warp] = in2;
        __syncthreads();
        uint32_t off = s_carry, tot = 0;
#pragma unroll
        for (int k = 0; k < PF_SCAN2_THREADS / 32; ++k) {
            const uint32_t t = s_warp[k];
            if ((uint32_t)k < warp) off += t;
            tot += t;
        }
        if (c < nchunks) f.cbase[c] = off + in2 - v;
        __syncthreads();
        if (tid == 0) s_carry += tot;
        __syncthreads();
    }
    __threadfence_block();
    __syncthreads();
    // frames: G at every boundary (cbase written by this block: read back through L2)
    if (warp == 0) {
        uint32_t carry = 0;
        for (int b0 = 0; b0 < p.B; b0 += 32) {
            const int bb = b0 + (int)lane;
            uint32_t m = 0;
            if (bb < p.B) {
                const uint32_t g0 = pf_prefix_at(f, (uint32_t)p.offsets[bb]), g1 = pf_prefix_at(f, (uint32_t)p.offsets[bb + 1]);
                const uint32_t raw = g1 - g0;
                f.frank0[bb] = g0;
                f.counts_raw[bb] = raw;
                m = min(raw, (uint32_t)p.V);
                p.voxel_counts[bb] = (int32_t)m;
            }
            uint32_t in3 = m;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, in3, d);
                if (lane >= (unsigned)d) in3 += o;
            }
            if (bb < p.B) f.base[bb] = (int32_t)(carry + in3 - m);
            carry += __shfl_sync(0xffffffffu, in3, 31);
        }
        if (lane == 0) f.base[p.B] = (int32_t)carry;
    }
}

// First-occurrence rank of the cell whose first point is fi (frame b); wv = wb[fi >> 5].
__device__ __forceinline__ uint32_t pf_rank(const PvF &f, const int b, const uint32_t fi, const uint2 wv)
{
    return __ldg(f.cbase + (fi >> (5 + PF_CHUNK_SHIFT))) + wv.x + __popc(wv.y & ((1u << (fi & 31u)) - 1u)) - __ldg(f.frank0 + b);
}

// ---------------------------------------------------------------------------------------------
// F4 -- finalize: stream over the map, one slot per thread.  grid = (slots / 256, B)
// Cells holding more than T points ("heavy", 2-5 % of the points) get coors / num_points here and
// are registered for F5, which selects their T smallest point indices and writes their features.
// ---------------------------------------------------------------------------------------------
// The map (first[], accumulator rows) is the part of the workspace that is touched at random; with
// PF_WS_KEEP its accesses carry an L2 evict_last policy (and the streams around it evict_first), so
// a workspace that comes round again soon is still in the 126 MB L2 and its rows neither have to be
// filled from nor written back to DRAM one 32-byte sector at a time.
#ifndef PF_WS_KEEP
#define PF_WS_KEEP 1
#endif
template <int NV>
__device__ __forceinline__ void pf_ld_row(const float *row, float (&r)[NV * 4])
{
    if constexpr (NV == 2) {        // one 256-bit load per 32-byte row
        if (PF_WS_KEEP) {
            const unsigned long long pol = pv_policy_evict_last();
            asm volatile("ld.global.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                         : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
                         : "l"(row), "l"(pol));
        } else
        asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
                     : "l"(row));
    } else {
        const unsigned long long pol = pv_policy_evict_last();
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const float4 v = PF_WS_KEEP ? pv_ld_keep(reinterpret_cast<const float4 *>(row) + q, pol)
                                        : __ldcg(reinterpret_cast<const float4 *>(row) + q);
            r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
        }
    }
}
// Restores a row: zeros, except words 0 and 1 which keep `keep0`, `keep1` -- for a heavy cell its
// candidate-range offset and (word 2, starting at zero) its arrival cursor, until F4 is done.
template <int NV>
__device__ __forceinline__ void pf_st_row_clean(float *row, float keep0, float keep1 = 0.0f)
{
    if constexpr (NV == 2) {
        if (PF_WS_KEEP) {
            const unsigned long long pol = pv_policy_evict_last();
            asm volatile("st.global.L2::cache_hint.v8.f32 [%0], {%1,%2,%3,%3,%3,%3,%3,%3}, %4;"
                         ::"l"(row), "f"(keep0), "f"(keep1), "f"(0.0f), "l"(pol) : "memory");
        } else
        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%3,%3,%3,%3,%3};" ::"l"(row), "f"(keep0), "f"(keep1), "f"(0.0f) : "memory");
    } else {
#pragma unroll
        for (int q = 0; q < NV; ++q)
            __stcg(reinterpret_cast<float4 *>(row) + q, make_float4(q == 0 ? keep0 : 0.f, q == 0 ? keep1 : 0.f, 0.f, 0.f));
    }
}
__device__ __forceinline__ uint32_t pf_ld_first(const uint32_t *p)
{
    if (!PF_WS_KEEP) return __ldcs(p);
    uint32_t v;
    const unsigned long long pol = pv_policy_evict_last();
    asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void pf_st_first(uint32_t *p, uint32_t v)
{
    if (!PF_WS_KEEP) { *p = v; return; }
    const unsigned long long pol = pv_policy_evict_last();
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}

// One occupied map slot s of frame b = cell (cz, cy, cx): rank lookup, per-voxel outputs, heavy
// registration, map restore.  m[] / dens receive what the dense canvas / density hold for the cell.
template <int NV, int CC, bool CANVAS, bool DENSE>
__device__ __forceinline__ void pf_finalize_cell(const PvParams &p, const PvF &f, const uint32_t s, const int b,
                                                 const uint32_t fi, const uint32_t cx, const uint32_t cy,
                                                 const uint32_t cz, const float (&r)[NV * 4], const uint2 wv,
                                                 float (&m)[CANVAS ? NV * 4 : 1], int32_t &dens)
{
    constexpr int CT = NV * 4;
    const int C = CC ? CC : p.C;
    float *rowp = f.acc + (size_t)s * f.rowf;
    const uint32_t rank = pf_rank(f, b, fi, wv);
    float keep0 = 0.0f, keep1 = 0.0f;
    if (rank < (uint32_t)p.V) {                                           // :60-61 max_voxels
        float cntf = 0.0f;
#pragma unroll
        for (int k = 0; k < CT; ++k) cntf = k == C ? r[k] : cntf;
        const uint32_t cnt = (uint32_t)cntf;
        const uint32_t T = (uint32_t)p.T;
        const uint32_t L = min(cnt, T);
        const int32_t vid = __ldg(f.base + b) + (int32_t)rank;
        const uint32_t cell = (cz * (uint32_t)p.grid[1] + cy) * (uint32_t)p.grid[0] + cx;
        if (KF_EXP != 4) __stcs(reinterpret_cast<int4 *>(p.coors) + vid, make_int4(b, (int)cz, (int)cy, (int)cx));   // write-once outputs: streaming
        if (KF_EXP != 4) __stcs(p.num_points + vid, (int32_t)L);
        dens = (int32_t)cnt;                                              // :70-71 un-capped count
        if (cnt > T) {
            // heavy: the row holds the sum over ALL points; F5 re-sums the T smallest indices
            const unsigned long long a = atomicAdd(reinterpret_cast<unsigned long long *>(f.ctrl + 4),
                                                   (1ull << 32) | (unsigned long long)cnt);
            const uint32_t hid = (uint32_t)(a >> 32), off = (uint32_t)a;
            if (hid < f.hmax) {
                uint4 *hi = f.hinfo + 2 * (size_t)hid;
                hi[0] = make_uint4(s, (uint32_t)vid, off, cnt);
                hi[1] = make_uint4(cell, (uint32_t)b, 0u, 0u);               // .z = arrival cursor
                keep0 = __uint_as_float(hid);                         // (informational)
                keep1 = __uint_as_float(off);                         // kf_heavy_points: hlist[off + cursor++]
                atomicOr(f.hbits + (s >> 5), 1u << (s & 31u));
            } else atomicOr(p.ws.ctrl + 1, 1u);
        } else {
            const float nf = (float)L, inv = __frcp_rn(nf);
            float mean[CT];
#pragma unroll
            for (int k = 0; k < CT; ++k) {
                mean[k] = k < C ? pv_div_count(r[k], nf, inv) : 0.0f;     // voxel_encoder.py:18-22
                if (CANVAS) m[k] = mean[k];
            }
            if (!CANVAS && !DENSE && p.canvas) {
                float *cv = p.canvas + (size_t)b * C * p.cells + cell;
#pragma unroll
                for (int k = 0; k < CT; ++k)
                    if (k < C) cv[(size_t)k * p.cells] = mean[k];
            }
            if (p.feats && KF_EXP != 4) pv_store_feats<CT>(p.feats, vid, C, mean);
        }
        if (!DENSE && p.density) p.density[(size_t)b * p.cells + cell] = (int32_t)cnt;
    }
    // restore the map -- after the loaded row was consumed (see kf_scan)
    if (KF_EXP != 2) pf_st_row_clean<NV>(rowp, keep0, keep1);
    if (KF_EXP != 2) pf_st_first(f.first + s, PV_INF);
    if (!DENSE) f.keys[s] = PV_INF;
}

// Hash maps: grid = (slots / 256, B) over the slot index; dense outputs were zero-filled by the host.
template <int NV, int CC>
__global__ void __launch_bounds__(256) kf_finalize(const __grid_constant__ PvParams p, const __grid_constant__ PvF f)
{
    pf_pdl_trigger();
    pf_pdl_wait();
    const int b = blockIdx.y;
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= f.capf) return;
    const uint32_t s = (uint32_t)b * f.capf + l;
    const uint32_t fi = __ldcs(f.first + s);
    if (fi == PV_INF) return;
    const uint32_t cell = __ldcg(f.keys + s), nx = p.grid[0], ny = p.grid[1];
    const uint32_t x = cell % nx, yz = cell / nx, cz = yz / ny, cy = yz - cz * ny;
    float m[1], r[NV * 4];
    int32_t dens;
    pf_ld_row<NV>(f.acc + (size_t)s * f.rowf, r);
    const uint2 wv = __ldg(f.wb + (fi >> 5));
    pf_finalize_cell<NV, CC, false, false>(p, f, s, b, fi, x, cy, cz, r, wv, m, dens);
}

// Direct maps: one block per patch of PF_PATCH azimuth x PF_PATCH range cells of one z layer.  The
// map is PHI FASTEST (the order LiDAR points arrive in), the canvas RHO FASTEST
// (pillar_encoder.py:211-217), so the block works in two phases around a shared-memory transpose:
//   1. a warp reads 32 azimuth-consecutive slots (coalesced; their rows, first-point words and --
//      because azimuth neighbours are mostly first hit by consecutive points -- their output rows
//      are neighbours too, so the per-voxel loads and stores of a warp share 128-byte lines),
//   2. a warp writes 32 range-consecutive canvas cells per channel (one full line per store),
//      every element exactly once, zeros included: no zero fill, no scatter.
// grid = (patches_y * nz, patches_x, B); each thread owns PF_PATCH / 8 cells of the patch.
// Tried and slower: fetching all rows of a thread up front with cp.async into shared memory
// (512 threads, 2 cells each: 37 vs 33 us), 16-byte canvas stores, unaligned per-channel feature
// stores -- the kernel is insensitive to its instruction count.
#define PF_PATCH 32
template <int NV, int CC, bool CANVAS>
#ifndef PF_FIN_WARPS
#define PF_FIN_WARPS 8           // measured: 8 warps 30.7 us, 16 warps 32.5 us, 4 warps 33.1 us
#endif
__global__ void __launch_bounds__(PF_FIN_WARPS * 32) kf_finalize_patch(const __grid_constant__ PvParams p, const __grid_constant__ PvF f)
{
    constexpr int CT = NV * 4;
    constexpr int KP = PF_PATCH / PF_FIN_WARPS;              // cells per thread
    extern __shared__ float s_t[];                           // [C (+1 density)][PF_PATCH rho][PF_PATCH + 1 phi]
    const int C = CC ? CC : p.C;
    const int b = blockIdx.z;
    const uint32_t nx = p.grid[0], ny = p.grid[1];
    const uint32_t py = (ny + PF_PATCH - 1) / PF_PATCH;
    const uint32_t pxi = blockIdx.y;
    uint32_t pyi = blockIdx.x, z = 0;
    if (p.grid[2] > 1) { z = pyi / py; pyi -= z * py; }
    const uint32_t lane = threadIdx.x & 31u, wq = threadIdx.x >> 5;
    int32_t *s_d = reinterpret_cast<int32_t *>(s_t) + (CANVAS ? (size_t)C : 0) * PF_PATCH * (PF_PATCH + 1);

    pf_pdl_trigger();
    pf_pdl_wait();                           // map, bitmap prefix and row bases of the earlier kernels
    // ---- phase 1: lane = azimuth ----
    const uint32_t y = pyi * PF_PATCH + lane;
    uint32_t fi[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        const uint32_t x = pxi * PF_PATCH + wq + (uint32_t)PF_FIN_WARPS * k;
        fi[k] = (x < nx && y < ny) ? pf_ld_first(f.first + (size_t)b * f.capf + (z * nx + x) * ny + y) : PV_INF;
    }
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        const uint32_t xl = wq + (uint32_t)PF_FIN_WARPS * k, x = pxi * PF_PATCH + xl;
        float m[CANVAS ? CT : 1];
        int32_t dens = 0;
        if (CANVAS) {
#pragma unroll
            for (int c = 0; c < CT; ++c) m[c] = 0.0f;
        }
        if (fi[k] != PV_INF) {
            const uint32_t s = (uint32_t)b * f.capf + (z * nx + x) * ny + y;
            float r[CT];
            pf_ld_row<NV>(f.acc + (size_t)s * f.rowf, r);
            const uint2 wv = __ldg(f.wb + (fi[k] >> 5));
            pf_finalize_cell<NV, CC, CANVAS, true>(p, f, s, b, fi[k], x, y, z, r, wv, m, dens);
        }
        if (CANVAS) {
#pragma unroll
            for (int c = 0; c < CT; ++c)
                if (c < C) s_t[((size_t)c * PF_PATCH + xl) * (PF_PATCH + 1) + lane] = m[c];
        }
        if (p.density) s_d[xl * (PF_PATCH + 1) + lane] = dens;
    }
    if (!CANVAS && !p.density) return;
    __syncthreads();
    // ---- phase 2: lane = range ----
    const uint32_t x2 = pxi * PF_PATCH + lane;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        const uint32_t yl = wq + (uint32_t)PF_FIN_WARPS * k, y2 = pyi * PF_PATCH + yl;
        if (x2 >= nx || y2 >= ny) continue;
        const size_t cell = ((size_t)z * ny + y2) * nx + x2;
        if (CANVAS) {
            float *cv = p.canvas + (size_t)b * C * p.cells + cell;
#pragma unroll
            for (int c = 0; c < CT; ++c)
                if (c < C) { __stcs(cv, s_t[((size_t)c * PF_PATCH + lane) * (PF_PATCH + 1) + yl]); cv += p.cells; }
        }
        if (p.density) p.density[(size_t)b * p.cells + cell] = s_d[lane * (PF_PATCH + 1) + yl];
    }
}

// ---------------------------------------------------------------------------------------------
// F3, pillar grids with a canvas (the headline path): one block per 32 x 32 patch like
// kf_finalize_patch, but
//   * the occupied cells of a warp's 128 cells (23-30 % on a LiDAR frame) are COMPACTED first
//     (ballot + popcount into a per-warp list), so the long per-voxel path runs on full warps:
//     1-2 dense iterations instead of 4 sparse ones;
//   * the patch's canvas tile [C][32 phi][32 rho] is assembled in shared memory (zero filled while
//     the predecessor kernel drains, means scattered in by the packed lanes) and leaves as ONE TMA
//     tensor store (cp.async.bulk.tensor.3d, 128-byte swizzle so the scatter is at most 4-way
//     conflicted): no per-element LDS / STG phase, every canvas element still written exactly once.
// ---------------------------------------------------------------------------------------------
#define PF_TMA_SWZ(row, col) ((uint32_t)(row) * 32u + (((((uint32_t)(col)) >> 2) ^ ((uint32_t)(row) & 7u)) << 2) + ((uint32_t)(col) & 3u))
template <int NV, int CC>
__global__ void __launch_bounds__(256) kf_finalize_tma(const __grid_constant__ PvParams p, const __grid_constant__ PvF f,
                                                       const __grid_constant__ CUtensorMap tmap)
{
    constexpr int CT = NV * 4;
    extern __shared__ __align__(1024) unsigned char s_dyn[];
    __shared__ uint32_t s_fi[8][128];
    __shared__ uint8_t s_id[8][128];
    const int C = CC ? CC : p.C;
    const int b = blockIdx.z;
    const uint32_t nx = p.grid[0], ny = p.grid[1];
    const uint32_t pxi = blockIdx.y, pyi = blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wq = tid >> 5;
    // the swizzle pattern repeats every 1024 bytes: align the tile in the shared window
    const uint32_t dyn0 = pf_smem_addr(s_dyn), tile0 = (dyn0 + 1023u) & ~1023u;
    float *s_t = reinterpret_cast<float *>(s_dyn + (tile0 - dyn0));

    pf_pdl_trigger();
    {   // zero fill: independent of the earlier kernels, overlaps their tail
        float4 *t4 = reinterpret_cast<float4 *>(s_t);
        for (int q = (int)tid; q < C * 256; q += 256) t4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    pf_pdl_wait();                           // map, bitmap prefix and row bases of the earlier kernels
    const uint32_t y = pyi * 32u + lane;
    uint32_t fi[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t x = pxi * 32u + wq + 8u * k;
        fi[k] = (x < nx && y < ny) ? pf_ld_first(f.first + (size_t)b * f.capf + x * ny + y) : PV_INF;
    }
    uint32_t cnt = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool occ = fi[k] != PV_INF;
        const uint32_t mask = __ballot_sync(0xffffffffu, occ);
        if (occ) {
            const uint32_t pos = cnt + __popc(mask & ((1u << lane) - 1u));
            s_fi[wq][pos] = fi[k];
            s_id[wq][pos] = (uint8_t)(k * 32 + (int)lane);
        }
        cnt += __popc(mask);
    }
    __syncthreads();                         // tile zeroed (all warps), lists visible
    for (uint32_t e = lane; e < cnt; e += 32u) {
        const uint32_t fiv = s_fi[wq][e], id = s_id[wq][e];
        const uint32_t yl = id & 31u, xl = wq + 8u * (id >> 5);
        const uint32_t x = pxi * 32u + xl, yy = pyi * 32u + yl;
        const uint32_t s = (uint32_t)b * f.capf + x * ny + yy;
        float r[CT], m[CT];
        if (KF_EXP == 1) {
#pragma unroll
            for (int c = 0; c < CT; ++c) r[c] = 1.0f;
        } else pf_ld_row<NV>(f.acc + (size_t)s * f.rowf, r);
        const uint2 wv = __ldg(f.wb + (fiv >> 5));
#pragma unroll
        for (int c = 0; c < CT; ++c) m[c] = 0.0f;
        int32_t dens = 0;
        pf_finalize_cell<NV, CC, true, true>(p, f, s, b, fiv, x, yy, 0u, r, wv, m, dens);
#pragma unroll
        for (int c = 0; c < CT; ++c)
            if (c < C) s_t[PF_TMA_SWZ(c * 32 + (int)yl, xl)] = m[c];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> the TMA's async proxy
    __syncthreads();
    if (tid == 0 && KF_EXP != 3) {
        const unsigned long long pol = pf_policy_evict_first();      // the canvas is write-once
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2, %3}], [%4], %5;"
                     ::"l"(reinterpret_cast<unsigned long long>(&tmap)), "r"((int)(pxi * 32u)), "r"((int)(pyi * 32u)),
                       "r"(b * C), "r"(tile0), "l"(pol) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // the tile must outlive the read
    }
}

// ---------------------------------------------------------------------------------------------
// F5a -- every point of a heavy cell appends its index to the cell's candidate range (one atomic
// cursor per cell; the range was sized from the exact count).  The bitmap lookup has warp
// locality (phi-fastest bit order).
// ---------------------------------------------------------------------------------------------
#define PF_HPT 8                         // points per thread in the heavy-point pass
__global__ void __launch_bounds__(PF_THREADS) kf_heavy_points(const __grid_constant__ PvParams p, const __grid_constant__ PvF f)
{
    pf_pdl_trigger();
    pf_pdl_wait();
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * PF_HPT;
    if (i0 >= p.n) return;
    if ((i0 & 31u) == 0) f.bits[i0 >> 5] = 0u;      // the first-point bitmap is consumed: back to its clean state
    // the slot words are requested before the "any heavy cell at all?" word is looked at: one round trip
    uint32_t sa[PF_HPT];
    if (i0 + PF_HPT <= p.n) {
#pragma unroll
        for (int h = 0; h < PF_HPT / 4; ++h) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(f.sa + i0) + h);
            sa[4 * h] = v.x; sa[4 * h + 1] = v.y; sa[4 * h + 2] = v.z; sa[4 * h + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < PF_HPT; ++j) sa[j] = i0 + j < p.n ? f.sa[i0 + j] : PV_INF;
    }
    if (__ldg(reinterpret_cast<const unsigned long long *>(f.ctrl + 4)) == 0ull) return;   // no heavy cell in this batch
    uint32_t w[PF_HPT];
#pragma unroll
    for (int j = 0; j < PF_HPT; ++j) w[j] = sa[j] != PV_INF ? __ldg(f.hbits + (sa[j] >> 5)) : 0u;
#pragma unroll
    for (int j = 0; j < PF_HPT; ++j) {
        if (!((w[j] >> (sa[j] & 31u)) & 1u)) continue;
        // the heavy cell's row holds {id, candidate-range offset, arrival cursor}: one line, no indirection
        uint32_t *row = reinterpret_cast<uint32_t *>(f.acc + (size_t)sa[j] * f.rowf);
        const uint32_t off = __ldcg(row + 1);
        const uint32_t pos = atomicAdd(row + 2, 1u);
        f.hlist[off + pos] = i0 + j;
    }
}

// ---------------------------------------------------------------------------------------------
// F5b -- one warp per heavy cell: T rounds of "smallest candidate above the previous pick"
// (REDUX.MIN across the warp) give the T smallest point indices in ascending order -- exactly the
// points the reference keeps (point_cloud_ops.py:66); their rows are gathered and summed (fixed
// tree, deterministic), the mean goes to the feature row and the canvas.
// ---------------------------------------------------------------------------------------------
#define PF_CAND_REGS 8
__global__ void __launch_bounds__(256) kf_heavy_cells(const __grid_constant__ PvParams p, const __grid_constant__ PvF f)
{
    pf_pdl_wait();
    const unsigned long long alloc = __ldcg(reinterpret_cast<const unsigned long long *>(f.ctrl + 4));
    const uint32_t nh = min((uint32_t)(alloc >> 32), f.hmax);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t T = (uint32_t)p.T;
    const int C = p.C, c_in = p.c_in;
    for (uint32_t hid = gw; hid < nh; hid += nwarps) {
        const uint4 h0 = __ldcg(f.hinfo + 2 * (size_t)hid);
        const uint4 h1 = __ldcg(f.hinfo + 2 * (size_t)hid + 1);
        const uint32_t s = h0.x, vid = h0.y, off = h0.z, cnt = h0.w, cell = h1.x, b = h1.y;
        const uint32_t *cand = f.hlist + off;
        uint32_t cr[PF_CAND_REGS];
        const bool in_regs = cnt <= 32u * PF_CAND_REGS;
#pragma unroll
        for (int q = 0; q < PF_CAND_REGS; ++q) cr[q] = (in_regs && lane + 32u * q < cnt) ? __ldcg(cand + lane + 32u * q) : PV_INF;
        float acc[PV_MAX_CHANNELS];
#pragma unroll
        for (int k = 0; k < PV_MAX_CHANNELS; ++k) acc[k] = 0.0f;
        uint32_t lower = 0;                     // candidates below `lower` are already picked
        for (uint32_t r0 = 0; r0 < T; r0 += 32) {
            uint32_t mine = PV_INF;             // lane k keeps pick r0 + k
            const uint32_t rounds = min(32u, T - r0);
            for (uint32_t r = 0; r < rounds; ++r) {
                uint32_t best = PV_INF;
                if (in_regs) {
#pragma unroll
                    for (int q = 0; q < PF_CAND_REGS; ++q) best = min(best, cr[q] >= lower ? cr[q] : PV_INF);
                } else {
                    for (uint32_t k = lane; k < cnt; k += 32) {
                        const uint32_t v = __ldcg(cand + k);
                        best = min(best, v >= lower ? v : PV_INF);
                    }
                }
                best = __reduce_min_sync(0xffffffffu, best);
                if (lane == r) mine = best;
                lower = best + 1u;
            }
            if (mine != PV_INF) {
                float row[PV_MAX_CHANNELS];
                pv_feature_row(p.pts, mine, c_in, p.cart, row);
#pragma unroll
                for (int q = 0; q < PV_MAX_CHANNELS; ++q) acc[q] = __fadd_rn(acc[q], row[q]);
            }
        }
        const float nf = (float)T;
#pragma unroll
        for (int q = 0; q < PV_MAX_CHANNELS; ++q) {
            if (q < C) {
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) acc[q] = __fadd_rn(acc[q], __shfl_xor_sync(0xffffffffu, acc[q], d));
                if (lane == 0) {
                    const float mean = __fdiv_rn(acc[q], nf);
                    if (p.feats) p.feats[(size_t)vid * C + q] = mean;
                    if (p.canvas) p.canvas[((size_t)b * C + q) * p.cells + cell] = mean;
                }
            }
        }
        if (lane == 0) {
            *reinterpret_cast<float4 *>(f.acc + (size_t)s * f.rowf) = make_float4(0.f, 0.f, 0.f, 0.f);   // the row's last dirty words
            const uint32_t hb = s;
            atomicAnd(f.hbits + (hb >> 5), ~(1u << (hb & 31u)));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Dynamic voxelization: Voxelization.voxelize_dynamic
// (datasets/pipelines/voxelization.py:169-172) + torch.unique(grid_ind, dim=0) + scatter_mean of
// DynamicVoxelEncoderV1 (models/readers/voxel_encoder.py:38-44) + DynamicPPScatter
// (models/readers/pillar_encoder.py:413-432).  No max_points / max_voxels caps, voxels ordered by
// (b, z, y, x) = cell order, so the rank of a cell is the popcount prefix of the CELL-order
// occupancy bitmap that kf_insert<PF_MODE_DYN> sets: insert -> scan -> finalize, three launches, and the
// finalize pass writes every per-voxel output in increasing row order.
// ---------------------------------------------------------------------------------------------
template <int NV, int CC, bool CANVAS>
__global__ void __launch_bounds__(256) kf_dyn_finalize(const __grid_constant__ PvParams p, const __grid_constant__ PvF f)
{
    constexpr int CT = NV * 4;
    pf_pdl_trigger();
    pf_pdl_wait();
    const int C = CC ? CC : p.C;
    const int b = blockIdx.z;
    const uint32_t nx = p.grid[0], ny = p.grid[1];
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, yz = blockIdx.y;
    if (x >= nx) return;
    const uint32_t l = yz * nx + x;
    const uint32_t s = (uint32_t)b * f.capf + l;
    const uint2 wv = __ldg(f.wb + (size_t)b * f.wcap + (l >> 5));
    float m[CANVAS ? CT : 1];
    if (CANVAS) {
#pragma unroll
        for (int k = 0; k < CT; ++k) m[k] = 0.0f;
    }
    if ((wv.y >> (l & 31u)) & 1u) {
        float *rowp = f.acc + (size_t)s * f.rowf;
        float r[CT];
        pf_ld_row<NV>(rowp, r);
        const uint32_t rank = wv.x + __popc(wv.y & ((1u << (l & 31u)) - 1u));
        const int32_t vid = __ldg(f.base + b) + (int32_t)rank;
        float cntf = 0.0f;
#pragma unroll
        for (int k = 0; k < CT; ++k) cntf = k == C ? r[k] : cntf;
        const uint32_t cz = yz / ny, cy = yz - cz * ny;                   // uniform
        reinterpret_cast<int4 *>(p.coors)[vid] = make_int4(b, (int)cz, (int)cy, (int)x);   // unq row
        if (p.num_points) p.num_points[vid] = (int32_t)cntf;                               // unq_cnt
        const float inv = __frcp_rn(cntf);
        float mean[CT];
#pragma unroll
        for (int k = 0; k < CT; ++k) {
            mean[k] = k < C ? pv_div_count(r[k], cntf, inv) : 0.0f;       // scatter_mean: sum / count
            if (CANVAS) m[k] = mean[k];
        }
        if (p.feats) pv_store_feats<CT>(p.feats, vid, C, mean);
        pf_st_row_clean<NV>(rowp, 0.0f);                                  // restore the row (after use)
    }
    if (CANVAS) {                                                         // DynamicPPScatter, zeros included
        float *cv = p.canvas + (size_t)b * C * p.cells + l;
#pragma unroll
        for (int k = 0; k < CT; ++k)
            if (k < C) { __stcs(cv, m[k]); cv += p.cells; }
    }
}

// The same for hash-map grids: one thread per map slot; the voxel's row = popcount prefix of its CELL in
// the cell-order bitmap (so rows are still in torch.unique order), its sums sit in the slot's row.
template <int NV, int CC>
__global__ void __launch_bounds__(256) kf_dyn_finalize_hash(const __grid_constant__ PvParams p, const __grid_constant__ PvF f)
{
    constexpr int CT = NV * 4;
    pf_pdl_trigger();
    pf_pdl_wait();
    const int C = CC ? CC : p.C;
    const int b = blockIdx.y;
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= f.capf) return;
    const uint32_t s = (uint32_t)b * f.capf + l;
    const uint32_t cell = __ldcs(f.keys + s);
    if (cell == PV_INF) return;
    const uint2 wv = __ldg(f.wb + (size_t)b * f.wcap + (cell >> 5));
    const uint32_t rank = wv.x + __popc(wv.y & ((1u << (cell & 31u)) - 1u));
    const int32_t vid = __ldg(f.base + b) + (int32_t)rank;
    float *rowp = f.acc + (size_t)s * f.rowf;
    float r[CT];
    pf_ld_row<NV>(rowp, r);
    float cntf = 0.0f;
#pragma unroll
    for (int k = 0; k < CT; ++k) cntf = k == C ? r[k] : cntf;
    const uint32_t nx = p.grid[0], ny = p.grid[1];
    const uint32_t x = cell % nx, yz = cell / nx, cz = yz / ny, cy = yz - cz * ny;
    reinterpret_cast<int4 *>(p.coors)[vid] = make_int4(b, (int)cz, (int)cy, (int)x);   // unq row
    if (p.num_points) p.num_points[vid] = (int32_t)cntf;                               // unq_cnt
    const float inv = __frcp_rn(cntf);
    float mean[CT];
#pragma unroll
    for (int k = 0; k < CT; ++k) mean[k] = k < C ? pv_div_count(r[k], cntf, inv) : 0.0f;   // scatter_mean: sum / count
    if (p.feats) pv_store_feats<CT>(p.feats, vid, C, mean);
    pf_st_row_clean<NV>(rowp, 0.0f);
    f.keys[s] = PV_INF;
}

// unq_inv[i] = row of point i's voxel (torch.unique's return_inverse)
__global__ void __launch_bounds__(PF_THREADS) kf_dyn_inverse(const __grid_constant__ PvParams p, const __grid_constant__ PvF f)
{
    pf_pdl_wait();
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * PF_PPT;
    if (i0 >= p.n) return;
    const uint32_t frame_bits = f.wcap * 32u;
#pragma unroll
    for (int j = 0; j < PF_PPT; ++j) {
        const uint32_t i = i0 + j;
        if (i >= p.n) break;
        const uint32_t sa = __ldcs(f.sa + i);
        int32_t v = -1;                                                   // rejected row of a caller-provided grid_ind
        if (sa != PV_INF) {
            const uint2 wv = __ldg(f.wb + (sa >> 5));
            const uint32_t b = sa / frame_bits;
            v = __ldg(f.base + b) + (int32_t)(wv.x + __popc(wv.y & ((1u << (sa & 31u)) - 1u)));
        }
        p.unq_inv[i] = v;
    }
}

// Binning only (Voxelization.voxelize_dynamic, voxelization.py:169-172, + the batch column of collate):
// grid_ind[i] = (b, z, y, x) = floor(clip((p - lo) / vs, 0, grid - 1)) for EVERY point, on any grid --
// no map, no bitmap, so the 3-D Waymo grid (94 M cells) is as good as a pillar grid.
__global__ void __launch_bounds__(256) kf_dyn_grid_ind(const __grid_constant__ PvParams p)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    float v[PV_MAX_CHANNELS];
    pv_feature_row(p.pts, i, p.c_in, p.cart, v);
    const int b = pv_frame_of(p.offsets, p.B, i);
    float c[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float q = pv_bin(v[d], p.lo[d], p.vs[d], p.inv_vs[d]);       // floor of the IEEE quotient
        c[d] = fminf(fmaxf(q, 0.0f), p.gridf[d] - 1.0f);                   // clip (NaN -> 0)
    }
    reinterpret_cast<int4 *>(p.grid_ind)[i] = make_int4(b, (int)c[2], (int)c[1], (int)c[0]);
}

int pvf_run_grid_ind(PvParams &p, cudaStream_t st)
{
    if (p.n == 0) return PV_OK;
    kf_dyn_grid_ind<<<(p.n + 255) / 256, 256, 0, st>>>(p);
    return pv_last_cuda_error();
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
static size_t pf_align(size_t v, size_t a) { return (v + a - 1) / a * a; }

int pvf_make_layout(const pv_config *cfg, int64_t n_cap, int32_t batch, int64_t frame_capacity,
                    int32_t max_channels, void *base, PvF *w)
{
    int rc = pv_check_config(cfg);
    if (rc) return rc;
    if (batch <= 0 || n_cap < 0 || frame_capacity < 0 || n_cap >= (1ll << 30)) return PV_ERR_BAD_ARGUMENT;
    if (max_channels < 3 || max_channels > PV_MAX_CHANNELS) return PV_ERR_BAD_ARGUMENT;
    if (frame_capacity > n_cap) frame_capacity = n_cap;
    const uint64_t cells = (uint64_t)cfg->grid[0] * cfg->grid[1] * cfg->grid[2];
    const bool dense = cells <= PV_DENSE_MAX_CELLS;
    uint64_t capf;
    if (dense) capf = (cells + 3) / 4 * 4;
    else {
        const uint64_t want = (uint64_t)frame_capacity + (uint64_t)frame_capacity / 4 + 1;
        capf = 1024;
        while (capf < want) capf <<= 1;
    }
    if (capf * (uint64_t)batch >= 0xFFFFFF00ull) return PV_ERR_BAD_ARGUMENT;
    const size_t n = (size_t)(n_cap > 0 ? n_cap : 1);
    const size_t fcap = (size_t)(frame_capacity > 0 ? frame_capacity : 1);
    const size_t slots = (size_t)capf * batch;
    w->capf = (uint32_t)capf;
    w->dense = dense ? 1u : 0u;
    // bitmap words per frame: one bit per point (static path) or per cell (dynamic path, direct maps)
    // dynamic voxelization keeps a CELL-order occupancy bitmap: direct maps always, hash-map grids up to
    // PV_DYN_MAX_CELLS cells per frame (the reference's 3-D cylinder grids: 640 x 640 x 40, 1024 x 1024 x 40)
    const bool dyn_cells = dense || cells <= PV_DYN_MAX_CELLS;
    w->dyn_ok = dyn_cells ? 1u : 0u;
    const size_t bit_items = dyn_cells && cells > fcap ? (size_t)cells : fcap;
    w->wcap = (uint32_t)(((bit_items + 31) / 32 + 3) / 4 * 4);
    w->rowf_cap = (uint32_t)((max_channels + 1 + 3) / 4 * 4);
    w->rowf = w->rowf_cap;
    w->hmax = (uint32_t)(n / ((size_t)cfg->max_points + 1) + 1);
    char *p0 = (char *)base;
    size_t o = 0;
    // ---- clean = 0 ----
    w->ctrl = (uint32_t *)(p0 + o);        o = pf_align(o + 16 * 4, 256);
    // static path: one bit per point of the batch (global index); dynamic path: per-frame cell bitmaps
    const size_t bit_words = std::max((size_t)w->wcap * batch, (n + 31) / 32 + 8);
    w->bits = (uint32_t *)(p0 + o);        o = pf_align(o + bit_words * 4, 256);
    w->hbits = (uint32_t *)(p0 + o);       o = pf_align(o + (slots / 32 + 2) * 4, 256);
    w->acc = (float *)(p0 + o);            o = pf_align(o + slots * w->rowf_cap * 4, 256);
    // ---- clean = all ones ----
    w->first = (uint32_t *)(p0 + o);       o = pf_align(o + slots * 4 + 16, 256);
    w->keys = (uint32_t *)(p0 + o);        if (!dense) o = pf_align(o + slots * 4, 256);
    // ---- no clean state ----
    w->hlist = (uint32_t *)(p0 + o);       o = pf_align(o + (n + 32) * 4, 256);
    w->hinfo = (uint4 *)(p0 + o);          o = pf_align(o + (size_t)w->hmax * 32, 256);
    w->base = (int32_t *)(p0 + o);         o = pf_align(o + (size_t)(batch + 1) * 4, 256);
    w->counts_raw = (uint32_t *)(p0 + o);  o = pf_align(o + (size_t)batch * 4, 256);
    w->sa = (uint32_t *)(p0 + o);          o = pf_align(o + n * 4 + 32, 256);
    w->wb = (uint2 *)(p0 + o);             o = pf_align(o + bit_words * 8, 256);
    w->max_chunks = (uint32_t)((n >> 5) / PF_CHUNK_WORDS + 2);
    w->cagg = (uint32_t *)(p0 + o);        o = pf_align(o + (size_t)w->max_chunks * 4, 256);
    w->cbase = (uint32_t *)(p0 + o);       o = pf_align(o + ((size_t)w->max_chunks + 1) * 4, 256);
    w->frank0 = (uint32_t *)(p0 + o);      o = pf_align(o + (size_t)batch * 4, 256);
    w->total_bytes = o;
    return PV_OK;
}

int pvf_init(const PvF &f, int32_t batch, int64_t n_cap, cudaStream_t st)
{
    (void)batch; (void)n_cap;
    const size_t zero_bytes = (size_t)((char *)f.first - (char *)f.ctrl);
    const size_t ff_bytes = (size_t)((char *)f.hlist - (char *)f.first);
    if (cudaMemsetAsync(f.ctrl, 0, zero_bytes, st) != cudaSuccess) return PV_ERR_CUDA;
    if (cudaMemsetAsync(f.first, 0xFF, ff_bytes, st) != cudaSuccess) return PV_ERR_CUDA;
    return PV_OK;
}

// Launch with programmatic stream serialization (see pf_pdl_wait): the kernel may start while the
// previous kernel of the stream is still draining; it synchronises itself with pf_pdl_wait.
template <typename K, typename... Args>
static int pf_launch_pdl(K kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const Args &...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...) == cudaSuccess ? PV_OK : PV_ERR_CUDA;
}

// Tensor map of the canvas [B * C][ny][nx] f32 for kf_finalize_tma's stores: box = C x 32 x 32,
// 128-byte swizzle.  Encoded on the host per call (no device work, no allocation); the driver entry
// point is looked up once (immutable after that).
typedef CUresult (*pf_encode_tiled_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static pf_encode_tiled_t pf_encode_tiled()
{
    static const pf_encode_tiled_t fn = [] {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            sym = nullptr;
        return reinterpret_cast<pf_encode_tiled_t>(sym);
    }();
    return fn;
}

static bool pf_canvas_tmap(const PvParams &p, CUtensorMap *tm)
{
    const pf_encode_tiled_t enc = pf_encode_tiled();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)p.grid[0], (cuuint64_t)p.grid[1], (cuuint64_t)p.B * (cuuint64_t)p.C};
    const cuuint64_t strides[2] = {(cuuint64_t)p.grid[0] * 4ull, (cuuint64_t)p.grid[0] * (cuuint64_t)p.grid[1] * 4ull};
    const cuuint32_t box[3] = {32u, 32u, (cuuint32_t)p.C};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.canvas, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool DENSE, int CIN, bool CART, int NV, int MODE = PF_MODE_FREE>
static int pf_launch_insert(const PvParams &p, const PvF &f, cudaStream_t st)
{
    const unsigned grid = (p.n + PF_TILE - 1) / PF_TILE;
    const size_t smem = (size_t)PF_TILE * p.c_in * sizeof(float);
    // the fused front end never asks for pc_grid_ind: compile that branch out of its kernels
    constexpr bool SPLIT = MODE == PF_MODE_FREE && CIN > 0;
    auto kern = (SPLIT && !p.grid_ind) ? kf_insert<DENSE, CIN, CART, NV, MODE, !SPLIT> : kf_insert<DENSE, CIN, CART, NV, MODE, true>;
    // the 48 KB default covers dynamic + static shared memory (the kernel has a few static words)
    if (smem + 1024 > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return PV_ERR_CUDA;
    kern<<<grid, PF_THREADS, smem, st>>>(p, f);
    return PV_OK;
}

// SM count of the current device (queried, not assumed; cached per device ordinal).
int pv_sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

template <int CIN, bool CART, int NV>
static int pf_launch_insert_stream(const PvParams &p, const PvF &f, cudaStream_t st)
{
    const uint32_t n_tiles = (p.n + 127u) >> 7;
    const uint32_t want = (n_tiles + KI_WARPS - 1) / KI_WARPS, cap = (uint32_t)pv_sm_count() * KI_BLOCKS_PER_SM;
    const uint32_t grid = want < cap ? want : cap;
    const size_t smem = (size_t)KI_WARPS * KI_STAGES * 128 * CIN * sizeof(float);
    kf_insert_lanes<CIN, CART, NV><<<grid, KI_WARPS * 32, smem, st>>>(p, f, n_tiles, grid * KI_WARPS);
    return PV_OK;
}

template <bool DENSE>
static int pf_dispatch_insert(const PvParams &p, const PvF &f, cudaStream_t st)
{
    const int nv = (int)f.rowf / 4;
    // the fused front end's hot path: direct map, no pc_grid_ind, 16-byte aligned rows
    if (DENSE && !p.grid_ind && (reinterpret_cast<uintptr_t>(p.pts) & 15u) == 0) {
        if (p.cart && p.c_in == 5) return pf_launch_insert_stream<5, true, 2>(p, f, st);    // nuScenes (x,y,z,i,dt)
        if (p.cart && p.c_in == 6) return pf_launch_insert_stream<6, true, 3>(p, f, st);    // Waymo (x,y,z,i,e,dt)
        if (p.cart && p.c_in == 4) return pf_launch_insert_stream<4, true, 2>(p, f, st);
        if (!p.cart && p.c_in == 7) return pf_launch_insert_stream<7, false, 2>(p, f, st);
        if (!p.cart && p.c_in == 8) return pf_launch_insert_stream<8, false, 3>(p, f, st);
    }
    if (p.cart && p.c_in == 5) return pf_launch_insert<DENSE, 5, true, 2>(p, f, st);     // nuScenes (x,y,z,i,dt)
    if (p.cart && p.c_in == 6) return pf_launch_insert<DENSE, 6, true, 3>(p, f, st);     // Waymo (x,y,z,i,e,dt)
    if (p.cart && p.c_in == 4) return pf_launch_insert<DENSE, 4, true, 2>(p, f, st);
    if (!p.cart && p.c_in == 7) return pf_launch_insert<DENSE, 7, false, 2>(p, f, st);
    if (!p.cart && p.c_in == 8) return pf_launch_insert<DENSE, 8, false, 3>(p, f, st);
    switch (nv) {
    case 1: return pf_launch_insert<DENSE, 0, false, 1>(p, f, st);
    case 2: return pf_launch_insert<DENSE, 0, false, 2>(p, f, st);
    case 3: return pf_launch_insert<DENSE, 0, false, 3>(p, f, st);
    case 4: return pf_launch_insert<DENSE, 0, false, 4>(p, f, st);
    default: return pf_launch_insert<DENSE, 0, false, 5>(p, f, st);
    }
}

// Binning front end of the list-based pipeline (voxelize.cu): only the three coordinates matter.
template <bool DENSE>
static int pf_dispatch_insert_lists(const PvParams &p, const PvF &f, cudaStream_t st)
{
    if (p.n == 0) return PV_OK;
    if (p.cart && p.c_in == 5) return pf_launch_insert<DENSE, 5, true, 2, PF_MODE_LISTS>(p, f, st);
    if (p.cart && p.c_in == 6) return pf_launch_insert<DENSE, 6, true, 2, PF_MODE_LISTS>(p, f, st);
    if (!p.cart && p.c_in == 7) return pf_launch_insert<DENSE, 7, false, 1, PF_MODE_LISTS>(p, f, st);
    if (!p.cart && p.c_in == 8) return pf_launch_insert<DENSE, 8, false, 1, PF_MODE_LISTS>(p, f, st);
    return p.cart ? pf_launch_insert<DENSE, 0, false, 2, PF_MODE_LISTS>(p, f, st)
                  : pf_launch_insert<DENSE, 0, false, 1, PF_MODE_LISTS>(p, f, st);
}

int pvf_insert_lists(PvParams &p, PvF &f, cudaStream_t st)
{
    return p.ws.dense ? pf_dispatch_insert_lists<true>(p, f, st) : pf_dispatch_insert_lists<false>(p, f, st);
}

static int pf_dispatch_insert_dyn_hash(const PvParams &p, const PvF &f, cudaStream_t st)
{
    switch ((int)f.rowf / 4) {
    case 1: return pf_launch_insert<false, 0, false, 1, PF_MODE_DYN>(p, f, st);
    case 2: return pf_launch_insert<false, 0, false, 2, PF_MODE_DYN>(p, f, st);
    case 3: return pf_launch_insert<false, 0, false, 3, PF_MODE_DYN>(p, f, st);
    case 4: return pf_launch_insert<false, 0, false, 4, PF_MODE_DYN>(p, f, st);
    default: return pf_launch_insert<false, 0, false, 5, PF_MODE_DYN>(p, f, st);
    }
}

template <int NV>
static int pf_launch_dyn_finalize_hash(const PvParams &p, const PvF &f, cudaStream_t st)
{
    return pf_launch_pdl(kf_dyn_finalize_hash<NV, 0>, dim3((f.capf + 255) / 256, (unsigned)p.B), dim3(256), 0, st, p, f);
}

static int pf_dispatch_insert_dyn(const PvParams &p, const PvF &f, cudaStream_t st)
{
    if (p.cart && p.c_in == 5) return pf_launch_insert<true, 5, true, 2, PF_MODE_DYN>(p, f, st);
    if (!p.cart && p.c_in == 7) return pf_launch_insert<true, 7, false, 2, PF_MODE_DYN>(p, f, st);
    switch ((int)f.rowf / 4) {
    case 1: return pf_launch_insert<true, 0, false, 1, PF_MODE_DYN>(p, f, st);
    case 2: return pf_launch_insert<true, 0, false, 2, PF_MODE_DYN>(p, f, st);
    case 3: return pf_launch_insert<true, 0, false, 3, PF_MODE_DYN>(p, f, st);
    case 4: return pf_launch_insert<true, 0, false, 4, PF_MODE_DYN>(p, f, st);
    default: return pf_launch_insert<true, 0, false, 5, PF_MODE_DYN>(p, f, st);
    }
}

template <int NV, int CC>
static void pf_launch_dyn_finalize(const PvParams &p, const PvF &f, cudaStream_t st)
{
    const dim3 grid(((unsigned)p.grid[0] + 255) / 256, (unsigned)p.grid[1] * (unsigned)p.grid[2], (unsigned)p.B);
    if (p.canvas) pf_launch_pdl(kf_dyn_finalize<NV, CC, true>, grid, dim3(256), 0, st, p, f);
    else pf_launch_pdl(kf_dyn_finalize<NV, CC, false>, grid, dim3(256), 0, st, p, f);
}

// Dynamic voxelization launch sequence (direct maps only).
int pvf_run_dynamic(PvParams &p, PvF &f, cudaStream_t st)
{
    f.rowf = (uint32_t)((p.C + 1 + 3) / 4 * 4);
    if (f.rowf > f.rowf_cap) return PV_ERR_WORKSPACE;
    if (!f.dyn_ok || p.B > 65535) return PV_ERR_UNSUPPORTED;            // more than PV_DYN_MAX_CELLS cells per frame
    if ((unsigned long long)p.B * f.wcap * 32ull >= 0xFFFFFF00ull) return PV_ERR_BAD_ARGUMENT;   // bit addresses are 32-bit
    if (!f.dense) {
        // 3-D cylinder grids (voxelnet_det_cylinder_singlehead.py, voxelnet_seg_cylinder.py): rows in the hash
        // map, voxel order from the cell-order bitmap; no dense canvas on these grids
        if (p.canvas) return PV_ERR_BAD_CONFIG;
        if (p.n > 0) {
            const int rc = pf_dispatch_insert_dyn_hash(p, f, st);
            if (rc) return rc;
        }
        if (pf_launch_pdl(kf_scan, dim3((unsigned)p.B), dim3(PF_SCAN_THREADS), 0, st, p, f)) return PV_ERR_CUDA;
        int rc;
        switch ((int)f.rowf / 4) {
        case 1: rc = pf_launch_dyn_finalize_hash<1>(p, f, st); break;
        case 2: rc = pf_launch_dyn_finalize_hash<2>(p, f, st); break;
        case 3: rc = pf_launch_dyn_finalize_hash<3>(p, f, st); break;
        case 4: rc = pf_launch_dyn_finalize_hash<4>(p, f, st); break;
        default: rc = pf_launch_dyn_finalize_hash<5>(p, f, st); break;
        }
        if (rc) return rc;
        if (p.unq_inv && p.n > 0 &&
            pf_launch_pdl(kf_dyn_inverse, dim3((p.n + PF_TILE - 1) / PF_TILE), dim3(PF_THREADS), 0, st, p, f)) return PV_ERR_CUDA;
        return pv_last_cuda_error();
    }
    if ((unsigned)p.grid[1] * (unsigned)p.grid[2] > 65535u) return PV_ERR_UNSUPPORTED;
    if (p.n > 0) {
        const int rc = pf_dispatch_insert_dyn(p, f, st);
        if (rc) return rc;
    }
    if (pf_launch_pdl(kf_scan, dim3((unsigned)p.B), dim3(PF_SCAN_THREADS), 0, st, p, f)) return PV_ERR_CUDA;
    switch ((int)f.rowf / 4) {
    case 1: pf_launch_dyn_finalize<1, 0>(p, f, st); break;
    case 2: if (p.C == 7) pf_launch_dyn_finalize<2, 7>(p, f, st); else pf_launch_dyn_finalize<2, 0>(p, f, st); break;
    case 3: pf_launch_dyn_finalize<3, 0>(p, f, st); break;
    case 4: pf_launch_dyn_finalize<4, 0>(p, f, st); break;
    default: pf_launch_dyn_finalize<5, 0>(p, f, st); break;
    }
    if (p.unq_inv && p.n > 0 &&
        pf_launch_pdl(kf_dyn_inverse, dim3((p.n + PF_TILE - 1) / PF_TILE), dim3(PF_THREADS), 0, st, p, f)) return PV_ERR_CUDA;
    return pv_last_cuda_error();
}

template <int NV, int CC>
static int pf_launch_finalize_nv(const PvParams &p, const PvF &f, cudaStream_t st)
{
    if (!f.dense) {
        return pf_launch_pdl(kf_finalize<NV, CC>, dim3((f.capf + 255) / 256, (unsigned)p.B), dim3(256), 0, st, p, f);
    }
    const unsigned px = ((unsigned)p.grid[0] + PF_PATCH - 1) / PF_PATCH, py = ((unsigned)p.grid[1] + PF_PATCH - 1) / PF_PATCH;
    const dim3 grid(py * (unsigned)p.grid[2], px, (unsigned)p.B);        // px <= 2^15 for a direct map
    if (p.B > 65535 || grid.y > 65535u) return PV_ERR_UNSUPPORTED;
    // canvas without density on a pillar grid: compacted cells + one TMA tensor store per patch
    if (p.canvas && !p.density && p.grid[2] == 1 && (p.grid[0] & 3) == 0 && (reinterpret_cast<uintptr_t>(p.canvas) & 15u) == 0) {
        CUtensorMap tm;
        if (pf_canvas_tmap(p, &tm)) {
            const size_t smem_t = (size_t)p.C * 32 * 32 * sizeof(float) + 1024;
            auto kt = kf_finalize_tma<NV, CC>;
            if (smem_t + 5 * 1024 > 48 * 1024 &&
                cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t) != cudaSuccess)
                return PV_ERR_CUDA;
            return pf_launch_pdl(kt, grid, dim3(256), smem_t, st, p, f, tm);
        }
    }
    const size_t smem = ((p.canvas ? (size_t)p.C : 0) + (p.density ? 1 : 0)) * PF_PATCH * (PF_PATCH + 1) * sizeof(float);
    auto kern = p.canvas ? kf_finalize_patch<NV, CC, true> : kf_finalize_patch<NV, CC, false>;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return PV_ERR_CUDA;
    return pf_launch_pdl(kern, grid, dim3(PF_FIN_WARPS * 32), smem, st, p, f);
}

static int pf_launch_finalize(const PvParams &p, const PvF &f, cudaStream_t st)
{
    switch ((int)f.rowf / 4) {
    case 1: return pf_launch_finalize_nv<1, 0>(p, f, st);
    case 2: return p.C == 7 ? pf_launch_finalize_nv<2, 7>(p, f, st) : pf_launch_finalize_nv<2, 0>(p, f, st);
    case 3: return p.C == 8 ? pf_launch_finalize_nv<3, 8>(p, f, st) : pf_launch_finalize_nv<3, 0>(p, f, st);
    case 4: return pf_launch_finalize_nv<4, 0>(p, f, st);
    default: return pf_launch_finalize_nv<5, 0>(p, f, st);
    }
}

#define PF_MARK(k) do { if (ev && cudaEventRecord(ev[k], st) != cudaSuccess) return PV_ERR_CUDA; } while (0)

// Stage boundaries: 0 insert, 1 cells, 2 scan, 3 finalize, 4 heavy (points + cells).
int pvf_run(PvParams &p, PvF &f, cudaStream_t st, cudaEvent_t *ev)
{
    f.rowf = (uint32_t)((p.C + 1 + 3) / 4 * 4);
    if (f.rowf > f.rowf_cap) return PV_ERR_WORKSPACE;
    if (!f.dense) {     // hash map: slot order is not cell order, dense outputs need a zero fill first
        if (p.canvas && cudaMemsetAsync(p.canvas, 0, (size_t)p.B * p.C * p.cells * sizeof(float), st) != cudaSuccess)
            return PV_ERR_CUDA;
        if (p.density && cudaMemsetAsync(p.density, 0, (size_t)p.B * p.cells * sizeof(int32_t), st) != cudaSuccess)
            return PV_ERR_CUDA;
    }
    PF_MARK(0);
    if (p.n > 0) {
        const int rc = f.dense ? pf_dispatch_insert<true>(p, f, st) : pf_dispatch_insert<false>(p, f, st);
        if (rc) return rc;
    }
    PF_MARK(1);
    PF_MARK(2);                              // (the first-point bitmap is built by the insert kernel)
    if (pf_launch_pdl(kf_scan2, dim3(((p.n >> 5) + PF_CHUNK_WORDS) / PF_CHUNK_WORDS), dim3(PF_SCAN2_THREADS), 0, st, p, f)) return PV_ERR_CUDA;
    PF_MARK(3);
    {
        const int rc = pf_launch_finalize(p, f, st);
        if (rc) return rc;
    }
    PF_MARK(4);
    if (p.n > 0 && KF_EXP != 9) {
        if (pf_launch_pdl(kf_heavy_points, dim3((p.n + PF_THREADS * PF_HPT - 1) / (PF_THREADS * PF_HPT)), dim3(PF_THREADS), 0, st, p, f)) return PV_ERR_CUDA;
        if (pf_launch_pdl(kf_heavy_cells, dim3(296), dim3(256), 0, st, p, f)) return PV_ERR_CUDA;
    }
    PF_MARK(5);
    return pv_last_cuda_error();
}
