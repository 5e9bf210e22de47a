// pfn_fused.cu -- PillarFeatureNet (two PFN layers, eval) as ONE warp-specialised kernel with the
// second layer's GEMM on the 5th-generation tensor cores (sm_100a: tcgen05.mma, accumulators in TMEM).
//
// Reference: PillarFeatureNet.forward / PFNLayer.forward_static, det3d/models/readers/pillar_encoder.py
// :131-169 / :49-61 -- decoration (cluster offset :137-140, pillar-centre offset :145-150, optional
// distance :154-156), padding mask :161-164, then per layer Linear (no bias) -> BatchNorm1d (eval,
// ATen order) -> ReLU -> max over ALL T slots of the voxel.  Padded slots enter as all-zero rows, so
// after the norm they hold relu(shift) != 0 and take part in the max; all padded rows of a voxel are
// identical, so ONE representative row per non-full voxel is evaluated ("useful rows").
//
// Two kernels.  k_pfn_rows<SRC> (full occupancy) gathers and decorates the useful rows from one of two sources and
// writes them as float4 planes + one descriptor per group of <= 32 rows:
//   SRC 0  the padded tensor voxels [M, T, C] of the drop-in reader (pv_pfn_forward), in slices of P2_SLICE voxels;
//   SRC 1  the per-voxel POINT LISTS of the list-based voxelizer (voxelize.cu: kept / vox_kg / vox_c / vox_cell)
//          + the raw point rows -- the [M, T, C] tensor is never materialised; also writes coors / num_points
//          and restores the lists (pv_forward_pfn_canvas).
// k_pfn_fused<C0Q>: one persistent CTA per SM, 25 warps:
//   warps 0-7   PRODUCERS, two sets of four (set = A-operand stage).  A warp takes one GROUP (lane = row): loads its
//               decorated rows (requested one group ahead), runs layer 0 (K = C + 5 <= 16, 4 % of the FLOPs) as
//               packed fp32 FMAs with the folded BatchNorm + ReLU -- all before it needs the operand stage --, then
//               writes the x0 half of the layer-1 operand row, takes the per-voxel maximum by segmented warp scans
//               in the same registers and writes the x_max0 half; both split into TF32 hi / lo parts, in the
//               canonical K-major UMMA layout.  Four groups = one 128-row MMA tile.
//   warp 24     ISSUER: one thread issues the 3 x K/8 tcgen05.mma.kind::tf32 of the tile (3xTF32 split:
//               lo.hi + hi.lo + hi.hi, fp32-accurate) into one of FOUR TMEM accumulator stages and commits to two
//               mbarriers (operand stage -> producers, accumulator -> epilogue team).  The GEMM is issued
//               TRANSPOSED, D[unit, row] = W1 . X^T (the weight is the M-side operand), so that TMEM lane = output
//               unit, column = row.
//   warps 8-23  EPILOGUE, two teams of eight (a team serves every other round): warp (e, half) owns TMEM lanes
//               [32 e, 32 e + 32) = 32 units and two groups of the tile.  A thread holds ONE unit: its BatchNorm
//               scale / shift sit in two registers, the rows of a voxel are consecutive columns, so the per-voxel
//               maximum is a running FMNMX down the row -- x -> relu(|s| x + b) is monotone (the sign of s went
//               into the weight row) -- and BatchNorm + ReLU are applied once per voxel.  One coalesced 128-byte
//               store per voxel.
// Pipelines on mbarriers (operand full / operand consumed / accumulator ready / records valid / accumulator free),
// no block-wide barrier after the prologue; see the barrier protocol in DESIGN.md section 3.5.
#include "tc_common.cuh"
#include "pfn_fused.cuh"

// warps 0-15 producers (two sets of eight: set = operand stage; TWO warps per group of 32 rows, each computing 16 of
// layer 0's 32 units for the group's rows) | warps 16-27 epilogue in three TEAMS of four (team = tile index mod 3; a
// warp = one TMEM quarter = 32 units, all four groups of the tile) | warp 28 issuer.
// Four TMEM accumulator stages (acc = tile & 3, tile = 2 * round + set) against two operand stages in shared memory:
// an operand stage is free again when its MMAs retire, an accumulator stage only when an epilogue team has drained it
// -- and a lone warp retires an instruction every ~8 cycles, so draining a tile takes a team of four ~4 us.  With two
// accumulator stages that time was on the critical path of every tile (traced: issue 1.4 us -> MMA + wake-up 2.0 us ->
// epilogue 3.0 us -> next issue); with four stages three teams drain three tiles at once (1.3 us per tile), and two
// warps per group halve the producers' serial instruction stream (the tile period was theirs: 2.8 us).
#ifdef P2_TRACE          // development builds only: per-warp time stamps of block 0 (tools/pfn_trace.py)
__device__ long long g_p2_trace[32 * 64 * 2 * 3];
__device__ long long g_p2_trace2[16 * 64 * 2 * 4];   // (unused by the three-team layout)
#define P2_STAMP(round, s, k) do { if (blockIdx.x == 0 && lane == 0 && (round) < 64u) g_p2_trace[((warp * 64 + (round)) * 2 + (s)) * 3 + (k)] = clock64(); } while (0)
#define P2_STAMP2(round, s, k) do { if (blockIdx.x == 0 && lane == 0 && (round) < 64u) g_p2_trace2[(((warp - P2_EPI_WARP0) * 64 + (round)) * 2 + (s)) * 4 + (k)] = clock64(); } while (0)
extern "C" int pv_debug_p2_trace(void *dst, size_t bytes)
{
    return cudaMemcpyFromSymbol(dst, g_p2_trace, bytes < sizeof(g_p2_trace) ? bytes : sizeof(g_p2_trace)) == cudaSuccess ? 0 : -4;
}
extern "C" int pv_debug_p2_trace2(void *dst, size_t bytes)
{
    return cudaMemcpyFromSymbol(dst, g_p2_trace2, bytes < sizeof(g_p2_trace2) ? bytes : sizeof(g_p2_trace2)) == cudaSuccess ? 0 : -4;
}
#else
#define P2_STAMP(round, s, k) do { } while (0)
#define P2_STAMP2(round, s, k) do { } while (0)
#endif
#define P2_NPROD 16
#define P2_EPI_WARP0 16
#define P2_ISSUER_WARP 28
#define P2_THREADS (29 * 32)
// (measured: 4 epilogue warps and 128 registers per thread: 1.60 ms vs 1.26 ms)
// (No setmaxnreg: it moves registers inside the pool the CTA was LAUNCHED with -- threads x the kernel's register
// count -- and an .inc that pool cannot satisfy blocks for ever.  Every role fits the launch count since the
// producers keep one live array of 32 values.)
#ifndef P2_WAIT_HINT_NS
#define P2_WAIT_HINT_NS 20000
#endif
#ifndef P2_MC
#define P2_MC 64               // voxels per mini-chunk (the unit a producer warp fetches); <= 64
#endif
#define P2_U0 32               // units of layer 0
#define P2_K 64                // K of layer 1 = 2 * P2_U0
#define P2_C0 16               // decorated input width, padded

struct P2Meta {                // what the epilogue needs to know about one group
    int32_t vid0;              // output row of the group's first voxel (the following voxels take the following rows)
    int32_t pad0[3];
    uint32_t lasts;            // bit k: row k is the last row of its voxel
    uint32_t done;             // the producer ran out of work: nothing to drain
    uint32_t pad[2];
};

// Barrier wait with a watchdog: a pipeline bug must not hang the GPU.  After ~0.2 s of waiting the
// block raises its abort flag, records which wait starved (tag) next to the chunk counter and every
// role loop drains; the block sets bit 2 of the status word, which pv_read_status reports as PV_ERR_INTERNAL.
__device__ __forceinline__ bool p2_mbar_wait(uint32_t bar, uint32_t parity, volatile uint32_t *abort_flag,
                                             unsigned int *diag, uint32_t tag)
{
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return true;
    // not there yet: sleep between polls (a spinning warp costs the other roles of the SM their issue
    // slots: 41 % of all executed instructions in the first profiles, also with a fixed 100 ns nap); the
    // watchdog looks at the clock every 64 polls
    const long long t0 = clock64();
    for (uint32_t spins = 1;; ++spins) {
        // try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes or the
        // hint runs out, instead of polling (polls were 27-41 % of all executed instructions)
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity), "r"((uint32_t)P2_WAIT_HINT_NS) : "memory");
        if (done) return true;
        if ((spins & 15u) == 0) {
            if (*abort_flag) { if (blockIdx.x == diag[1]) diag[2 + (threadIdx.x >> 5)] = tag; return false; }
            if (clock64() - t0 > 400000000ll) {
                if (atomicCAS(diag, 0u, tag) == 0u) diag[1] = blockIdx.x;      // first block to starve: its warps report where they wait
                __threadfence();
                *abort_flag = 1u;
                if (blockIdx.x == diag[1]) diag[2 + (threadIdx.x >> 5)] = tag;
                return false;
            }
        }
    }
}
__device__ __forceinline__ void p2_mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// packed fp32 pairs (sm_100: FFMA2 -- two IEEE fp32 FMAs per instruction, bit-identical to two scalar FMAs)
__device__ __forceinline__ unsigned long long p2_pack(float a, float b)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void p2_unpack(unsigned long long v, float &a, float &b)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long p2_fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// segmented inclusive scan step helpers: flags bit d set <=> lane >= 2^d and lanes (lane - 2^d, lane]
// hold no first row of a voxel, i.e. lane - 2^d belongs to the same voxel
// (whole arrays at a time: the NV shuffles of a step are independent, so they pipeline)
template <int NV>
__device__ __forceinline__ void p2_seg_max(float (&v)[NV], uint32_t flags, uint32_t nsteps)
{
    for (uint32_t d = 0; d < nsteps; ++d) {
        const bool take = (flags >> d) & 1u;
        float o[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) o[k] = __shfl_up_sync(0xffffffffu, v[k], 1u << d);
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] = take ? fmaxf(v[k], o[k]) : v[k];
    }
}
template <int NV>
__device__ __forceinline__ void p2_seg_sum(float (&v)[NV], uint32_t flags, uint32_t nsteps)
{
    for (uint32_t d = 0; d < nsteps; ++d) {
        const bool take = (flags >> d) & 1u;
        float o[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) o[k] = __shfl_up_sync(0xffffffffu, v[k], 1u << d);
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] = take ? __fadd_rn(v[k], o[k]) : v[k];
    }
}

// ---------------------------------------------------------------------------------------------
// k_pfn_rows -- the gather half of the fused front end (pv_forward_pfn_canvas), at full occupancy.
// The dependent chain voxel record -> list entry -> point row -> cluster mean -> decoration took the
// eight producer warps of k_pfn_fused ~11 us per 128 rows (one block per SM, nothing to hide the
// three round trips behind); here every SM runs 48 warps of it.  A warp takes a mini-chunk of 64
// voxels, packs whole voxels into groups of <= 32 rows (lane = row) exactly as the tensor-core
// kernel wants them, and writes per group
//   drows[k4][row_start + lane]  the decorated row, (C + 5 (+1)) floats as up to four float4 PLANES (a
//                            warp's store / load covers 512 contiguous bytes); the representative
//                            padded row of a non-full voxel is all zeros (:161-164)
//   desc[chunk * 64 + g]     {row_start, first output row, head mask, total | scan steps << 8}
// and per chunk ngroups[chunk].  Rows are allocated per chunk with one atomic on counter[40].
// Also writes coors / num_points of the voxelizer and restores its point lists.
// ---------------------------------------------------------------------------------------------
#define P2_ROWS_THREADS 256
#ifndef P2R_EXP
#define P2R_EXP 0      // timing experiments (never defined in product builds)
#endif
#ifndef P2_ROWS_BLOCKS
#define P2_ROWS_BLOCKS 5
#endif
template <int SRC>          // 1: point lists of the list-based voxelizer (fused front end); 0: the padded tensor [M, T, C] (pv_pfn_forward)
__global__ void __launch_bounds__(P2_ROWS_THREADS, P2_ROWS_BLOCKS) k_pfn_rows(const __grid_constant__ P2Args a)
{
    const int lane = threadIdx.x & 31;
    const uint32_t n_warps = gridDim.x * (P2_ROWS_THREADS / 32);
    const int T = a.t, C = a.c, c0q = (a.c0 + 3) >> 2;
    for (uint32_t id = blockIdx.x * (P2_ROWS_THREADS / 32) + (threadIdx.x >> 5); id < a.n_chunks; id += n_warps) {
        // SRC 1: chunk id -> (frame b, voxels [r0, r0 + 64) of the frame); SRC 0: one "frame" = voxels [v0, v0 + m) of the tensor
        const int b = SRC ? (int)(id / a.chunks_per_frame) : 0;
        const uint32_t r0 = SRC ? (id - (uint32_t)b * a.chunks_per_frame) * P2_MC : (uint32_t)a.v0 + id * P2_MC;
        const uint32_t cnt = SRC ? (uint32_t)__ldg(a.voxel_counts + b) : (uint32_t)(a.v0 + a.m);
        const uint32_t v_end = min(cnt, r0 + P2_MC);
        uint32_t v_next = r0, ng = 0;
        if (v_next >= v_end) {
            if (lane == 0) a.ngroups_out[id] = 0u;
            continue;
        }
        // records of the chunk's 64 voxels (two windows), rows of the chunk -> one allocation
        int n_w[2];
        uint32_t kg_w[2], cell_w[2];
        int rows_chunk = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t vi = r0 + 32u * h + lane;
            n_w[h] = 0; kg_w[h] = 0; cell_w[h] = 0;
            if (vi < v_end) {
                if (SRC) {
                    const size_t v = (size_t)b * a.fcap + vi;
                    n_w[h] = (int)min(__ldcs(a.vox_c + v), (uint32_t)T);
                    kg_w[h] = __ldcs(a.vox_kg + v);
                    cell_w[h] = __ldcs(a.vox_cell + v);
                } else {                                   // kg_w / cell_w carry the pillar's (y, x) of coors (b, z, y, x)
                    n_w[h] = min(max(__ldg(a.num + vi), 0), T);
                    const int4 co = __ldg(reinterpret_cast<const int4 *>(a.coors_in) + vi);
                    kg_w[h] = (uint32_t)co.z; cell_w[h] = (uint32_t)co.w;
                }
                rows_chunk += n_w[h] < T ? n_w[h] + 1 : T;
            }
        }
        rows_chunk = __reduce_add_sync(0xffffffffu, rows_chunk);
        uint32_t row_cursor = 0;
        if (lane == 0) row_cursor = atomicAdd(a.counter + 40, (uint32_t)rows_chunk);
        row_cursor = __shfl_sync(0xffffffffu, row_cursor, 0);
        const int vid0_chunk = SRC ? __ldg(a.base + b) + (int)r0 : (int)r0;
        while (v_next < v_end) {
            // ---- pack whole voxels into <= 32 rows: lane i looks at voxel v_next + i ----
            const uint32_t rel = v_next - r0 + lane;                          // < 64 + 31
            const int n_lo = __shfl_sync(0xffffffffu, n_w[0], rel & 31u), n_hi = __shfl_sync(0xffffffffu, n_w[1], rel & 31u);
            const uint32_t kg_lo = __shfl_sync(0xffffffffu, kg_w[0], rel & 31u), kg_hi = __shfl_sync(0xffffffffu, kg_w[1], rel & 31u);
            const uint32_t ce_lo = __shfl_sync(0xffffffffu, cell_w[0], rel & 31u), ce_hi = __shfl_sync(0xffffffffu, cell_w[1], rel & 31u);
            const uint32_t vi = v_next + lane;
            const bool in_chunk = vi < v_end;
            const int n_i = in_chunk ? (rel < 32u ? n_lo : n_hi) : 0;
            const uint32_t kg_i = rel < 32u ? kg_lo : kg_hi, cell_i = rel < 32u ? ce_lo : ce_hi;
            int4 co_i;
            if (SRC) {
                const uint32_t cxi = cell_i % (uint32_t)a.nx, yz = cell_i / (uint32_t)a.nx;
                co_i = make_int4(b, (int)(yz / (uint32_t)a.ny), (int)(yz % (uint32_t)a.ny), (int)cxi);
            } else co_i = make_int4(0, 0, (int)kg_i, (int)cell_i);
            const int rows_i = in_chunk ? (n_i < T ? n_i + 1 : T) : 0;
            int incl = rows_i;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            const uint32_t fits = __ballot_sync(0xffffffffu, in_chunk && incl <= 32);
            const int nv = __popc(fits);                                      // >= 1: a voxel has at most 32 rows
            const int start_i = incl - rows_i;
            const uint32_t heads = __reduce_or_sync(0xffffffffu, lane < nv ? 1u << start_i : 0u);
            const int total = __shfl_sync(0xffffffffu, incl, nv - 1);
            const int maxlen = __reduce_max_sync(0xffffffffu, lane < nv ? rows_i : 0);
            uint32_t nsteps = 0;
            while ((1 << nsteps) < maxlen) ++nsteps;
            // ---- lane = row: which voxel, which slot ----
            const bool row_ok = lane < total;
            const int j = __popc(heads & (0xFFFFFFFFu >> (31 - lane))) - 1;
            const int jj = row_ok ? j : 0;
            const int start_j = __shfl_sync(0xffffffffu, start_i, jj);
            const int n_j = __shfl_sync(0xffffffffu, n_i, jj);
            const int rows_j = __shfl_sync(0xffffffffu, rows_i, jj);
            const uint32_t kg_j = __shfl_sync(0xffffffffu, kg_i, jj);
            const int cx_j = __shfl_sync(0xffffffffu, co_i.w, jj), cy_j = __shfl_sync(0xffffffffu, co_i.z, jj);
            const int q = lane - start_j;
            const bool valid = row_ok && q < n_j;                             // a real point (not the padded representative)
            uint32_t flags = 0;
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                const int dist = 1 << d;
                const uint32_t span = lane >= dist ? (0xFFFFFFFFu >> (31 - lane)) & ~(0xFFFFFFFFu >> (31 - (lane - dist))) : 0xFFFFFFFFu;
                if (row_ok && lane >= dist && (heads & span) == 0u) flags |= 1u << d;
            }
            const int vid_i = vid0_chunk + (int)(v_next - r0) + lane;
            const int last_lane = start_j + rows_j - 1;
            float f[PV_MAX_CHANNELS];
#pragma unroll
            for (int k = 0; k < PV_MAX_CHANNELS; ++k) f[k] = 0.0f;
            if (valid) {
                if (SRC) {
                    uint32_t idx = __ldcg(a.kept + kg_j + q);
                    if (P2R_EXP == 1) idx = kg_j + q;                         // experiment: no random gather
                    pv_feature_row(a.pts, idx, a.c_in, a.cart, f);
                } else {
                    const float *src = a.voxels + ((size_t)(vid_i - lane + j) * T + q) * C;   // voxel of this row = first voxel of the group + j
#pragma unroll
                    for (int k = 0; k < PV_MAX_CHANNELS; ++k)
                        if (k < C) f[k] = __ldg(src + k);
                }
            }
            if (SRC && lane < nv) {                                           // per-voxel outputs of the voxelizer
                __stcs(reinterpret_cast<int4 *>(a.coors_out) + vid_i, co_i);
                __stcs(a.num_out + vid_i, n_i);
            }
            // cluster mean (:137-139): sum over the voxel's rows / num
            float sm[3] = {f[0], f[1], f[2]};
            p2_seg_sum<3>(sm, flags, nsteps);
            const float sx = __shfl_sync(0xffffffffu, sm[0], last_lane & 31), sy = __shfl_sync(0xffffffffu, sm[1], last_lane & 31),
                        sz = __shfl_sync(0xffffffffu, sm[2], last_lane & 31);
            const float nf = (float)n_j;
            const float mx = __fdiv_rn(sx, nf), my = __fdiv_rn(sy, nf), mz = __fdiv_rn(sz, nf);
            const float pcx = __fadd_rn(__fmul_rn((float)cx_j, a.vx), a.x_off);   // :146-147
            const float pcy = __fadd_rn(__fmul_rn((float)cy_j, a.vy), a.y_off);   // :149-150
            float in[P2_C0];
#pragma unroll
            for (int k = 0; k < P2_C0; ++k) {
                float val = 0.0f;
                if (k < C) val = f[k < PV_MAX_CHANNELS ? k : 0];
                else if (k == C) val = __fsub_rn(f[0], mx);                   // :140
                else if (k == C + 1) val = __fsub_rn(f[1], my);
                else if (k == C + 2) val = __fsub_rn(f[2], mz);
                else if (k == C + 3) val = __fsub_rn(f[0], pcx);
                else if (k == C + 4) val = __fsub_rn(f[1], pcy);
                else if (k == C + 5 && a.with_distance)
                    val = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(f[0], f[0]), __fmul_rn(f[1], f[1])), __fmul_rn(f[2], f[2])));   // :155
                in[k] = valid ? val : 0.0f;                                   // :161-164 mask: padded rows are zero
            }
            if (row_ok && P2R_EXP != 2) {
                // four planes of float4 columns: the 32 lanes of a store write 512 contiguous bytes
                float4 *dst = a.drows_out + (size_t)row_cursor + lane;
#pragma unroll
                for (int k4 = 0; k4 < P2_C0 / 4; ++k4)
                    if (k4 < c0q) __stcg(dst + (size_t)k4 * a.drow_stride, make_float4(in[4 * k4], in[4 * k4 + 1], in[4 * k4 + 2], in[4 * k4 + 3]));
            }
            // restore the list for the next call -- only now: a store to an address whose load is still in
            // flight stalls the load/store unit for the whole round trip (measured here: 440 -> 240 us)
            if (SRC && valid && P2R_EXP != 3) a.kept[kg_j + q] = PV_INF;
            if (lane == 0)
                __stcg(a.desc_out + (size_t)id * P2_MC + ng, make_uint4(row_cursor, (uint32_t)vid_i, heads, (uint32_t)total | (nsteps << 8)));
            row_cursor += (uint32_t)total;
            ++ng;
            v_next += (uint32_t)nv;
        }
        if (lane == 0) a.ngroups_out[id] = ng;
    }
}

template <int C0Q>          // float4s per decorated row (compile-time: the row registers are live across a hand-off)
__global__ void __launch_bounds__(P2_THREADS, 1) k_pfn_fused(const __grid_constant__ P2Args a)
{
    extern __shared__ __align__(128) float smem[];
    const int N = a.n1;
    float *a_st = smem;                                    // [2 stages][hi | lo][128 x 64] canonical
    float *b_hi = a_st + 2 * 2 * TC_M * P2_K;              // [128 x 64] canonical: layer-1 weight, units >= N are zero
    float *b_lo = b_hi + TC_M * P2_K;
    float *w0t = b_lo + TC_M * P2_K;                       // [P2_C0][P2_U0]: layer-0 weight, transposed
    float *bn0 = w0t + P2_C0 * P2_U0;                      // mean, invstd, gamma, beta: 4 x 32
    float *bn1 = bn0 + 4 * P2_U0;                          // 4 x N
    P2Meta *meta = reinterpret_cast<P2Meta *>(bn1 + 4 * N);   // [2 sets][4 rounds in flight][4 groups]
    uint4 *s_desc = reinterpret_cast<uint4 *>(meta + 32);      // [8 producer pairs][2][64]: group descriptors of the pair's chunk
    __shared__ __align__(8) unsigned long long s_full[2], s_mma[2], s_acc[4], s_free[4], s_rec[4];
    __shared__ uint32_t s_tmem, s_exit[2], s_chunk[8][2];
    __shared__ uint32_t s_abort;
    unsigned int *diag = a.counter + 2;     // diagnostic words of the watchdog (0 = healthy)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t ncols = 512;                             // four accumulator stages x 128 rows (columns): all of TMEM

    // ---- prologue: weights, BatchNorm constants, barriers, TMEM ----
    if (warp == P2_ISSUER_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc_smem_u32(&s_tmem)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            tc_mbar_init(tc_smem_u32(&s_full[s]), 8);
            tc_mbar_init(tc_smem_u32(&s_mma[s]), 1);
            s_exit[s] = s == 0 ? 0u : 0xFFFFFFFFu;      // [0] producers: stop; [1] epilogue: first tile index that is not real
        }
        for (int q = 0; q < 4; ++q) {
            tc_mbar_init(tc_smem_u32(&s_acc[q]), 1);
            tc_mbar_init(tc_smem_u32(&s_free[q]), 4);
            tc_mbar_init(tc_smem_u32(&s_rec[q]), 1);
        }
        s_abort = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = tid; e < TC_M * P2_K; e += P2_THREADS) {  // linear.weight of layer 1 is [N, 64], K-major already
        const int r = e / P2_K, k = e - r * P2_K;
        float hi = 0.0f, lo = 0.0f;
        // units whose folded BatchNorm scale is negative get their weight row negated: the epilogue then
        // always tracks the MAXIMUM of the raw accumulator (relu(s a + b) = relu(|s| (-a) + b) for s < 0)
        if (r < N) tc_split(__ldg(a.gamma1 + r) < 0.0f ? -__ldg(a.w1 + e) : __ldg(a.w1 + e), hi, lo);
        const uint32_t o = tc_canon(r, k, TC_M);
        b_hi[o] = hi; b_lo[o] = lo;
    }
    for (int e = tid; e < P2_C0 * P2_U0; e += P2_THREADS) {
        const int k = e / P2_U0, u = e - k * P2_U0;
        w0t[e] = k < a.c0 ? __ldg(a.w0 + u * a.c0 + k) : 0.0f;
    }
    // both BatchNorms are folded to one FMA per element, y = x * s + b with s = invstd * gamma and
    // b = beta - mean * s (moves y by ulps against ATen's (x - mean) * invstd * gamma + beta)
    for (int o = tid; o < P2_U0; o += P2_THREADS) {
        const float sc = __fmul_rn(__fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(a.var0[o], a.eps))), a.gamma0[o]);
        bn0[o] = sc;
        bn0[P2_U0 + o] = __fsub_rn(a.beta0[o], __fmul_rn(a.mean0[o], sc));
    }
    for (int o = tid; o < N; o += P2_THREADS) {
        const float sc = __fmul_rn(__fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(a.var1[o], a.eps))), a.gamma1[o]);
        bn1[2 * o] = fabsf(sc);                              // the sign went into the weight row
        bn1[2 * o + 1] = __fsub_rn(a.beta1[o], __fmul_rn(a.mean1[o], sc));
    }
    for (int o = tid; o < 32; o += P2_THREADS) { meta[o].lasts = 0u; meta[o].done = 1u; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    // (No setmaxnreg: it moves registers inside the pool the CTA was LAUNCHED with -- threads x the kernel's register
    // count -- and an .inc that pool cannot satisfy blocks for ever.  Every role fits the launch count.)
    if (warp < P2_EPI_WARP0) {
        // =====================================================================================
        // PRODUCER warp: set = operand stage, g = group inside the tile, uh = first of its 16 layer-0 units.
        // The two warps of a pair walk the same groups (same chunk, same descriptors, same rows).
        // =====================================================================================
        const int set = warp >> 3, g = (warp >> 1) & 3, pair = warp >> 1, uh = (warp & 1) * (P2_U0 / 2);
        constexpr int UH = P2_U0 / 2;
        float *a_hi = a_st + (size_t)set * 2 * TC_M * P2_K, *a_lo = a_hi + TC_M * P2_K;
        const uint32_t bar_full = tc_smem_u32(&s_full[set]), bar_mma = tc_smem_u32(&s_mma[set]);
        uint4 *my_desc = s_desc + pair * 2 * P2_MC;       // descriptors of the pair's chunk (written by its first warp), double-buffered:
                                                          // the first warp may stage chunk k + 1 while its partner still reads chunk k
        uint32_t ng = 0, gi = 0, n_fetch = 0;
        // the NEXT group's descriptor and rows: requested one group ahead, right after layer 0 of the current
        // group has consumed the row registers, so that the loads travel while the warp waits for its operand
        // stage and writes it (traced: 1-2 us of exposed load latency per group without this)
        // (measured and rejected: ONE queue of groups instead of 64-voxel chunks per warp -- ticket, descriptor and
        // rows become three dependent round trips per group: fused front end 1.674 vs 1.658 ms)
        uint4 dsc_n = make_uint4(0, 0, 0, 0);
        float4 in4[C0Q];
        bool oow_n = false;
        auto fetch_next = [&]() {
            while (!oow_n && gi >= ng) {
                // the pair's first warp takes the next chunk from the queue and stages its descriptors; a named
                // barrier of the pair's 64 threads publishes them (slot parity: the partner may still read the old id)
                const uint32_t slot = n_fetch & 1u;
                ++n_fetch;
                my_desc = s_desc + (pair * 2 + slot) * P2_MC;
                if ((warp & 1) == 0) {
                    uint32_t id = 0;
                    if (lane == 0) id = atomicAdd(a.counter, 1u);
                    id = __shfl_sync(0xffffffffu, id, 0);
                    uint32_t cnt = 0;
                    if (id < a.n_chunks) {
                        cnt = __ldcs(a.ngroups + id);
                        if ((uint32_t)lane < cnt) my_desc[lane] = __ldcs(a.desc + (size_t)id * P2_MC + lane);       // (only what k_pfn_rows wrote)
                        if (32u + lane < cnt) my_desc[32 + lane] = __ldcs(a.desc + (size_t)id * P2_MC + 32 + lane);
                    } else cnt = 0xFFFFFFFFu;               // the queue is empty
                    if (lane == 0) s_chunk[pair][slot] = cnt;
                }
                asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
                const uint32_t cnt = *reinterpret_cast<volatile uint32_t *>(&s_chunk[pair][slot]);
                if (cnt == 0xFFFFFFFFu) { oow_n = true; break; }
                ng = cnt;
                gi = 0;
            }
            dsc_n = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int k4 = 0; k4 < C0Q; ++k4) in4[k4] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (!oow_n) {
                dsc_n = my_desc[gi];
                ++gi;
                if ((uint32_t)lane < (dsc_n.w & 0xffu)) {
                    const float4 *src4 = a.drows + (size_t)dsc_n.x + lane;
#pragma unroll
                    for (int k4 = 0; k4 < C0Q; ++k4) in4[k4] = __ldg(src4 + (size_t)k4 * a.drow_stride);   // (the partner reads the same rows)
                }
            }
        };
        fetch_next();
        for (uint32_t round = 0;; ++round) {
            const uint4 dsc = dsc_n;
            const bool out_of_work = oow_n;
            // ---- everything that does not touch the operand stage comes BEFORE the wait for it: the group's
            // segment structure and this warp's half of layer 0 (Linear as packed FFMA2 pairs -> folded BatchNorm -> ReLU) ----
            const uint32_t heads = dsc.z, total = dsc.w & 0xffu, nsteps = (dsc.w >> 8) & 0xffu;
            const bool row_ok = (uint32_t)lane < total;
            const uint32_t above = lane < 31 ? heads & (0xFFFFFFFEu << lane) : 0u;   // heads in lanes > lane
            const int last_lane = (above ? __ffs(above) - 1 : (int)total) - 1;
            const bool seg_last = row_ok && lane == last_lane;
            uint32_t flags = 0;
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                const int dist = 1 << d;
                const uint32_t span = lane >= dist ? (0xFFFFFFFFu >> (31 - lane)) & ~(0xFFFFFFFFu >> (31 - (lane - dist))) : 0xFFFFFFFFu;
                if (row_ok && lane >= dist && (heads & span) == 0u) flags |= 1u << d;
            }
            float x0[UH];
            {
                unsigned long long acc[UH / 2];
#pragma unroll
                for (int u = 0; u < UH / 2; ++u) acc[u] = 0ull;
#pragma unroll
                for (int k4 = 0; k4 < C0Q; ++k4) {
                    const float in[4] = {in4[k4].x, in4[k4].y, in4[k4].z, in4[k4].w};
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const ulonglong2 *wr = reinterpret_cast<const ulonglong2 *>(w0t + (4 * k4 + kk) * P2_U0 + uh);
                        const unsigned long long xin = p2_pack(in[kk], in[kk]);
#pragma unroll
                        for (int u4 = 0; u4 < UH / 4; ++u4) {
                            const ulonglong2 w = wr[u4];
                            acc[2 * u4] = p2_fma2(xin, w.x, acc[2 * u4]);
                            acc[2 * u4 + 1] = p2_fma2(xin, w.y, acc[2 * u4 + 1]);
                        }
                    }
                }
                const ulonglong2 *bsc = reinterpret_cast<const ulonglong2 *>(bn0 + uh), *bsh = reinterpret_cast<const ulonglong2 *>(bn0 + P2_U0 + uh);
#pragma unroll
                for (int u4 = 0; u4 < UH / 4; ++u4) {
                    const ulonglong2 sc2 = bsc[u4], sh2 = bsh[u4];
                    float y0, y1, y2, y3;
                    p2_unpack(p2_fma2(acc[2 * u4], sc2.x, sh2.x), y0, y1);
                    p2_unpack(p2_fma2(acc[2 * u4 + 1], sc2.y, sh2.y), y2, y3);
                    x0[4 * u4] = row_ok ? fmaxf(y0, 0.0f) : 0.0f; x0[4 * u4 + 1] = row_ok ? fmaxf(y1, 0.0f) : 0.0f;
                    x0[4 * u4 + 2] = row_ok ? fmaxf(y2, 0.0f) : 0.0f; x0[4 * u4 + 3] = row_ok ? fmaxf(y3, 0.0f) : 0.0f;
                }
            }
            if (!out_of_work) fetch_next();                  // the row registers are free again
            P2_STAMP(round, 0, 0);
            if (round > 0) {
                if (!p2_mbar_wait(bar_mma, (round - 1) & 1u, &s_abort, diag, 0x100u | (set << 4) | g | (round << 16))) break;   // the stage's previous tile has been consumed
                if (*reinterpret_cast<volatile uint32_t *>(&s_exit[0])) break;
            }
            P2_STAMP(round, 0, 1);
            P2Meta *mt = meta + ((set * 4 + (round & 3u)) * 4 + g);
            if (out_of_work) {
                if (lane == 0 && uh == 0) mt->done = 1u;
                __syncwarp();
                if (lane == 0) p2_mbar_arrive(bar_full);
                continue;
            }
            const int row = g * 32 + lane;
            // the x0 part of the operand row leaves BEFORE the per-voxel maximum is taken in the same registers
#pragma unroll
            for (int u4 = 0; u4 < UH / 4; ++u4) {
                float4 hi, lo;
                tc_split(x0[4 * u4], hi.x, lo.x); tc_split(x0[4 * u4 + 1], hi.y, lo.y); tc_split(x0[4 * u4 + 2], hi.z, lo.z); tc_split(x0[4 * u4 + 3], hi.w, lo.w);
                const uint32_t o = tc_canon(row, uh + 4 * u4, TC_M);
                *reinterpret_cast<float4 *>(a_hi + o) = hi; *reinterpret_cast<float4 *>(a_lo + o) = lo;
            }
            p2_seg_max<UH>(x0, flags, nsteps);
#pragma unroll
            for (int u = 0; u < UH; ++u) x0[u] = __shfl_sync(0xffffffffu, x0[u], last_lane & 31);
#pragma unroll
            for (int u4 = 0; u4 < UH / 4; ++u4) {
                float4 hi, lo;
                tc_split(x0[4 * u4], hi.x, lo.x); tc_split(x0[4 * u4 + 1], hi.y, lo.y); tc_split(x0[4 * u4 + 2], hi.z, lo.z); tc_split(x0[4 * u4 + 3], hi.w, lo.w);
                const uint32_t o = tc_canon(row, P2_U0 + uh + 4 * u4, TC_M);
                *reinterpret_cast<float4 *>(a_hi + o) = hi; *reinterpret_cast<float4 *>(a_lo + o) = lo;
            }
            if (uh == 0) {                                   // the group's record is written by the pair's first warp
                const uint32_t lasts = __ballot_sync(0xffffffffu, seg_last);
                if (lane == 0) { mt->vid0 = (int)dsc.y; mt->lasts = lasts; mt->done = 0u; }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> the tensor core's async proxy
            __syncwarp();
            if (lane == 0) p2_mbar_arrive(bar_full);
            P2_STAMP(round, 0, 2);
        }
    } else if (warp == P2_ISSUER_WARP) {
        // =====================================================================================
        // ISSUER: the whole warp walks the tiles in order (tile = 2 * round + set; barrier waits are warp-wide, the warp
        // stays converged for the block barrier and the TMEM release at the end); lane 0 issues.  A set whose producers
        // are out of work keeps handing in empty tiles (all four records "done"), which are passed on unissued, so
        // every tile index is released in order until BOTH sets are empty in the same round.
        // =====================================================================================
        const uint32_t idesc = tc_idesc_tf32(TC_M, TC_M);     // D[unit (M = 128, padded), row (N = 128)]
        const uint32_t a_k = (TC_M / 8) * 128, mn = 128;
        // shared-memory descriptors of k-step 0, built once: a k-step advances the 14-bit start-address field by
        // 2 * a_k bytes >> 4 (no carry: shared memory ends below 256 KB).  The issuing thread is one serial instruction
        // stream (~8 cycles per instruction next to 28 other warps): rebuilding four descriptors per k-step cost
        // ~400 instructions = 1.9 us per tile, more than the tensor pipe needs for the tile (traced)
        const unsigned long long d_wh = tc_desc(tc_smem_u32(b_hi), a_k, mn), d_wl = tc_desc(tc_smem_u32(b_lo), a_k, mn);
        const unsigned long long d_xh0 = tc_desc(tc_smem_u32(a_st), a_k, mn), d_xl0 = tc_desc(tc_smem_u32(a_st + TC_M * P2_K), a_k, mn);
        const unsigned long long d_set = (unsigned long long)((2u * TC_M * P2_K * 4u) >> 4), d_ks = (unsigned long long)((2u * a_k) >> 4);
        uint32_t empty0 = 0;
        bool pending0 = false;                                 // set 0 handed in an empty tile this round: its producers are released
                                                               // together with the decision of tile (round, 1) -- go on, or stop
        for (uint32_t tile = 0;; ++tile) {
            const uint32_t s = tile & 1u, round = tile >> 1, acc = tile & 3u;
            if (!p2_mbar_wait(tc_smem_u32(&s_full[s]), round & 1u, &s_abort, diag, 0x200u | (s << 4) | (round << 16))) break;
            P2_STAMP(round, s, 0);
            const P2Meta *mt = meta + (s * 4 + (round & 3u)) * 4;
            const uint32_t all_done = *reinterpret_cast<const volatile uint32_t *>(&mt[0].done) &
                                      *reinterpret_cast<const volatile uint32_t *>(&mt[1].done) &
                                      *reinterpret_cast<const volatile uint32_t *>(&mt[2].done) &
                                      *reinterpret_cast<const volatile uint32_t *>(&mt[3].done);
            // the accumulator stage has been drained by its team (tile - 4).  A parity wait tells two phases apart, so no
            // barrier may run two phases ahead of its slowest waiter: every arrival on s_acc / s_rec [acc] below comes
            // after this wait
            if (tile >= 4 && !p2_mbar_wait(tc_smem_u32(&s_free[acc]), ((tile >> 2) - 1u) & 1u, &s_abort, diag, 0x300u | (acc << 4) | (round << 16))) break;
            P2_STAMP(round, s, 1);
            if (s == 0) empty0 = all_done;
            const bool finish = s == 1 && all_done && empty0;          // both sets handed in an empty tile this round
            if (finish && lane == 0) {
                *reinterpret_cast<volatile uint32_t *>(&s_exit[1]) = tile;   // teams: tiles from here on are not real
                *reinterpret_cast<volatile uint32_t *>(&s_exit[0]) = 1u;     // producers: stop after this round's s_mma
                __threadfence_block();
            }
            __syncwarp();
            if (s == 1 && pending0) {                                   // set 0's empty tile of this round: release its producers now
                if (lane == 0) p2_mbar_arrive(tc_smem_u32(&s_mma[0]));
                pending0 = false;
            }
            if (all_done) {                                             // an empty tile: nothing to issue, the team skips it
                if (lane == 0) {
                    p2_mbar_arrive(tc_smem_u32(&s_rec[acc]));
                    p2_mbar_arrive(tc_smem_u32(&s_acc[acc]));
                    if (s == 1) p2_mbar_arrive(tc_smem_u32(&s_mma[1]));   // set 1: go on handing in empty tiles, or (finish) stop
                }
                if (s == 0) pending0 = true;
                __syncwarp();
                if (!finish) continue;
                // ---- the end: the two other teams wait for the next two tile indices (both sets' producers have been
                // released above and stop when they see the flag) ----
                bool ok = true;
                for (uint32_t tt = tile + 1; tt <= tile + 2 && ok; ++tt) {
                    const uint32_t a2 = tt & 3u;
                    if (tt >= 4) ok = p2_mbar_wait(tc_smem_u32(&s_free[a2]), ((tt >> 2) - 1u) & 1u, &s_abort, diag, 0x600u | (a2 << 4) | (round << 16));
                    if (ok && lane == 0) { p2_mbar_arrive(tc_smem_u32(&s_rec[a2])); p2_mbar_arrive(tc_smem_u32(&s_acc[a2])); }
                    __syncwarp();
                }
                break;
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                // hands the producers' group records (acquired with the full barrier) on to the epilogue:
                // a plain release / acquire chain, independent of the tensor core's commit
                p2_mbar_arrive(tc_smem_u32(&s_rec[acc]));
                const uint32_t d_tmem = tmem + acc * (uint32_t)TC_M;
                const unsigned long long dxh = d_xh0 + s * d_set, dxl = d_xl0 + s * d_set;
#pragma unroll
                for (int ks = 0; ks < P2_K / 8; ++ks) {
                    tc_mma_tf32(d_tmem, d_wl + ks * d_ks, dxh + ks * d_ks, idesc, ks > 0 ? 1u : 0u);     // small terms first
                    tc_mma_tf32(d_tmem, d_wh + ks * d_ks, dxl + ks * d_ks, idesc, 1u);
                    tc_mma_tf32(d_tmem, d_wh + ks * d_ks, dxh + ks * d_ks, idesc, 1u);
                }
                tc_commit(tc_smem_u32(&s_mma[s]));      // operand stage free (producers) ...
                tc_commit(tc_smem_u32(&s_acc[acc]));    // ... and accumulator ready (epilogue team), when the MMAs retire
            }
            __syncwarp();
            P2_STAMP(round, s, 2);
        }
    } else {
        // =====================================================================================
        // EPILOGUE warp: team = tile index mod 3 it serves, e = TMEM quarter (lanes [32 e, 32 e + 32) = 32 units),
        // all four groups of the tile
        // =====================================================================================
        const int ew = warp - P2_EPI_WARP0, team = ew >> 2, e = warp & 3;
        const int unit = e * 32 + lane;
        const bool has_units = e * 32 < N;                                   // warp-uniform
        const float sc = unit < N ? bn1[2 * unit] : 0.0f, sh = unit < N ? bn1[2 * unit + 1] : 0.0f;
        const float neutral = __int_as_float(0xff800000);                    // -inf: sc >= 0 here, so the maximum of the raw accumulator decides
        for (uint32_t tile = (uint32_t)team;; tile += 3) {
            const uint32_t s = tile & 1u, round = tile >> 1, acc = tile & 3u, par = (tile >> 2) & 1u;
            if (!p2_mbar_wait(tc_smem_u32(&s_rec[acc]), par, &s_abort, diag, 0x500u | (acc << 4) | (ew & 3) | (round << 16))) break;
            if (!p2_mbar_wait(tc_smem_u32(&s_acc[acc]), par, &s_abort, diag, 0x400u | (acc << 4) | (ew & 3) | (round << 16))) break;
            // (compute-sanitizer racecheck reports this read against the issuer's write of s_exit[1]: benign by design -- a real
            // tile is below the final value and below the initial 0xFFFFFFFF alike; a tile at or past the end is only
            // released, through the barriers above, after the write)
            if (tile >= *reinterpret_cast<volatile uint32_t *>(&s_exit[1])) break;
            P2_STAMP(round, s, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int gg = 0; gg < 4 && has_units; ++gg) {
                const P2Meta *mt = meta + ((s * 4 + (round & 3u)) * 4 + gg);
                if (*reinterpret_cast<const volatile uint32_t *>(&mt->done)) continue;
                const uint32_t lasts = mt->lasts;
                float v[32];
                tc_ld_32x32(tmem + ((uint32_t)(e * 32) << 16) + acc * (uint32_t)TC_M + (uint32_t)(gg * 32), v);
                // the voxels of a group take consecutive output rows (first one = vid0): the store offset advances
                // by one row at every voxel end
                float run = neutral;
                float *const dst = a.out + (size_t)mt->vid0 * N + unit;   // units are a multiple of 32: every lane of the warp owns one
                uint32_t off = 0;
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const bool last = (lasts >> k) & 1u;                 // warp-uniform: the voxel ends at row k
                    run = fmaxf(run, v[k]);
                    if (last) __stcs(dst + off, fmaxf(__fmaf_rn(run, sc, sh), 0.0f));
                    off += last ? (uint32_t)N : 0u;
                    run = last ? neutral : run;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) p2_mbar_arrive(tc_smem_u32(&s_free[acc]));
            P2_STAMP(round, s, 1);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0 && s_abort) {                       // a wait starved: tell the host (pv_read_status)
        atomicOr(a.counter + 1, 4u);
        if (a.status) atomicOr(a.status, 4u);
    }
    if (warp == P2_ISSUER_WARP)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(ncols) : "memory");
}

// Shapes the fused kernel covers: two layers, 32 units in the first (K = 64), 32 | units of the last
// <= 128, decorated width <= 16, T <= 32.
bool pv_pfn_fused_supported(const pv_pfn_layer *layers, int n_layers, int t, int c, int with_distance)
{
    if (n_layers != 2 || t < 1 || t > 32 || c < 3 || c > PV_MAX_CHANNELS) return false;
    const int c0 = c + 5 + (with_distance ? 1 : 0);
    if (c0 > P2_C0 || layers[0].in_channels != c0 || layers[0].units != P2_U0) return false;
    if (layers[1].in_channels != 2 * P2_U0 || layers[1].units % 32 != 0 || layers[1].units < 32 || layers[1].units > 128) return false;
    return true;
}

size_t pv_pfn_fused_smem(int n1)
{
    return sizeof(float) * (2 * 2 * (size_t)TC_M * P2_K + 2 * (size_t)TC_M * P2_K + P2_C0 * P2_U0 + 4 * P2_U0 + 4 * (size_t)n1) +
           sizeof(P2Meta) * 32 + 16 * P2_MC * sizeof(uint4) + 128;
}

// counter: 64 words (256 bytes) of device scratch, zeroed before every launch: [0] the dynamic mini-chunk queue,
// [1] status bits in the layout pv_read_status expects (bit 2: the watchdog fired), [2] watchdog tag (0 = healthy),
// [3] the block that starved first, [4 .. 28] where each of its 25 warps was waiting, [40] row allocator of k_pfn_rows.
int pv_pfn_fused_launch(P2Args &a, const pv_pfn_layer *layers, int batch_frames, long long voxels_per_frame_cap,
                        cudaStream_t st)
{
    a.w0 = layers[0].weight; a.mean0 = layers[0].bn_mean; a.var0 = layers[0].bn_var; a.gamma0 = layers[0].bn_gamma; a.beta0 = layers[0].bn_beta;
    a.w1 = layers[1].weight; a.mean1 = layers[1].bn_mean; a.var1 = layers[1].bn_var; a.gamma1 = layers[1].bn_gamma; a.beta1 = layers[1].bn_beta;
    a.n1 = layers[1].units;
    a.chunks_per_frame = (uint32_t)((voxels_per_frame_cap + P2_MC - 1) / P2_MC);
    if (a.chunks_per_frame == 0) a.chunks_per_frame = 1;
    a.n_chunks = a.chunks_per_frame * (uint32_t)batch_frames;
    const size_t smem = pv_pfn_fused_smem(a.n1);
    if (cudaMemsetAsync(a.counter, 0, 64 * sizeof(unsigned int), st) != cudaSuccess) return PV_ERR_CUDA;   // queues + watchdog words
    const unsigned want = (a.n_chunks + 7) / 8;
    const unsigned sms = (unsigned)pv_sm_count();
    // gather + decorate first, at full occupancy; then layer 0 / tcgen05 layer 1 / per-voxel max
    a.drows = a.drows_out; a.desc = a.desc_out; a.ngroups = a.ngroups_out;
    const unsigned blocks = (a.n_chunks + P2_ROWS_THREADS / 32 - 1) / (P2_ROWS_THREADS / 32);
    const unsigned rgrid = blocks < sms * 2 * P2_ROWS_BLOCKS ? blocks : sms * 2 * P2_ROWS_BLOCKS;
    if (a.mode == 1) k_pfn_rows<1><<<rgrid, P2_ROWS_THREADS, 0, st>>>(a);
    else k_pfn_rows<0><<<rgrid, P2_ROWS_THREADS, 0, st>>>(a);
    const unsigned grid = want < sms ? want : sms;
    const int c0q = (a.c0 + 3) >> 2;
    if (c0q <= 2) {
        if (cudaFuncSetAttribute(k_pfn_fused<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PV_ERR_CUDA;
        k_pfn_fused<2><<<grid, P2_THREADS, smem, st>>>(a);
    } else if (c0q == 3) {
        if (cudaFuncSetAttribute(k_pfn_fused<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PV_ERR_CUDA;
        k_pfn_fused<3><<<grid, P2_THREADS, smem, st>>>(a);
    } else {
        if (cudaFuncSetAttribute(k_pfn_fused<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PV_ERR_CUDA;
        k_pfn_fused<4><<<grid, P2_THREADS, smem, st>>>(a);
    }
    return pv_last_cuda_error();
}

// rows per float4 plane of the pre-pass's row store (a multiple of 16 rows = 256 bytes)
long long pv_pfn_rows_stride(long long rows) { return (rows + 64 + 15) & ~15ll; }

// Rows / descriptors of the pre-pass (mode 1): bytes for `rows` decorated rows and `voxel_cap` voxels in `batch` frames.
size_t pv_pfn_rows_bytes(long long rows, long long voxels_per_frame_cap, int batch_frames, size_t *desc_off, size_t *ng_off)
{
    size_t chunks = (size_t)((voxels_per_frame_cap + P2_MC - 1) / P2_MC);
    if (chunks == 0) chunks = 1;
    chunks *= (size_t)batch_frames;
    size_t o = (size_t)pv_pfn_rows_stride(rows) * P2_C0 * sizeof(float);
    if (desc_off) *desc_off = o;
    o += (chunks * P2_MC * sizeof(uint4) + 255) & ~(size_t)255;
    if (ng_off) *ng_off = o;
    o += (chunks * sizeof(uint32_t) + 255) & ~(size_t)255;
    return o;
}

// The drop-in call on the padded tensor runs in slices of P2_SLICE voxels, so that the decorated rows of a slice
// (at most T per voxel, 64 bytes each) fit a bounded scratch whatever M is.
#define P2_SLICE 524288ll     // (131072: 2.0 ms instead of 1.2 for 960 k voxels -- 2048 chunks per launch leave the persistent producers unbalanced)
size_t pv_pfn_fused_tensor_bytes(long long m, int t)
{
    const long long ms = m < P2_SLICE ? m : P2_SLICE;
    return 256 + pv_pfn_rows_bytes(ms * t, ms, 1, nullptr, nullptr);
}

int pv_pfn_fused_tensor(const float *voxels, const int32_t *num_points, const int32_t *coors, int64_t m, int32_t t,
                        int32_t c, int32_t with_distance, float vx, float vy, float x_off, float y_off,
                        const pv_pfn_layer *layers, float eps, void *workspace, float *out, cudaStream_t st)
{
    P2Args a = {};
    a.mode = 0;
    a.voxels = voxels; a.num = num_points; a.coors_in = coors;
    a.t = t; a.c = c; a.with_distance = with_distance ? 1 : 0; a.c0 = c + 5 + a.with_distance;
    a.vx = vx; a.vy = vy; a.x_off = x_off; a.y_off = y_off; a.eps = eps;
    a.counter = reinterpret_cast<unsigned int *>(workspace); a.out = out;
    const long long ms_max = m < P2_SLICE ? m : P2_SLICE;
    size_t desc_off = 0, ng_off = 0;
    pv_pfn_rows_bytes(ms_max * t, ms_max, 1, &desc_off, &ng_off);
    char *rows0 = reinterpret_cast<char *>(workspace) + 256;
    a.drows_out = reinterpret_cast<float4 *>(rows0);
    a.drow_stride = pv_pfn_rows_stride(ms_max * t);
    a.desc_out = reinterpret_cast<uint4 *>(rows0 + desc_off);
    a.ngroups_out = reinterpret_cast<uint32_t *>(rows0 + ng_off);
    for (long long v0 = 0; v0 < m; v0 += P2_SLICE) {           // same stream: a slice's kernels finish before the next one reuses the scratch
        a.v0 = v0;
        a.m = m - v0 < P2_SLICE ? m - v0 : P2_SLICE;
        const int rc = pv_pfn_fused_launch(a, layers, 1, a.m, st);
        if (rc) return rc;
    }
    return PV_OK;
}
