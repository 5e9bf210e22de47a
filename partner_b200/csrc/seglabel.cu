// seglabel.cu -- segmentation voxel labels and the point <-> voxel map consumers (sm_100a).
//
// Reference:
//   Voxelization.get_grid_ind, train branch      det3d/datasets/pipelines/voxelization.py:40-60
//       valid = pc_label >= 0; the valid (z, y, x, label) rows are LEXSORTED by cell and handed to
//   AssignLabel.assign_voxel_labels              det3d/datasets/pipelines/preprocess.py:170-191
//       a sequential numba loop over the sorted rows with a 256-entry uint16 counter per run:
//       voxel_labels[z, y, x] = argmax(counter) (ties -> smallest label), 0 where no point fell;
//       the unsorted valid rows become `valid_grid_ind`.
//   SegHead.predict                              det3d/models/seg_heads/seg_head.py:171-193
//       preds = pred_labels[count][z, y, x] (or [0, y, x] for a 2-D map) per valid point.
//
// The sort only groups the rows of a cell; the result is the per-cell MAJORITY VOTE, which is
// order free.  Here:
//   G1 k_seg_insert   per point: validity ballot word (stable compaction) + count[(frame, cell, label)]++
//                     in an open-addressing table (64-bit keys, CAS claim, 2 x n slots)
//   G2 k_seg_scan     one block: popcount prefix over the validity words, per-frame offsets
//   G3 k_seg_emit     valid row i -> valid_grid_ind[prefix(i)]
//   G4 k_seg_vote     per table slot: atomicMax(best[cell], (count mod 2^16) << 8 | (255 - label)) -- the
//                     counter is uint16 in the reference and wraps; the complemented label makes the
//                     maximum break ties towards the smallest label, as np.argmax does
//   G5 k_seg_write    dense pass: voxel_labels[cell] = decoded winner (0 for empty cells), int64
// All integer work, bit-exact.
#include "pv_common.cuh"

#define SG_EMPTY 0xFFFFFFFFFFFFFFFFull
#define SG_SCAN_THREADS 1024

struct SgParams {
    const int32_t *gi;          // [n, 3] (z, y, x)
    const int32_t *label;       // [n]
    const int32_t *offsets;     // [B + 1]
    int32_t B;
    uint32_t n;
    int32_t grid[3];            // nx, ny, nz
    unsigned long long cells;
    unsigned long long *keys;   // [cap]  ((frame * cells + cell) << 8) | label      (clean = all ones)
    uint32_t *counts;           // [cap]                                             (clean = 0)
    uint32_t cap_mask;
    uint32_t *best;             // [B * cells]                                       (clean = 0)
    uint32_t *words;            // [nw] validity ballots
    uint32_t *pre;              // [nw] exclusive popcount prefix
    uint32_t nw;
    uint32_t *status;
    long long *voxel_labels;    // [B, nz, ny, nx]
    int32_t *valid_gi;          // [n, 3]
    int32_t *valid_offsets;     // [B + 1]
};

__device__ __forceinline__ unsigned long long sg_hash(unsigned long long k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}

__global__ void __launch_bounds__(256) k_seg_insert(const __grid_constant__ SgParams q)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < q.n;
    const int32_t lab = live ? __ldg(q.label + i) : -1;
    bool valid = lab >= 0;                                   // voxelization.py:44
    if (lab > 255) { valid = false; atomicOr(q.status, 2u); } // counter has 256 entries (preprocess.py:178)
    int32_t z = 0, y = 0, x = 0;
    if (valid) {
        z = __ldg(q.gi + (size_t)i * 3); y = __ldg(q.gi + (size_t)i * 3 + 1); x = __ldg(q.gi + (size_t)i * 3 + 2);
        if ((unsigned)z >= (unsigned)q.grid[2] || (unsigned)y >= (unsigned)q.grid[1] || (unsigned)x >= (unsigned)q.grid[0]) {
            valid = false;
            atomicOr(q.status, 2u);
        }
    }
    const unsigned word = __ballot_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31u) == 0 && (i >> 5) < q.nw) q.words[i >> 5] = word;
    if (!valid) return;
    const int b = pv_frame_of(q.offsets, q.B, i);
    const unsigned long long cell = ((unsigned long long)z * q.grid[1] + y) * q.grid[0] + x;
    const unsigned long long key = (((unsigned long long)b * q.cells + cell) << 8) | (unsigned long long)lab;
    uint32_t h = (uint32_t)sg_hash(key) & q.cap_mask;
    for (uint32_t probe = 0; probe <= q.cap_mask; ++probe) {
        const unsigned long long old = atomicCAS(q.keys + h, SG_EMPTY, key);
        if (old == SG_EMPTY || old == key) { atomicAdd(q.counts + h, 1u); return; }
        h = (h + 1) & q.cap_mask;
    }
    atomicOr(q.status, 1u);
}

// P(i) = number of valid rows before row i (i <= n).
__device__ __forceinline__ uint32_t sg_rank(const SgParams &q, uint32_t i, uint32_t total)
{
    if (i >= q.n) return total;
    return q.pre[i >> 5] + __popc(q.words[i >> 5] & ((1u << (i & 31u)) - 1u));
}

__global__ void __launch_bounds__(SG_SCAN_THREADS) k_seg_scan(const __grid_constant__ SgParams q)
{
    __shared__ uint32_t s_warp[SG_SCAN_THREADS / 32];
    __shared__ uint32_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t w0 = 0; w0 < q.nw; w0 += SG_SCAN_THREADS) {
        const uint32_t w = w0 + tid;
        const uint32_t c = w < q.nw ? __popc(q.words[w]) : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (unsigned)d) incl += o;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t off = s_carry;
        for (uint32_t j = 0; j < warp; ++j) off += s_warp[j];
        if (w < q.nw) q.pre[w] = off + incl - c;
        __syncthreads();
        if (tid == SG_SCAN_THREADS - 1) s_carry = off + incl;
        __syncthreads();
    }
    __threadfence_block();
    const uint32_t total = s_carry;
    for (int b = (int)tid; b <= q.B; b += SG_SCAN_THREADS)
        q.valid_offsets[b] = (int32_t)sg_rank(q, (uint32_t)q.offsets[b], total);
}

__global__ void __launch_bounds__(256) k_seg_emit(const __grid_constant__ SgParams q)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q.n) return;
    const uint32_t word = __ldg(q.words + (i >> 5));
    if (!((word >> (i & 31u)) & 1u)) return;
    const uint32_t dst = __ldg(q.pre + (i >> 5)) + __popc(word & ((1u << (i & 31u)) - 1u));
    q.valid_gi[(size_t)dst * 3] = __ldg(q.gi + (size_t)i * 3);
    q.valid_gi[(size_t)dst * 3 + 1] = __ldg(q.gi + (size_t)i * 3 + 1);
    q.valid_gi[(size_t)dst * 3 + 2] = __ldg(q.gi + (size_t)i * 3 + 2);
}

__global__ void __launch_bounds__(256) k_seg_vote(const __grid_constant__ SgParams q)
{
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h > q.cap_mask) return;
    const unsigned long long key = q.keys[h];
    if (key == SG_EMPTY) return;
    const uint32_t cnt = q.counts[h] & 0xFFFFu;                              // uint16 counter (preprocess.py:178)
    atomicMax(q.best + (key >> 8), (cnt << 8) | (255u - (uint32_t)(key & 255u)));
}

__global__ void __launch_bounds__(256) k_seg_write(const __grid_constant__ SgParams q)
{
    const unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (unsigned long long)q.B * q.cells) return;
    const uint32_t v = q.best[j];
    q.voxel_labels[j] = (v >> 8) ? (long long)(255u - (v & 255u)) : 0ll;      // all counts zero -> argmax = 0
}

struct SgLayout { size_t keys, counts, best, words, pre, status, total; uint32_t cap; };

static SgLayout sg_layout(unsigned long long cells, long long n, int batch)
{
    SgLayout L;
    uint32_t cap = 1024;
    while ((long long)cap < 2 * n) cap <<= 1;
    L.cap = cap;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t nw = (size_t)(n + 31) / 32 + 1;
    size_t o = 0;
    L.keys = o;   o = al(o + (size_t)cap * 8);
    L.counts = o; o = al(o + (size_t)cap * 4);
    L.best = o;   o = al(o + (size_t)batch * cells * 4);
    L.words = o;  o = al(o + nw * 4);
    L.pre = o;    o = al(o + nw * 4);
    L.status = o; o = al(o + 16);
    L.total = o;
    return L;
}

extern "C" {

size_t pv_seg_workspace_bytes(const pv_config *cfg, int64_t n_total, int32_t batch)
{
    if (pv_check_config(cfg) || n_total < 0 || n_total >= (1ll << 30) || batch <= 0) return 0;
    const unsigned long long cells = (unsigned long long)cfg->grid[0] * cfg->grid[1] * cfg->grid[2];
    return sg_layout(cells, n_total, batch).total;
}

int pv_seg_voxel_labels(const pv_config *cfg, const int32_t *pc_grid_ind, const int32_t *pc_label,
                        const int32_t *frame_offsets, int32_t batch, int64_t n_total, void *workspace,
                        size_t workspace_bytes, int64_t *voxel_labels, int32_t *valid_grid_ind,
                        int32_t *valid_offsets, int32_t *status, pv_stream_t stream)
{
    int rc = pv_check_config(cfg);
    if (rc) return rc;
    if (batch <= 0 || n_total < 0 || n_total >= (1ll << 30)) return PV_ERR_BAD_ARGUMENT;
    if (!frame_offsets || !workspace || !voxel_labels || !valid_grid_ind || !valid_offsets || !status) return PV_ERR_BAD_ARGUMENT;
    if (n_total > 0 && (!pc_grid_ind || !pc_label)) return PV_ERR_BAD_ARGUMENT;
    const unsigned long long cells = (unsigned long long)cfg->grid[0] * cfg->grid[1] * cfg->grid[2];
    const SgLayout L = sg_layout(cells, n_total, batch);
    if (workspace_bytes < L.total || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return PV_ERR_WORKSPACE;
    const unsigned long long dense = (unsigned long long)batch * cells;
    if ((dense + 255) / 256 > 0x7FFFFFFFull) return PV_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = reinterpret_cast<char *>(workspace);
    SgParams q;
    q.gi = pc_grid_ind; q.label = pc_label; q.offsets = frame_offsets; q.B = batch; q.n = (uint32_t)n_total;
    q.grid[0] = cfg->grid[0]; q.grid[1] = cfg->grid[1]; q.grid[2] = cfg->grid[2];
    q.cells = cells;
    q.keys = reinterpret_cast<unsigned long long *>(ws + L.keys);
    q.counts = reinterpret_cast<uint32_t *>(ws + L.counts);
    q.cap_mask = L.cap - 1;
    q.best = reinterpret_cast<uint32_t *>(ws + L.best);
    q.words = reinterpret_cast<uint32_t *>(ws + L.words);
    q.pre = reinterpret_cast<uint32_t *>(ws + L.pre);
    q.nw = (uint32_t)((n_total + 31) / 32);
    q.status = reinterpret_cast<uint32_t *>(status);
    q.voxel_labels = reinterpret_cast<long long *>(voxel_labels);
    q.valid_gi = valid_grid_ind; q.valid_offsets = valid_offsets;
    // stateless: the call prepares what it uses (table 12 bytes per slot, 4 bytes per cell)
    if (cudaMemsetAsync(q.keys, 0xFF, (size_t)L.cap * 8, st) != cudaSuccess) return PV_ERR_CUDA;
    if (cudaMemsetAsync(q.counts, 0, L.best - L.counts + (size_t)batch * cells * 4, st) != cudaSuccess) return PV_ERR_CUDA;
    if (cudaMemsetAsync(status, 0, sizeof(int32_t), st) != cudaSuccess) return PV_ERR_CUDA;
    if (n_total > 0) k_seg_insert<<<(unsigned)((q.nw * 32ull + 255) / 256), 256, 0, st>>>(q);
    k_seg_scan<<<1, SG_SCAN_THREADS, 0, st>>>(q);
    if (n_total > 0) {
        k_seg_emit<<<(unsigned)((n_total + 255) / 256), 256, 0, st>>>(q);
        k_seg_vote<<<(L.cap + 255) / 256, 256, 0, st>>>(q);
    }
    k_seg_write<<<(unsigned)((dense + 255) / 256), 256, 0, st>>>(q);
    return pv_last_cuda_error();
}

// SegHead.predict (seg_heads/seg_head.py:171-193): out[i] = pred_labels[frame(i)][z, y, x] for every valid
// point; nz_pred == 0 selects the 2-D form pred_labels[frame][y, x] (the z index is ignored, :188).
__global__ void __launch_bounds__(256) k_seg_gather(const long long *__restrict__ pred, int nz, int ny, int nx,
                                                    const int32_t *__restrict__ gi, const int32_t *__restrict__ voff,
                                                    int B, long long n, long long *__restrict__ out, uint32_t *status)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = pv_frame_of(voff, B, (uint32_t)i);
    const int z = nz ? gi[i * 3] : 0, y = gi[i * 3 + 1], x = gi[i * 3 + 2];
    const int nzz = nz ? nz : 1;
    if ((unsigned)z >= (unsigned)nzz || (unsigned)y >= (unsigned)ny || (unsigned)x >= (unsigned)nx) {
        atomicOr(status, 2u);
        out[i] = 0;
        return;
    }
    out[i] = pred[(((long long)b * nzz + z) * ny + y) * nx + x];
}

int pv_seg_gather_points(const int64_t *pred_labels, int32_t nz_pred, int32_t ny, int32_t nx,
                         const int32_t *valid_grid_ind, const int32_t *valid_offsets, int32_t batch,
                         int64_t n_valid, int64_t *out, int32_t *status, pv_stream_t stream)
{
    if (!pred_labels || !valid_offsets || !status || batch <= 0 || n_valid < 0 || n_valid >= (1ll << 31) || nz_pred < 0 || ny <= 0 || nx <= 0)
        return PV_ERR_BAD_ARGUMENT;
    if (n_valid > 0 && (!valid_grid_ind || !out)) return PV_ERR_BAD_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(status, 0, sizeof(int32_t), st) != cudaSuccess) return PV_ERR_CUDA;
    if (n_valid > 0)
        k_seg_gather<<<(unsigned)((n_valid + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<const long long *>(pred_labels), nz_pred, ny, nx, valid_grid_ind, valid_offsets, batch,
            n_valid, reinterpret_cast<long long *>(out), reinterpret_cast<uint32_t *>(status));
    return pv_last_cuda_error();
}

}  // extern "C"
