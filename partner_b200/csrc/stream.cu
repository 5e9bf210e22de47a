// stream.cu -- azimuth-sector streaming of a polar point cloud (sm_100a).
//
// Reference: Voxelization.voxelize_streaming_polar, det3d/datasets/pipelines/voxelization.py:305-371
// (PolarStream's on-the-fly sectors): for each of `nsectors` azimuth wedges the reference selects
// the wedge's points with np.where (stable order), shifts their azimuth to the first wedge,
// recomputes x, y = rho * cos / sin(phi) and takes the clamped grid index on a grid whose azimuth
// extent is ny / nsectors.  nsectors boolean masks + nsectors fancy-index copies on the CPU.
//
// Here it is one stable multi-way partition: every point belongs to exactly one wedge, so
//   S1  k_sector_words   per warp and sector one 32-bit membership word (ballot) -- coalesced, no atomics
//   S2  k_sector_scan    one block per sector: popcount prefix over the N/32 words; sector sizes, bases
//   S3  k_sector_emit    per point: destination = base[sector] + prefix + popcount(lower lanes);
//                        transformed row, grid index and source index written sector-major
// Selection, order, shifted azimuth and grid index are bit-exact; x, y use cosf / sinf (numpy's
// float32 cos / sin are vendor SIMD routines, a few ulp apart).
#include "pv_common.cuh"

#define ST_MAX_SECTORS 64
#define ST_SCAN_THREADS 1024

struct StParams {
    const float *pts;           // [n, c] cylinder rows (rho, phi, z, x, y, ...)
    long long n;
    int c, ns;
    float lo[3], vs[3];
    float topf[3];              // cur_grid - 1 as float
    float min_az, interval;
    uint32_t nw;                // words per sector = ceil(n / 32)
    uint32_t *words;            // [ns][nw]
    uint32_t *pre;              // [ns][nw] exclusive popcount prefix inside the sector
    uint32_t *base;             // [ns + 1] first output row of each sector
    uint32_t *ticket;
    float *out_pts;
    int32_t *out_gi, *out_idx, *counts;
};

// Sector of azimuth phi (:350-358): the first wedge is open below, the last one open above; NaN -> none.
__device__ __forceinline__ int st_sector(const StParams &q, float phi)
{
    int s = -1;
    for (int k = 0; k < q.ns; ++k) {
        const float lo = __fadd_rn(q.min_az, __fmul_rn((float)k, q.interval));
        const float hi = __fadd_rn(q.min_az, __fmul_rn((float)(k + 1), q.interval));
        const bool in = k == 0 ? phi < hi : (k == q.ns - 1 ? phi >= lo : (phi >= lo && phi < hi));
        if (in && s < 0) s = k;
    }
    return s;
}

__global__ void __launch_bounds__(256) k_sector_words(const __grid_constant__ StParams q)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long w = i >> 5;
    if (w >= q.nw) return;
    const int s = i < q.n ? st_sector(q, __ldg(q.pts + i * q.c + 1)) : -1;
    for (int k = 0; k < q.ns; ++k) {
        const unsigned word = __ballot_sync(0xffffffffu, s == k);
        if ((threadIdx.x & 31) == 0) q.words[(size_t)k * q.nw + w] = word;
    }
}

__global__ void __launch_bounds__(ST_SCAN_THREADS) k_sector_scan(const __grid_constant__ StParams q)
{
    __shared__ uint32_t s_warp[ST_SCAN_THREADS / 32];
    __shared__ uint32_t s_carry, s_last;
    const int k = blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t *words = q.words + (size_t)k * q.nw;
    uint32_t *pre = q.pre + (size_t)k * q.nw;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t w0 = 0; w0 < q.nw; w0 += ST_SCAN_THREADS) {
        const uint32_t w = w0 + tid;
        const uint32_t c = w < q.nw ? __popc(words[w]) : 0u;
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
        if (w < q.nw) pre[w] = off + incl - c;
        __syncthreads();
        if (tid == ST_SCAN_THREADS - 1) s_carry = off + incl;
        __syncthreads();
    }
    if (tid == 0) {
        q.counts[k] = (int32_t)s_carry;
        __threadfence();
        const uint32_t done = atomicAdd(q.ticket, 1u);
        s_last = done == gridDim.x - 1 ? 1u : 0u;
        if (s_last) *q.ticket = 0u;
    }
    __syncthreads();
    if (s_last && tid == 0) {               // sector bases (nsectors is small)
        __threadfence();
        uint32_t acc = 0;
        for (int j = 0; j < q.ns; ++j) { q.base[j] = acc; acc += (uint32_t)__ldcg(q.counts + j); }
        q.base[q.ns] = acc;
    }
}

__global__ void __launch_bounds__(256) k_sector_emit(const __grid_constant__ StParams q)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q.n) return;
    const float *row = q.pts + i * q.c;
    float v[PV_MAX_CHANNELS];
#pragma unroll
    for (int k = 0; k < PV_MAX_CHANNELS; ++k) v[k] = k < q.c ? __ldg(row + k) : 0.0f;
    const int s = st_sector(q, v[1]);
    if (s < 0) return;
    const size_t wi = (size_t)s * q.nw + (size_t)(i >> 5);
    const uint32_t word = __ldg(q.words + wi);
    const uint32_t dst = __ldg(q.base + s) + __ldg(q.pre + wi) + __popc(word & ((1u << (i & 31)) - 1u));
    const float lo_s = __fadd_rn(q.min_az, __fmul_rn((float)s, q.interval));
    v[1] = __fsub_rn(v[1], __fsub_rn(lo_s, q.min_az));                 // :360
    v[3] = __fmul_rn(v[0], cosf(v[1]));                                // :361
    v[4] = __fmul_rn(v[0], sinf(v[1]));                                // :362
    float *o = q.out_pts + (size_t)dst * q.c;
#pragma unroll
    for (int k = 0; k < PV_MAX_CHANNELS; ++k)
        if (k < q.c) o[k] = v[k];
    int32_t *g = q.out_gi + (size_t)dst * 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) {                                      // :366-368, same arithmetic as pv_dynamic_voxelize
        float t = __fdiv_rn(__fsub_rn(v[j], q.lo[j]), q.vs[j]);
        t = fminf(fmaxf(t, 0.0f), q.topf[j]);                          // NaN -> 0
        g[2 - j] = (int32_t)floorf(t);
    }
    q.out_idx[dst] = (int32_t)i;
}

static size_t st_align(size_t v) { return (v + 255) / 256 * 256; }

extern "C" {

size_t pv_stream_workspace_bytes(int64_t n, int32_t nsectors)
{
    if (n < 0 || nsectors < 1 || nsectors > ST_MAX_SECTORS) return 0;
    const size_t nw = (size_t)(n + 31) / 32 + 1;
    return 2 * st_align(nw * nsectors * 4) + st_align((size_t)(nsectors + 1) * 4) + 256;
}

int pv_stream_sectors(const pv_config *cfg, const float *points, int64_t n, int32_t c, int32_t nsectors,
                      float max_azimuth, void *workspace, size_t workspace_bytes, float *points_out,
                      int32_t *grid_ind_out, int32_t *point_index, int32_t *sector_counts, pv_stream_t stream)
{
    int rc = pv_check_config(cfg);
    if (rc) return rc;
    if (n < 0 || n >= (1ll << 31) || c < 5 || c > PV_MAX_CHANNELS || nsectors < 1 || nsectors > ST_MAX_SECTORS)
        return PV_ERR_BAD_ARGUMENT;
    if (!sector_counts || !workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return PV_ERR_BAD_ARGUMENT;
    if (workspace_bytes < pv_stream_workspace_bytes(n, nsectors)) return PV_ERR_WORKSPACE;
    if (n > 0 && (!points || !points_out || !grid_ind_out || !point_index)) return PV_ERR_BAD_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
    StParams q;
    q.pts = points; q.n = n; q.c = c; q.ns = nsectors;
    const int cur_grid[3] = {cfg->grid[0], cfg->grid[1] / nsectors, cfg->grid[2]};      // :316-317
    for (int j = 0; j < 3; ++j) { q.lo[j] = cfg->lo[j]; q.vs[j] = cfg->vs[j]; q.topf[j] = (float)(cur_grid[j] - 1); }
    q.min_az = cfg->lo[1];
    {   // interval = (max_az - min_az) / nsectors in float32 (:315); volatile: no excess precision
        volatile float d = max_azimuth - cfg->lo[1];
        volatile float iv = d / (float)nsectors;
        q.interval = iv;
    }
    q.nw = (uint32_t)((n + 31) / 32);
    char *w = (char *)workspace;
    q.words = (uint32_t *)w;                 w += st_align(((size_t)q.nw + 1) * nsectors * 4);
    q.pre = (uint32_t *)w;                   w += st_align(((size_t)q.nw + 1) * nsectors * 4);
    q.base = (uint32_t *)w;                  w += st_align((size_t)(nsectors + 1) * 4);
    q.ticket = (uint32_t *)w;
    q.out_pts = points_out; q.out_gi = grid_ind_out; q.out_idx = point_index; q.counts = sector_counts;
    if (cudaMemsetAsync(q.ticket, 0, 4, st) != cudaSuccess) return PV_ERR_CUDA;
    if (n == 0) {
        if (cudaMemsetAsync(sector_counts, 0, (size_t)nsectors * 4, st) != cudaSuccess) return PV_ERR_CUDA;
        return PV_OK;
    }
    const unsigned blocks = (unsigned)(((long long)q.nw * 32 + 255) / 256);
    k_sector_words<<<blocks, 256, 0, st>>>(q);
    k_sector_scan<<<(unsigned)nsectors, ST_SCAN_THREADS, 0, st>>>(q);
    k_sector_emit<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(q);
    return pv_last_cuda_error();
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Rigid warp of the xyz columns between sweeps (voxelize_streaming_by_sweep, voxelization.py:442-447:
// points[:, :3] = (hstack(points[:, :3], 1) @ tm.T)[:, :3] in float64, stored back as float32) and
// the time-lag fix of the last column (:439).  m = rows 0..2 of the 4 x 4 matrix, row-major.
// ---------------------------------------------------------------------------------------------
struct AffParams { double m[12]; float t_shift; };

__global__ void __launch_bounds__(256) k_affine_points(const float *__restrict__ in, long long n, int c,
                                                       const __grid_constant__ AffParams a, float *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *row = in + i * c;
    float *o = out + i * c;
    const double x = (double)row[0], y = (double)row[1], z = (double)row[2];
#pragma unroll
    for (int r = 0; r < 3; ++r)      // dot product in the order numpy's matmul accumulates a length-4 row
        o[r] = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, a.m[4 * r]), __dmul_rn(y, a.m[4 * r + 1])),
                                          __dmul_rn(z, a.m[4 * r + 2])), a.m[4 * r + 3]);
    for (int k = 3; k < c; ++k) o[k] = k == c - 1 ? __fsub_rn(row[k], a.t_shift) : row[k];
}

extern "C" int pv_affine_points(const float *in, int64_t n, int32_t c, const double *matrix3x4, float t_shift,
                                float *out, pv_stream_t stream)
{
    if (n < 0 || c < 3 || !matrix3x4 || (n > 0 && (!in || !out))) return PV_ERR_BAD_ARGUMENT;
    if (n == 0) return PV_OK;
    AffParams a;
    for (int k = 0; k < 12; ++k) a.m[k] = matrix3x4[k];
    a.t_shift = t_shift;
    k_affine_points<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, n, c, a, out);
    return pv_last_cuda_error();
}
