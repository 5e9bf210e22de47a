// dynpfn.cu -- DynamicPFNet.forward on the outputs of pv_dynamic_voxelize (sm_100a).
//
// Reference: det3d/models/readers/pillar_encoder.py:262-411 (DynamicPFNet), PFNLayer.forward_dynamic
// :63-71, get_cluster :228-238, polar2cart / cart2polar :240-260.  Per point: feature decoration
// (offsets from the voxel's mean and from the voxel's centre), then n layers of Linear (no bias) +
// ReLU + scatter_max over the voxel; non-last layers feed cat([x, x_max[unq_inv]]) to the next one.
// There is no normalisation in the dynamic forward.
//
// The reference runs this point-major: two [N, .] GEMMs and two scatter_max passes with one atomic
// per (point, channel).  Here it runs VOXEL-major: points are counting-sorted by voxel row (counts
// come from the voxelizer, one cursor atomic per point), a block owns 32 consecutive voxels and
// streams their points through shared memory in chunks of 64; the segmented maxima live in shared
// memory (non-negative floats compare as integers), so the only global traffic is one gather of
// every point row per layer pass and one coalesced write of the [M, units] result.  fp32 FFMA with
// exact accumulation order k = 0..K-1 (as the oracle); register tile 4 points x 4 units.
#include <stdlib.h>

#include "pv_common.cuh"

#define DP_VB 32          // voxels per block
#define DP_PC 128         // points per chunk (a group of 32 voxels holds ~100 points: usually one chunk)
#define DP_THREADS 256
#define DP_MAX_C0 32      // decorated row width
#define DP_MAX_U1 64      // units of the non-last layer (input of the last one: 2 * U1)
#define DP_MAX_U2 128     // units of the last layer

struct DpParams {
    const float *points;          // [n, c]
    const int32_t *unq, *unq_inv, *unq_cnt;
    const float *mean;            // [m, c] per-voxel mean of the rows
    const uint32_t *offs;         // [m + 1] exclusive prefix of unq_cnt
    const uint32_t *perm;         // [n] point indices sorted by voxel row
    const float *w1, *w2;         // layer weights [units, in] (w2 = NULL for a single layer)
    float *out;                   // [m, u_last]
    long long n, m;
    int c, c0, u1, u2;            // u1 = units of the first layer; u2 = units of the second (0 if none)
    int cylinder, flags;
    float vx, vy, x_off, y_off;
};

// ---------------------------------------------------------------------------------------------
// exclusive scan of the per-voxel point counts (three small launches) and the counting sort
// ---------------------------------------------------------------------------------------------
#define SC_ITEMS 8
#define SC_TILE (256 * SC_ITEMS)

__global__ void __launch_bounds__(256) k_dp_block_sums(const int32_t *__restrict__ cnt, long long m, uint32_t *__restrict__ sums)
{
    __shared__ uint32_t s_w[8];
    const long long base = (long long)blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
    uint32_t t = 0;
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) t += base + j < m ? (uint32_t)cnt[base + j] : 0u;
    t = __reduce_add_sync(0xffffffffu, t);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t a = 0;
        for (int k = 0; k < 8; ++k) a += s_w[k];
        sums[blockIdx.x] = a;
    }
}

__global__ void __launch_bounds__(1024) k_dp_scan_sums(uint32_t *sums, int nb, uint32_t *total)
{
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_carry;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const uint32_t v = i < nb ? sums[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (unsigned)d) incl += o;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        uint32_t off = s_carry;
        for (uint32_t k = 0; k < warp; ++k) off += s_w[k];
        if (i < nb) sums[i] = off + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = off + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}

__global__ void __launch_bounds__(256) k_dp_apply(const int32_t *__restrict__ cnt, long long m, const uint32_t *__restrict__ sums,
                                                  uint32_t *__restrict__ offs, uint32_t *__restrict__ cursor)
{
    __shared__ uint32_t s_w[8];
    const long long base = (long long)blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
    uint32_t v[SC_ITEMS], t = 0;
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) { v[j] = base + j < m ? (uint32_t)cnt[base + j] : 0u; t += v[j]; }
    uint32_t incl = t;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (unsigned)d) incl += o;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t run = sums[blockIdx.x] + incl - t;
    for (uint32_t k = 0; k < warp; ++k) run += s_w[k];
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) {
        if (base + j < m) { offs[base + j] = run; cursor[base + j] = 0u; }
        run += v[j];
    }
    if (base <= m && m < base + SC_ITEMS) {     // the thread that owns element m writes the total
        uint32_t tot = sums[blockIdx.x] + incl - t;
        for (uint32_t k = 0; k < warp; ++k) tot += s_w[k];
        for (int j = 0; j < SC_ITEMS && base + j < m; ++j) tot += v[j];
        offs[m] = tot;
    }
}

__global__ void __launch_bounds__(256) k_dp_sort(const int32_t *__restrict__ unq_inv, long long n, long long m,
                                                 const uint32_t *__restrict__ offs, uint32_t *cursor, uint32_t *__restrict__ perm)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t v = unq_inv[i];
    if (v < 0 || v >= m) return;
    const uint32_t k = atomicAdd(cursor + v, 1u);
    perm[offs[v] + k] = (uint32_t)i;
}

// ---------------------------------------------------------------------------------------------
// the voxel-major network
// ---------------------------------------------------------------------------------------------
// y[p][o] = relu(sum_k x[p][k] * wt[k][o]) for the chunk's points, register tile 4 points x 4 units;
// STORE: write y to ys[p][o]; every output is max-reduced into smax[vl[p]][o] (integer compare).
// init (optional): per-voxel start value of the accumulators, init[vl[p]][o] (the x_max half of the
// last layer's input is the same for every point of a voxel: its product is taken once per voxel).
template <bool STORE, bool MAX = true>
__device__ __forceinline__ void dp_layer(const float *__restrict__ xs, int xs_ld, int K, const float *__restrict__ wt, int U,
                                         int npts, const int *__restrict__ vl, float *__restrict__ ys, int ys_ld, int *__restrict__ smax,
                                         const float *__restrict__ init = nullptr)
{
    const int ugroups = U >> 2, tiles = (DP_PC / 4) * ugroups;
    for (int t = threadIdx.x; t < tiles; t += DP_THREADS) {
        const int pg = t / ugroups, ug = t - pg * ugroups;
        const int p0 = pg * 4;
        if (p0 >= npts) continue;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            if (init) {
                const float4 i4 = *reinterpret_cast<const float4 *>(init + vl[min(p0 + a, npts - 1)] * U + ug * 4);
                acc[a][0] = i4.x; acc[a][1] = i4.y; acc[a][2] = i4.z; acc[a][3] = i4.w;
            } else {
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
            }
        }
        const float *x0 = xs + (size_t)p0 * xs_ld;
        for (int k = 0; k < K; k += 4) {
            float4 xv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) xv[a] = *reinterpret_cast<const float4 *>(x0 + a * xs_ld + k);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 w = *reinterpret_cast<const float4 *>(wt + (size_t)(k + kk) * U + ug * 4);
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const float x = kk == 0 ? xv[a].x : kk == 1 ? xv[a].y : kk == 2 ? xv[a].z : xv[a].w;
                    acc[a][0] = __fmaf_rn(x, w.x, acc[a][0]);
                    acc[a][1] = __fmaf_rn(x, w.y, acc[a][1]);
                    acc[a][2] = __fmaf_rn(x, w.z, acc[a][2]);
                    acc[a][3] = __fmaf_rn(x, w.w, acc[a][3]);
                }
            }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            if (p0 + a >= npts) break;
            if (!MAX) {                                                   // plain product rows (no activation)
                if (STORE) *reinterpret_cast<float4 *>(ys + (size_t)(p0 + a) * ys_ld + ug * 4) =
                    make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
                continue;
            }
            int *mx = smax + vl[p0 + a] * U + ug * 4;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const float y = fmaxf(acc[a][b], 0.0f);                   // ReLU (:65)
                if (STORE) ys[(size_t)(p0 + a) * ys_ld + ug * 4 + b] = y;
                atomicMax(mx + b, __float_as_int(y));                     // scatter_max (:66); y >= 0
            }
        }
    }
}

__global__ void __launch_bounds__(DP_THREADS) k_dyn_pfn(const __grid_constant__ DpParams q)
{
    extern __shared__ __align__(16) float smem[];
    const int c = q.c, c0 = q.c0, c0p = (q.c0 + 3) & ~3, u1 = q.u1, u2 = q.u2;
    const int in2 = 2 * u1;
    // carve-up
    float *w1t = smem;                                    // [c0p][u1]
    float *w2t = w1t + c0p * u1;                          // [in2][u2]
    float *meta = w2t + (u2 ? in2 * u2 : 0);              // [VB][9]: mean x,y,z,r,a ; xc, yc, rc, ac
    int *max1 = reinterpret_cast<int *>(meta + DP_VB * 12);   // [VB][u1]
    int *max2 = max1 + DP_VB * u1;                        // [VB][u2]
    float *D = reinterpret_cast<float *>(max2 + DP_VB * (u2 ? u2 : 0));   // [PC][c0p]
    float *Y = D + DP_PC * c0p;                           // [PC][u1] first-layer outputs of the chunk
    float *P = Y + DP_PC * u1;                            // [VB][u2] per-voxel half of the last layer: W2[:, u1:] . x_max
    __shared__ int s_vl[DP_PC];

    const int tid = threadIdx.x;

    // ---- weights, transposed to [k][o]; padded input rows are zero.  PERSISTENT blocks: the
    // transposition (34 strided loads per thread for the 64 -> 128 layer) is paid once per block,
    // not once per 32 voxels (a group of 32 voxels holds ~100 points: the weights used to cost as
    // much as the arithmetic) ----
    for (int e = tid; e < c0p * u1; e += DP_THREADS) {
        const int k = e / u1, o = e - k * u1;
        w1t[e] = k < c0 ? __ldg(q.w1 + (size_t)o * c0 + k) : 0.0f;
    }
    for (int e = tid; e < (u2 ? in2 * u2 : 0); e += DP_THREADS) {
        const int k = e / u2, o = e - k * u2;
        w2t[e] = __ldg(q.w2 + (size_t)o * in2 + k);
    }
    const bool xyz_cluster = q.flags & 1, raz_cluster = q.flags & 2, xy_center = q.flags & 4, ra_center = q.flags & 8;
    const int xi[3] = {q.cylinder ? 3 : 0, q.cylinder ? 4 : 1, 2};
    const int ri[2] = {q.cylinder ? 0 : c - 2, q.cylinder ? 1 : c - 1};
    const long long n_groups = (q.m + DP_VB - 1) / DP_VB;
    for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const long long v0 = grp * DP_VB;
    const int nv = (int)min((long long)DP_VB, q.m - v0);
    const uint32_t pbeg = q.offs[v0], pend = q.offs[v0 + nv];
    __syncthreads();                                     // the previous group's maxima have been written out
    for (int e = tid; e < DP_VB * u1; e += DP_THREADS) max1[e] = 0;
    for (int e = tid; e < DP_VB * u2; e += DP_THREADS) max2[e] = 0;
    // ---- per-voxel constants: the means get_cluster subtracts, the cell centre in both frames ----
    if (tid < nv) {
        const long long v = v0 + tid;
        const float *mv = q.mean + (size_t)v * c;
        const int xi0 = q.cylinder ? 3 : 0, xi1 = q.cylinder ? 4 : 1;
        const int ri0 = q.cylinder ? 0 : c - 2, ri1 = q.cylinder ? 1 : c - 1;
        float *mt = meta + tid * 12;
        mt[0] = mv[xi0]; mt[1] = mv[xi1]; mt[2] = mv[2]; mt[3] = mv[ri0]; mt[4] = mv[ri1];
        const int4 u = reinterpret_cast<const int4 *>(q.unq)[v];                       // (b, z, y, x)
        const float c1 = __fadd_rn(__fmul_rn((float)u.w, q.vx), q.x_off);              // :350
        const float c2 = __fadd_rn(__fmul_rn((float)u.z, q.vy), q.y_off);              // :351
        float xc = c1, yc = c2, rc = c1, ac = c2;
        if (q.cylinder) { xc = __fmul_rn(c1, cosf(c2)); yc = __fmul_rn(c1, sinf(c2)); }          // polar2cart :240-249
        else { rc = __fsqrt_rn(__fadd_rn(__fmul_rn(c1, c1), __fmul_rn(c2, c2))); ac = atan2f(c2, c1); }   // cart2polar :251-260
        mt[5] = xc; mt[6] = yc; mt[7] = rc; mt[8] = ac;
    }
    __syncthreads();

    // decorated rows of one chunk -> D, voxel-local ids -> s_vl.  Two threads per point (all 256
    // threads busy: the others would only wait at the barrier): one copies the raw columns, the
    // other computes the decorations.
    auto build = [&](uint32_t p0, int npts) {
        const int pt = tid >> 1, half = tid & 1;
        if (pt < DP_PC) {
            float *row = D + (size_t)pt * c0p;
            if (pt < npts) {
                const uint32_t i = q.perm[p0 + pt];
                const float *pr = q.points + (size_t)i * c;
                const int vl = (int)(q.unq_inv[i] - v0);
                float pv[PV_MAX_CHANNELS];
#pragma unroll
                for (int k = 0; k < PV_MAX_CHANNELS; ++k) pv[k] = k < c ? __ldg(pr + k) : 0.0f;
                auto col = [&](int k) { float r = 0.0f;
#pragma unroll
                    for (int j = 0; j < PV_MAX_CHANNELS; ++j) r = j == k ? pv[j] : r;
                    return r; };
                if (half == 0) {
                    s_vl[pt] = vl;
                    for (int k = 0; k < c; ++k) row[k] = col(k);
                } else {
                    const float *mt = meta + vl * 12;
                    int o = c;
                    if (xyz_cluster) for (int k = 0; k < 3; ++k) row[o++] = __fsub_rn(col(xi[k]), mt[k]);       // :363-364
                    if (xy_center) { row[o++] = __fsub_rn(col(xi[0]), mt[5]); row[o++] = __fsub_rn(col(xi[1]), mt[6]); }   // :366-372
                    if (raz_cluster) {                                                                           // :373-384
                        row[o++] = __fsub_rn(col(ri[0]), mt[3]); row[o++] = __fsub_rn(col(ri[1]), mt[4]);
                        if (!xyz_cluster) row[o++] = __fsub_rn(col(2), mt[2]);
                    }
                    if (ra_center) { row[o++] = __fsub_rn(col(ri[0]), mt[7]); row[o++] = __fsub_rn(col(ri[1]), mt[8]); }   // :385-391
                    for (; o < c0p; ++o) row[o] = 0.0f;
                }
            } else if (half == 0) {
                s_vl[pt] = 0;
                for (int o = 0; o < c0p; ++o) row[o] = 0.0f;
            }
        }
    };

    if (u2 == 0) {
        // single (last) layer: decoration -> Linear -> ReLU -> max
        for (uint32_t p0 = pbeg; p0 < pend; p0 += DP_PC) {
            const int npts = (int)min((uint32_t)DP_PC, pend - p0);
            build(p0, npts);
            __syncthreads();
            dp_layer<false>(D, c0p, c0p, w1t, u1, npts, s_vl, nullptr, 0, max1);
            __syncthreads();
        }
        for (int e = tid; e < nv * u1; e += DP_THREADS) q.out[(size_t)v0 * u1 + e] = __int_as_float(max1[e]);
        continue;
    }
    // pass A: first-layer maxima of every voxel of the block
    for (uint32_t p0 = pbeg; p0 < pend; p0 += DP_PC) {
        const int npts = (int)min((uint32_t)DP_PC, pend - p0);
        build(p0, npts);
        __syncthreads();
        dp_layer<false>(D, c0p, c0p, w1t, u1, npts, s_vl, nullptr, 0, max1);
        __syncthreads();
    }
    // the last layer's input is cat([x, x_max[unq_inv]]) (:70): its x_max half is the same for all
    // points of a voxel, so W2[:, u1:] . x_max is taken once per voxel and seeds the accumulators
    // (fp32 sum in the order x_max terms first, then the point's own terms)
    for (int e = tid; e < DP_PC; e += DP_THREADS) s_vl[e] = e < DP_VB ? e : 0;
    __syncthreads();
    dp_layer<true, false>(reinterpret_cast<const float *>(max1), u1, u1, w2t + (size_t)u1 * u2, u2, nv, s_vl, P, u2, nullptr);
    __syncthreads();
    // pass B: recompute layer 1 (cheap), run the last layer on the point's own half
    for (uint32_t p0 = pbeg; p0 < pend; p0 += DP_PC) {
        const int npts = (int)min((uint32_t)DP_PC, pend - p0);
        build(p0, npts);
        __syncthreads();
        dp_layer<true>(D, c0p, c0p, w1t, u1, npts, s_vl, Y, u1, max1);     // maxima are final already: the atomics are no-ops
        __syncthreads();
        dp_layer<false>(Y, u1, u1, w2t, u2, npts, s_vl, nullptr, 0, max2, P);
        __syncthreads();
    }
    for (int e = tid; e < nv * u2; e += DP_THREADS) q.out[(size_t)v0 * u2 + e] = __int_as_float(max2[e]);
    }
}

static size_t dp_align(size_t v) { return (v + 255) / 256 * 256; }

extern "C" {

size_t pv_dynamic_pfn_workspace_bytes(int64_t n, int64_t m)
{
    if (n < 0 || m < 0) return 0;
    const size_t nb = (size_t)(m + 1 + SC_TILE - 1) / SC_TILE + 1;
    return dp_align((size_t)(m + 1) * 4) + dp_align((size_t)(m + 1) * 4) + dp_align((size_t)(n + 1) * 4) + dp_align(nb * 4 + 16);
}

int pv_dynamic_pfn(const float *points, const int32_t *unq, const int32_t *unq_inv, const int32_t *unq_cnt,
                   const float *voxel_mean, int64_t n, int64_t m, int32_t c, int32_t cylinder, int32_t flags,
                   float vx, float vy, float x_off, float y_off, const pv_pfn_layer *layers, int32_t n_layers,
                   void *workspace, size_t workspace_bytes, float *out, pv_stream_t stream)
{
    if (n < 0 || m < 0 || c < 3 || c > PV_MAX_CHANNELS || !layers) return PV_ERR_BAD_ARGUMENT;
    if (n_layers < 1 || n_layers > 2) return PV_ERR_UNSUPPORTED;
    if (m == 0) return PV_OK;
    if (!points || !unq || !unq_inv || !unq_cnt || !voxel_mean || !out || !workspace) return PV_ERR_BAD_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(unq) & 15u) != 0 || (reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return PV_ERR_BAD_ARGUMENT;
    if (workspace_bytes < pv_dynamic_pfn_workspace_bytes(n, m)) return PV_ERR_WORKSPACE;
    const int xyz_cluster = flags & 1, raz_cluster = flags & 2, xy_center = flags & 4, ra_center = flags & 8;
    const int c0 = c + (xyz_cluster ? 3 : 0) + (xy_center ? 2 : 0) + (raz_cluster ? (xyz_cluster ? 2 : 3) : 0) + (ra_center ? 2 : 0);
    if (c0 > DP_MAX_C0 || layers[0].in_channels != c0) return PV_ERR_UNSUPPORTED;
    DpParams q;
    q.points = points; q.unq = unq; q.unq_inv = unq_inv; q.unq_cnt = unq_cnt; q.mean = voxel_mean;
    q.n = n; q.m = m; q.c = c; q.c0 = c0; q.cylinder = cylinder ? 1 : 0; q.flags = flags;
    q.vx = vx; q.vy = vy; q.x_off = x_off; q.y_off = y_off; q.out = out;
    q.w1 = layers[0].weight; q.u1 = layers[0].units; q.w2 = nullptr; q.u2 = 0;
    if (q.u1 <= 0 || (q.u1 & 3) || !q.w1) return PV_ERR_UNSUPPORTED;
    if (n_layers == 2) {
        q.w2 = layers[1].weight; q.u2 = layers[1].units;
        if (q.u1 > DP_MAX_U1 || q.u2 <= 0 || (q.u2 & 3) || q.u2 > DP_MAX_U2 || layers[1].in_channels != 2 * q.u1 || !q.w2)
            return PV_ERR_UNSUPPORTED;
    } else if (q.u1 > DP_MAX_U2) return PV_ERR_UNSUPPORTED;
    char *w = (char *)workspace;
    uint32_t *offs = (uint32_t *)w;            w += dp_align((size_t)(m + 1) * 4);
    uint32_t *cursor = (uint32_t *)w;          w += dp_align((size_t)(m + 1) * 4);
    uint32_t *perm = (uint32_t *)w;            w += dp_align((size_t)(n + 1) * 4);
    uint32_t *sums = (uint32_t *)w;
    q.offs = offs; q.perm = perm;
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = (int)((m + 1 + SC_TILE - 1) / SC_TILE);
    k_dp_block_sums<<<nb, 256, 0, st>>>(unq_cnt, m, sums);
    k_dp_scan_sums<<<1, 1024, 0, st>>>(sums, nb, sums + nb);
    k_dp_apply<<<nb, 256, 0, st>>>(unq_cnt, m, sums, offs, cursor);
    if (n > 0) k_dp_sort<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(unq_inv, n, m, offs, cursor, perm);
    const int c0p = (c0 + 3) & ~3, in2 = 2 * q.u1;
    const size_t smem = sizeof(float) * ((size_t)c0p * q.u1 + (q.u2 ? (size_t)in2 * q.u2 : 0) + DP_VB * 12 +
                                         (size_t)DP_VB * q.u1 + (size_t)DP_VB * q.u2 + (size_t)DP_PC * c0p + (size_t)DP_PC * q.u1 +
                                         (size_t)DP_VB * q.u2);
    if (smem > 200 * 1024) return PV_ERR_UNSUPPORTED;
    if (smem + 1024 > 48 * 1024 &&
        cudaFuncSetAttribute(k_dyn_pfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return PV_ERR_CUDA;
    const long long groups = (m + DP_VB - 1) / DP_VB;
    const long long resident = 148ll * (smem > 110 * 1024 ? 1 : 2);   // persistent blocks (117 registers: two per SM)
    k_dyn_pfn<<<(unsigned)(groups < resident ? groups : resident), DP_THREADS, smem, st>>>(q);
    return pv_last_cuda_error();
}

}  // extern "C"
