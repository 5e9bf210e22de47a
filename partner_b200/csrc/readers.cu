// readers.cu -- drop-in kernels behind det3d's reader / scatter modules (sm_100a).
//
//   pv_vfe_mean     VoxelFeatureExtractorV3.forward   det3d/models/readers/voxel_encoder.py:15-22
//   pv_pfn_forward  PillarFeatureNet.forward (eval)   det3d/models/readers/pillar_encoder.py:131-169
//   pv_scatter      PointPillarsScatter.forward       det3d/models/readers/pillar_encoder.py:189-225
#include <stdlib.h>

#include "pv_common.cuh"

// ---------------------------------------------------------------------------------------------
// mean VFE on a padded [m, t, c] tensor: one thread per (voxel, channel); sum in slot order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_vfe_mean(const float *__restrict__ voxels,
                                                  const int32_t *__restrict__ num, long long m,
                                                  int t, int c, float *__restrict__ out)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * c) return;
    const long long v = idx / c;
    const int k = (int)(idx - v * c);
    const float *src = voxels + v * t * c + k;
    float s = 0.0f;
    for (int j = 0; j < t; ++j) s = __fadd_rn(s, __ldg(src + (size_t)j * c));
    out[idx] = __fdiv_rn(s, (float)num[v]);
}

// ---------------------------------------------------------------------------------------------
// scatter: (1) BEV index map  map[b, y*nx+x] = max row id (last duplicate wins, as index_put_
// does on an ordered loop);  (2) canvas[b, ch, y, x] = map >= 0 ? feats[row, ch] : 0 -- the zero
// fill is fused into the only write of every canvas element.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bev_index(const int32_t *__restrict__ coors, long long m,
                                                   int batch, int ny, int nx,
                                                   int32_t *__restrict__ map,
                                                   long long *__restrict__ bev_index)
{
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= m) return;
    const int4 c = reinterpret_cast<const int4 *>(coors)[v];          // (b, z, y, x)
    const long long idx = (long long)c.z * nx + c.w;                  // :211
    if (bev_index) bev_index[v] = idx;
    if (c.x < 0 || c.x >= batch) return;                              // :207 batch mask
    if (c.z < 0 || c.z >= ny || c.w < 0 || c.w >= nx) return;
    atomicMax(map + (size_t)c.x * ny * nx + idx, (int32_t)v);
}

// the same with the number of valid rows read on the device (fused front ends: no host sync)
__global__ void __launch_bounds__(256) k_bev_index_dev(const int32_t *__restrict__ coors, const int32_t *__restrict__ total_rows,
                                                       long long cap, int batch, int ny, int nx, int32_t *__restrict__ map)
{
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= cap || v >= (long long)__ldg(total_rows)) return;
    const int4 c = reinterpret_cast<const int4 *>(coors)[v];          // (b, z, y, x)
    if (c.x < 0 || c.x >= batch || c.z < 0 || c.z >= ny || c.w < 0 || c.w >= nx) return;
    atomicMax(map + (size_t)c.x * ny * nx + (size_t)c.z * nx + c.w, (int32_t)v);
}

template <int VEC>
__global__ void __launch_bounds__(256) k_canvas_from_index(const float *__restrict__ feats,
                                                           const int32_t *__restrict__ map, int c,
                                                           uint32_t cells, float *__restrict__ canvas)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    const uint32_t cell0 = q * VEC;
    if (cell0 >= cells) return;
    int32_t row[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) row[k] = map[(size_t)b * cells + cell0 + k];
    float *dst = canvas + (size_t)b * c * cells + cell0;
    if (VEC == 4 && (c & 3) == 0 && (reinterpret_cast<uintptr_t>(feats) & 15u) == 0) {
        // four channels at a time: one 16-byte gather per cell (a gathered load costs one slot per
        // lane whatever its width), a 4 x 4 transpose in registers, one 16-byte store per channel
        for (int ch = 0; ch < c; ch += 4) {
            float4 r[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                r[k] = row[k] >= 0 ? __ldg(reinterpret_cast<const float4 *>(feats + (size_t)row[k] * c + ch)) : make_float4(0.f, 0.f, 0.f, 0.f);
            float *d = dst + (size_t)ch * cells;
            __stcs(reinterpret_cast<float4 *>(d), make_float4(r[0].x, r[1 % VEC].x, r[2 % VEC].x, r[3 % VEC].x));
            __stcs(reinterpret_cast<float4 *>(d + cells), make_float4(r[0].y, r[1 % VEC].y, r[2 % VEC].y, r[3 % VEC].y));
            __stcs(reinterpret_cast<float4 *>(d + 2 * (size_t)cells), make_float4(r[0].z, r[1 % VEC].z, r[2 % VEC].z, r[3 % VEC].z));
            __stcs(reinterpret_cast<float4 *>(d + 3 * (size_t)cells), make_float4(r[0].w, r[1 % VEC].w, r[2 % VEC].w, r[3 % VEC].w));
        }
        return;
    }
    for (int ch = 0; ch < c; ++ch) {
        float v[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] = row[k] >= 0 ? __ldg(feats + (size_t)row[k] * c + ch) : 0.0f;
        if (VEC == 4) __stcs(reinterpret_cast<float4 *>(dst + (size_t)ch * cells), make_float4(v[0], v[1], v[2], v[3]));
        else dst[(size_t)ch * cells] = v[0];
    }
}

// ---------------------------------------------------------------------------------------------
// PFN, fp32 SIMT baseline: one 128-thread block walks voxels; all layer weights live in shared
// memory (transposed, [k][o]); only the useful rows are evaluated -- the num valid rows plus ONE
// representative padded row when num < T (all padded rows of a voxel are identical, and the
// reference's max runs over them too: pillar_encoder.py:55).
// ---------------------------------------------------------------------------------------------
#define PFN_THREADS 128
#define PFN_MAX_T 32
#define PFN_MAX_W 128

struct PfnArgs {
    const float *voxels; const int32_t *num; const int32_t *coors;
    long long m; int t, c, with_distance;
    float vx, vy, x_off, y_off, eps;
    int n_layers;
    const float *w[PV_MAX_PFN_LAYERS], *mean[PV_MAX_PFN_LAYERS], *var[PV_MAX_PFN_LAYERS],
        *gamma[PV_MAX_PFN_LAYERS], *beta[PV_MAX_PFN_LAYERS];
    int in_w[PV_MAX_PFN_LAYERS], units[PV_MAX_PFN_LAYERS], w_off[PV_MAX_PFN_LAYERS];
    int w_total, stride;
    float *out;
};

__global__ void __launch_bounds__(PFN_THREADS) k_pfn_simt(const __grid_constant__ PfnArgs a)
{
    extern __shared__ __align__(16) float smem[];
    float *s_w = smem;                                  // all layers, transposed [k][o]
    float *s_bn = s_w + a.w_total;                      // per layer 4 * PFN_MAX_W: mean, invstd, gamma, beta
    float *s_a = s_bn + a.n_layers * 4 * PFN_MAX_W;     // [PFN_MAX_T][stride]
    float *s_b = s_a + PFN_MAX_T * a.stride;            // [PFN_MAX_T][stride]
    int *s_max = reinterpret_cast<int *>(s_b + PFN_MAX_T * a.stride);  // [PFN_MAX_W]
    __shared__ float s_mean[3];
    const int tid = threadIdx.x;

    for (int l = 0; l < a.n_layers; ++l) {
        const int K = a.in_w[l], U = a.units[l];
        for (int e = tid; e < K * U; e += PFN_THREADS) {
            const int o = e / K, k = e - o * K;          // weight is [U, K]
            s_w[a.w_off[l] + k * U + o] = a.w[l][e];
        }
        for (int o = tid; o < U; o += PFN_THREADS) {
            float *bn = s_bn + l * 4 * PFN_MAX_W;
            bn[o] = a.mean[l][o];
            bn[PFN_MAX_W + o] = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(a.var[l][o], a.eps)));
            bn[2 * PFN_MAX_W + o] = a.gamma[l][o];
            bn[3 * PFN_MAX_W + o] = a.beta[l][o];
        }
    }
    __syncthreads();

    const int c = a.c, t = a.t;
    const int c0 = c + 5 + (a.with_distance ? 1 : 0);
    for (long long v = blockIdx.x; v < a.m; v += gridDim.x) {
        const float *f = a.voxels + v * t * c;
        const int n = min(max(a.num[v], 0), t);
        const int R = n < t ? n + 1 : t;                 // rows evaluated (last one = padding)
        if (tid < 3) {                                   // :137-139 sum over all T slots / num
            float s = 0.0f;
            for (int j = 0; j < t; ++j) s = __fadd_rn(s, __ldg(f + j * c + tid));
            s_mean[tid] = __fdiv_rn(s, (float)a.num[v]);
        }
        __syncthreads();
        const float cx = __fadd_rn(__fmul_rn((float)a.coors[v * 4 + 3], a.vx), a.x_off);   // :146-147
        const float cy = __fadd_rn(__fmul_rn((float)a.coors[v * 4 + 2], a.vy), a.y_off);   // :149-150
        for (int e = tid; e < R * c0; e += PFN_THREADS) {
            const int r = e / c0, k = e - r * c0;
            float val = 0.0f;
            if (r < n) {                                 // padded row stays zero (:161-164 mask)
                const float *p = f + r * c;
                if (k < c) val = __ldg(p + k);
                else if (k < c + 3) val = __fsub_rn(__ldg(p + k - c), s_mean[k - c]);       // :140
                else if (k == c + 3) val = __fsub_rn(__ldg(p), cx);
                else if (k == c + 4) val = __fsub_rn(__ldg(p + 1), cy);
                else {
                    const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);           // :155
                    val = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
                }
            }
            s_a[r * a.stride + k] = val;
        }
        float *cur = s_a, *nxt = s_b;
        for (int l = 0; l < a.n_layers; ++l) {
            const int K = a.in_w[l], U = a.units[l];
            const bool last = (l == a.n_layers - 1);
            for (int o = tid; o < U; o += PFN_THREADS) s_max[o] = 0;   // relu output >= 0
            __syncthreads();
            const float *W = s_w + a.w_off[l];
            const float *bn = s_bn + l * 4 * PFN_MAX_W;
            const int groups = max(1, PFN_THREADS / U);
            const int grp = tid / U, o = tid - grp * U;
            if (grp < groups) {
                for (int oo = o; oo < U; oo += PFN_THREADS) {   // U > 128 never happens; loop runs once
                    const float mu = bn[oo], is = bn[PFN_MAX_W + oo], ga = bn[2 * PFN_MAX_W + oo],
                                be = bn[3 * PFN_MAX_W + oo];
                    float mx = 0.0f;
                    for (int r0 = grp; r0 < R; r0 += 4 * groups) {
                        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                        for (int k = 0; k < K; ++k) {
                            const float w = W[k * U + oo];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int r = r0 + q * groups;
                                if (r < R) acc[q] = __fmaf_rn(cur[r * a.stride + k], w, acc[q]);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int r = r0 + q * groups;
                            if (r < R) {
                                // ATen eval batch norm: (x - mean) * invstd * gamma + beta, then ReLU
                                float y = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(acc[q], mu), is), ga), be);
                                y = fmaxf(y, 0.0f);
                                if (!last) nxt[r * a.stride + oo] = y;
                                mx = fmaxf(mx, y);
                            }
                        }
                    }
                    atomicMax(s_max + oo, __float_as_int(mx));
                }
            }
            __syncthreads();
            if (last) {
                for (int oo = tid; oo < U; oo += PFN_THREADS) a.out[v * U + oo] = __int_as_float(s_max[oo]);
            } else {
                for (int e = tid; e < R * U; e += PFN_THREADS) {   // :59-60 concat the repeated max
                    const int r = e / U, oo = e - r * U;
                    nxt[r * a.stride + U + oo] = __int_as_float(s_max[oo]);
                }
                float *tmp = cur; cur = nxt; nxt = tmp;
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// PFN, tiled fp32 kernel.  Persistent 256-thread blocks; each packs whole voxels into tiles of at
// most 64 useful rows (valid points + one representative padded row per non-full voxel) and runs
// every layer on the tile out of shared memory:
//     Linear([x, x_max]) = Wa . x (per row)  +  Wb . x_max (per voxel)
// so the concatenated activation of pillar_encoder.py:59-61 is never formed.  GEMMs are register
// tiled: a warp owns 8 rows, a lane owns outputs lane, lane+32, ... (8 x N/32 accumulators),
// activations are read as broadcast float4 along K, weights conflict-free along N.
// Epilogue: ATen eval BatchNorm order, ReLU, max over the voxel's rows via shared atomicMax
// (post-ReLU values are >= 0, so integer order == float order).
// ---------------------------------------------------------------------------------------------
#define PT_THREADS 256
#define PT_ROWS 64           // rows per tile
#define PT_VOX 32            // voxels per tile (each contributes >= 2 rows, or exactly T)
#define PT_MAX_IN 24         // decorated input width, padded to a multiple of 4
#define PT_CHUNK 256         // voxels per dynamically scheduled chunk (rows per voxel vary a lot)

struct PtArgs {
    const float *voxels; const int32_t *num; const int32_t *coors;
    long long m; int t, c, with_distance;
    float vx, vy, x_off, y_off, eps;
    int n_layers;
    const float *w[PV_MAX_PFN_LAYERS], *mean[PV_MAX_PFN_LAYERS], *var[PV_MAX_PFN_LAYERS],
        *gamma[PV_MAX_PFN_LAYERS], *beta[PV_MAX_PFN_LAYERS];
    int in_w[PV_MAX_PFN_LAYERS], units[PV_MAX_PFN_LAYERS];
    int wa_off[PV_MAX_PFN_LAYERS], wb_off[PV_MAX_PFN_LAYERS];   // offsets of Wa^T / Wb^T in shared memory
    int ka[PV_MAX_PFN_LAYERS];                                   // per-row K (padded to 4)
    int w_total, xs;                                             // xs = activation row stride (floats)
    const float4 *prep;     // [m][2]: (mean x, mean y, mean z, num) and (pillar centre x, y, -, -)
    unsigned int *chunk_counter;   // dynamic scheduling: next chunk of PT_CHUNK voxels
    float *out;
};

// Per-voxel statistics, fully parallel (one thread per voxel): cluster mean of xyz summed over
// ALL T slots in slot order / num (pillar_encoder.py:137-139) and the pillar centre (:145-150).
__global__ void __launch_bounds__(256) k_pfn_prep(const float *__restrict__ voxels, const int32_t *__restrict__ num,
                                                  const int32_t *__restrict__ coors, long long m, int t, int c,
                                                  float vx, float vy, float x_off, float y_off,
                                                  float4 *__restrict__ prep, uint32_t *__restrict__ chunk_rows)
{
    __shared__ uint32_t s_rows[8];
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v == 0) *reinterpret_cast<unsigned int *>(prep + 2 * m) = 0u;    // chunk counter of k_pfn_tiled
    // useful rows of this block's PT_CHUNK voxels (block size == PT_CHUNK)
    {
        uint32_t rows = 0;
        if (v < m) { const int n = min(max(num[v], 0), t); rows = (uint32_t)(n < t ? n + 1 : t); }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, d);
        if ((threadIdx.x & 31) == 0) s_rows[threadIdx.x >> 5] = rows;
        __syncthreads();
        if (threadIdx.x == 0 && chunk_rows) {
            uint32_t tot = 0;
            for (int k = 0; k < 8; ++k) tot += s_rows[k];
            chunk_rows[blockIdx.x] = tot;
        }
    }
    if (v >= m) return;
    const float *f = voxels + v * t * c;
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    for (int j = 0; j < t; ++j) {
        sx = __fadd_rn(sx, __ldg(f + j * c));
        sy = __fadd_rn(sy, __ldg(f + j * c + 1));
        sz = __fadd_rn(sz, __ldg(f + j * c + 2));
    }
    const int n = num[v];
    const float nf = (float)n;
    const int4 co = reinterpret_cast<const int4 *>(coors)[v];
    prep[2 * v] = make_float4(__fdiv_rn(sx, nf), __fdiv_rn(sy, nf), __fdiv_rn(sz, nf), __int_as_float(n));
    prep[2 * v + 1] = make_float4(__fadd_rn(__fmul_rn((float)co.w, vx), x_off),
                                  __fadd_rn(__fmul_rn((float)co.z, vy), y_off), 0.0f, 0.0f);
}

// acc[i][j] += sum_k A[row0 + i][k] * Wt[k][J * lane + j]   for the 8 rows of this warp.
// A rows are read as broadcast float4 along K, the lane's J contiguous weights as one vector.
template <int J>
__device__ __forceinline__ void pt_fma_row(float av, const float (&w)[J], float (&acc)[J])
{
#pragma unroll
    for (int j = 0; j < J; ++j) acc[j] = __fmaf_rn(av, w[j], acc[j]);
}

template <int J>
__device__ __forceinline__ void pt_load_w(const float *p, float (&w)[J])
{
    if (J == 4) { const float4 v = *reinterpret_cast<const float4 *>(p); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
    else if (J == 2) { const float2 v = *reinterpret_cast<const float2 *>(p); w[0] = v.x; w[1] = v.y; }
    else {
#pragma unroll
        for (int j = 0; j < J; ++j) w[j] = p[j];
    }
}

template <int J>
__device__ __forceinline__ void pt_gemm(const float *__restrict__ A, int astride, int K, int row0,
                                        const float *__restrict__ Wt, int N, int lane, float (&acc)[8][J])
{
    const float *wp = Wt + J * lane;
    for (int k4 = 0; k4 < K; k4 += 4) {
        float4 a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4 *>(A + (row0 + i) * astride + k4);
        float w[J];
        pt_load_w<J>(wp + (k4 + 0) * N, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) pt_fma_row<J>(a[i].x, w, acc[i]);
        pt_load_w<J>(wp + (k4 + 1) * N, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) pt_fma_row<J>(a[i].y, w, acc[i]);
        pt_load_w<J>(wp + (k4 + 2) * N, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) pt_fma_row<J>(a[i].z, w, acc[i]);
        pt_load_w<J>(wp + (k4 + 3) * N, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) pt_fma_row<J>(a[i].w, w, acc[i]);
    }
}

// One layer on the current tile.  xin: [PT_ROWS][xs] per-row input (K = ka), vmax_in: [PT_VOX][128]
// per-voxel max of the previous layer (layer > 0).  Writes xout (non-last layers) and vmax_out.
template <int J>
__device__ __forceinline__ void pt_layer(const PtArgs &a, int l, int n_rows, int n_vox, const float *s_w,
                                         const float *s_bn, const float *xin, const float *vmax_in,
                                         float *xout, int *vmax_out, float *s_p, const int *s_row_vox)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = a.units[l];
    const bool last = l == a.n_layers - 1;
    const float *bn = s_bn + l * 4 * PFN_MAX_W;
    // ---- per-voxel part: P[v][o] = Wb . vmax_in[v]  (layers > 0) ----
    if (l > 0) {
        const int Kb = a.units[l - 1];
        const int v0 = warp * 8;
        if (v0 < n_vox) {
            float acc[8][J];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < J; ++j) acc[i][j] = 0.0f;
            pt_gemm<J>(vmax_in, PFN_MAX_W, Kb, v0, s_w + a.wb_off[l], N, lane, acc);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < J; ++j) s_p[(v0 + i) * PFN_MAX_W + J * lane + j] = acc[i][j];
        }
    }
    for (int vl = warp; vl < n_vox; vl += PT_THREADS / 32) {
#pragma unroll
        for (int j = 0; j < J; ++j) vmax_out[vl * PFN_MAX_W + J * lane + j] = 0;
    }
    __syncthreads();
    // ---- per-row part + epilogue ----
    const int row0 = warp * 8;
    if (row0 < n_rows) {
        float acc[8][J];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < J; ++j) acc[i][j] = 0.0f;
        pt_gemm<J>(xin, a.xs, a.ka[l], row0, s_w + a.wa_off[l], N, lane, acc);
        float mu[J], is[J], ga[J], be[J], run[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int o = J * lane + j;
            mu[j] = bn[o]; is[j] = bn[PFN_MAX_W + o]; ga[j] = bn[2 * PFN_MAX_W + o]; be[j] = bn[3 * PFN_MAX_W + o];
            run[j] = 0.0f;
        }
        int cur = s_row_vox[row0];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = row0 + i;
            if (r < n_rows) {                              // warp-uniform
                const int v = s_row_vox[r];
                if (v != cur) {                            // rows are grouped by voxel: flush the finished run
#pragma unroll
                    for (int j = 0; j < J; ++j) { atomicMax(vmax_out + cur * PFN_MAX_W + J * lane + j, __float_as_int(run[j])); run[j] = 0.0f; }
                    cur = v;
                }
                float y[J];
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    float x = acc[i][j];
                    if (l > 0) x = __fadd_rn(x, s_p[v * PFN_MAX_W + J * lane + j]);
                    // ATen eval batch norm: (x - mean) * invstd * gamma + beta, then ReLU
                    y[j] = fmaxf(__fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(x, mu[j]), is[j]), ga[j]), be[j]), 0.0f);
                    run[j] = fmaxf(run[j], y[j]);
                }
                if (!last) {
#pragma unroll
                    for (int j = 0; j < J; ++j) xout[r * a.xs + J * lane + j] = y[j];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < J; ++j) atomicMax(vmax_out + cur * PFN_MAX_W + J * lane + j, __float_as_int(run[j]));
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PT_THREADS, 2) k_pfn_tiled(const __grid_constant__ PtArgs a)
{
    extern __shared__ __align__(16) float smem[];
    float *s_w = smem;                                             // all layers: Wa^T then Wb^T, [k][o]
    float *s_bn = s_w + a.w_total;                                 // per layer 4 * PFN_MAX_W
    float *s_xa = s_bn + a.n_layers * 4 * PFN_MAX_W;               // [PT_ROWS][xs]
    float *s_xb = s_xa + PT_ROWS * a.xs;                           // [PT_ROWS][xs]
    float *s_va = s_xb + PT_ROWS * a.xs;                           // [PT_VOX][PFN_MAX_W] voxel max (ping)
    float *s_vb = s_va + PT_VOX * PFN_MAX_W;                       // [PT_VOX][PFN_MAX_W] voxel max (pong)
    float *s_p = s_vb + PT_VOX * PFN_MAX_W;                        // [PT_VOX][PFN_MAX_W] per-voxel term
    int *s_row_vox = reinterpret_cast<int *>(s_p + PT_VOX * PFN_MAX_W);   // [PT_ROWS]
    int *s_row_t = s_row_vox + PT_ROWS;                            // [PT_ROWS]
    float *s_mean = reinterpret_cast<float *>(s_row_t + PT_ROWS);  // [PT_VOX][8]: mean xyz, n, centre xy
    int *s_vrows = reinterpret_cast<int *>(s_mean + PT_VOX * 8);   // [PT_VOX + 1] first row of each voxel
    int *s_vn = s_vrows + PT_VOX + 4;                              // [PT_VOX] valid points per voxel
    __shared__ int s_nvox, s_nrows;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = a.c, t = a.t;
    const int c0 = c + 5 + (a.with_distance ? 1 : 0);

    // ---- weights: Linear weight is [U, K]; stored transposed, split into the per-row part
    // (columns < in_w - units[l-1] for l > 0, all columns for l = 0) and the per-voxel part ----
    for (int l = 0; l < a.n_layers; ++l) {
        const int K = a.in_w[l], U = a.units[l];
        const int Ka = l == 0 ? K : K - a.units[l - 1];
        for (int e = tid; e < a.ka[l] * U; e += PT_THREADS) {
            const int k = e / U, o = e - k * U;
            s_w[a.wa_off[l] + e] = k < Ka ? a.w[l][o * K + k] : 0.0f;
        }
        if (l > 0) {
            const int Kb = a.units[l - 1];
            for (int e = tid; e < Kb * U; e += PT_THREADS) {
                const int k = e / U, o = e - k * U;
                s_w[a.wb_off[l] + e] = a.w[l][o * K + Ka + k];
            }
        }
        for (int o = tid; o < U; o += PT_THREADS) {
            float *bn = s_bn + l * 4 * PFN_MAX_W;
            bn[o] = a.mean[l][o];
            bn[PFN_MAX_W + o] = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(a.var[l][o], a.eps)));
            bn[2 * PFN_MAX_W + o] = a.gamma[l][o];
            bn[3 * PFN_MAX_W + o] = a.beta[l][o];
        }
    }
    __syncthreads();

    // ---- chunks of PT_CHUNK consecutive voxels, handed out dynamically ----
    __shared__ unsigned int s_chunk;
    while (true) {
    if (tid == 0) s_chunk = atomicAdd(a.chunk_counter, 1u);
    __syncthreads();
    long long v_next = (long long)s_chunk * PT_CHUNK;
    if (v_next >= a.m) break;
    const long long v_end = min(a.m, v_next + (long long)PT_CHUNK);
    // packing info is fetched one tile ahead (warp 0: lane i looks at voxel v_next + i)
    int n_ahead = 0;
    if (warp == 0 && v_next + lane < v_end) n_ahead = __float_as_int(__ldg(&a.prep[2 * (v_next + lane)].w));
    while (v_next < v_end) {
        // pack whole voxels greedily into <= PT_ROWS rows (warp 0)
        if (warp == 0) {
            const long long v = v_next + lane;
            int rows = 0;
            if (v < v_end) {
                const int n = min(max(n_ahead, 0), t);
                rows = n < t ? n + 1 : t;
            }
            int incl = rows;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            const unsigned fits = __ballot_sync(0xffffffffu, v < v_end && incl <= PT_ROWS);
            const int nv = __popc(fits);                  // prefix property: fits is a run of low bits
            if (lane < nv) { s_vrows[lane] = incl - rows; s_vn[lane] = min(max(n_ahead, 0), t); }
            if (lane == nv - 1) { s_vrows[nv] = incl; s_nrows = incl; }
            if (lane == 0) s_nvox = nv;
            // prefetch the packing info of the next tile
            const long long vn = v_next + nv + lane;
            n_ahead = vn < v_end ? __float_as_int(__ldg(&a.prep[2 * vn].w)) : 0;
        }
        __syncthreads();
        const int n_vox = s_nvox, n_rows = s_nrows;
        const long long vbase = v_next;
        // per-voxel statistics and row maps
        for (int vl = warp; vl < n_vox; vl += PT_THREADS / 32) {
            if (lane < 2) {
                const float4 q = __ldg(&a.prep[2 * (vbase + vl) + lane]);
                reinterpret_cast<float4 *>(s_mean)[vl * 2 + lane] = q;
            }
            const int r0 = s_vrows[vl], r1 = s_vrows[vl + 1];
            for (int r = r0 + lane; r < r1; r += 32) { s_row_vox[r] = vl; s_row_t[r] = r - r0; }
        }
        __syncthreads();
        // decorated input rows (:140-164): warp per row, lane per column; rows >= n_rows and the
        // padded representatives are zero
        for (int r = warp; r < PT_ROWS; r += PT_THREADS / 32) {
            float val = 0.0f;
            if (r < n_rows && lane < c0) {
                const int vl = s_row_vox[r], tt = s_row_t[r];
                if (tt < s_vn[vl]) {
                    const float *p = a.voxels + ((vbase + vl) * t + tt) * c;
                    const int k = lane;
                    if (k < c) val = __ldg(p + k);
                    else if (k < c + 3) val = __fsub_rn(__ldg(p + k - c), s_mean[vl * 8 + k - c]);       // :140
                    else if (k == c + 3) val = __fsub_rn(__ldg(p), s_mean[vl * 8 + 4]);                    // :146-147
                    else if (k == c + 4) val = __fsub_rn(__ldg(p + 1), s_mean[vl * 8 + 5]);                // :149-150
                    else {
                        const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);                      // :155
                        val = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
                    }
                }
            }
            if (lane < a.ka[0]) s_xa[r * a.xs + lane] = val;
        }
        __syncthreads();
        // ---- layers ----
        float *xin = s_xa, *xout = s_xb;
        float *vin = s_va, *vout = s_vb;
        for (int l = 0; l < a.n_layers; ++l) {
            switch (a.units[l] >> 5) {
            case 1: pt_layer<1>(a, l, n_rows, n_vox, s_w, s_bn, xin, vin, xout, reinterpret_cast<int *>(vout), s_p, s_row_vox); break;
            case 2: pt_layer<2>(a, l, n_rows, n_vox, s_w, s_bn, xin, vin, xout, reinterpret_cast<int *>(vout), s_p, s_row_vox); break;
            case 3: pt_layer<3>(a, l, n_rows, n_vox, s_w, s_bn, xin, vin, xout, reinterpret_cast<int *>(vout), s_p, s_row_vox); break;
            default: pt_layer<4>(a, l, n_rows, n_vox, s_w, s_bn, xin, vin, xout, reinterpret_cast<int *>(vout), s_p, s_row_vox); break;
            }
            float *tmp = xin; xin = xout; xout = tmp;
            tmp = vin; vin = vout; vout = tmp;
        }
        // ---- output: max of the last layer (vin after the swap) ----
        const int U = a.units[a.n_layers - 1];
        for (int vl = warp; vl < n_vox; vl += PT_THREADS / 32) {
            float *dst = a.out + (vbase + vl) * U;
            for (int o = lane; o < U; o += 32) __stcs(dst + o, vin[vl * PFN_MAX_W + o]);
        }
        __syncthreads();
        v_next += n_vox;
    }
    }
}

// pfn_fused.cu: the warp-specialised kernel with the second layer on tcgen05 (default for two-layer nets)
struct P2Args;
bool pv_pfn_fused_supported(const pv_pfn_layer *layers, int n_layers, int t, int c, int with_distance);
int pv_pfn_fused_tensor(const float *voxels, const int32_t *num_points, const int32_t *coors, int64_t m, int32_t t,
                        int32_t c, int32_t with_distance, float vx, float vy, float x_off, float y_off,
                        const pv_pfn_layer *layers, float eps, void *workspace, float *out, cudaStream_t st);
size_t pv_pfn_fused_tensor_bytes(long long m, int t);     // scratch of the call above (queue words + decorated rows of one slice)

static int pfn_tiled_supported(const pv_pfn_layer *layers, int n_layers, int t)
{
    if (t < 2 || t > PT_ROWS) return 0;
    for (int l = 0; l < n_layers; ++l)
        if (layers[l].units % 32 != 0 || layers[l].units > PFN_MAX_W) return 0;
    return 1;
}

extern "C" size_t pv_scatter_workspace_bytes(int32_t batch, int32_t ny, int32_t nx);

// PointPillarsScatter for a fused front end: rows [0, *total_rows) of feats / coors are valid.
int pv_scatter_dev(const float *feats, const int32_t *coors, const int32_t *total_rows, int64_t cap, int32_t c,
                   int32_t batch, int32_t ny, int32_t nx, void *workspace, float *canvas, cudaStream_t st)
{
    const size_t need = pv_scatter_workspace_bytes(batch, ny, nx);
    int32_t *map = (int32_t *)workspace;
    if (cudaMemsetAsync(map, 0xFF, need, st) != cudaSuccess) return PV_ERR_CUDA;
    if (cap > 0) k_bev_index_dev<<<(unsigned)((cap + 255) / 256), 256, 0, st>>>(coors, total_rows, cap, batch, ny, nx, map);
    const uint32_t cells = (uint32_t)ny * (uint32_t)nx;
    const bool vec = (cells % 4 == 0) && ((reinterpret_cast<uintptr_t>(canvas) & 15u) == 0);
    if (vec) k_canvas_from_index<4><<<dim3((cells / 4 + 255) / 256, (unsigned)batch), 256, 0, st>>>(feats, map, c, cells, canvas);
    else k_canvas_from_index<1><<<dim3((cells + 255) / 256, (unsigned)batch), 256, 0, st>>>(feats, map, c, cells, canvas);
    return pv_last_cuda_error();
}

extern "C" {

int pv_vfe_mean(const float *voxels, const int32_t *num_points, int64_t m, int32_t t, int32_t c,
                float *out, pv_stream_t stream)
{
    if (m < 0 || t <= 0 || c <= 0) return PV_ERR_BAD_ARGUMENT;
    if (m == 0) return PV_OK;
    if (!voxels || !num_points || !out) return PV_ERR_BAD_ARGUMENT;
    const long long total = (long long)m * c;
    k_vfe_mean<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(voxels, num_points, m, t, c, out);
    return pv_last_cuda_error();
}

size_t pv_scatter_workspace_bytes(int32_t batch, int32_t ny, int32_t nx)
{
    if (batch <= 0 || ny <= 0 || nx <= 0) return 0;
    return (size_t)batch * ny * nx * sizeof(int32_t);
}

int pv_scatter(const float *feats, const int32_t *coors, int64_t m, int32_t c, int32_t batch,
               int32_t ny, int32_t nx, void *workspace, size_t workspace_bytes, float *canvas,
               int64_t *bev_index, pv_stream_t stream)
{
    if (m < 0 || c <= 0 || batch <= 0 || ny <= 0 || nx <= 0 || !canvas || !workspace) return PV_ERR_BAD_ARGUMENT;
    if (m > 0 && (!feats || !coors)) return PV_ERR_BAD_ARGUMENT;
    if (m >= (1ll << 31)) return PV_ERR_BAD_ARGUMENT;
    const size_t need = pv_scatter_workspace_bytes(batch, ny, nx);
    if (workspace_bytes < need) return PV_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(coors) & 15u) != 0) return PV_ERR_BAD_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *map = (int32_t *)workspace;
    if (cudaMemsetAsync(map, 0xFF, need, st) != cudaSuccess) return PV_ERR_CUDA;
    if (m > 0)
        k_bev_index<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(coors, m, batch, ny, nx, map, (long long *)bev_index);
    const uint32_t cells = (uint32_t)ny * (uint32_t)nx;
    const bool vec = (cells % 4 == 0) && ((reinterpret_cast<uintptr_t>(canvas) & 15u) == 0);
    if (vec) {
        dim3 grid((cells / 4 + 255) / 256, (unsigned)batch);
        k_canvas_from_index<4><<<grid, 256, 0, st>>>(feats, map, c, cells, canvas);
    } else {
        dim3 grid((cells + 255) / 256, (unsigned)batch);
        k_canvas_from_index<1><<<grid, 256, 0, st>>>(feats, map, c, cells, canvas);
    }
    return pv_last_cuda_error();
}

struct PfnWsLayout { size_t prep, chunk_rows, total; size_t chunks; };
static PfnWsLayout pfn_ws_layout(int64_t m, int32_t t)
{
    (void)t;
    PfnWsLayout L;
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    L.chunks = (size_t)(m + PT_CHUNK - 1) / PT_CHUNK;
    size_t o = 0;
    L.prep = o;        o = up(o + ((size_t)m * 2 + 1) * sizeof(float4));     // also holds the chunk counters (last 16 bytes)
    L.chunk_rows = o;  o = up(o + (L.chunks + 1) * 4);
    L.total = o;
    return L;
}

size_t pv_pfn_workspace_bytes(int64_t m, int32_t t)
{
    if (m <= 0 || t <= 0) return 0;
    const size_t tiled = pfn_ws_layout(m, t).total, fused = pv_pfn_fused_tensor_bytes(m, t);
    return tiled > fused ? tiled : fused;
}

int pv_pfn_forward(const float *voxels, const int32_t *num_points, const int32_t *coors, int64_t m,
                   int32_t t, int32_t c, int32_t with_distance, float vx, float vy, float x_off,
                   float y_off, const pv_pfn_layer *layers, int32_t n_layers, float eps,
                   void *workspace, size_t workspace_bytes, float *out, pv_stream_t stream)
{
    if (m < 0 || t <= 0 || c < 3 || !layers || n_layers <= 0) return PV_ERR_BAD_ARGUMENT;
    if (n_layers > PV_MAX_PFN_LAYERS || t > PFN_MAX_T) return PV_ERR_UNSUPPORTED;
    if (m == 0) return PV_OK;
    if (!voxels || !num_points || !coors || !out) return PV_ERR_BAD_ARGUMENT;
    for (int l = 0; l < n_layers; ++l)
        if (!layers[l].weight || !layers[l].bn_mean || !layers[l].bn_var || !layers[l].bn_gamma || !layers[l].bn_beta)
            return PV_ERR_BAD_ARGUMENT;
    // two-layer nets (every PFN the reference's configs build): the warp-specialised kernel, second
    // layer on the tensor cores -- no environment switch, no global state
    if (pv_pfn_fused_supported(layers, n_layers, t, c, with_distance) && workspace && workspace_bytes >= pv_pfn_fused_tensor_bytes(m, t) &&
        (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0 && (reinterpret_cast<uintptr_t>(coors) & 15u) == 0 &&
        (reinterpret_cast<uintptr_t>(out) & 15u) == 0)
    {
        const int rc = pv_pfn_fused_tensor(voxels, num_points, coors, m, t, c, with_distance, vx, vy, x_off, y_off, layers, eps,
                                           workspace, out, (cudaStream_t)stream);
        if (rc != PV_ERR_UNSUPPORTED) return rc;
    }
    if (pfn_tiled_supported(layers, n_layers, t) && c + 5 + (with_distance ? 1 : 0) <= PT_MAX_IN) {
        PtArgs q;
        q.voxels = voxels; q.num = num_points; q.coors = coors; q.m = m; q.t = t; q.c = c;
        q.with_distance = with_distance ? 1 : 0;
        q.vx = vx; q.vy = vy; q.x_off = x_off; q.y_off = y_off; q.eps = eps; q.n_layers = n_layers; q.out = out;
        int width = c + 5 + q.with_distance, off = 0, xs = 0;
        for (int l = 0; l < n_layers; ++l) {
            const pv_pfn_layer &L = layers[l];
            if (L.in_channels != width) return PV_ERR_BAD_ARGUMENT;
            if (!L.weight || !L.bn_mean || !L.bn_var || !L.bn_gamma || !L.bn_beta) return PV_ERR_BAD_ARGUMENT;
            q.w[l] = L.weight; q.mean[l] = L.bn_mean; q.var[l] = L.bn_var; q.gamma[l] = L.bn_gamma; q.beta[l] = L.bn_beta;
            q.in_w[l] = L.in_channels; q.units[l] = L.units;
            const int ka = l == 0 ? L.in_channels : L.in_channels - layers[l - 1].units;   // per-row K
            q.ka[l] = (ka + 3) & ~3;
            q.wa_off[l] = off; off += q.ka[l] * L.units;
            q.wb_off[l] = off; if (l > 0) off += layers[l - 1].units * L.units;
            if (q.ka[l] > xs) xs = q.ka[l];
            if (l < n_layers - 1 && L.units > xs) xs = L.units;
            width = (l == n_layers - 1) ? L.units : 2 * L.units;
        }
        q.w_total = (off + 3) & ~3;
        q.xs = xs + 4;                                      // +4: rows start in different banks
        const size_t smem = sizeof(float) * ((size_t)q.w_total + (size_t)n_layers * 4 * PFN_MAX_W +
                                             2 * (size_t)PT_ROWS * q.xs + 3 * (size_t)PT_VOX * PFN_MAX_W + PT_VOX * 8) +
                            sizeof(int) * (2 * PT_ROWS + 2 * PT_VOX + 8) + 64;
        if (smem <= 110 * 1024 && workspace && workspace_bytes >= pv_pfn_workspace_bytes(m, t) &&
            (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0 && (reinterpret_cast<uintptr_t>(coors) & 15u) == 0) {
            const PfnWsLayout W = pfn_ws_layout(m, t);
            char *ws = reinterpret_cast<char *>(workspace);
            cudaStream_t st = (cudaStream_t)stream;
            q.prep = reinterpret_cast<const float4 *>(ws + W.prep);
            q.chunk_counter = reinterpret_cast<unsigned int *>(reinterpret_cast<float4 *>(ws + W.prep) + 2 * m);
            uint32_t *chunk_rows = reinterpret_cast<uint32_t *>(ws + W.chunk_rows);
            k_pfn_prep<<<(unsigned)W.chunks, PT_CHUNK, 0, st>>>(
                voxels, num_points, coors, m, t, c, vx, vy, x_off, y_off, reinterpret_cast<float4 *>(ws + W.prep), chunk_rows);
            if (cudaFuncSetAttribute(k_pfn_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                return PV_ERR_CUDA;
            const long long want = (m + PT_CHUNK - 1) / PT_CHUNK;
            const long long cap = 2ll * pv_sm_count();            // two resident blocks per SM
            const unsigned grid = (unsigned)(want < cap ? want : cap);
            k_pfn_tiled<<<grid, PT_THREADS, smem, st>>>(q);
            return pv_last_cuda_error();
        }
    }
    PfnArgs a;
    a.voxels = voxels; a.num = num_points; a.coors = coors; a.m = m; a.t = t; a.c = c;
    a.with_distance = with_distance ? 1 : 0;
    a.vx = vx; a.vy = vy; a.x_off = x_off; a.y_off = y_off; a.eps = eps; a.n_layers = n_layers;
    a.out = out;
    int width = c + 5 + a.with_distance, off = 0, stride = width;
    for (int l = 0; l < n_layers; ++l) {
        const pv_pfn_layer &L = layers[l];
        if (L.in_channels != width || L.units <= 0 || L.units > PFN_MAX_W || L.in_channels > PFN_MAX_W)
            return L.in_channels != width ? PV_ERR_BAD_ARGUMENT : PV_ERR_UNSUPPORTED;
        if (!L.weight || !L.bn_mean || !L.bn_var || !L.bn_gamma || !L.bn_beta) return PV_ERR_BAD_ARGUMENT;
        a.w[l] = L.weight; a.mean[l] = L.bn_mean; a.var[l] = L.bn_var; a.gamma[l] = L.bn_gamma; a.beta[l] = L.bn_beta;
        a.in_w[l] = L.in_channels; a.units[l] = L.units; a.w_off[l] = off;
        off += L.in_channels * L.units;
        width = (l == n_layers - 1) ? L.units : 2 * L.units;
        if (l < n_layers - 1 && width > PFN_MAX_W) return PV_ERR_UNSUPPORTED;
        if (width > stride) stride = width;
    }
    a.w_total = (off + 3) & ~3;
    a.stride = stride | 1;   // odd stride: rows land in different banks
    const size_t smem = sizeof(float) * ((size_t)a.w_total + (size_t)n_layers * 4 * PFN_MAX_W +
                                         2 * (size_t)PFN_MAX_T * a.stride) + sizeof(int) * PFN_MAX_W;
    if (smem > 200 * 1024) return PV_ERR_UNSUPPORTED;
    if (cudaFuncSetAttribute(k_pfn_simt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return PV_ERR_CUDA;
    const unsigned grid = (unsigned)min((long long)pv_sm_count() * 3, (long long)m);
    k_pfn_simt<<<grid, PFN_THREADS, smem, (cudaStream_t)stream>>>(a);
    return pv_last_cuda_error();
}

}  // extern "C"
