// readers.cu -- drop-in kernels behind det3d's reader / scatter modules (sm_100a).
//
//   pv_vfe_mean     VoxelFeatureExtractorV3.forward   det3d/models/readers/voxel_encoder.py:15-22
//   pv_pfn_forward  PillarFeatureNet.forward (eval)   det3d/models/readers/pillar_encoder.py:131-169
//   pv_scatter      PointPillarsScatter.forward       det3d/models/readers/pillar_encoder.py:189-225
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

int pv_pfn_forward(const float *voxels, const int32_t *num_points, const int32_t *coors, int64_t m,
                   int32_t t, int32_t c, int32_t with_distance, float vx, float vy, float x_off,
                   float y_off, const pv_pfn_layer *layers, int32_t n_layers, float eps,
                   float *out, pv_stream_t stream)
{
    if (m < 0 || t <= 0 || c < 3 || !layers || n_layers <= 0) return PV_ERR_BAD_ARGUMENT;
    if (n_layers > PV_MAX_PFN_LAYERS || t > PFN_MAX_T) return PV_ERR_UNSUPPORTED;
    if (m == 0) return PV_OK;
    if (!voxels || !num_points || !coors || !out) return PV_ERR_BAD_ARGUMENT;
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
    const unsigned grid = (unsigned)min((long long)148 * 3, (long long)m);
    k_pfn_simt<<<grid, PFN_THREADS, smem, (cudaStream_t)stream>>>(a);
    return pv_last_cuda_error();
}

}  // extern "C"
