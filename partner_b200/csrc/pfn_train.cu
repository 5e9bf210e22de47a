// pfn_train.cu -- PillarFeatureNet in TRAINING mode: forward with batch statistics and backward (sm_100a).
//
// Reference: PFNLayer.forward_static with self.norm in training mode, det3d/models/readers/pillar_encoder.py
// :37-38,49-61 (BatchNorm1d, eps 1e-3, momentum 0.01), inside PillarFeatureNet.forward :131-169, and what
// torch.autograd derives from it.  The layer sees the PADDED tensor: BatchNorm1d normalises over ALL M * T
// rows, padded slots included (their layer-0 input is zero, so they pull the mean towards 0; in later
// layers they carry [relu(shift), x_max]), running statistics are updated with the unbiased variance, and
// the maximum runs over all T slots.  To mirror that exactly -- including which rows the statistics and
// the gradients see -- these kernels work on the padded formulation, row r = voxel * T + slot; training
// batches are small next to the inference batches the tensor-core kernel serves, and every buffer the
// backward pass needs (layer inputs X_l, pre-norm outputs Z_l, batch mean / invstd, arg-max rows) is kept
// in the caller's workspace, which the autograd function holds on to.
//
//   forward, per layer l:   Z = X W^T;  mu, var = batch statistics over the M * T rows (two passes);
//                           Y = relu((Z - mu) * invstd * gamma + beta);  x_max[v] = max_t Y[v, t], arg-max row;
//                           running = (1 - m) running + m * {mu, var * N / (N - 1)};
//                           X_{l+1} = [Y | x_max repeated over t]           (not for the last layer)
//   backward, l = last .. 0: G = dY: the x_max gradient routed to its arg-max row (+ the direct gradient of
//                           X_{l+1}'s Y half), masked by Y > 0;   dbeta = sum G, dgamma = sum G xhat;
//                           dZ = gamma invstd (G - dbeta / N - xhat dgamma / N);   dW = dZ^T X;
//                           dX = dZ W  ->  its Y half is the next G, its x_max half summed over t.
#include <algorithm>

#include "pv_common.cuh"

#define PTR_THREADS 256

struct PtrLayer {
    const float *w, *gamma, *beta;
    float *run_mean, *run_var;
    int k, u;                  // input width, units
    float *x, *z;              // [R, k], [R, u]
    float *stat;               // [4 * u]: mean, invstd, (backward) dbeta, dgamma
    double *sums;              // [4 * u]: the column sums behind them, accumulated in double
    float *xmax;               // [M, u]
    int32_t *arg;              // [M, u] slot of the maximum
};

// decorated, masked input rows of layer 0 (:137-164)
__global__ void __launch_bounds__(PTR_THREADS) k_tr_decorate(const float *__restrict__ voxels, const int32_t *__restrict__ num,
                                                             const int32_t *__restrict__ coors, long long m, int t, int c,
                                                             int with_distance, float vx, float vy, float x_off, float y_off,
                                                             float *__restrict__ x0, int c0)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m * t) return;
    const long long v = r / t;
    const int q = (int)(r - v * t);
    const int n = num[v];
    float *o = x0 + r * c0;
    if (q >= n) {                                           // :161-164 mask
        for (int k = 0; k < c0; ++k) o[k] = 0.0f;
        return;
    }
    const float *f = voxels + v * t * c;
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;                   // :137-139 sum over all T slots / num
    for (int j = 0; j < t; ++j) {
        sx = __fadd_rn(sx, f[j * c]); sy = __fadd_rn(sy, f[j * c + 1]); sz = __fadd_rn(sz, f[j * c + 2]);
    }
    const float nf = (float)n;
    const float *p = f + q * c;
    for (int k = 0; k < c; ++k) o[k] = p[k];
    o[c] = __fsub_rn(p[0], __fdiv_rn(sx, nf));
    o[c + 1] = __fsub_rn(p[1], __fdiv_rn(sy, nf));
    o[c + 2] = __fsub_rn(p[2], __fdiv_rn(sz, nf));
    o[c + 3] = __fsub_rn(p[0], __fadd_rn(__fmul_rn((float)coors[v * 4 + 3], vx), x_off));
    o[c + 4] = __fsub_rn(p[1], __fadd_rn(__fmul_rn((float)coors[v * 4 + 2], vy), y_off));
    if (with_distance) o[c + 5] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(p[0], p[0]), __fmul_rn(p[1], p[1])), __fmul_rn(p[2], p[2])));
}

// Z[r, o] = sum_k X[r, k] W[o, k]: one thread per (row, unit); W through shared memory
__global__ void __launch_bounds__(PTR_THREADS) k_tr_linear(const float *__restrict__ x, const float *__restrict__ w, long long rows,
                                                           int k, int u, float *__restrict__ z)
{
    extern __shared__ float s_w[];                           // [u][k + 1]
    for (int e = threadIdx.x; e < u * k; e += blockDim.x) s_w[(e / k) * (k + 1) + e % k] = w[e];
    __syncthreads();
    const int rpb = blockDim.x / u;                          // rows per block pass (u divides the block size or is larger)
    if (rpb == 0) return;
    const int o = threadIdx.x % u, rl = threadIdx.x / u;
    if (rl >= rpb) return;
    for (long long r = (long long)blockIdx.x * rpb + rl; r < rows; r += (long long)gridDim.x * rpb) {
        const float *xr = x + r * k;
        float acc = 0.0f;
        for (int j = 0; j < k; ++j) acc = __fmaf_rn(xr[j], s_w[o * (k + 1) + j], acc);
        z[r * u + o] = acc;
    }
}

// column sums of f(z) over the rows, accumulated into out[u] (zeroed by the caller).  MODE 0: z;  1: (z - mean)^2
template <int MODE>
__global__ void __launch_bounds__(PTR_THREADS) k_tr_colsum(const float *__restrict__ z, long long rows, int u,
                                                           const float *__restrict__ mean, double *__restrict__ out)
{
    const int o = threadIdx.x % u, rl = threadIdx.x / u, rpb = blockDim.x / u;
    if (rpb == 0 || rl >= rpb) return;
    const double mu = MODE == 1 ? (double)mean[o] : 0.0;
    double acc = 0.0;                                        // 10^4..10^6 rows per column: float sums lose 1e-4 of the variance
    for (long long r = (long long)blockIdx.x * rpb + rl; r < rows; r += (long long)gridDim.x * rpb) {
        const double v = (double)z[r * u + o];
        acc += MODE == 1 ? (v - mu) * (v - mu) : v;
    }
    atomicAdd(out + o, acc);
}

// sums -> mean / invstd, running statistics (BatchNorm1d training: momentum m, unbiased running variance)
__global__ void k_tr_finish_stats(float *stat, const double *sums, int u, double n_rows, float eps, float momentum, float *run_mean,
                                  float *run_var, int pass)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= u) return;
    if (pass == 0) { stat[o] = (float)(sums[o] / n_rows); return; }                   // mean
    const float var = (float)(sums[u + o] / n_rows);                                  // biased: what normalises
    const float unb = n_rows > 1.0 ? (float)(sums[u + o] / (n_rows - 1.0)) : var;
    run_mean[o] = (1.0f - momentum) * run_mean[o] + momentum * stat[o];
    run_var[o] = (1.0f - momentum) * run_var[o] + momentum * unb;
    stat[u + o] = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, eps)));                  // invstd
}

// backward: dbeta, dgamma from their double sums
__global__ void k_tr_finish_grads(float *stat, const double *sums, int u)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= u) return;
    stat[2 * u + o] = (float)sums[2 * u + o];
    stat[3 * u + o] = (float)sums[3 * u + o];
}

// Y = relu(bn(Z)), per-voxel max + arg-max; writes the next layer's input [Y | x_max] when x_next != NULL
__global__ void __launch_bounds__(PTR_THREADS) k_tr_bn_relu_max(const float *__restrict__ z, const float *__restrict__ stat,
                                                                const float *__restrict__ gamma, const float *__restrict__ beta,
                                                                long long m, int t, int u, float *__restrict__ xmax,
                                                                int32_t *__restrict__ arg, float *__restrict__ x_next)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m * u) return;
    const long long v = e / u;
    const int o = (int)(e - v * u);
    const float mu = stat[o], is = stat[u + o], ga = gamma[o], be = beta[o];
    float best = -1.0f;
    int bi = 0;
    for (int q = 0; q < t; ++q) {
        const float y = fmaxf(__fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(z[(v * t + q) * u + o], mu), is), ga), be), 0.0f);
        if (x_next) x_next[(v * t + q) * (2 * u) + o] = y;
        if (y > best) { best = y; bi = q; }                  // first maximum, like torch.max
    }
    xmax[e] = best;
    arg[e] = bi;
    if (x_next)
        for (int q = 0; q < t; ++q) x_next[(v * t + q) * (2 * u) + u + o] = best;
}

// G[r, o] = (x_max gradient if r is the arg-max row) + direct gradient (dX_next's Y half), masked by Y > 0.
// Also accumulates dbeta = sum G and dgamma = sum G * xhat into stat[2u..], stat[3u..] (zeroed by the caller).
__global__ void __launch_bounds__(PTR_THREADS) k_tr_grad_y(const float *__restrict__ z, const float *__restrict__ stat, double *__restrict__ sums,
                                                           const float *__restrict__ gamma, const float *__restrict__ beta,
                                                           const float *__restrict__ d_xmax, const int32_t *__restrict__ arg,
                                                           const float *__restrict__ d_next, int next_stride,
                                                           long long m, int t, int u, float *__restrict__ g)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m * u) return;
    const long long v = e / u;
    const int o = (int)(e - v * u);
    const float mu = stat[o], is = stat[u + o], ga = gamma[o], be = beta[o];
    const float dm = d_xmax[e];
    const int am = arg[e];
    double sb = 0.0, sg = 0.0;
    for (int q = 0; q < t; ++q) {
        const long long r = v * t + q;
        const float xh = __fmul_rn(__fsub_rn(z[r * u + o], mu), is);
        const float y = __fadd_rn(__fmul_rn(xh, ga), be);
        float gy = (q == am ? dm : 0.0f) + (d_next ? d_next[r * next_stride + o] : 0.0f);
        gy = y > 0.0f ? gy : 0.0f;
        g[r * u + o] = gy;
        sb += (double)gy; sg += (double)gy * (double)xh;
    }
    atomicAdd(sums + 2 * u + o, sb);
    atomicAdd(sums + 3 * u + o, sg);
}

// dZ = gamma invstd (G - dbeta / N - xhat dgamma / N), in place over G
__global__ void __launch_bounds__(PTR_THREADS) k_tr_grad_z(const float *__restrict__ z, const float *__restrict__ stat,
                                                           const float *__restrict__ gamma, long long rows, int u, double n_rows,
                                                           float *__restrict__ g)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * u) return;
    const int o = (int)(e % u);
    const float mu = stat[o], is = stat[u + o];
    const float xh = (z[e] - mu) * is;
    const float inv_n = (float)(1.0 / n_rows);
    g[e] = gamma[o] * is * (g[e] - stat[2 * u + o] * inv_n - xh * stat[3 * u + o] * inv_n);
}

// dW[o, j] += sum_r dZ[r, o] X[r, j]: block-level partial sums over a slab of rows, one atomic per element
__global__ void __launch_bounds__(PTR_THREADS) k_tr_grad_w(const float *__restrict__ dz, const float *__restrict__ x, long long rows,
                                                           int k, int u, float *__restrict__ dw)
{
    const long long slab = (rows + gridDim.x - 1) / gridDim.x;
    const long long r0 = (long long)blockIdx.x * slab, r1 = min(rows, r0 + slab);
    for (int e = threadIdx.x; e < u * k; e += blockDim.x) {
        const int o = e / k, j = e - o * k;
        float acc = 0.0f;
        for (long long r = r0; r < r1; ++r) acc = __fmaf_rn(dz[r * u + o], x[r * k + j], acc);
        atomicAdd(dw + e, acc);
    }
}

// dX[r, j] = sum_o dZ[r, o] W[o, j] for layer l >= 1: the Y half (j < up) goes to d_prev[r, j]; the x_max half is
// summed over the voxel's T rows into d_xmax_prev[v, j - up]
__global__ void __launch_bounds__(PTR_THREADS) k_tr_grad_x(const float *__restrict__ dz, const float *__restrict__ w, long long m, int t,
                                                           int k, int u, float *__restrict__ d_prev, float *__restrict__ d_xmax_prev)
{
    extern __shared__ float s_w[];                           // [u][k]
    for (int e = threadIdx.x; e < u * k; e += blockDim.x) s_w[e] = w[e];
    __syncthreads();
    const int up = k / 2;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m * up) return;
    const long long v = e / up;
    const int j = (int)(e - v * up);
    float sum_max = 0.0f;
    for (int q = 0; q < t; ++q) {
        const float *dr = dz + (v * t + q) * u;
        float a = 0.0f, b = 0.0f;
        for (int o = 0; o < u; ++o) { a = __fmaf_rn(dr[o], s_w[o * k + j], a); b = __fmaf_rn(dr[o], s_w[o * k + up + j], b); }
        d_prev[(v * t + q) * up + j] = a;
        sum_max += b;
    }
    d_xmax_prev[e] = sum_max;
}

static size_t tr_up(size_t v) { return (v + 255) / 256 * 256; }

// workspace: per layer X, Z, stat, xmax, arg; plus two gradient scratch buffers of the widest [R, u]
static size_t tr_layout(int64_t m, int32_t t, const int *kk, const int *uu, int n_layers, char *base, PtrLayer *L, float **g0,
                        float **g1, float **dxm0, float **dxm1)
{
    const size_t R = (size_t)m * t;
    size_t o = 0;
    int umax = 0;
    for (int l = 0; l < n_layers; ++l) {
        if (L) { L[l].x = (float *)(base + o); } o = tr_up(o + R * kk[l] * 4);
        if (L) { L[l].z = (float *)(base + o); } o = tr_up(o + R * uu[l] * 4);
        if (L) { L[l].stat = (float *)(base + o); } o = tr_up(o + 4 * (size_t)uu[l] * 4);
        if (L) { L[l].sums = (double *)(base + o); } o = tr_up(o + 4 * (size_t)uu[l] * 8);
        if (L) { L[l].xmax = (float *)(base + o); } o = tr_up(o + (size_t)m * uu[l] * 4);
        if (L) { L[l].arg = (int32_t *)(base + o); } o = tr_up(o + (size_t)m * uu[l] * 4);
        umax = uu[l] > umax ? uu[l] : umax;
    }
    if (g0) *g0 = (float *)(base + o); o = tr_up(o + R * umax * 4);
    if (g1) *g1 = (float *)(base + o); o = tr_up(o + R * umax * 4);
    if (dxm0) *dxm0 = (float *)(base + o); o = tr_up(o + (size_t)m * umax * 4);
    if (dxm1) *dxm1 = (float *)(base + o); o = tr_up(o + (size_t)m * umax * 4);
    return o;
}

static int tr_shapes(const pv_pfn_layer *layers, int n_layers, int c, int with_distance, int *kk, int *uu)
{
    if (!layers || n_layers <= 0 || n_layers > PV_MAX_PFN_LAYERS) return PV_ERR_BAD_ARGUMENT;
    int width = c + 5 + (with_distance ? 1 : 0);
    for (int l = 0; l < n_layers; ++l) {
        if (layers[l].in_channels != width) return PV_ERR_BAD_ARGUMENT;
        if (layers[l].units <= 0 || layers[l].units > PTR_THREADS || PTR_THREADS % layers[l].units != 0 || width > 256) return PV_ERR_UNSUPPORTED;
        kk[l] = width; uu[l] = layers[l].units;
        width = 2 * layers[l].units;
    }
    return PV_OK;
}

extern "C" {

size_t pv_pfn_train_workspace_bytes(int64_t m, int32_t t, int32_t c, int32_t with_distance, const pv_pfn_layer *layers,
                                    int32_t n_layers)
{
    int kk[PV_MAX_PFN_LAYERS], uu[PV_MAX_PFN_LAYERS];
    if (m <= 0 || t <= 0 || tr_shapes(layers, n_layers, c, with_distance, kk, uu) != PV_OK) return 0;
    return tr_layout(m, t, kk, uu, n_layers, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int pv_pfn_train_forward(const float *voxels, const int32_t *num_points, const int32_t *coors, int64_t m, int32_t t,
                         int32_t c, int32_t with_distance, float vx, float vy, float x_off, float y_off,
                         const pv_pfn_layer *layers, int32_t n_layers, float eps, float momentum, void *workspace,
                         size_t workspace_bytes, float *out, pv_stream_t stream)
{
    int kk[PV_MAX_PFN_LAYERS], uu[PV_MAX_PFN_LAYERS];
    int rc = tr_shapes(layers, n_layers, c, with_distance, kk, uu);
    if (rc) return rc;
    if (m <= 0 || t <= 0 || c < 3 || !voxels || !num_points || !coors || !workspace || !out) return PV_ERR_BAD_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return PV_ERR_BAD_ARGUMENT;
    PtrLayer L[PV_MAX_PFN_LAYERS];
    if (tr_layout(m, t, kk, uu, n_layers, (char *)workspace, L, nullptr, nullptr, nullptr, nullptr) > workspace_bytes) return PV_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const long long R = (long long)m * t;
    const int sms = pv_sm_count();
    k_tr_decorate<<<(unsigned)((R + PTR_THREADS - 1) / PTR_THREADS), PTR_THREADS, 0, st>>>(
        voxels, num_points, coors, m, t, c, with_distance ? 1 : 0, vx, vy, x_off, y_off, L[0].x, kk[0]);
    for (int l = 0; l < n_layers; ++l) {
        const pv_pfn_layer &P = layers[l];
        if (!P.weight || !P.bn_mean || !P.bn_var || !P.bn_gamma || !P.bn_beta) return PV_ERR_BAD_ARGUMENT;
        const int k = kk[l], u = uu[l];
        const size_t smem = (size_t)u * (k + 1) * 4;
        if (smem > 48 * 1024 && cudaFuncSetAttribute(k_tr_linear, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return PV_ERR_CUDA;
        const unsigned grid = (unsigned)std::min<long long>((R * u + PTR_THREADS - 1) / PTR_THREADS, (long long)sms * 16);
        k_tr_linear<<<grid, PTR_THREADS, smem, st>>>(L[l].x, P.weight, R, k, u, L[l].z);
        if (cudaMemsetAsync(L[l].sums, 0, 4 * (size_t)u * 8, st) != cudaSuccess) return PV_ERR_CUDA;
        k_tr_colsum<0><<<grid, PTR_THREADS, 0, st>>>(L[l].z, R, u, nullptr, L[l].sums);
        k_tr_finish_stats<<<(u + 127) / 128, 128, 0, st>>>(L[l].stat, L[l].sums, u, (double)R, eps, momentum, nullptr, nullptr, 0);
        k_tr_colsum<1><<<grid, PTR_THREADS, 0, st>>>(L[l].z, R, u, L[l].stat, L[l].sums + u);
        k_tr_finish_stats<<<(u + 127) / 128, 128, 0, st>>>(L[l].stat, L[l].sums, u, (double)R, eps, momentum, const_cast<float *>(P.bn_mean),
                                                           const_cast<float *>(P.bn_var), 1);
        const bool last = l == n_layers - 1;
        k_tr_bn_relu_max<<<(unsigned)((m * u + PTR_THREADS - 1) / PTR_THREADS), PTR_THREADS, 0, st>>>(
            L[l].z, L[l].stat, P.bn_gamma, P.bn_beta, m, t, u, last ? out : L[l].xmax, L[l].arg, last ? nullptr : L[l + 1].x);
    }
    return pv_last_cuda_error();
}

int pv_pfn_train_backward(const float *d_out, int64_t m, int32_t t, int32_t c, int32_t with_distance,
                          const pv_pfn_layer *layers, int32_t n_layers, void *workspace, size_t workspace_bytes,
                          float *const *d_weight, float *const *d_gamma, float *const *d_beta, pv_stream_t stream)
{
    int kk[PV_MAX_PFN_LAYERS], uu[PV_MAX_PFN_LAYERS];
    int rc = tr_shapes(layers, n_layers, c, with_distance, kk, uu);
    if (rc) return rc;
    if (m <= 0 || t <= 0 || !d_out || !workspace || !d_weight || !d_gamma || !d_beta) return PV_ERR_BAD_ARGUMENT;
    PtrLayer L[PV_MAX_PFN_LAYERS];
    float *g0, *g1, *dxm0, *dxm1;
    if (tr_layout(m, t, kk, uu, n_layers, (char *)workspace, L, &g0, &g1, &dxm0, &dxm1) > workspace_bytes) return PV_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const long long R = (long long)m * t;
    const int sms = pv_sm_count();
    const float *d_xmax = d_out;          // gradient of the layer's x_max output
    const float *d_next = nullptr;        // direct gradient of the layer's Y output (from the next layer's input)
    float *g = g0, *other = g1, *dxm = dxm0, *dxm_other = dxm1;
    for (int l = n_layers - 1; l >= 0; --l) {
        const pv_pfn_layer &P = layers[l];
        const int k = kk[l], u = uu[l];
        if (!d_weight[l] || !d_gamma[l] || !d_beta[l]) return PV_ERR_BAD_ARGUMENT;
        if (cudaMemsetAsync(L[l].sums + 2 * u, 0, 2 * (size_t)u * 8, st) != cudaSuccess) return PV_ERR_CUDA;
        k_tr_grad_y<<<(unsigned)((m * u + PTR_THREADS - 1) / PTR_THREADS), PTR_THREADS, 0, st>>>(
            L[l].z, L[l].stat, L[l].sums, P.bn_gamma, P.bn_beta, d_xmax, L[l].arg, d_next, u, m, t, u, g);
        k_tr_finish_grads<<<(u + 127) / 128, 128, 0, st>>>(L[l].stat, L[l].sums, u);
        if (cudaMemcpyAsync(d_beta[l], L[l].stat + 2 * u, (size_t)u * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return PV_ERR_CUDA;
        if (cudaMemcpyAsync(d_gamma[l], L[l].stat + 3 * u, (size_t)u * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return PV_ERR_CUDA;
        k_tr_grad_z<<<(unsigned)((R * u + PTR_THREADS - 1) / PTR_THREADS), PTR_THREADS, 0, st>>>(L[l].z, L[l].stat, P.bn_gamma, R, u, (double)R, g);
        if (cudaMemsetAsync(d_weight[l], 0, (size_t)u * k * 4, st) != cudaSuccess) return PV_ERR_CUDA;
        k_tr_grad_w<<<(unsigned)std::min<long long>(R, (long long)sms * 8), PTR_THREADS, 0, st>>>(g, L[l].x, R, k, u, d_weight[l]);
        if (l > 0) {
            const size_t smem = (size_t)u * k * 4;
            if (smem > 48 * 1024 && cudaFuncSetAttribute(k_tr_grad_x, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                return PV_ERR_CUDA;
            const int up = k / 2;
            k_tr_grad_x<<<(unsigned)((m * up + PTR_THREADS - 1) / PTR_THREADS), PTR_THREADS, smem, st>>>(g, P.weight, m, t, k, u, other, dxm);
            // g (G, then dZ of this layer) is consumed by now and free for the next layer's G; `other` carries the
            // direct gradient down to the next layer's k_tr_grad_y, which has read it before its own k_tr_grad_x
            // overwrites it; the two x_max gradient buffers alternate
            d_next = other; d_xmax = dxm;
            float *tmp = dxm; dxm = dxm_other; dxm_other = tmp;
        }
    }
    return pv_last_cuda_error();
}

}  // extern "C"
