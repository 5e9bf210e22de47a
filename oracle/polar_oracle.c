/*
 * polar_oracle.c -- CPU restatement of PARTNER's point -> polar-grid front end.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under partner_b200/ may import, link or
 * execute this file.  Legitimate users: tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs, where it is the checker or
 * the timed CPU baseline -- never the product path.
 *
 * Parity status: PINNED.  tests/golden/make_golden.py imports the reference's
 * own numba / torch functions from /root/reference in the build container and
 * freezes their outputs under tests/golden/*.npz; tests/test_oracle_golden.py
 * checks every function below against those vectors (bit-exact for all
 * integer / gather outputs, 1e-5 for the float reductions).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference checkout).  Plain C99, scalar, single threaded, compiled with
 * -ffp-contract=off so that each float operation is rounded separately, as
 * numpy / numba / eager torch do.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* phi = arctan2(y, x)  (det3d/datasets/pipelines/utils.py:41)                */
/*                                                                           */
/* numpy's float32 arctan2 is a vendor SIMD routine (<= 3.4 ulp, differs from */
/* glibc and from CUDA), so it cannot be a bit-exact target.  The front end   */
/* therefore DEFINES phi by this fixed sequence of IEEE-754 binary32          */
/* operations (div, mul, fma, add; max error 1.66 ulp, measured <= 4 ulp from */
/* numpy on 2e7 points).  Only correctly rounded primitives are used, so the  */
/* CUDA kernel reproduces it bit for bit.                                     */
/* ------------------------------------------------------------------------- */
static const float PO_ATAN_C[9] = {
    -0x1.55553ep-2f, 0x1.9991fep-3f, -0x1.2421b4p-3f, 0x1.c099fap-4f, -0x1.583482p-4f,
    0x1.dac9b4p-5f,  -0x1.fed102p-6f, 0x1.65a5f8p-7f,  -0x1.d62f3cp-10f};

float po_atan2f(float y, float x)
{
    if (x != x || y != y) return NAN;
    float ax = fabsf(x), ay = fabsf(y);
    float mx = ax > ay ? ax : ay;
    float mn = ax > ay ? ay : ax;
    float a = mn / mx;
    if (mx == 0.0f) a = 0.0f;
    if (isinf(mn)) a = 1.0f;
    float s = a * a;
    float p = PO_ATAN_C[8];
    for (int i = 7; i >= 0; --i) p = fmaf(p, s, PO_ATAN_C[i]);
    float r = fmaf(a * s, p, a);
    if (ay > ax) r = (0x1.921fb6p+0f - r) + (-0x1.777a5cp-25f);  /* pi/2 = hi + lo */
    if (signbit(x)) r = (0x1.921fb6p+1f - r) + (-0x1.777a5cp-24f); /* pi   = hi + lo */
    return copysignf(r, y);
}

/* ------------------------------------------------------------------------- */
/* transform_points -- det3d/datasets/pipelines/utils.py:34-47                */
/* rho = sqrt(x**2 + y**2) : four separately rounded f32 ops (utils.py:40).   */
/* cylinder: [rho, phi, z, x, y, feat3..]   (utils.py:42-44)                  */
/* cuboid  : [x, y, z, feat3.., rho, phi]   (utils.py:45-47)                  */
/* ------------------------------------------------------------------------- */
void po_transform_points(const float *in, int64_t n, int c_in, int cylinder, float *out)
{
    const int c = c_in + 2;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = in + i * c_in;
        float *q = out + i * c;
        float xx = p[0] * p[0];
        float yy = p[1] * p[1];
        float rho = sqrtf(xx + yy);
        float phi = po_atan2f(p[1], p[0]);
        if (cylinder) {
            q[0] = rho; q[1] = phi; q[2] = p[2]; q[3] = p[0]; q[4] = p[1];
            for (int k = 3; k < c_in; ++k) q[k + 2] = p[k];
        } else {
            for (int k = 0; k < c_in; ++k) q[k] = p[k];
            q[c_in] = rho; q[c_in + 1] = phi;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* grid_size -- det3d/core/input/voxel_generator.py:10-11 and                 */
/* det3d/ops/point_cloud/point_cloud_ops.py:27-31: (hi - lo) / vs in f32,     */
/* np.round (half to even), xyz order.                                        */
/* ------------------------------------------------------------------------- */
void po_grid_size(const float *voxel_size, const float *range, int32_t *grid)
{
    for (int j = 0; j < 3; ++j) {
        float g = (range[3 + j] - range[j]) / voxel_size[j];
        grid[j] = (int32_t)rintf(g);
    }
}

/* ------------------------------------------------------------------------- */
/* points_to_voxel(reverse_index=True) -- point_cloud_ops.py:146-224 wrapping */
/* _points_to_voxel_reverse_kernel :7-72.                                     */
/*                                                                           */
/* Caller provides voxels[max_voxels*T*C], coors[max_voxels*3],               */
/* num_points[max_voxels]; they are cleared here exactly as the reference     */
/* allocates them zeroed (:185-190).  The dense coor_to_voxelidx map          */
/* (nz*ny*nx int32 = -1, :186) is allocated and filled per call, as the       */
/* reference does -- that cost is part of the reference path.                 */
/* pc_grid_ind [n,3] and density [nz,ny,nx] may be NULL (return_* = False).   */
/* Returns voxel_num, or -1 if the map allocation fails.                      */
/*                                                                           */
/* NaN coordinates are undefined behaviour in the reference (the int32 cast   */
/* of NaN indexes out of bounds); here, as in the CUDA build, a NaN bin is    */
/* treated as below range (dropped, grid index clamped to 0).                 */
/* ------------------------------------------------------------------------- */
int64_t po_points_to_voxel(const float *points, int64_t n, int c,
                           const float *voxel_size, const float *range,
                           int max_points, int max_voxels,
                           float *voxels, int32_t *coors, int32_t *num_points,
                           int32_t *pc_grid_ind, int32_t *density)
{
    int32_t grid[3];
    po_grid_size(voxel_size, range, grid);
    const int64_t nx = grid[0], ny = grid[1], nz = grid[2];
    const int64_t cells = nx * ny * nz;
    int32_t *map = (int32_t *)malloc((size_t)cells * sizeof(int32_t));
    if (!map) return -1;
    for (int64_t k = 0; k < cells; ++k) map[k] = -1;                    /* :186 */
    memset(num_points, 0, (size_t)max_voxels * sizeof(int32_t));        /* :185 */
    memset(voxels, 0, (size_t)max_voxels * max_points * c * sizeof(float)); /* :187 */
    memset(coors, 0, (size_t)max_voxels * 3 * sizeof(int32_t));         /* :190 */
    if (pc_grid_ind) memset(pc_grid_ind, 0, (size_t)n * 3 * sizeof(int32_t)); /* :36 */
    if (density) memset(density, 0, (size_t)cells * sizeof(int32_t));   /* :40 */

    int32_t coor[3] = {0, 0, 0};   /* persists across iterations, like :32 */
    int64_t voxel_num = 0;
    const int want_ind = pc_grid_ind != NULL;
    for (int64_t i = 0; i < n; ++i) {                                   /* :42 */
        int failed = 0;
        for (int j = 0; j < 3; ++j) {
            float cf = floorf((points[i * c + j] - range[j]) / voxel_size[j]); /* :45 */
            int32_t ci;
            if (cf != cf) { failed = 1; if (!want_ind) break; ci = 0; }
            else if (cf < 0.0f || cf >= (float)grid[j]) {               /* :46 */
                failed = 1;
                if (!want_ind) break;                                   /* :51 */
                ci = cf < 0.0f ? 0 : grid[j] - 1;                       /* :49 */
            } else ci = (int32_t)cf;
            coor[2 - j] = ci;                                           /* :52 */
        }
        if (want_ind) {                                                 /* :53-54 */
            pc_grid_ind[i * 3 + 0] = coor[0];
            pc_grid_ind[i * 3 + 1] = coor[1];
            pc_grid_ind[i * 3 + 2] = coor[2];
        }
        if (failed) continue;                                           /* :55-56 */
        int64_t cell = ((int64_t)coor[0] * ny + coor[1]) * nx + coor[2];
        int32_t vid = map[cell];                                        /* :57 */
        if (vid == -1) {
            vid = (int32_t)voxel_num;
            if (voxel_num >= max_voxels) continue;                      /* :60-61 */
            voxel_num += 1;
            map[cell] = vid;
            coors[vid * 3 + 0] = coor[0];
            coors[vid * 3 + 1] = coor[1];
            coors[vid * 3 + 2] = coor[2];
        }
        int32_t num = num_points[vid];
        if (num < max_points) {                                         /* :66-68 */
            memcpy(voxels + ((int64_t)vid * max_points + num) * c, points + i * c,
                   (size_t)c * sizeof(float));
            num_points[vid] = num + 1;
        }
        if (density) density[cell] += 1;                                /* :70-71 */
    }
    free(map);
    return voxel_num;
}

/* ------------------------------------------------------------------------- */
/* VoxelFeatureExtractorV3.forward -- det3d/models/readers/voxel_encoder.py   */
/* :15-22: sum over the T slots (zero padding included) / num_points.         */
/* ------------------------------------------------------------------------- */
void po_vfe_mean(const float *voxels, const int32_t *num_points, int64_t m, int t, int c,
                 float *out)
{
    for (int64_t v = 0; v < m; ++v) {
        float nf = (float)num_points[v];
        for (int k = 0; k < c; ++k) {
            float s = 0.0f;
            for (int j = 0; j < t; ++j) s += voxels[(v * t + j) * c + k];
            out[v * c + k] = s / nf;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* PillarFeatureNet.forward (eval) -- det3d/models/readers/pillar_encoder.py  */
/* :131-169 with PFNLayer.forward_static :49-61.                              */
/*                                                                           */
/* n_layers PFN layers; layer l has weight W_l [units_l, in_l] (nn.Linear,    */
/* no bias, :41), BatchNorm1d running stats + affine (eval: (x-mu)*invstd*g+b */
/* with invstd = 1/sqrt(var+eps)), ReLU, max over ALL T slots -- padded slots */
/* included (:55), which is the reference's behaviour.  Non-last layers       */
/* output cat([x, repeat(x_max)]) (:59-61).  units[l] is the Linear's output  */
/* width (already halved for non-last layers, :34-36).                        */
/* coors is [m,4] (b,z,y,x).  out is [m, units[n_layers-1]].                  */
/* vx, vy, x_off, y_off are the f32 images of the Python doubles (:123-126).  */
/* ------------------------------------------------------------------------- */
void po_pfn_forward(const float *voxels, const int32_t *num_points, const int32_t *coors,
                    int64_t m, int t, int c, int with_distance,
                    float vx, float vy, float x_off, float y_off,
                    int n_layers, const int32_t *units,
                    const float *const *weight, const float *const *bn_mean,
                    const float *const *bn_var, const float *const *bn_gamma,
                    const float *const *bn_beta, float eps, float *out)
{
    const int c0 = c + 5 + (with_distance ? 1 : 0);
    int max_w = c0;
    for (int l = 0; l < n_layers; ++l) if (2 * units[l] > max_w) max_w = 2 * units[l];
    float *cur = (float *)malloc((size_t)t * max_w * sizeof(float));
    float *nxt = (float *)malloc((size_t)t * max_w * sizeof(float));
    float *xmax = (float *)malloc((size_t)max_w * sizeof(float));
    for (int64_t v = 0; v < m; ++v) {
        const float *f = voxels + v * t * c;
        const float nf = (float)num_points[v];
        float mean[3];
        for (int k = 0; k < 3; ++k) {                                   /* :137-139 */
            float s = 0.0f;
            for (int j = 0; j < t; ++j) s += f[j * c + k];
            mean[k] = s / nf;
        }
        const float cx = (float)coors[v * 4 + 3] * vx + x_off;          /* :146-147 */
        const float cy = (float)coors[v * 4 + 2] * vy + y_off;          /* :149-150 */
        for (int j = 0; j < t; ++j) {
            float *row = cur + j * c0;
            const float *p = f + j * c;
            for (int k = 0; k < c; ++k) row[k] = p[k];
            for (int k = 0; k < 3; ++k) row[c + k] = p[k] - mean[k];    /* :140 */
            row[c + 3] = p[0] - cx;
            row[c + 4] = p[1] - cy;
            if (with_distance)                                          /* :155 */
                row[c + 5] = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
            const float mask = j < num_points[v] ? 1.0f : 0.0f;         /* :161-164 */
            for (int k = 0; k < c0; ++k) row[k] *= mask;
        }
        int in_w = c0;
        for (int l = 0; l < n_layers; ++l) {
            const int u = units[l];
            const int last = l == n_layers - 1;
            const int out_w = last ? u : 2 * u;
            for (int o = 0; o < u; ++o) xmax[o] = -INFINITY;
            for (int j = 0; j < t; ++j) {
                for (int o = 0; o < u; ++o) {
                    float acc = 0.0f;
                    const float *w = weight[l] + (size_t)o * in_w;
                    for (int k = 0; k < in_w; ++k) acc += cur[j * in_w + k] * w[k]; /* :50 */
                    const float invstd = 1.0f / sqrtf(bn_var[l][o] + eps);
                    float y = (acc - bn_mean[l][o]) * invstd * bn_gamma[l][o] + bn_beta[l][o]; /* :52 */
                    y = y > 0.0f ? y : 0.0f;                            /* :54 */
                    if (!last) nxt[j * out_w + o] = y;
                    if (y > xmax[o]) xmax[o] = y;                       /* :55 */
                }
            }
            if (last) {
                for (int o = 0; o < u; ++o) out[v * u + o] = xmax[o];   /* :56-57 */
            } else {
                for (int j = 0; j < t; ++j)
                    for (int o = 0; o < u; ++o) nxt[j * out_w + u + o] = xmax[o]; /* :59-60 */
                float *tmp = cur; cur = nxt; nxt = tmp;
                in_w = out_w;
            }
        }
    }
    free(cur); free(nxt); free(xmax);
}

/* ------------------------------------------------------------------------- */
/* PointPillarsScatter.forward -- pillar_encoder.py:189-225.                  */
/* canvas [batch, c, ny, nx] zeroed; canvas[b, :, y*nx + x] = feats[v, :].    */
/* bev_index (optional, [m] int64) receives the BEV index map y*nx + x (:211).*/
/* Later duplicates overwrite earlier ones, as index_put_ on CPU does.        */
/* ------------------------------------------------------------------------- */
void po_scatter(const float *feats, const int32_t *coors, int64_t m, int c, int batch,
                int ny, int nx, float *canvas, int64_t *bev_index)
{
    const int64_t plane = (int64_t)ny * nx;
    memset(canvas, 0, (size_t)batch * c * plane * sizeof(float));
    for (int64_t v = 0; v < m; ++v) {
        const int b = coors[v * 4 + 0];
        const int64_t idx = (int64_t)coors[v * 4 + 2] * nx + coors[v * 4 + 3];
        if (bev_index) bev_index[v] = idx;
        if (b < 0 || b >= batch) continue;                              /* :207 mask */
        for (int k = 0; k < c; ++k) canvas[((int64_t)b * c + k) * plane + idx] = feats[v * c + k];
    }
}

/* ------------------------------------------------------------------------- */
/* Dynamic voxelization ("next" row f1 of SURVEY.md section 8)                 */
/* ------------------------------------------------------------------------- */

/* voxelize_dynamic -- det3d/datasets/pipelines/voxelization.py:169-172:
 *   pc_grid_ind = floor(clip((points[:, :3] - pc_range[:3]) / voxel_size, 0, grid_size - 1))[:, ::-1]
 * float32 subtract and divide (numpy), clamp, floor; EVERY point gets a cell (out-of-range points
 * are clamped into the border cells, nothing is dropped).  NaN is undefined behaviour in the
 * reference (astype(int) of NaN); here it maps to cell 0 like any value below the range.
 * out: int32 [n, 3] in (z, y, x) order. */
void po_dynamic_grid_ind(const float *points, int64_t n, int c, const float *voxel_size,
                         const float *range, int32_t *out)
{
    int32_t grid[3];
    po_grid_size(voxel_size, range, grid);
    for (int64_t i = 0; i < n; ++i) {
        for (int j = 0; j < 3; ++j) {
            float q = (points[i * c + j] - range[j]) / voxel_size[j];
            float hi = (float)(grid[j] - 1);
            if (!(q >= 0.0f)) q = 0.0f;      /* below range or NaN */
            if (q > hi) q = hi;
            out[i * 3 + (2 - j)] = (int32_t)floorf(q);
        }
    }
}

typedef struct { int32_t k[4]; int64_t idx; } po_key4;
static int po_cmp_key4(const void *a, const void *b)
{
    const po_key4 *x = (const po_key4 *)a, *y = (const po_key4 *)b;
    for (int j = 0; j < 4; ++j)
        if (x->k[j] != y->k[j]) return x->k[j] < y->k[j] ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

/* DynamicVoxelEncoderV1.forward -- det3d/models/readers/voxel_encoder.py:38-44:
 *   unq, unq_inv, unq_cnt = torch.unique(grid_ind, return_inverse=True, return_counts=True, dim=0)
 *   features = torch_scatter.scatter_mean(features, unq_inv, dim=0)
 * torch.unique(dim=0) sorts the rows lexicographically, so voxels come out in (b, z, y, x) order;
 * scatter_mean (third-party torch_scatter, not vendored by the reference; documented semantics):
 * per-voxel fp32 sum in point order divided by the count.  No max_points / max_voxels caps.
 * grid_ind int32 [n, 4] (b, z, y, x); returns M; unq [M, 4], inv [n], cnt [M], mean [M, c]. */
int64_t po_dynamic_mean(const int32_t *grid_ind, const float *feats, int64_t n, int c,
                        int32_t *unq, int64_t *inv, int64_t *cnt, float *mean)
{
    if (n == 0) return 0;
    po_key4 *keys = (po_key4 *)malloc((size_t)n * sizeof(po_key4));
    if (!keys) return -1;
    for (int64_t i = 0; i < n; ++i) {
        memcpy(keys[i].k, grid_ind + i * 4, 4 * sizeof(int32_t));
        keys[i].idx = i;
    }
    qsort(keys, (size_t)n, sizeof(po_key4), po_cmp_key4);
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (i == 0 || memcmp(keys[i].k, keys[i - 1].k, 4 * sizeof(int32_t)) != 0) {
            memcpy(unq + m * 4, keys[i].k, 4 * sizeof(int32_t));
            cnt[m] = 0;
            ++m;
        }
        inv[keys[i].idx] = m - 1;
        cnt[m - 1] += 1;
    }
    free(keys);
    memset(mean, 0, (size_t)m * c * sizeof(float));
    for (int64_t i = 0; i < n; ++i)                 /* scatter_sum in point order */
        for (int k = 0; k < c; ++k) mean[inv[i] * c + k] += feats[i * c + k];
    for (int64_t v = 0; v < m; ++v)
        for (int k = 0; k < c; ++k) mean[v * c + k] /= (float)cnt[v];
    return m;
}

/* ------------------------------------------------------------------------- */
/* DynamicPFNet.forward -- det3d/models/readers/pillar_encoder.py:338-411,    */
/* PFNLayer.forward_dynamic :63-71, get_cluster :228-238, polar2cart /         */
/* cart2polar :240-260.                                                        */
/*                                                                           */
/* points [n, c]; unq_inv [n] voxel row of every point; unq [m, 4] (b,z,y,x). */
/* flags: bit0 xyz_cluster, bit1 raz_cluster, bit2 xy_center, bit3 ra_center. */
/* cylinder != 0: voxel_shape == 'cylinder' (xyz = cols [3,4,2], ra = cols     */
/* [0,1]); else 'cuboid' (xyz = cols [0,1,2], ra = the last two cols) -- the  */
/* reference's configs leave the reader at its 'cuboid' default (:269).       */
/* Decoration order (:347-391): points, xyz - mean(xyz), xyz[:2] - centre,    */
/* ra(z) - mean, ra - centre.  Layers: Linear (no bias), ReLU -- NO norm in    */
/* the dynamic forward (:64-65) -- scatter_max over the voxel; non-last        */
/* layers output cat([x, x_max[unq_inv]]).  scatter_mean / scatter_max are     */
/* torch_scatter's (third party, not vendored): mean = fp32 sum in point      */
/* order / count, max = plain maximum.  vx, vy, x_off, y_off: f32 images of   */
/* the Python doubles (:331-334).  out [m, units[n_layers-1]].                */
/* ------------------------------------------------------------------------- */
int po_dynamic_pfn(const float *points, const int64_t *unq_inv, const int32_t *unq, int64_t n,
                   int64_t m, int c, int cylinder, int flags, float vx, float vy, float x_off,
                   float y_off, int n_layers, const int32_t *units, const float *const *weight,
                   float *out)
{
    const int xyz_cluster = flags & 1, raz_cluster = flags & 2, xy_center = flags & 4, ra_center = flags & 8;
    const int xi[3] = {cylinder ? 3 : 0, cylinder ? 4 : 1, 2};
    const int ri[2] = {cylinder ? 0 : c - 2, cylinder ? 1 : c - 1};
    int c0 = c + (xyz_cluster ? 3 : 0) + (xy_center ? 2 : 0) + (raz_cluster ? (xyz_cluster ? 2 : 3) : 0) + (ra_center ? 2 : 0);
    /* per-voxel means of the columns get_cluster needs (scatter_mean: sum in point order / count) */
    float *mean = (float *)calloc((size_t)m * 5, sizeof(float));    /* x, y, z, r, a */
    int64_t *cnt = (int64_t *)calloc((size_t)m, sizeof(int64_t));
    if (!mean || !cnt) return -1;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = points + i * c;
        float *mv = mean + unq_inv[i] * 5;
        for (int k = 0; k < 3; ++k) mv[k] += p[xi[k]];
        for (int k = 0; k < 2; ++k) mv[3 + k] += p[ri[k]];
        cnt[unq_inv[i]] += 1;
    }
    for (int64_t v = 0; v < m; ++v)
        for (int k = 0; k < 5; ++k) mean[v * 5 + k] /= (float)(cnt[v] > 0 ? cnt[v] : 1);
    int max_w = c0;
    for (int l = 0; l < n_layers; ++l) if (2 * units[l] > max_w) max_w = 2 * units[l];
    float *cur = (float *)malloc((size_t)n * max_w * sizeof(float));
    float *nxt = (float *)malloc((size_t)n * max_w * sizeof(float));
    float *xmax = (float *)malloc((size_t)m * max_w * sizeof(float));
    if (!cur || !nxt || !xmax) return -1;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = points + i * c;
        const int64_t v = unq_inv[i];
        const float *mv = mean + v * 5;
        float *row = cur + i * c0;
        int o = 0;
        for (int k = 0; k < c; ++k) row[o++] = p[k];
        const float center1 = (float)unq[v * 4 + 3] * vx + x_off;      /* :350 */
        const float center2 = (float)unq[v * 4 + 2] * vy + y_off;      /* :351 */
        if (xyz_cluster) for (int k = 0; k < 3; ++k) row[o++] = p[xi[k]] - mv[k];
        if (xy_center) {
            float xc = center1, yc = center2;
            if (cylinder) { xc = center1 * cosf(center2); yc = center1 * sinf(center2); }   /* polar2cart */
            row[o++] = p[xi[0]] - xc;
            row[o++] = p[xi[1]] - yc;
        }
        if (raz_cluster) {
            row[o++] = p[ri[0]] - mv[3];
            row[o++] = p[ri[1]] - mv[4];
            if (!xyz_cluster) row[o++] = p[2] - mv[2];
        }
        if (ra_center) {
            float rc = center1, ac = center2;
            if (!cylinder) { rc = sqrtf(center1 * center1 + center2 * center2); ac = atan2f(center2, center1); }  /* cart2polar */
            row[o++] = p[ri[0]] - rc;
            row[o++] = p[ri[1]] - ac;
        }
    }
    int in_w = c0;
    for (int l = 0; l < n_layers; ++l) {
        const int u = units[l], last = l == n_layers - 1, out_w = last ? u : 2 * u;
        for (int64_t k = 0; k < m * u; ++k) xmax[k] = -INFINITY;
        for (int64_t i = 0; i < n; ++i) {
            for (int o = 0; o < u; ++o) {
                float acc = 0.0f;
                const float *w = weight[l] + (size_t)o * in_w;
                for (int k = 0; k < in_w; ++k) acc += cur[i * in_w + k] * w[k];
                const float y = acc > 0.0f ? acc : 0.0f;
                if (!last) nxt[i * out_w + o] = y;
                float *xm = xmax + unq_inv[i] * u + o;
                if (y > *xm) *xm = y;
            }
        }
        if (last) memcpy(out, xmax, (size_t)m * u * sizeof(float));
        else {
            for (int64_t i = 0; i < n; ++i)
                for (int o = 0; o < u; ++o) nxt[i * out_w + u + o] = xmax[unq_inv[i] * u + o];
            float *tmp = cur; cur = nxt; nxt = tmp;
            in_w = out_w;
        }
    }
    free(mean); free(cnt); free(cur); free(nxt); free(xmax);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Voxelization.voxelize_streaming_polar -- the per-sector point selection,   */
/* azimuth shift and grid index of det3d/datasets/pipelines/voxelization.py   */
/* :305-371 (evaluation path; the training-time ground-truth rotation is      */
/* annotation work, not part of the point path).                             */
/*   interval = (max_az - min_az) / nsectors          (float32, numpy >= 2)   */
/*   sector i keeps phi in [min_az + i*interval, min_az + (i+1)*interval),    */
/*   the first sector is open below, the last one open above (:352-357);      */
/*   phi -= (lo_i - min_az); x = rho*cos(phi); y = rho*sin(phi)   (:360-362)   */
/*   grid_ind = floor(clip((p[:3] - lo) / vs, 0, cur_grid - 1))[::-1] with     */
/*   cur_grid[1] = grid[1] // nsectors                             (:366-368)  */
/* points [n, c] are cylinder rows (rho, phi, z, x, y, ...).  Outputs are      */
/* sector-major, original order inside a sector: out_points [n, c],           */
/* out_gi [n, 3] (z, y, x), out_index [n] original row, counts [nsectors].    */
/* Returns the number of rows written (NaN azimuths belong to no sector).     */
/* ------------------------------------------------------------------------- */
int64_t po_stream_polar(const float *points, int64_t n, int c, const float *voxel_size,
                        const float *range, int nsectors, float *out_points, int32_t *out_gi,
                        int32_t *out_index, int64_t *counts)
{
    int32_t grid[3];
    po_grid_size(voxel_size, range, grid);
    const float min_az = range[1], max_az = range[4];
    const float interval = (max_az - min_az) / (float)nsectors;
    int32_t cur_grid[3] = {grid[0], grid[1] / nsectors, grid[2]};
    int64_t w = 0;
    for (int s = 0; s < nsectors; ++s) {
        const float lo = min_az + (float)s * interval;
        const float hi = min_az + (float)(s + 1) * interval;
        const float shift = lo - min_az;
        counts[s] = 0;
        for (int64_t i = 0; i < n; ++i) {
            const float phi = points[i * c + 1];
            int keep;
            if (s == 0) keep = phi < hi;                       /* :352-353 (also when nsectors == 1) */
            else if (s == nsectors - 1) keep = phi >= lo;      /* :354-355 */
            else keep = phi >= lo && phi < hi;                 /* :356-357 */
            if (!keep) continue;
            float *o = out_points + w * c;
            memcpy(o, points + i * c, (size_t)c * sizeof(float));
            o[1] = phi - shift;                                /* :360 */
            o[3] = o[0] * cosf(o[1]);                          /* :361 */
            o[4] = o[0] * sinf(o[1]);                          /* :362 */
            for (int j = 0; j < 3; ++j) {
                float q = (o[j] - range[j]) / voxel_size[j];
                const float top = (float)(cur_grid[j] - 1);
                if (!(q >= 0.0f)) q = 0.0f;
                if (q > top) q = top;
                out_gi[w * 3 + (2 - j)] = (int32_t)floorf(q);
            }
            out_index[w] = (int32_t)i;
            counts[s] += 1;
            ++w;
        }
    }
    return w;
}

/* ------------------------------------------------------------------------- */
/* Voxelization.get_grid_ind, train branch (det3d/datasets/pipelines/        */
/* voxelization.py:40-60) + AssignLabel.assign_voxel_labels (det3d/datasets/  */
/* pipelines/preprocess.py:170-191).                                          */
/*   valid = pc_label >= 0                                           (:44)     */
/*   rows (z, y, x, label) of the valid points, lexsorted with x as the       */
/*   primary key, then y, then z (np.lexsort((g0, g1, g2)), :47)               */
/*   sequential scan with a 256-entry uint16 counter per run of equal cells;   */
/*   voxel_labels[z, y, x] = argmax(counter), first maximum   (preprocess.py   */
/*   :178-191); cells without a valid point stay 0            (:50)            */
/* pc_grid_ind [n, 3] (z, y, x), pc_label [n]; voxel_labels int64 [nz, ny, nx] */
/* (zeroed here), valid_grid_ind [n, 3] receives the valid rows in their       */
/* original order (:45, :58).  Returns the number of valid rows, -1 on         */
/* allocation failure.                                                        */
/* ------------------------------------------------------------------------- */
typedef struct { int32_t z, y, x, l; } po_pair;

static int po_pair_cmp(const void *a, const void *b)
{
    const po_pair *p = (const po_pair *)a, *q = (const po_pair *)b;
    if (p->x != q->x) return p->x < q->x ? -1 : 1;
    if (p->y != q->y) return p->y < q->y ? -1 : 1;
    if (p->z != q->z) return p->z < q->z ? -1 : 1;
    return 0;       /* rows of one cell: their order does not change the counter */
}

static int64_t po_argmax_u16(const uint16_t *c)
{
    int best = 0;
    for (int k = 1; k < 256; ++k)
        if (c[k] > c[best]) best = k;
    return best;
}

int64_t po_seg_voxel_labels(const int32_t *pc_grid_ind, const int32_t *pc_label, int64_t n, int nz,
                            int ny, int nx, int64_t *voxel_labels, int32_t *valid_grid_ind)
{
    memset(voxel_labels, 0, (size_t)nz * ny * nx * sizeof(int64_t));
    po_pair *pairs = (po_pair *)malloc((size_t)(n > 0 ? n : 1) * sizeof(po_pair));
    if (!pairs) return -1;
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (pc_label[i] < 0) continue;
        pairs[m].z = pc_grid_ind[i * 3]; pairs[m].y = pc_grid_ind[i * 3 + 1]; pairs[m].x = pc_grid_ind[i * 3 + 2];
        pairs[m].l = pc_label[i];
        memcpy(valid_grid_ind + m * 3, pc_grid_ind + i * 3, 3 * sizeof(int32_t));
        ++m;
    }
    if (m > 0) {
        qsort(pairs, (size_t)m, sizeof(po_pair), po_pair_cmp);
        uint16_t counter[256];
        memset(counter, 0, sizeof counter);
        counter[pairs[0].l] = 1;                                         /* :179 */
        po_pair cur = pairs[0];
        for (int64_t i = 1; i < m; ++i) {
            if (pairs[i].z != cur.z || pairs[i].y != cur.y || pairs[i].x != cur.x) {       /* :185 */
                voxel_labels[((int64_t)cur.z * ny + cur.y) * nx + cur.x] = po_argmax_u16(counter);
                memset(counter, 0, sizeof counter);
                cur = pairs[i];
            }
            counter[pairs[i].l] = (uint16_t)(counter[pairs[i].l] + 1);    /* uint16: wraps at 65536 */
        }
        voxel_labels[((int64_t)cur.z * ny + cur.y) * nx + cur.x] = po_argmax_u16(counter);
    }
    free(pairs);
    return m;
}
