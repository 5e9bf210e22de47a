/*
 * polar_voxel_b200.h -- C ABI of the B200 (sm_100a) polar front end.
 *
 * The reference (fudan-zvg/PARTNER, det3d lineage) has no FFI on this path: it
 * calls numba-compiled Python and eager torch ops directly.  This header is the
 * boundary a det3d maintainer binds instead (ctypes stub in INTEGRATION.md), in
 * the style the reference uses for its other native ops: launchers that take
 * raw device pointers + sizes (det3d/ops/iou3d_nms/src/iou3d_nms.cpp:47-66).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless noted HOST;
 *   - the library never allocates, frees or synchronises; every launch goes to
 *     the caller's stream, so every entry point is CUDA-graph capturable;
 *   - one opaque caller-provided workspace (size from pv_workspace_bytes).  It is
 *     SELF-CLEANING: pv_workspace_init() prepares it once per (config, capacities);
 *     every call restores whatever it touched, so steady-state calls issue no
 *     memset.  Re-run pv_workspace_init after an error or when config, batch or a
 *     capacity changes;
 *   - return value: PV_OK (0) or a negative PV_ERR_* code.  Conditions only
 *     detectable on the device (hash table overflow) set a status word that
 *     pv_read_status() copies back.  Nothing here calls exit().
 *   - all floating point is IEEE binary32 without contraction, so integer
 *     outputs are bit-identical to the reference's numba loop on the same input.
 *
 * Point rows are float32 [n, c_in].  With is_cartesian != 0 rows are
 * (x, y, z, feat...) and the kernels apply the cylinder transform of
 * det3d/datasets/pipelines/utils.py:34-44 on the fly, producing
 * (rho, phi, z, x, y, feat...) with C = c_in + 2 channels; otherwise rows are
 * already polar (what VoxelGenerator.generate receives) and C = c_in.
 */
#ifndef POLAR_VOXEL_B200_H_
#define POLAR_VOXEL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *pv_stream_t; /* == cudaStream_t */

#define PV_OK 0
#define PV_ERR_BAD_CONFIG (-1)    /* non-positive grid / T / V, grid too large        */
#define PV_ERR_BAD_ARGUMENT (-2)  /* null pointer, channel count out of range, ...    */
#define PV_ERR_WORKSPACE (-3)     /* workspace smaller than pv_workspace_bytes says    */
#define PV_ERR_CUDA (-4)          /* a CUDA runtime call failed (cudaGetLastError)     */
#define PV_ERR_TABLE_FULL (-5)    /* device status: a frame had more points than frame_capacity */
#define PV_ERR_UNSUPPORTED (-6)   /* layer shape outside what the kernels cover        */
#define PV_ERR_INTERNAL (-7)      /* device status: a pipeline wait of the tensor-core PFN kernel starved (watchdog) */

#define PV_MAX_CHANNELS 16        /* C (after the transform) <= 16                     */
#define PV_MAX_PFN_LAYERS 4

/* Replaces the state of VoxelGenerator (det3d/core/input/voxel_generator.py:6-17):
 * voxel_size / point_cloud_range already rounded to f32, grid = round((hi-lo)/vs). */
typedef struct pv_config {
    float lo[3];        /* range lower bound, (rho, phi, z) order             */
    float vs[3];        /* voxel size                                         */
    int32_t grid[3];    /* nx, ny, nz                                         */
    int32_t max_points; /* T = max_points_in_voxel                            */
    int32_t max_voxels; /* V = max_voxels per frame                           */
    int32_t pipeline;   /* which of the two pipelines serves calls that do not ask for the padded
                           voxels tensor: 0 auto (list-free on direct-map grids of <= 2^20 cells per
                           frame, list-based on hash-map grids), 1 list-based, 2 list-free.  Part of
                           the configuration, not process state: 1 / 2 exist for measurements and tests */
} pv_config;

/* One PFNLayer (det3d/models/readers/pillar_encoder.py:19-61) in eval mode. */
typedef struct pv_pfn_layer {
    const float *weight;   /* linear.weight [units, in_channels], no bias     */
    const float *bn_mean;  /* norm.running_mean [units]                       */
    const float *bn_var;   /* norm.running_var  [units]                       */
    const float *bn_gamma; /* norm.weight       [units]                       */
    const float *bn_beta;  /* norm.bias         [units]                       */
    int32_t in_channels;
    int32_t units;         /* Linear output width (already halved if not last) */
} pv_pfn_layer;

int pv_version(void);
const char *pv_error_string(int code);

/* Which pipeline (1 lists, 2 list-free) pv_forward_mean_canvas / pv_profile_mean_canvas run for
 * this configuration (cfg->pipeline resolved); names the stages pv_profile_mean_canvas reports. */
int pv_profile_pipeline(const pv_config *cfg);

/* Bytes of workspace for a batch of `batch` frames holding at most
 * `max_points_total` points, no frame larger than `frame_capacity` points, voxelized with at
 * most `channels` feature channels C (c_in + 2 for Cartesian input, c_in otherwise). */
size_t pv_workspace_bytes(const pv_config *cfg, int64_t max_points_total, int32_t batch,
                          int64_t frame_capacity, int32_t channels);

/* One-time preparation of a workspace for (cfg, capacities); see "SELF-CLEANING" above. */
int pv_workspace_init(const pv_config *cfg, int64_t max_points_total, int32_t batch,
                      int64_t frame_capacity, int32_t channels, void *workspace,
                      size_t workspace_bytes, pv_stream_t stream);

/* transform_points (det3d/datasets/pipelines/utils.py:34-47).  out is [n, c_in+2]. */
int pv_transform_points(const float *in, int64_t n, int32_t c_in, int32_t cylinder, float *out,
                        pv_stream_t stream);

/*
 * Hard voxelization of a batch of frames: replaces points_to_voxel
 * (det3d/ops/point_cloud/point_cloud_ops.py:146-224, reverse_index=True) for
 * every frame plus the batch assembly of collate_kitti
 * (det3d/torchie/parallel/collate.py:157-164: batch index prepended, frames
 * concatenated).  Frame f owns points [frame_offsets[f], frame_offsets[f+1]).
 *
 * Outputs (rows are in frame order, then first-occurrence order; SM = sum of
 * per-frame voxel counts <= min(batch * V, n_total)):
 *   coors        int32 [SM, 4]  (b, z, y, x)                     required
 *   num_points   int32 [SM]                                       required
 *   voxel_counts int32 [batch]   M per frame                      required
 *   voxels       f32   [SM, T, C] zero padded                     or NULL
 *   mean_feats   f32   [SM, C]   mean over kept points (VFE V3)   or NULL
 *   pc_grid_ind  int32 [n_total, 3] clamped (z, y, x)             or NULL
 *   density      int32 [batch, nz, ny, nx] un-capped counts of kept voxels, or NULL
 * Output buffers must have capacity for min(batch * V, n_total) rows.
 * max_points_total / frame_capacity are the CAPACITIES the workspace was initialised with
 * (n_total <= max_points_total, every frame <= frame_capacity points).
 *
 * Two pipelines sit behind this entry point.  With voxels == NULL the call runs LIST-FREE
 * (fused.cu): per-cell accumulator rows updated with 16-byte fp32 reductions, no per-voxel
 * point lists; integer outputs are bit-exact, mean_feats is the sum of the same addends in an
 * unspecified order (relative error ~1e-7, not run-to-run bit-reproducible).  With voxels != NULL
 * the list-based pipeline (voxelize.cu) runs; its sums are in point-index order and reproducible.
 */
int pv_voxelize(const pv_config *cfg, const float *points, const int32_t *frame_offsets,
                int32_t batch, int64_t n_total, int32_t c_in, int32_t is_cartesian,
                int64_t max_points_total, int64_t frame_capacity, void *workspace,
                size_t workspace_bytes, int32_t *coors, int32_t *num_points, int32_t *voxel_counts,
                float *voxels, float *mean_feats, int32_t *pc_grid_ind, int32_t *density,
                pv_stream_t stream);

/* Fused front end for pillar grids (nz == 1): pv_voxelize (mean_feats) followed by the
 * scatter of PointPillarsScatter (pillar_encoder.py:189-225) into canvas
 * f32 [batch, C, ny, nx]; every canvas element is written exactly once. */
int pv_forward_mean_canvas(const pv_config *cfg, const float *points, const int32_t *frame_offsets,
                           int32_t batch, int64_t n_total, int32_t c_in, int32_t is_cartesian,
                           int64_t max_points_total, int64_t frame_capacity, void *workspace,
                           size_t workspace_bytes, int32_t *coors, int32_t *num_points,
                           int32_t *voxel_counts, float *mean_feats, float *canvas,
                           pv_stream_t stream);

/* Fused PILLAR-FEATURE-NET front end (BASELINE config 3): replaces, for a batch of frames,
 *   VoxelGenerator.generate + collate (as pv_voxelize)  ->  PillarFeatureNet.forward in eval mode
 *   (det3d/models/readers/pillar_encoder.py:131-169, PFNLayer.forward_static :49-61)  ->
 *   PointPillarsScatter.forward (:189-225),
 * i.e. PointPillars.extract_feat_static (det3d/models/detectors/point_pillars.py:28-35) on raw points.
 * The padded voxels tensor [M, T, C] is never materialised: the PFN kernel gathers the kept points of
 * every voxel through the voxelizer's point lists (fused cylinder transform for Cartesian input), runs
 * layer 0 in fp32 FMAs and layer 1 on the tensor cores (tcgen05.mma kind::tf32, 3xTF32 split, fp32
 * accumulators in TMEM), BatchNorm (eval, ATen order) + ReLU + the per-voxel maximum over all T slots
 * (padded-slot quirk included) in the epilogue.
 *   layers (HOST array), vx, vy, x_off, y_off, eps as for pv_pfn_forward; two layers, 32 units in the
 *   first, 32 | units of the last <= 128 (every PillarFeatureNet the reference's configs build),
 *   C + 5 (+1 with_distance) <= 16, max_points <= 32; PV_ERR_UNSUPPORTED otherwise.
 * Outputs: coors [SM, 4], num_points [SM], voxel_counts [batch] as pv_voxelize; pfn_feats f32 [SM, U]
 * (capacity min(batch * V, n_total) rows); canvas f32 [batch, U, ny, nx] or NULL (pillar grids).
 * aux_workspace: pv_pfn_canvas_workspace_bytes(batch, ny, nx, max_points_total, cfg->max_voxels) bytes,
 * 256-byte aligned (chunk queues, watchdog words, BEV index map, and the decorated rows the gather
 * pre-pass hands to the tensor-core kernel: 64 B per kept point + 64 B per voxel); workspace as for
 * pv_voxelize. */
size_t pv_pfn_canvas_workspace_bytes(int32_t batch, int32_t ny, int32_t nx, int64_t max_points_total,
                                     int32_t max_voxels);
int pv_forward_pfn_canvas(const pv_config *cfg, const float *points, const int32_t *frame_offsets,
                          int32_t batch, int64_t n_total, int32_t c_in, int32_t is_cartesian,
                          int64_t max_points_total, int64_t frame_capacity, void *workspace,
                          size_t workspace_bytes, void *aux_workspace, size_t aux_bytes,
                          const pv_pfn_layer *layers, int32_t n_layers, int32_t with_distance, float vx, float vy,
                          float x_off, float y_off, float eps, int32_t *coors, int32_t *num_points,
                          int32_t *voxel_counts, float *pfn_feats, float *canvas, pv_stream_t stream);

/* Measurement aid for bench.py: runs pv_forward_mean_canvas (canvas may be NULL for 3-D grids)
 * `iters` times with CUDA events recorded on `stream` between the stages and returns the average
 * milliseconds per stage in stage_ms (HOST, PV_PROFILE_STAGES floats).  List-free pipeline:
 * 0 insert (builds the first-point bitmap), 1 (empty; was the cells pass), 2 scan, 3 finalize (+ canvas),
 * 4 heavy cells; list-based pipeline:
 * 0 bin_insert, 1 cell_flags, 2 scan, 3 place, 4 emit (see pv_profile_pipeline).  Synchronises. */
#define PV_PROFILE_STAGES 5
int pv_profile_mean_canvas(const pv_config *cfg, const float *points, const int32_t *frame_offsets,
                           int32_t batch, int64_t n_total, int32_t c_in, int32_t is_cartesian,
                           int64_t max_points_total, int64_t frame_capacity, void *workspace,
                           size_t workspace_bytes, int32_t *coors, int32_t *num_points,
                           int32_t *voxel_counts, float *mean_feats, float *canvas,
                           pv_stream_t stream, int32_t iters, float *stage_ms);

/*
 * Dynamic voxelization of a batch ("dynamic=True" configs of the reference): replaces
 *   - the grid index of Voxelization.voxelize_dynamic (det3d/datasets/pipelines/voxelization.py:169-172:
 *     floor(clip((p - lo) / vs, 0, grid - 1)), every point gets a cell, nothing is dropped) and the
 *     batch-index padding of collate_kitti (torchie/parallel/collate.py:157-164),
 *   - torch.unique(grid_ind, return_inverse=True, return_counts=True, dim=0) + torch_scatter.scatter_mean
 *     in DynamicVoxelEncoderV1.forward (det3d/models/readers/voxel_encoder.py:38-44),
 *   - DynamicPPScatter.forward (det3d/models/readers/pillar_encoder.py:413-432) when canvas != NULL.
 * Voxels are ordered by (b, z, y, x) as torch.unique sorts them; there is no max_points / max_voxels cap.
 * Two input modes: grid_ind_in == NULL -> the points are binned here (frame_offsets required, Cartesian
 * rows are transformed on the fly as in pv_voxelize); grid_ind_in != NULL -> int32 [n, 4] (b, z, y, x)
 * rows computed by the caller are used as they are (frame_offsets may be NULL; rows outside the grid
 * set a status bit and get unq_inv = -1).
 * Outputs (M = sum of voxel_counts <= min(n_total, batch * cells) rows):
 *   grid_ind_out int32 [n, 4] or NULL      unq      int32 [M, 4] (b, z, y, x)   required
 *   unq_inv      int32 [n]   or NULL       unq_cnt  int32 [M]  or NULL
 *   voxel_counts int32 [batch]  required   mean_feats f32 [M, C] or NULL
 *   canvas       f32 [batch, C, ny, nx] or NULL (pillar grids)
 * Grids up to 2^26 cells per frame: direct maps (<= 2^20 cells, the pillar grids) keep the rows per cell,
 * larger grids (the reference's 3-D cylinder grids of voxelnet_det_cylinder_singlehead.py /
 * voxelnet_seg_cylinder.py: 1024 x 1024 x 40, 640 x 640 x 40) keep them in the hash map while the voxel
 * order still comes from a cell-order occupancy bitmap (no canvas there); PV_ERR_UNSUPPORTED beyond.
 * Integer outputs are bit-exact; means as in the list-free pipeline.
 */
int pv_dynamic_voxelize(const pv_config *cfg, const float *points, const int32_t *frame_offsets,
                        const int32_t *grid_ind_in, int32_t batch, int64_t n_total, int32_t c_in,
                        int32_t is_cartesian, int64_t max_points_total, int64_t frame_capacity,
                        void *workspace, size_t workspace_bytes, int32_t *grid_ind_out, int32_t *unq,
                        int32_t *unq_inv, int32_t *unq_cnt, int32_t *voxel_counts, float *mean_feats,
                        float *canvas, pv_stream_t stream);

/* Binning only: the grid index of Voxelization.voxelize_dynamic (voxelization.py:169-172) with the batch
 * column of collate_kitti, grid_ind_out int32 [n, 4] (b, z, y, x) = floor(clip((p - lo) / vs, 0, grid - 1))
 * for every point -- on ANY grid (no map is built), no workspace. */
int pv_dynamic_grid_ind(const pv_config *cfg, const float *points, const int32_t *frame_offsets, int32_t batch,
                        int64_t n_total, int32_t c_in, int32_t is_cartesian, int32_t *grid_ind_out, pv_stream_t stream);

/*
 * DynamicPFNet.forward (det3d/models/readers/pillar_encoder.py:262-411; PFNLayer.forward_dynamic :63-71)
 * on the outputs of pv_dynamic_voxelize: feature decoration (get_cluster :228-238, cell centres
 * :350-351 through polar2cart / cart2polar :240-260), then n_layers x (Linear without bias, ReLU,
 * scatter_max over the voxel); non-last layers pass cat([x, x_max[unq_inv]]) on.  No normalisation
 * (the dynamic forward never calls PFNLayer.norm; only `weight` of pv_pfn_layer is read).
 *   points [n, c] feature rows; unq [m, 4] (b, z, y, x); unq_inv [n] (-1 = skip); unq_cnt [m];
 *   voxel_mean [m, c] = scatter_mean of the rows (pv_dynamic_voxelize's mean_feats);
 *   cylinder != 0: voxel_shape == 'cylinder' (xyz = columns 3, 4, 2; ra = columns 0, 1), else
 *   'cuboid' (xyz = columns 0..2; ra = the last two columns) -- the reference's configs leave the
 *   reader at 'cuboid'; flags: bit 0 xyz_cluster, 1 raz_cluster, 2 xy_center, 3 ra_center;
 *   vx, vy, x_off, y_off as computed at pillar_encoder.py:331-334;  out [m, units of the last layer].
 * 1 or 2 layers, decorated width <= 32, units multiples of 4 (first <= 64 when a second layer
 * follows, last <= 128); PV_ERR_UNSUPPORTED otherwise.  `layers` is a HOST array.
 */
size_t pv_dynamic_pfn_workspace_bytes(int64_t n, int64_t m);
int pv_dynamic_pfn(const float *points, const int32_t *unq, const int32_t *unq_inv, const int32_t *unq_cnt,
                   const float *voxel_mean, int64_t n, int64_t m, int32_t c, int32_t cylinder, int32_t flags,
                   float vx, float vy, float x_off, float y_off, const pv_pfn_layer *layers, int32_t n_layers,
                   void *workspace, size_t workspace_bytes, float *out, pv_stream_t stream);

/*
 * Azimuth-sector streaming of ONE polar point cloud: the point path of
 * Voxelization.voxelize_streaming_polar (det3d/datasets/pipelines/voxelization.py:305-371) for all
 * `nsectors` wedges in one stable partition.  points [n, c >= 5] are cylinder rows
 * (rho, phi, z, x, y, ...); wedge i keeps phi in [lo + i * iv, lo + (i+1) * iv) with
 * iv = (max_azimuth - lo) / nsectors in float32 (first wedge open below, last open above, :350-358),
 * shifts phi to the first wedge (:360), recomputes x, y = rho * cos / sin(phi) (:361-362) and takes
 * floor(clip((p - lo) / vs, 0, cur_grid - 1)) with cur_grid[1] = ny / nsectors (:366-368).
 * Outputs are sector-major, original order inside a sector (what np.where gives):
 *   points_out f32 [n, c], grid_ind_out int32 [n, 3] (z, y, x), point_index int32 [n] (source row),
 *   sector_counts int32 [nsectors]; rows [0, sum(sector_counts)) are valid (NaN azimuths are in no wedge).
 * max_azimuth = pc_range[4] (cfg keeps only the lower bounds).
 */
size_t pv_stream_workspace_bytes(int64_t n, int32_t nsectors);
int pv_stream_sectors(const pv_config *cfg, const float *points, int64_t n, int32_t c, int32_t nsectors,
                      float max_azimuth, void *workspace, size_t workspace_bytes, float *points_out,
                      int32_t *grid_ind_out, int32_t *point_index, int32_t *sector_counts, pv_stream_t stream);

/* Rigid warp between sweeps of Voxelization.voxelize_streaming_by_sweep
 * (det3d/datasets/pipelines/voxelization.py:439-447): out[:, :3] = float32(M[:3, :3] . xyz + M[:3, 3])
 * evaluated in float64, out[:, c-1] = in[:, c-1] - t_shift (the time-lag fix), other columns copied.
 * matrix3x4: HOST pointer to the first three rows of the 4 x 4 transform, row-major. */
int pv_affine_points(const float *in, int64_t n, int32_t c, const double *matrix3x4, float t_shift,
                     float *out, pv_stream_t stream);

/* Segmentation voxel labels: the train branch of
 * Voxelization.get_grid_ind (det3d/datasets/pipelines/voxelization.py:40-60) followed by
 * AssignLabel.assign_voxel_labels (det3d/datasets/pipelines/preprocess.py:170-191), batched.
 *   pc_grid_ind int32 [n, 3] (z, y, x): the clamped per-point grid index pv_voxelize returns
 *   pc_label    int32 [n]: semantic label of every point; < 0 = unlabelled (dropped, :44); must be < 256
 * Outputs:
 *   voxel_labels   int64 [batch, nz, ny, nx]: per cell the label held by most of its valid points
 *                  (ties -> smallest label, np.argmax; the reference's uint16 counter wraps at 65536
 *                  and so does this), 0 for cells without valid points; every element is written
 *   valid_grid_ind int32 [n, 3]: the rows of the valid points in their original order (:45, :58);
 *                  frame b owns rows [valid_offsets[b], valid_offsets[b + 1])
 *   valid_offsets  int32 [batch + 1]
 *   status         int32 [1] (device): 0, bit 0 = table full, bit 1 = a label > 255 or a grid index
 *                  outside the grid was skipped (undefined behaviour in the reference)
 * The reference sorts the rows by cell only to group them; the vote is order free (bit-exact). */
size_t pv_seg_workspace_bytes(const pv_config *cfg, int64_t n_total, int32_t batch);
int pv_seg_voxel_labels(const pv_config *cfg, const int32_t *pc_grid_ind, const int32_t *pc_label,
                        const int32_t *frame_offsets, int32_t batch, int64_t n_total, void *workspace,
                        size_t workspace_bytes, int64_t *voxel_labels, int32_t *valid_grid_ind,
                        int32_t *valid_offsets, int32_t *status, pv_stream_t stream);

/* SegHead.predict (det3d/models/seg_heads/seg_head.py:171-193): out[i] = pred_labels[b][z, y, x]
 * for every valid point i of frame b; pred_labels int64 [batch, nz_pred, ny, nx].  nz_pred == 0
 * selects the 2-D form pred_labels [batch, ny, nx] indexed [y, x] (:188).  status as above (bit 1 =
 * an index outside the map; that row reads 0). */
int pv_seg_gather_points(const int64_t *pred_labels, int32_t nz_pred, int32_t ny, int32_t nx,
                         const int32_t *valid_grid_ind, const int32_t *valid_offsets, int32_t batch,
                         int64_t n_valid, int64_t *out, int32_t *status, pv_stream_t stream);

/* Copies the device status word of the last call on `workspace` to the host (synchronises
 * `stream`).  Returns PV_OK or PV_ERR_TABLE_FULL / PV_ERR_BAD_ARGUMENT / PV_ERR_CUDA, and
 * PV_ERR_INTERNAL when the barrier watchdog of the tensor-core PFN kernel fired (its results are
 * then undefined).  Also accepts the workspace of pv_pfn_forward. */
int pv_read_status(const void *workspace, pv_stream_t stream);

/* VoxelFeatureExtractorV3.forward (det3d/models/readers/voxel_encoder.py:15-22):
 * out[m, c] = sum_t voxels[m, t, c] / num_points[m]. */
int pv_vfe_mean(const float *voxels, const int32_t *num_points, int64_t m, int32_t t, int32_t c,
                float *out, pv_stream_t stream);

/* PillarFeatureNet.forward in eval mode (pillar_encoder.py:131-169, PFNLayer :49-61):
 * decoration (cluster offset, pillar-centre offset, optional distance), padding mask,
 * n_layers x (Linear, BatchNorm1d eval, ReLU, max over all T slots incl. padding).
 * voxels [m, t, c], num_points [m], coors [m, 4] (b, z, y, x); out [m, units of last layer].
 * `layers` is a HOST array. vx, vy, x_off, y_off as computed at pillar_encoder.py:123-126.
 * Padding slots (t >= num_points) must hold zeros, as points_to_voxel produces them. */
int pv_pfn_forward(const float *voxels, const int32_t *num_points, const int32_t *coors, int64_t m,
                   int32_t t, int32_t c, int32_t with_distance, float vx, float vy, float x_off,
                   float y_off, const pv_pfn_layer *layers, int32_t n_layers, float eps,
                   void *workspace, size_t workspace_bytes, float *out, pv_stream_t stream);

/* PillarFeatureNet in TRAINING mode (pillar_encoder.py:37-38,49-61 with self.norm in training mode, and the
 * backward pass torch.autograd derives from it).  BatchNorm1d uses the batch statistics over ALL m * t rows,
 * padded slots included, and updates running_mean / running_var IN PLACE through layers[l].bn_mean / bn_var
 * (momentum `momentum`, unbiased variance); the maximum runs over all t slots.
 *   forward : out [m, units of the last layer]; everything the backward pass needs stays in `workspace`
 *             (pv_pfn_train_workspace_bytes), which the caller keeps until pv_pfn_train_backward.
 *   backward: d_out [m, units of last layer] -> d_weight[l] [units, in_channels], d_gamma[l], d_beta[l] [units]
 *             (HOST arrays of device pointers, overwritten).  No gradient flows to the point features.
 * Units must divide 256; `layers` is a HOST array (same layout as pv_pfn_forward). */
size_t pv_pfn_train_workspace_bytes(int64_t m, int32_t t, int32_t c, int32_t with_distance, const pv_pfn_layer *layers,
                                    int32_t n_layers);
int pv_pfn_train_forward(const float *voxels, const int32_t *num_points, const int32_t *coors, int64_t m, int32_t t,
                         int32_t c, int32_t with_distance, float vx, float vy, float x_off, float y_off,
                         const pv_pfn_layer *layers, int32_t n_layers, float eps, float momentum, void *workspace,
                         size_t workspace_bytes, float *out, pv_stream_t stream);
int pv_pfn_train_backward(const float *d_out, int64_t m, int32_t t, int32_t c, int32_t with_distance,
                          const pv_pfn_layer *layers, int32_t n_layers, void *workspace, size_t workspace_bytes,
                          float *const *d_weight, float *const *d_gamma, float *const *d_beta, pv_stream_t stream);

/* Scratch bytes pv_pfn_forward needs for m voxels of t slots (per-voxel statistics and, for
 * two-layer nets whose second layer runs on tcgen05, the layer-0 rows). */
size_t pv_pfn_workspace_bytes(int64_t m, int32_t t);

/* Tensor-core building block of the PFN linear layers (PFNLayer.linear, pillar_encoder.py:41,50):
 * d[m, n] = a[m, k] . b[n, k]^T in fp32 via tcgen05.mma.kind::tf32 with the 3xTF32 split
 * (fp32-accurate), accumulators in TMEM.  n multiple of 16 in [16, 256], k multiple of 8, <= 64.
 * `variant` must be 0 (bit 0 swaps the descriptor stride roles; bring-up aid). */
int pv_tc_gemm_tf32x3(const float *a, const float *b, int32_t m, int32_t n, int32_t k, float *d,
                      int32_t variant, pv_stream_t stream);

/* Bytes of workspace pv_scatter needs (the BEV index map). */
size_t pv_scatter_workspace_bytes(int32_t batch, int32_t ny, int32_t nx);

/* PointPillarsScatter.forward (pillar_encoder.py:189-225): canvas [batch, c, ny, nx] =
 * zeros, canvas[b, :, y, x] = feats[v, :] for coors[v] = (b, z, y, x).  Rows whose batch
 * index is outside [0, batch) are ignored (:207); duplicate cells keep the last row.
 * bev_index (optional, int64 [m]) receives y * nx + x (:211). */
int pv_scatter(const float *feats, const int32_t *coors, int64_t m, int32_t c, int32_t batch,
               int32_t ny, int32_t nx, void *workspace, size_t workspace_bytes, float *canvas,
               int64_t *bev_index, pv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* POLAR_VOXEL_B200_H_ */
