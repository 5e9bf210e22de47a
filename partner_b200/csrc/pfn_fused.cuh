// pfn_fused.cuh -- host-side interface of the warp-specialised tensor-core PFN kernel (pfn_fused.cu).
#pragma once
#include "pv_common.cuh"

struct P2Args {
    int mode;
    // mode 0: padded tensor
    const float *voxels; const int32_t *num; const int32_t *coors_in; long long m, v0;   // voxels [v0, v0 + m) of the tensor (one slice)
    // mode 1: point lists of the list-based voxelizer
    const float *pts; int c_in, cart;
    const uint32_t *vox_cell, *vox_kg, *vox_c; uint32_t *kept;
    const int32_t *base; const int32_t *voxel_counts;
    uint32_t fcap; int32_t nx, ny;
    int32_t *coors_out; int32_t *num_out;
    // mode 1: decorated rows + group descriptors written by k_pfn_rows, read by k_pfn_fused<2>
    float4 *drows_out; uint4 *desc_out; uint32_t *ngroups_out; long long drow_stride;   // rows per float4 plane
    const float4 *drows; const uint4 *desc; const uint32_t *ngroups;
    // common
    int t, c, c0, with_distance;
    float vx, vy, x_off, y_off, eps;
    const float *w0, *mean0, *var0, *gamma0, *beta0;
    const float *w1, *mean1, *var1, *gamma1, *beta1;
    int n1;
    uint32_t chunks_per_frame, n_chunks;
    unsigned int *counter;
    unsigned int *status;      // optional second status word (the voxelizer workspace's, fused front end); may be null
    float *out;
};


// Shapes the kernel covers: two layers, 32 units in the first, 32 | units of the last <= 128,
// decorated width <= 16, T <= 32.
bool pv_pfn_fused_supported(const pv_pfn_layer *layers, int n_layers, int t, int c, int with_distance);
// counter: 64 words (256 bytes) of device scratch (chunk queue, status, watchdog diagnostics, row allocator), zeroed by the launch.
// workspace of the pre-pass: [rows x 64 B | descriptors | groups per chunk]; offsets of the last two are returned
long long pv_pfn_rows_stride(long long rows);
size_t pv_pfn_rows_bytes(long long rows, long long voxels_per_frame_cap, int batch_frames, size_t *desc_off, size_t *ng_off);
int pv_pfn_fused_launch(P2Args &a, const pv_pfn_layer *layers, int batch_frames, long long voxels_per_frame_cap,
                        cudaStream_t st);
