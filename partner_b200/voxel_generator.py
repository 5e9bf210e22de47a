"""Drop-in for det3d/core/input/voxel_generator.py:5-48 running on the B200.

Same constructor, ``generate`` signature, return tuple and properties as the reference class.
``generate`` accepts what the reference accepts (a float32 numpy array [N, C] of already-polar
points, returning numpy arrays) and additionally CUDA tensors (returning CUDA tensors, no host
round trip).  ``generate_batch`` voxelizes several frames in one launch sequence and returns them
collated the way collate_kitti does (det3d/torchie/parallel/collate.py:157-164).
"""
import numpy as np
import torch

from . import functional as F


class VoxelGenerator:
    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000, device=None):
        self._cfg, self._voxel_size, self._point_cloud_range, self._grid_size = F.make_config(
            voxel_size, point_cloud_range, max_num_points, max_voxels)
        self._max_num_points = max_num_points
        self._max_voxels = max_voxels
        self._device = torch.device(device) if device is not None else None

    # ---- reference API -------------------------------------------------------------------
    def generate(self, points, max_voxels=-1, return_pc_grid_ind=False, return_density=False):
        """voxel_generator.py:19-32 -> (voxels, coors [M,3] zyx, num_points, pc_grid_ind, density)."""
        as_numpy = isinstance(points, np.ndarray)
        pts = self._to_device(points)
        vb, _ = self._run([pts], max_voxels, True, return_pc_grid_ind, return_density, cartesian=False)
        m = vb.total()
        F.read_status(vb)
        voxels = vb.voxels[:m]
        coors = vb.coors[:m, 1:].contiguous()
        num = vb.num_points[:m]
        ind = vb.pc_grid_ind
        den = vb.density[0] if vb.density is not None else None
        if as_numpy:
            return F.to_numpy(voxels, coors, num, ind, den)     # page-locked staging, one synchronisation
        return voxels, coors, num, ind, den

    @property
    def voxel_size(self):
        return self._voxel_size

    @property
    def max_num_points_per_voxel(self):
        return self._max_num_points

    @property
    def point_cloud_range(self):
        return self._point_cloud_range

    @property
    def grid_size(self):
        return self._grid_size

    # ---- batched extension ---------------------------------------------------------------
    def generate_batch(self, frames, max_voxels=-1, return_pc_grid_ind=False, return_density=False,
                       cartesian=False, return_voxels=True, return_mean=False):
        """Voxelize a list of frames in one pass; outputs collated like collate_kitti.

        Returns dict(voxels [SM,T,C], coordinates [SM,4] (b,z,y,x), num_points [SM],
        num_voxels [B] int64, shape, and the optional outputs).  ``cartesian=True`` fuses
        transform_points (pipelines/utils.py:34-44) into the voxelizer.
        """
        pts = [self._to_device(p) for p in frames]
        vb, offsets = self._run(pts, max_voxels, return_voxels, return_pc_grid_ind, return_density,
                                cartesian=cartesian, want_mean=return_mean)
        counts = vb.voxel_counts.cpu()
        m = int(counts.sum())
        F.read_status(vb)
        out = dict(coordinates=vb.coors[:m], num_points=vb.num_points[:m],
                   num_voxels=counts.to(torch.int64), shape=self._grid_size)
        if vb.voxels is not None:
            out["voxels"] = vb.voxels[:m]
        if vb.mean_feats is not None:
            out["mean_features"] = vb.mean_feats[:m]
        if vb.pc_grid_ind is not None:
            out["pc_grid_ind"] = vb.pc_grid_ind
            out["point_offsets"] = offsets
        if vb.density is not None:
            out["n_points"] = vb.density
        return out

    # ---- internals -----------------------------------------------------------------------
    def _to_device(self, points):
        if isinstance(points, np.ndarray):
            if points.dtype != np.float32:
                # the reference would silently JIT a float64 specialisation with different bins
                raise ValueError("points must be float32, got %s" % points.dtype)
            if points.ndim != 2:
                raise ValueError("points must be [N, C]")
            dev = self._device or torch.device("cuda", torch.cuda.current_device())
            src = torch.from_numpy(np.ascontiguousarray(points))
            if src.numel() == 0:
                return src.to(dev)
            stage = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)     # cached page-locked staging
            stage.copy_(src)
            return stage.to(dev, non_blocking=True)
        if isinstance(points, torch.Tensor):
            if not points.is_cuda:
                raise ValueError("tensor input must live on the GPU (pass numpy for host data)")
            if points.dtype != torch.float32 or points.dim() != 2:
                raise ValueError("points must be float32 [N, C]")
            return points.contiguous()
        raise TypeError("points must be a numpy array or a CUDA tensor")

    def _run(self, pts, max_voxels, want_voxels, want_ind, want_den, cartesian, want_mean=False):
        if max_voxels == -1:
            max_voxels = self._max_voxels
        cfg = F.PvConfig()
        cfg.lo, cfg.vs, cfg.grid = self._cfg.lo, self._cfg.vs, self._cfg.grid
        cfg.max_points = self._cfg.max_points
        cfg.max_voxels = int(max_voxels)
        cfg.pipeline = self._cfg.pipeline
        sizes = [int(p.shape[0]) for p in pts]
        if len({p.shape[1] for p in pts}) != 1:
            raise ValueError("all frames need the same number of columns")
        offsets = np.zeros(len(pts) + 1, dtype=np.int32)
        np.cumsum(sizes, out=offsets[1:])
        dev = pts[0].device
        allp = pts[0] if len(pts) == 1 else torch.cat(pts, dim=0)
        off_dev = torch.from_numpy(offsets).to(dev)
        vb = F.voxelize(cfg, allp, off_dev, len(pts), max(sizes) if sizes else 0, cartesian,
                        want_voxels=want_voxels, want_mean=want_mean, want_grid_ind=want_ind,
                        want_density=want_den)
        return vb, offsets
