"""Fused, batched front end: raw Cartesian sweeps -> voxels -> reader -> BEV canvas on one GPU.

This is the whole hot path of SURVEY.md section 3.4 as one launch sequence per *batch*
(transform_points + VoxelGenerator.generate per frame + collate + reader + scatter in the
reference), with device-side voxel counts (no host sync) so it can be replayed as a CUDA graph.
Frames are independent, so multi-GPU use is plain frame sharding (``shard_range``); there is no
collective on this path.
"""
import numpy as np
import torch

from . import functional as F


from .sharding import shard_range  # noqa: F401,E402  (re-exported)


class PolarFrontEnd:
    """voxelize (+ fused cylinder transform) -> mean VFE -> dense polar BEV canvas.

    Parameters follow the reference's ``voxel_generator`` config dict
    (det3d/datasets/pipelines/voxelization.py:14-36): range, voxel_size, max_points_in_voxel,
    max_voxel_num.
    """

    def __init__(self, voxel_size, point_cloud_range, max_points_in_voxel, max_voxel_num,
                 cartesian=True, canvas=None, device=None, workspace_tag=0):
        self.cfg, self.voxel_size, self.point_cloud_range, self.grid_size = F.make_config(
            voxel_size, point_cloud_range, max_points_in_voxel, max_voxel_num)
        self.cartesian = bool(cartesian)
        self.pillar = int(self.grid_size[2]) == 1
        self.canvas = self.pillar if canvas is None else bool(canvas)
        if self.canvas and not self.pillar:
            raise ValueError("a dense BEV canvas needs a pillar grid (nz == 1)")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.workspace_tag = workspace_tag     # distinct tags -> instances may run on concurrent streams
        self._graph = None
        self._static = None

    # ---- device-resident path (what `value` in bench.py times) -----------------------------
    def forward_device(self, points, frame_offsets, batch, frame_capacity, out=None):
        """points [N, c_in] f32 CUDA, frame_offsets [batch+1] int32 CUDA -> VoxelBatch (no sync)."""
        return F.voxelize(self.cfg, points, frame_offsets, batch, frame_capacity, self.cartesian,
                          want_mean=True, canvas=self.canvas, out=out, ws_tag=self.workspace_tag)

    def capture(self, points, frame_offsets, batch, frame_capacity):
        """Record the launch sequence on static buffers as a CUDA graph; returns the VoxelBatch
        the graph writes.  Replay with ``replay()`` after refreshing ``points`` in place."""
        out = self.forward_device(points, frame_offsets, batch, frame_capacity)   # warm-up + allocation
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.forward_device(points, frame_offsets, batch, frame_capacity, out=out)
        self._graph, self._static = g, out
        return out

    def replay(self):
        self._graph.replay()
        return self._static

    # ---- host-buffer path (what `e2e` in bench.py times) ------------------------------------
    def make_host_io(self, n_total, batch, c_in):
        """Pinned staging buffers + device mirrors for ``forward_host``."""
        C = c_in + 2 if self.cartesian else c_in
        rows = max(1, min(batch * self.cfg.max_voxels, n_total))
        pin = lambda *s, dt=torch.float32: torch.empty(s, dtype=dt).pin_memory()   # noqa: E731
        io = dict(
            h_points=pin(n_total, c_in), h_offsets=pin(batch + 1, dt=torch.int32),
            d_points=torch.empty((n_total, c_in), dtype=torch.float32, device=self.device),
            d_offsets=torch.empty((batch + 1,), dtype=torch.int32, device=self.device),
            h_counts=pin(batch, dt=torch.int32), h_coors=pin(rows, 4, dt=torch.int32),
            h_num=pin(rows, dt=torch.int32), h_feats=pin(rows, C),
            h_canvas=pin(batch, C, int(self.grid_size[1]), int(self.grid_size[0])) if self.canvas else None,
            vb=None, batch=batch, n_total=n_total)
        return io

    def forward_host(self, io, frame_capacity):
        """H2D of io['h_points'/'h_offsets'] -> kernels -> D2H of every output.  Returns
        (sum of voxel counts, bytes copied H2D, bytes copied D2H); outputs land in io['h_*']."""
        io["d_points"].copy_(io["h_points"], non_blocking=True)
        io["d_offsets"].copy_(io["h_offsets"], non_blocking=True)
        vb = self.forward_device(io["d_points"], io["d_offsets"], io["batch"], frame_capacity, out=io["vb"])
        io["vb"] = vb
        io["h_counts"].copy_(vb.voxel_counts, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        m = int(io["h_counts"].sum())
        io["h_coors"][:m].copy_(vb.coors[:m], non_blocking=True)
        io["h_num"][:m].copy_(vb.num_points[:m], non_blocking=True)
        io["h_feats"][:m].copy_(vb.mean_feats[:m], non_blocking=True)
        d2h = io["h_counts"].numel() * 4 + m * (16 + 4 + 4 * vb.mean_feats.shape[1])
        if self.canvas:
            io["h_canvas"].copy_(vb.canvas, non_blocking=True)
            d2h += io["h_canvas"].numel() * 4
        torch.cuda.current_stream(self.device).synchronize()
        h2d = io["h_points"].numel() * 4 + io["h_offsets"].numel() * 4
        return m, h2d, d2h

    def forward_host_async(self, io, frame_capacity):
        """Like forward_host but without any host synchronisation: H2D, kernels and the D2H of
        every output buffer at full capacity are enqueued on the current stream, so steps issued
        on different streams overlap their copies and kernels.  Rows beyond sum(h_counts) are
        undefined.  Returns (bytes H2D, bytes D2H)."""
        io["d_points"].copy_(io["h_points"], non_blocking=True)
        io["d_offsets"].copy_(io["h_offsets"], non_blocking=True)
        vb = self.forward_device(io["d_points"], io["d_offsets"], io["batch"], frame_capacity, out=io["vb"])
        io["vb"] = vb
        io["h_counts"].copy_(vb.voxel_counts, non_blocking=True)
        io["h_coors"].copy_(vb.coors, non_blocking=True)
        io["h_num"].copy_(vb.num_points, non_blocking=True)
        io["h_feats"].copy_(vb.mean_feats, non_blocking=True)
        d2h = 4 * (io["h_counts"].numel() + io["h_coors"].numel() + io["h_num"].numel() + io["h_feats"].numel())
        if self.canvas:
            io["h_canvas"].copy_(vb.canvas, non_blocking=True)
            d2h += io["h_canvas"].numel() * 4
        return io["h_points"].numel() * 4 + io["h_offsets"].numel() * 4, d2h

    # ---- convenience ---------------------------------------------------------------------
    def __call__(self, frames):
        """List of float32 numpy frames [N_f, c_in] -> dict of numpy outputs (syncs)."""
        sizes = [int(f.shape[0]) for f in frames]
        offsets = np.zeros(len(frames) + 1, dtype=np.int32)
        np.cumsum(sizes, out=offsets[1:])
        pts = torch.from_numpy(np.ascontiguousarray(np.concatenate(frames, axis=0), dtype=np.float32)).to(self.device)
        vb = self.forward_device(pts, torch.from_numpy(offsets).to(self.device), len(frames), max(sizes))
        counts = vb.voxel_counts.cpu().numpy()
        F.read_status(vb)
        m = int(counts.sum())
        coors, num, feats, canvas = F.to_numpy(vb.coors[:m], vb.num_points[:m], vb.mean_feats[:m],
                                               vb.canvas if self.canvas else None)      # page-locked staging
        out = dict(coordinates=coors, num_points=num, num_voxels=counts.astype(np.int64), features=feats)
        if self.canvas:
            out["canvas"] = canvas
        return out


class PillarFrontEnd:
    """PointPillars.extract_feat_static (det3d/models/detectors/point_pillars.py:28-35) on raw sweeps:
    voxelize (+ fused cylinder transform) -> PillarFeatureNet (eval) -> PointPillarsScatter, as ONE
    launch sequence per batch (pv_forward_pfn_canvas).  The padded voxels tensor is never built; the
    PFN's second layer runs on the tensor cores.

    ``reader`` is a ``partner_b200.PillarFeatureNet`` (reference checkpoints load into it unchanged)
    with two PFN layers, e.g. ``num_filters=(64, 128)`` or ``(64, 64)``."""

    def __init__(self, voxel_size, point_cloud_range, max_points_in_voxel, max_voxel_num, reader, cartesian=True,
                 canvas=True, device=None, workspace_tag=0):
        self.cfg, self.voxel_size, self.point_cloud_range, self.grid_size = F.make_config(
            voxel_size, point_cloud_range, max_points_in_voxel, max_voxel_num)
        if int(self.grid_size[2]) != 1:
            raise ValueError("the pillar front end needs a pillar grid (nz == 1)")
        if reader.training:
            raise RuntimeError("PillarFrontEnd implements eval-mode BatchNorm only; call reader.eval() first")
        self.reader = reader
        self.cartesian, self.canvas = bool(cartesian), bool(canvas)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.workspace_tag = workspace_tag

    def _layers(self):
        eps = {l.norm.eps for l in self.reader.pfn_layers}
        if len(eps) != 1:
            raise ValueError("all PFN layers must share one BatchNorm eps")
        return [(l.linear.weight.detach().contiguous(), l.norm.running_mean, l.norm.running_var,
                 l.norm.weight.detach(), l.norm.bias.detach()) for l in self.reader.pfn_layers], eps.pop()

    def forward_device(self, points, frame_offsets, batch, frame_capacity, out=None):
        """points [N, c_in] f32 CUDA, frame_offsets [batch+1] int32 CUDA -> VoxelBatch (no sync):
        ``mean_feats`` holds the PFN features [capacity, U], ``canvas`` the BEV canvas [B, U, ny, nx]."""
        layers, eps = self._layers()
        rd = self.reader
        return F.forward_pfn_canvas(self.cfg, points, frame_offsets, batch, frame_capacity, self.cartesian, layers,
                                    rd.vx, rd.vy, rd.x_offset, rd.y_offset, rd._with_distance, eps,
                                    canvas=self.canvas, out=out, ws_tag=self.workspace_tag)

    def __call__(self, frames):
        """List of float32 numpy frames -> dict(coordinates, num_points, num_voxels, features, canvas) (syncs)."""
        sizes = [int(f.shape[0]) for f in frames]
        offsets = np.zeros(len(frames) + 1, dtype=np.int32)
        np.cumsum(sizes, out=offsets[1:])
        pts = torch.from_numpy(np.ascontiguousarray(np.concatenate(frames, axis=0), dtype=np.float32)).to(self.device)
        vb = self.forward_device(pts, torch.from_numpy(offsets).to(self.device), len(frames), max(sizes))
        counts = vb.voxel_counts.cpu().numpy()
        F.read_status(vb)
        m = int(counts.sum())
        coors, num, feats, canvas = F.to_numpy(vb.coors[:m], vb.num_points[:m], vb.mean_feats[:m],
                                               vb.canvas if self.canvas else None)
        out = dict(coordinates=coors, num_points=num, num_voxels=counts.astype(np.int64), features=feats)
        if self.canvas:
            out["canvas"] = canvas
        return out
