"""Tensor-level wrappers over the C ABI (torch is only the allocator / stream provider).

Every function takes CUDA float32/int32 tensors, validates them the way the reference's native
ops do (is_cuda / is_contiguous / dtype -- det3d/ops/iou3d_nms/src/iou3d_nms.cpp:14-27, but raising
instead of exiting), allocates outputs with torch and launches on the current stream.
"""
import numpy as np
import torch

from . import _lib
from ._lib import PvConfig, PvPfnLayer, check, current_stream, ptr

_workspaces = {}


_default_pipeline = 0


def set_default_pipeline(mode):
    """Test / measurement aid: pv_config.pipeline of every configuration made from now on
    (0 auto, 1 list-based, 2 list-free).  Host-side only; the library itself keeps no state."""
    global _default_pipeline
    if mode not in (0, 1, 2):
        raise ValueError("pipeline must be 0 (auto), 1 (list-based) or 2 (list-free)")
    _default_pipeline = mode


def make_config(voxel_size, point_cloud_range, max_points, max_voxels, pipeline=None):
    """VoxelGenerator.__init__ arithmetic (det3d/core/input/voxel_generator.py:6-11)."""
    rng = np.array(point_cloud_range, dtype=np.float32)
    vs = np.array(voxel_size, dtype=np.float32)
    grid = np.round((rng[3:] - rng[:3]) / vs).astype(np.int64)
    cfg = PvConfig()
    for j in range(3):
        cfg.lo[j] = float(rng[j])
        cfg.vs[j] = float(vs[j])
        cfg.grid[j] = int(grid[j])
    cfg.max_points = int(max_points)
    cfg.max_voxels = int(max_voxels)
    cfg.pipeline = _default_pipeline if pipeline is None else int(pipeline)
    return cfg, vs, rng, grid


def _need(t, dtype, name, ndim=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (there is no CPU path)" % name)
    if t.dtype != dtype:
        raise ValueError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    if ndim is not None and t.dim() != ndim:
        raise ValueError("%s must have %d dims" % (name, ndim))


def workspace(nbytes, device, tag="main"):
    """Grow-only byte buffer per (device, current stream, tag); the library itself never allocates.
    Keyed by the stream so that callers working concurrently on different streams never share (or,
    when one of them grows it, free) each other's scratch memory."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream, tag)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _bucket(n, quantum=65536):
    """Capacities are rounded up so that frames of varying size share one initialised workspace."""
    return max(quantum, (int(n) + quantum - 1) // quantum * quantum)


_voxel_ws = {}          # layout key -> initialised workspace tensor (LRU, self-cleaning: see header)
_VOXEL_WS_KEEP = 12


def voxel_workspace(cfg, n_cap, batch, frame_cap, channels, device, tag=0):
    """Workspace initialised (pv_workspace_init) for exactly this config + capacities.

    ``tag`` separates workspaces of callers that run concurrently on different streams."""
    dev_index = device.index if device.index is not None else torch.cuda.current_device()
    key = (dev_index, tag, tuple(cfg.lo), tuple(cfg.vs), tuple(cfg.grid), cfg.max_points, n_cap, batch, frame_cap, channels)   # both pipelines share one layout
    ws = _voxel_ws.pop(key, None)
    if ws is None:
        lib = _lib.load()
        nbytes = lib.pv_workspace_bytes(cfg, n_cap, batch, frame_cap, channels)
        if nbytes == 0:
            check(-1, "pv_workspace_bytes")
        while len(_voxel_ws) >= _VOXEL_WS_KEEP:
            _voxel_ws.pop(next(iter(_voxel_ws)))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        check(lib.pv_workspace_init(cfg, n_cap, batch, frame_cap, channels, ptr(ws), ws.numel(), current_stream(device)),
              "pv_workspace_init")
    _voxel_ws[key] = ws       # most recently used last
    return ws


def drop_voxel_workspaces():
    """Forget every cached workspace (call after a device-side error left one dirty)."""
    _voxel_ws.clear()


def transform_points(points, voxel_shape="cylinder"):
    """pv_transform_points: pipelines/utils.py:34-47 on the device."""
    _need(points, torch.float32, "points", 2)
    n, c_in = points.shape
    if c_in < 3:
        raise ValueError("points need at least x, y, z columns")
    if voxel_shape not in ("cylinder", "cuboid"):
        raise ValueError("voxel_shape must be 'cylinder' or 'cuboid'")
    out = torch.empty((n, c_in + 2), dtype=torch.float32, device=points.device)
    check(_lib.load().pv_transform_points(ptr(points), n, c_in, 1 if voxel_shape == "cylinder" else 0,
                                          ptr(out), current_stream(points.device)), "pv_transform_points")
    return out


class VoxelBatch:
    """Outputs of one batched voxelization, in capacity layout until ``counts`` is read."""

    __slots__ = ("coors", "num_points", "voxel_counts", "voxels", "mean_feats", "pc_grid_ind",
                 "density", "canvas", "ws", "cfg", "n_cap", "f_cap")

    def total(self):
        """Sum of per-frame voxel counts (one device->host read)."""
        return int(self.voxel_counts.sum().item())


def _check_out(r, dev, rows, batch, want):
    """``out=`` reuses a previous call's buffers: they must fit THIS call (rows = min(batch * V, n), batch,
    channel counts, the outputs asked for) -- the kernels write `rows` rows whatever the tensors hold."""
    def fits(t, shape, dtype, what):
        if t is None:
            raise ValueError("out= has no %s buffer, but this call produces one" % what)
        if t.device != dev or t.dtype != dtype or not t.is_contiguous():
            raise ValueError("out.%s must be a contiguous %s tensor on %s" % (what, dtype, dev))
        if t.dim() != len(shape) or t.shape[0] < shape[0] or tuple(t.shape[1:]) != tuple(shape[1:]):
            raise ValueError("out.%s has shape %s, this call needs at least %s" % (what, tuple(t.shape), tuple(shape)))
    fits(r.coors, (rows, 4), torch.int32, "coors")
    fits(r.num_points, (rows,), torch.int32, "num_points")
    fits(r.voxel_counts, (batch,), torch.int32, "voxel_counts")
    if r.voxel_counts.shape[0] != batch:
        raise ValueError("out.voxel_counts has %d entries, this call has %d frames" % (r.voxel_counts.shape[0], batch))
    for name, (shape, dtype) in want.items():
        fits(getattr(r, name), shape, dtype, name)


def voxelize(cfg, points, frame_offsets, batch, frame_capacity, is_cartesian, want_voxels=False,
             want_mean=False, want_grid_ind=False, want_density=False, canvas=False, out=None, ws_tag=0):
    """pv_voxelize / pv_forward_mean_canvas on CUDA tensors.

    points [N, c_in] f32, frame_offsets [batch+1] int32 (device).  Returns a VoxelBatch whose row
    tensors have capacity min(batch*V, N); rows [0, sum(voxel_counts)) are valid.
    ``out`` may carry a previous VoxelBatch whose buffers are reused (static-shape serving).
    """
    _need(points, torch.float32, "points", 2)
    _need(frame_offsets, torch.int32, "frame_offsets", 1)
    if frame_offsets.numel() != batch + 1:
        raise ValueError("frame_offsets must have batch+1 entries")
    n, c_in = points.shape
    C = c_in + 2 if is_cartesian else c_in
    dev = points.device
    lib = _lib.load()
    n_cap = _bucket(n)
    f_cap = min(n_cap, _bucket(frame_capacity))
    ws = voxel_workspace(cfg, n_cap, batch, f_cap, C, dev, ws_tag)
    rows = max(1, min(batch * cfg.max_voxels, n))
    T = cfg.max_points
    r = out if out is not None else VoxelBatch()
    if out is None:
        r.coors = torch.empty((rows, 4), dtype=torch.int32, device=dev)
        r.num_points = torch.empty((rows,), dtype=torch.int32, device=dev)
        r.voxel_counts = torch.empty((batch,), dtype=torch.int32, device=dev)
        r.voxels = torch.empty((rows, T, C), dtype=torch.float32, device=dev) if want_voxels else None
        r.mean_feats = torch.empty((rows, C), dtype=torch.float32, device=dev) if (want_mean or canvas) else None
        r.pc_grid_ind = torch.empty((n, 3), dtype=torch.int32, device=dev) if want_grid_ind else None
        r.density = (torch.empty((batch, cfg.grid[2], cfg.grid[1], cfg.grid[0]), dtype=torch.int32, device=dev)
                     if want_density else None)
        r.canvas = (torch.empty((batch, C, cfg.grid[1], cfg.grid[0]), dtype=torch.float32, device=dev)
                    if canvas else None)
    else:
        want = {}
        if want_voxels:
            want["voxels"] = ((rows, T, C), torch.float32)
        if want_mean or canvas:
            want["mean_feats"] = ((rows, C), torch.float32)
        if want_grid_ind:
            want["pc_grid_ind"] = ((n, 3), torch.int32)
        if want_density:
            want["density"] = ((batch, int(cfg.grid[2]), int(cfg.grid[1]), int(cfg.grid[0])), torch.int32)
        if canvas:
            want["canvas"] = ((batch, C, int(cfg.grid[1]), int(cfg.grid[0])), torch.float32)
        _check_out(r, dev, rows, batch, want)
        if not want_voxels:
            r.voxels = None
    r.ws = ws
    r.cfg = cfg
    r.n_cap, r.f_cap = n_cap, f_cap
    st = current_stream(dev)
    if canvas:
        check(lib.pv_forward_mean_canvas(cfg, ptr(points), ptr(frame_offsets), batch, n, c_in,
                                         1 if is_cartesian else 0, n_cap, f_cap, ptr(ws), ws.numel(),
                                         ptr(r.coors), ptr(r.num_points), ptr(r.voxel_counts),
                                         ptr(r.mean_feats), ptr(r.canvas), st), "pv_forward_mean_canvas")
    else:
        check(lib.pv_voxelize(cfg, ptr(points), ptr(frame_offsets), batch, n, c_in,
                              1 if is_cartesian else 0, n_cap, f_cap, ptr(ws), ws.numel(),
                              ptr(r.coors), ptr(r.num_points), ptr(r.voxel_counts), ptr(r.voxels),
                              ptr(r.mean_feats), ptr(r.pc_grid_ind), ptr(r.density), st), "pv_voxelize")
    return r


def _pfn_layer_array(layers):
    arr = (PvPfnLayer * len(layers))()
    for i, (w, mean, var, gamma, beta) in enumerate(layers):
        for nm, x in (("weight", w), ("mean", mean), ("var", var), ("gamma", gamma), ("beta", beta)):
            _need(x, torch.float32, "pfn_layers.%d.%s" % (i, nm))
        arr[i].weight, arr[i].bn_mean, arr[i].bn_var = w.data_ptr(), mean.data_ptr(), var.data_ptr()
        arr[i].bn_gamma, arr[i].bn_beta = gamma.data_ptr(), beta.data_ptr()
        arr[i].units, arr[i].in_channels = w.shape[0], w.shape[1]
    return arr


def forward_pfn_canvas(cfg, points, frame_offsets, batch, frame_capacity, is_cartesian, layers, vx, vy, x_off, y_off,
                       with_distance, eps, canvas=True, out=None, ws_tag=0):
    """pv_forward_pfn_canvas: voxelize -> PillarFeatureNet (eval) -> PointPillarsScatter in one launch
    sequence, without the padded [M, T, C] tensor.  layers as for pfn_forward.  Returns a VoxelBatch
    whose ``mean_feats`` slot holds the PFN features [capacity, U] and ``canvas`` [B, U, ny, nx]."""
    _need(points, torch.float32, "points", 2)
    _need(frame_offsets, torch.int32, "frame_offsets", 1)
    if frame_offsets.numel() != batch + 1:
        raise ValueError("frame_offsets must have batch+1 entries")
    n, c_in = points.shape
    C = c_in + 2 if is_cartesian else c_in
    dev = points.device
    lib = _lib.load()
    n_cap = _bucket(n)
    f_cap = min(n_cap, _bucket(frame_capacity))
    ws = voxel_workspace(cfg, n_cap, batch, f_cap, C, dev, ws_tag)
    rows = max(1, min(batch * cfg.max_voxels, n))
    units = layers[-1][0].shape[0]
    ny, nx = int(cfg.grid[1]), int(cfg.grid[0])
    r = out if out is not None else VoxelBatch()
    if out is None:
        r.coors = torch.empty((rows, 4), dtype=torch.int32, device=dev)
        r.num_points = torch.empty((rows,), dtype=torch.int32, device=dev)
        r.voxel_counts = torch.empty((batch,), dtype=torch.int32, device=dev)
        r.voxels = r.pc_grid_ind = r.density = None
        r.mean_feats = torch.empty((rows, units), dtype=torch.float32, device=dev)
        r.canvas = torch.empty((batch, units, ny, nx), dtype=torch.float32, device=dev) if canvas else None
    else:
        want = {"mean_feats": ((rows, units), torch.float32)}
        if canvas:
            want["canvas"] = ((batch, units, ny, nx), torch.float32)
        _check_out(r, dev, rows, batch, want)
    r.ws, r.cfg, r.n_cap, r.f_cap = ws, cfg, n_cap, f_cap
    aux = workspace(lib.pv_pfn_canvas_workspace_bytes(batch, ny, nx, n_cap, int(cfg.max_voxels)), dev, ("pfn_canvas", ws_tag))
    arr = _pfn_layer_array(layers)
    check(lib.pv_forward_pfn_canvas(cfg, ptr(points), ptr(frame_offsets), batch, n, c_in, 1 if is_cartesian else 0,
                                    n_cap, f_cap, ptr(ws), ws.numel(), ptr(aux), aux.numel(), arr, len(layers),
                                    1 if with_distance else 0, vx, vy, x_off, y_off, eps, ptr(r.coors), ptr(r.num_points),
                                    ptr(r.voxel_counts), ptr(r.mean_feats), ptr(r.canvas), current_stream(dev)),
          "pv_forward_pfn_canvas")
    return r


class DynamicBatch:
    """Outputs of pv_dynamic_voxelize in capacity layout (rows [0, sum(voxel_counts)) are valid)."""

    __slots__ = ("unq", "unq_inv", "unq_cnt", "voxel_counts", "mean_feats", "grid_ind", "canvas", "ws", "cfg")

    def total(self):
        return int(self.voxel_counts.sum().item())


def dynamic_voxelize(cfg, points, frame_offsets, batch, frame_capacity, is_cartesian=False, grid_ind=None,
                     want_inverse=True, want_counts=True, want_grid_ind=False, canvas=False, ws_tag=0, want_mean=True):
    """pv_dynamic_voxelize on CUDA tensors: dynamic voxelization + unique + scatter_mean (+ scatter).

    points [N, c_in] f32; either frame_offsets [batch+1] int32 (the points are binned here) or
    grid_ind [N, 4] int32 (b, z, y, x) computed by the caller.  Returns a DynamicBatch."""
    _need(points, torch.float32, "points", 2)
    n, c_in = points.shape
    if grid_ind is not None:
        _need(grid_ind, torch.int32, "grid_ind", 2)
        if grid_ind.shape != (n, 4):
            raise ValueError("grid_ind must be [N, 4] (b, z, y, x)")
    else:
        _need(frame_offsets, torch.int32, "frame_offsets", 1)
        if frame_offsets.numel() != batch + 1:
            raise ValueError("frame_offsets must have batch+1 entries")
    C = c_in + 2 if is_cartesian else c_in
    dev = points.device
    lib = _lib.load()
    n_cap = _bucket(n)
    f_cap = min(n_cap, _bucket(frame_capacity if frame_capacity else n))
    ws = voxel_workspace(cfg, n_cap, batch, f_cap, C, dev, ws_tag)
    cells = cfg.grid[0] * cfg.grid[1] * cfg.grid[2]
    rows = max(1, min(batch * cells, n))
    r = DynamicBatch()
    r.unq = torch.empty((rows, 4), dtype=torch.int32, device=dev)
    r.unq_inv = torch.empty((n,), dtype=torch.int32, device=dev) if want_inverse else None
    r.unq_cnt = torch.empty((rows,), dtype=torch.int32, device=dev) if want_counts else None
    r.voxel_counts = torch.empty((batch,), dtype=torch.int32, device=dev)
    r.mean_feats = torch.empty((rows, C), dtype=torch.float32, device=dev) if (want_mean or canvas) else None
    r.grid_ind = torch.empty((n, 4), dtype=torch.int32, device=dev) if want_grid_ind else None
    r.canvas = (torch.empty((batch, C, cfg.grid[1], cfg.grid[0]), dtype=torch.float32, device=dev)
                if canvas else None)
    r.ws, r.cfg = ws, cfg
    check(lib.pv_dynamic_voxelize(cfg, ptr(points), ptr(frame_offsets) if grid_ind is None else ptr(None),
                                  ptr(grid_ind), batch, n, c_in, 1 if is_cartesian else 0, n_cap, f_cap,
                                  ptr(ws), ws.numel(), ptr(r.grid_ind), ptr(r.unq), ptr(r.unq_inv), ptr(r.unq_cnt),
                                  ptr(r.voxel_counts), ptr(r.mean_feats), ptr(r.canvas), current_stream(dev)),
          "pv_dynamic_voxelize")
    return r


def dynamic_grid_ind(cfg, points, frame_offsets, batch, is_cartesian=False):
    """pv_dynamic_grid_ind: [N, 4] int32 (b, z, y, x) clamped grid index of every point, on any grid."""
    _need(points, torch.float32, "points", 2)
    _need(frame_offsets, torch.int32, "frame_offsets", 1)
    if frame_offsets.numel() != batch + 1:
        raise ValueError("frame_offsets must have batch+1 entries")
    n, c_in = points.shape
    gi = torch.empty((max(n, 1), 4), dtype=torch.int32, device=points.device)
    check(_lib.load().pv_dynamic_grid_ind(cfg, ptr(points), ptr(frame_offsets), batch, n, c_in, 1 if is_cartesian else 0,
                                          ptr(gi), current_stream(points.device)), "pv_dynamic_grid_ind")
    return gi[:n]


def dynamic_pfn(points, batch, m, weights, vx, vy, x_off, y_off, cylinder, xyz_cluster, raz_cluster, xy_center,
                ra_center):
    """pv_dynamic_pfn on a DynamicBatch (`batch`) with `m` valid voxel rows; weights: list of [U, K] CUDA f32."""
    _need(points, torch.float32, "points", 2)
    n, c = points.shape
    arr = (PvPfnLayer * len(weights))()
    for i, w in enumerate(weights):
        _need(w, torch.float32, "pfn_layers.%d.linear.weight" % i, 2)
        arr[i].weight = w.data_ptr()
        arr[i].units, arr[i].in_channels = w.shape[0], w.shape[1]
    dev = points.device
    lib = _lib.load()
    out = torch.empty((m, weights[-1].shape[0]), dtype=torch.float32, device=dev)
    ws = workspace(max(256, lib.pv_dynamic_pfn_workspace_bytes(n, m)), dev, "dynpfn")
    flags = (1 if xyz_cluster else 0) | (2 if raz_cluster else 0) | (4 if xy_center else 0) | (8 if ra_center else 0)
    check(lib.pv_dynamic_pfn(ptr(points), ptr(batch.unq), ptr(batch.unq_inv), ptr(batch.unq_cnt), ptr(batch.mean_feats),
                             n, m, c, 1 if cylinder else 0, flags, vx, vy, x_off, y_off, arr, len(weights),
                             ptr(ws), ws.numel(), ptr(out), current_stream(dev)), "pv_dynamic_pfn")
    return out


def stream_sectors(cfg, points, nsectors, max_azimuth):
    """pv_stream_sectors: (points_out [N, C], grid_ind [N, 3] (z, y, x), point_index [N], counts [nsectors])."""
    _need(points, torch.float32, "points", 2)
    n, c = points.shape
    dev = points.device
    lib = _lib.load()
    nbytes = lib.pv_stream_workspace_bytes(n, nsectors)
    if nbytes == 0:
        raise ValueError("nsectors must be in [1, 64]")
    ws = workspace(nbytes, dev, "stream")
    out = torch.empty((max(n, 1), c), dtype=torch.float32, device=dev)
    gi = torch.empty((max(n, 1), 3), dtype=torch.int32, device=dev)
    idx = torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
    counts = torch.empty((nsectors,), dtype=torch.int32, device=dev)
    check(lib.pv_stream_sectors(cfg, ptr(points), n, c, nsectors, float(np.float32(max_azimuth)), ptr(ws), ws.numel(),
                                ptr(out), ptr(gi), ptr(idx), ptr(counts), current_stream(dev)), "pv_stream_sectors")
    return out, gi, idx, counts


def seg_voxel_labels(cfg, pc_grid_ind, pc_label, frame_offsets, batch):
    """pv_seg_voxel_labels: (voxel_labels int64 [batch, nz, ny, nx], valid_grid_ind int32 [n_valid, 3],
    valid_offsets int32 [batch + 1]).  Synchronises once to read the status word and the valid count."""
    _need(pc_grid_ind, torch.int32, "pc_grid_ind", 2)
    _need(pc_label, torch.int32, "pc_label", 1)
    _need(frame_offsets, torch.int32, "frame_offsets", 1)
    n = pc_grid_ind.shape[0]
    if pc_grid_ind.shape[1] != 3 or pc_label.shape[0] != n or frame_offsets.shape[0] != batch + 1:
        raise ValueError("pc_grid_ind must be [n, 3], pc_label [n], frame_offsets [batch + 1]")
    dev = pc_grid_ind.device
    lib = _lib.load()
    nbytes = lib.pv_seg_workspace_bytes(cfg, n, batch)
    if nbytes == 0:
        raise ValueError("bad configuration for pv_seg_voxel_labels")
    ws = workspace(nbytes, dev, "seg")
    nx, ny, nz = (int(v) for v in cfg.grid)
    labels = torch.empty((batch, nz, ny, nx), dtype=torch.int64, device=dev)
    valid = torch.empty((max(n, 1), 3), dtype=torch.int32, device=dev)
    voff = torch.empty((batch + 1,), dtype=torch.int32, device=dev)
    status = torch.empty((1,), dtype=torch.int32, device=dev)
    check(lib.pv_seg_voxel_labels(cfg, ptr(pc_grid_ind), ptr(pc_label), ptr(frame_offsets), batch, n, ptr(ws), ws.numel(),
                                  ptr(labels), ptr(valid), ptr(voff), ptr(status), current_stream(dev)),
          "pv_seg_voxel_labels")
    st, n_valid = int(status.item()), int(voff[-1].item())
    if st & 1:
        raise RuntimeError("pv_seg_voxel_labels: label table full")
    if st & 2:
        raise ValueError("pv_seg_voxel_labels: a label > 255 or a grid index outside the grid")
    return labels, valid[:n_valid], voff


def seg_gather_points(pred_labels, valid_grid_ind, valid_offsets):
    """pv_seg_gather_points: per-point predictions pred_labels[b][z, y, x] (pred_labels [B, nz, ny, nx]) or
    pred_labels[b][y, x] (pred_labels [B, ny, nx]) -- SegHead.predict, seg_head.py:171-193."""
    _need(pred_labels, torch.int64, "pred_labels")
    _need(valid_grid_ind, torch.int32, "valid_grid_ind", 2)
    _need(valid_offsets, torch.int32, "valid_offsets", 1)
    if pred_labels.dim() not in (3, 4):
        raise ValueError("pred_labels must be [B, ny, nx] or [B, nz, ny, nx]")
    batch = pred_labels.shape[0]
    nz = pred_labels.shape[1] if pred_labels.dim() == 4 else 0
    ny, nx = int(pred_labels.shape[-2]), int(pred_labels.shape[-1])
    if valid_offsets.shape[0] != batch + 1 or valid_grid_ind.shape[1] != 3:
        raise ValueError("valid_offsets must be [B + 1], valid_grid_ind [n, 3]")
    n = valid_grid_ind.shape[0]
    dev = pred_labels.device
    out = torch.empty((n,), dtype=torch.int64, device=dev)
    status = torch.empty((1,), dtype=torch.int32, device=dev)
    check(_lib.load().pv_seg_gather_points(ptr(pred_labels), nz, ny, nx, ptr(valid_grid_ind), ptr(valid_offsets), batch,
                                           n, ptr(out), ptr(status), current_stream(dev)), "pv_seg_gather_points")
    if int(status.item()) & 2:
        raise ValueError("pv_seg_gather_points: a grid index outside the prediction map")
    return out


def to_numpy(*tensors):
    """Device tensors -> numpy arrays through page-locked buffers of torch's caching host allocator:
    all copies in flight together, one stream synchronisation (a pageable ``.cpu()`` of a 34 MB
    voxels tensor alone takes ~8 ms).  ``None`` entries pass through."""
    hs = []
    dev = None
    for t in tensors:
        if t is None or not t.is_cuda:
            hs.append(t)
            continue
        dev = t.device
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        hs.append(h)
    if dev is not None:
        torch.cuda.current_stream(dev).synchronize()
    return tuple(None if h is None else h.numpy() for h in hs)


def affine_points(points, matrix, t_shift=0.0):
    """pv_affine_points: xyz <- M[:3, :3] . xyz + M[:3, 3] (float64 arithmetic), last column -= t_shift."""
    import ctypes
    _need(points, torch.float32, "points", 2)
    n, c = points.shape
    m = np.ascontiguousarray(np.asarray(matrix, dtype=np.float64)[:3, :4])
    if m.shape != (3, 4):
        raise ValueError("matrix must be 4 x 4 (or 3 x 4)")
    out = torch.empty_like(points)
    check(_lib.load().pv_affine_points(ptr(points), n, c, m.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                       float(np.float32(t_shift)), ptr(out), current_stream(points.device)),
          "pv_affine_points")
    return out


def read_status(vb):
    rc = _lib.load().pv_read_status(ptr(vb.ws), current_stream(vb.ws.device))
    if rc != 0:
        drop_voxel_workspaces()      # the failed call left its workspace dirty
    check(rc, "device status")


def vfe_mean(features, num_voxels):
    _need(features, torch.float32, "features", 3)
    _need(num_voxels, torch.int32, "num_voxels", 1)
    m, t, c = features.shape
    if num_voxels.numel() != m:
        raise ValueError("num_voxels must have one entry per voxel")
    out = torch.empty((m, c), dtype=torch.float32, device=features.device)
    check(_lib.load().pv_vfe_mean(ptr(features), ptr(num_voxels), m, t, c, ptr(out),
                                  current_stream(features.device)), "pv_vfe_mean")
    return out


def pfn_forward(features, num_voxels, coors, layers, vx, vy, x_off, y_off, with_distance, eps):
    """layers: list of (weight [U,K], running_mean, running_var, gamma, beta) CUDA f32 tensors."""
    _need(features, torch.float32, "features", 3)
    _need(num_voxels, torch.int32, "num_voxels", 1)
    _need(coors, torch.int32, "coors", 2)
    m, t, c = features.shape
    if coors.shape != (m, 4) or num_voxels.numel() != m:
        raise ValueError("coors must be [M,4] and num_voxels [M]")
    arr = _pfn_layer_array(layers)
    out = torch.empty((m, layers[-1][0].shape[0]), dtype=torch.float32, device=features.device)
    lib = _lib.load()
    ws = workspace(max(256, lib.pv_pfn_workspace_bytes(m, t)), features.device, "pfn")
    check(lib.pv_pfn_forward(ptr(features), ptr(num_voxels), ptr(coors), m, t, c,
                             1 if with_distance else 0, vx, vy, x_off, y_off, arr, len(layers),
                             eps, ptr(ws), ws.numel(), ptr(out), current_stream(features.device)), "pv_pfn_forward")
    return out


def pfn_train_forward(features, num_voxels, coors, layers, vx, vy, x_off, y_off, with_distance, eps, momentum):
    """pv_pfn_train_forward.  layers as for pfn_forward (running_mean / running_var are UPDATED in place).
    Returns (out [M, U], saved) where ``saved`` holds what pfn_train_backward needs."""
    _need(features, torch.float32, "features", 3)
    _need(num_voxels, torch.int32, "num_voxels", 1)
    _need(coors, torch.int32, "coors", 2)
    m, t, c = features.shape
    if coors.shape != (m, 4) or num_voxels.numel() != m:
        raise ValueError("coors must be [M,4] and num_voxels [M]")
    arr = _pfn_layer_array(layers)
    lib = _lib.load()
    nbytes = lib.pv_pfn_train_workspace_bytes(m, t, c, 1 if with_distance else 0, arr, len(layers))
    if nbytes == 0:
        raise ValueError("PFN training kernels: unsupported layer shapes (units must divide 256)")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=features.device)      # lives until the backward pass
    out = torch.empty((m, layers[-1][0].shape[0]), dtype=torch.float32, device=features.device)
    check(lib.pv_pfn_train_forward(ptr(features), ptr(num_voxels), ptr(coors), m, t, c, 1 if with_distance else 0,
                                   vx, vy, x_off, y_off, arr, len(layers), eps, momentum, ptr(ws), ws.numel(), ptr(out),
                                   current_stream(features.device)), "pv_pfn_train_forward")
    return out, (ws, m, t, c, bool(with_distance))


def pfn_train_backward(d_out, saved, layers):
    """pv_pfn_train_backward -> lists (d_weight, d_gamma, d_beta), one entry per layer."""
    import ctypes
    ws, m, t, c, with_distance = saved
    _need(d_out, torch.float32, "d_out", 2)
    arr = _pfn_layer_array(layers)
    dev = d_out.device
    dw = [torch.empty_like(l[0]) for l in layers]
    dg = [torch.empty_like(l[3]) for l in layers]
    db = [torch.empty_like(l[4]) for l in layers]
    vec = lambda ts: (ctypes.c_void_p * len(ts))(*[x.data_ptr() for x in ts])      # noqa: E731
    check(_lib.load().pv_pfn_train_backward(ptr(d_out), m, t, c, 1 if with_distance else 0, arr, len(layers), ptr(ws),
                                            ws.numel(), vec(dw), vec(dg), vec(db), current_stream(dev)),
          "pv_pfn_train_backward")
    return dw, dg, db


def scatter(voxel_features, coords, batch_size, ny, nx, want_bev_index=False):
    _need(voxel_features, torch.float32, "voxel_features", 2)
    _need(coords, torch.int32, "coords", 2)
    m, c = voxel_features.shape
    if coords.shape != (m, 4):
        raise ValueError("coords must be [M,4] (b,z,y,x)")
    dev = voxel_features.device
    lib = _lib.load()
    nbytes = lib.pv_scatter_workspace_bytes(batch_size, ny, nx)
    ws = workspace(nbytes, dev, "scatter")
    canvas = torch.empty((batch_size, c, ny, nx), dtype=torch.float32, device=dev)
    bev = torch.empty((m,), dtype=torch.int64, device=dev) if want_bev_index else None
    check(lib.pv_scatter(ptr(voxel_features), ptr(coords), m, c, batch_size, ny, nx, ptr(ws),
                         ws.numel(), ptr(canvas), ptr(bev), current_stream(dev)), "pv_scatter")
    return (canvas, bev) if want_bev_index else canvas
