"""CPU oracle for the polar front end -- TEST INFRASTRUCTURE ONLY.

ctypes binding of ``oracle/polar_oracle.c`` (a plain-C restatement of the
reference's numba / numpy / torch code for this path; every C function cites
the reference file:line it follows).  Only ``tests/``, ``__graft_entry__.smoke``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this
package; ``partner_b200`` never does.

Parity status: pinned against the reference's own functions through the
golden vectors in ``tests/golden`` (see ``tests/golden/make_golden.py``).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpolar_oracle.so")
_lib = None


def build(force=False):
    """Compile polar_oracle.c with gcc (seconds)."""
    src = os.path.join(_HERE, "polar_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libpolar_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.po_points_to_voxel.restype = ctypes.c_int64
        L.po_atan2f.restype = ctypes.c_float
        L.po_atan2f.argtypes = [ctypes.c_float, ctypes.c_float]
        _lib = L
    return _lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


def _f32c(a):
    a = np.ascontiguousarray(a)
    if a.dtype != np.float32:
        raise TypeError("oracle expects float32, got %s" % a.dtype)
    return a


def grid_size(voxel_size, point_cloud_range):
    """voxel_generator.py:10-11 -> int64[3] in xyz order."""
    vs = np.asarray(voxel_size, dtype=np.float32)
    rg = np.asarray(point_cloud_range, dtype=np.float32)
    g = np.zeros(3, np.int32)
    lib().po_grid_size(_p(vs), _p(rg), _p(g))
    return g.astype(np.int64)


def transform_points(input_pc, voxel_shape="cylinder"):
    """pipelines/utils.py:34-47."""
    pc = _f32c(input_pc)
    n, c_in = pc.shape
    out = np.empty((n, c_in + 2), np.float32)
    lib().po_transform_points(_p(pc), ctypes.c_int64(n), ctypes.c_int(c_in),
                              ctypes.c_int(1 if voxel_shape == "cylinder" else 0), _p(out))
    return out


def points_to_voxel(points, voxel_size, coors_range, max_points=35, reverse_index=True,
                    max_voxels=20000, return_pc_grid_ind=False, return_density=False):
    """point_cloud_ops.py:146-224 (reverse_index=True only, the path det3d uses)."""
    if not reverse_index:
        raise NotImplementedError("the front end only uses reverse_index=True")
    pts = _f32c(points)
    n, c = pts.shape
    vs = np.asarray(voxel_size, dtype=np.float32)
    rg = np.asarray(coors_range, dtype=np.float32)
    g = grid_size(vs, rg)
    voxels = np.empty((max_voxels, max_points, c), np.float32)
    coors = np.empty((max_voxels, 3), np.int32)
    num = np.empty((max_voxels,), np.int32)
    ind = np.empty((n, 3), np.int32) if return_pc_grid_ind else None
    den = np.empty((int(g[2]), int(g[1]), int(g[0])), np.int32) if return_density else None
    m = lib().po_points_to_voxel(_p(pts), ctypes.c_int64(n), ctypes.c_int(c), _p(vs), _p(rg),
                                 ctypes.c_int(max_points), ctypes.c_int(max_voxels),
                                 _p(voxels), _p(coors), _p(num), _p(ind), _p(den))
    if m < 0:
        raise MemoryError("oracle: dense voxel map allocation failed")
    return voxels[:m], coors[:m], num[:m], ind, den


class VoxelGenerator:
    """core/input/voxel_generator.py:5-48."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000):
        self._point_cloud_range = np.array(point_cloud_range, dtype=np.float32)
        self._voxel_size = np.array(voxel_size, dtype=np.float32)
        self._grid_size = grid_size(self._voxel_size, self._point_cloud_range)
        self._max_num_points = max_num_points
        self._max_voxels = max_voxels

    def generate(self, points, max_voxels=-1, return_pc_grid_ind=False, return_density=False):
        if max_voxels == -1:
            max_voxels = self._max_voxels
        return points_to_voxel(points, self._voxel_size, self._point_cloud_range,
                               self._max_num_points, True, max_voxels,
                               return_pc_grid_ind, return_density)

    voxel_size = property(lambda s: s._voxel_size)
    max_num_points_per_voxel = property(lambda s: s._max_num_points)
    point_cloud_range = property(lambda s: s._point_cloud_range)
    grid_size = property(lambda s: s._grid_size)


def collate(frames):
    """torchie/parallel/collate.py:157-164 + concat of voxels/num_points/num_voxels.

    frames: list of (voxels, coors[M,3], num_points) -> voxels [SM,T,C],
    coordinates [SM,4] (b,z,y,x) int32, num_points [SM], num_voxels [B] int64.
    """
    vox = np.concatenate([f[0] for f in frames], axis=0)
    coor = np.concatenate(
        [np.pad(f[1], ((0, 0), (1, 0)), mode="constant", constant_values=i)
         for i, f in enumerate(frames)], axis=0).astype(np.int32)
    num = np.concatenate([f[2] for f in frames], axis=0)
    nv = np.array([f[0].shape[0] for f in frames], dtype=np.int64)
    return vox, coor, num, nv


def vfe_mean(voxels, num_points):
    """voxel_encoder.py:15-22."""
    v = _f32c(voxels)
    n = np.ascontiguousarray(num_points, dtype=np.int32)
    m, t, c = v.shape
    out = np.empty((m, c), np.float32)
    lib().po_vfe_mean(_p(v), _p(n), ctypes.c_int64(m), ctypes.c_int(t), ctypes.c_int(c), _p(out))
    return out


def pfn_forward(voxels, num_points, coors, layers, voxel_size, pc_range, with_distance=False,
                eps=1e-3):
    """pillar_encoder.py:131-169 in eval mode.

    layers: list of dicts {weight [U,K], mean, var, gamma, beta} (float32).
    """
    v = _f32c(voxels)
    n = np.ascontiguousarray(num_points, dtype=np.int32)
    co = np.ascontiguousarray(coors, dtype=np.int32)
    m, t, c = v.shape
    vx, vy = float(voxel_size[0]), float(voxel_size[1])
    x_off = vx / 2 + float(pc_range[0])        # pillar_encoder.py:125-126 (Python doubles)
    y_off = vy / 2 + float(pc_range[1])
    nl = len(layers)
    units = np.array([l["weight"].shape[0] for l in layers], np.int32)
    keep = []

    def arr(key):
        a = (ctypes.c_void_p * nl)()
        for i, l in enumerate(layers):
            x = _f32c(l[key])
            keep.append(x)
            a[i] = x.ctypes.data
        return a

    out = np.empty((m, int(units[-1])), np.float32)
    lib().po_pfn_forward(_p(v), _p(n), _p(co), ctypes.c_int64(m), ctypes.c_int(t), ctypes.c_int(c),
                         ctypes.c_int(1 if with_distance else 0),
                         ctypes.c_float(vx), ctypes.c_float(vy), ctypes.c_float(x_off),
                         ctypes.c_float(y_off), ctypes.c_int(nl), _p(units),
                         arr("weight"), arr("mean"), arr("var"), arr("gamma"), arr("beta"),
                         ctypes.c_float(eps), _p(out))
    return out


def scatter(voxel_features, coords, batch_size, input_shape):
    """pillar_encoder.py:189-225 -> (canvas [B,C,ny,nx], bev_index [M] int64)."""
    f = _f32c(voxel_features)
    co = np.ascontiguousarray(coords, dtype=np.int32)
    m, c = f.shape
    nx, ny = int(input_shape[0]), int(input_shape[1])
    canvas = np.empty((batch_size, c, ny, nx), np.float32)
    idx = np.empty((m,), np.int64)
    lib().po_scatter(_p(f), _p(co), ctypes.c_int64(m), ctypes.c_int(c), ctypes.c_int(batch_size),
                     ctypes.c_int(ny), ctypes.c_int(nx), _p(canvas), _p(idx))
    return canvas, idx


def dynamic_grid_ind(points, voxel_size, point_cloud_range):
    """voxelize_dynamic, datasets/pipelines/voxelization.py:169-172 -> int32 [N, 3] (z, y, x)."""
    pts = _f32c(points)
    n, c = pts.shape
    vs = np.asarray(voxel_size, dtype=np.float32)
    rg = np.asarray(point_cloud_range, dtype=np.float32)
    out = np.empty((n, 3), np.int32)
    lib().po_dynamic_grid_ind(_p(pts), ctypes.c_int64(n), ctypes.c_int(c), _p(vs), _p(rg), _p(out))
    return out


def dynamic_mean(grid_ind, features):
    """DynamicVoxelEncoderV1.forward, models/readers/voxel_encoder.py:38-44.

    grid_ind int [N, 4] (b, z, y, x), features f32 [N, C] ->
    (mean [M, C], unq int32 [M, 4], unq_inv int64 [N], unq_cnt int64 [M])."""
    gi = np.ascontiguousarray(grid_ind, dtype=np.int32)
    f = _f32c(features)
    n, c = f.shape
    unq = np.empty((max(n, 1), 4), np.int32)
    inv = np.empty((n,), np.int64)
    cnt = np.empty((max(n, 1),), np.int64)
    mean = np.empty((max(n, 1), c), np.float32)
    L = lib()
    L.po_dynamic_mean.restype = ctypes.c_int64
    m = L.po_dynamic_mean(_p(gi), _p(f), ctypes.c_int64(n), ctypes.c_int(c), _p(unq), _p(inv), _p(cnt), _p(mean))
    if m < 0:
        raise MemoryError("oracle: allocation failed")
    return mean[:m], unq[:m], inv, cnt[:m]


def dynamic_pfn(points, unq_inv, unq, weights, voxel_size, pc_range, voxel_shape="cuboid", xyz_cluster=False,
                raz_cluster=False, xy_center=False, ra_center=False):
    """DynamicPFNet.forward, models/readers/pillar_encoder.py:338-411 (weights: list of [U, K] f32)."""
    p = _f32c(points)
    inv = np.ascontiguousarray(unq_inv, dtype=np.int64)
    u4 = np.ascontiguousarray(unq, dtype=np.int32)
    n, c = p.shape
    m = u4.shape[0]
    vx, vy = float(voxel_size[0]), float(voxel_size[1])
    x_off = vx / 2 + float(pc_range[0])
    y_off = vy / 2 + float(pc_range[1])
    units = np.array([w.shape[0] for w in weights], np.int32)
    keep = [_f32c(w) for w in weights]
    arr = (ctypes.c_void_p * len(keep))(*[w.ctypes.data for w in keep])
    out = np.empty((m, int(units[-1])), np.float32)
    flags = (1 if xyz_cluster else 0) | (2 if raz_cluster else 0) | (4 if xy_center else 0) | (8 if ra_center else 0)
    rc = lib().po_dynamic_pfn(_p(p), _p(inv), _p(u4), ctypes.c_int64(n), ctypes.c_int64(m), ctypes.c_int(c),
                              ctypes.c_int(1 if voxel_shape != "cuboid" else 0), ctypes.c_int(flags),
                              ctypes.c_float(vx), ctypes.c_float(vy), ctypes.c_float(x_off), ctypes.c_float(y_off),
                              ctypes.c_int(len(keep)), _p(units), arr, _p(out))
    if rc:
        raise MemoryError("oracle: allocation failed")
    return out


def stream_polar(points, voxel_size, point_cloud_range, nsectors):
    """Voxelization.voxelize_streaming_polar (voxelization.py:305-371, evaluation path).

    -> list of (points [n_s, C], grid_ind int32 [n_s, 3] (z, y, x), point_index int32 [n_s]) per sector."""
    pts = _f32c(points)
    n, c = pts.shape
    vs = np.asarray(voxel_size, dtype=np.float32)
    rg = np.asarray(point_cloud_range, dtype=np.float32)
    op = np.empty((max(n, 1), c), np.float32)
    gi = np.empty((max(n, 1), 3), np.int32)
    ix = np.empty((max(n, 1),), np.int32)
    cnt = np.zeros((nsectors,), np.int64)
    L = lib()
    L.po_stream_polar.restype = ctypes.c_int64
    L.po_stream_polar(_p(pts), ctypes.c_int64(n), ctypes.c_int(c), _p(vs), _p(rg), ctypes.c_int(nsectors),
                      _p(op), _p(gi), _p(ix), _p(cnt))
    out, o = [], 0
    for s in range(nsectors):
        k = int(cnt[s])
        out.append((op[o:o + k], gi[o:o + k], ix[o:o + k]))
        o += k
    return out


def seg_voxel_labels(pc_grid_ind, pc_label, grid_size):
    """Voxelization.get_grid_ind, train branch (voxelization.py:40-60) + AssignLabel.assign_voxel_labels
    (preprocess.py:170-191).  grid_size is (nx, ny, nz) as VoxelGenerator.grid_size.

    -> (voxel_labels int64 [1, nz, ny, nx], valid_grid_ind int32 [n_valid, 3])."""
    gi = np.ascontiguousarray(pc_grid_ind, dtype=np.int32)
    lab = np.ascontiguousarray(np.asarray(pc_label).reshape(-1), dtype=np.int32)
    n = gi.shape[0]
    nx, ny, nz = (int(v) for v in grid_size)
    labels = np.empty((1, nz, ny, nx), np.int64)
    valid = np.empty((max(n, 1), 3), np.int32)
    L = lib()
    L.po_seg_voxel_labels.restype = ctypes.c_int64
    m = L.po_seg_voxel_labels(_p(gi), _p(lab), ctypes.c_int64(n), ctypes.c_int(nz), ctypes.c_int(ny), ctypes.c_int(nx),
                              _p(labels), _p(valid))
    if m < 0:
        raise MemoryError("oracle: allocation failed")
    return labels, valid[:m]


def seg_gather_points(pred_labels, valid_grid_ind):
    """SegHead.predict (seg_heads/seg_head.py:184-191) for one sample: pred_labels [nz, ny, nx] or [ny, nx]."""
    g = np.asarray(valid_grid_ind)
    if pred_labels.ndim == 2:
        return pred_labels[g[:, 1], g[:, 2]]
    return pred_labels[g[:, 0], g[:, 1], g[:, 2]]
