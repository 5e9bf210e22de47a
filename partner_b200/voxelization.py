"""Drop-in for det3d's ``Voxelization`` pipeline step (det3d/datasets/pipelines/voxelization.py:12-181,
462-476) for the non-streaming paths: ``voxelize_hard`` (incl. double-flip test-time augmentation,
:98-144) and ``voxelize_dynamic`` (:148-181).  Same constructor (``cfg=...``, ``super_tasks=...``),
same ``__call__(res, info)`` and the same keys written into ``res["lidar"]``; numpy in, numpy out.

The four voxelizations of a double-flip sample run as ONE batched launch sequence on the GPU.
Polar azimuth-sector streaming (``nsectors > 1``, :305-371) runs as one stable GPU partition
(evaluation path); sweep streaming with bidirectional padding (``transform_type == 'feature'``,
cylinder branch of :393-460) composes the GPU warp, cylinder transform and sector partition.
Training mode is covered as well: ``filter_gt`` (polar branch, utils.py:11-27), the per-sector ground truth of
the streaming paths (filter + rotation into the first wedge, :332-349), the per-sector point labels (:374-377),
the label hand-over of sweep streaming (:451-453) and the label assignment of ``get_grid_ind`` (:40-60, on the
GPU).  Cartesian sector / sweep streaming (:183-303, :404-422) and ``AssignLabel.assign_part_2d`` are outside the
polar front end and raise ``NotImplementedError``.
"""
import numpy as np

from . import functional as F
from .voxel_generator import VoxelGenerator


def _get(cfg, key, default=None, required=False):
    if isinstance(cfg, dict):
        if required and key not in cfg:
            raise KeyError(key)
        return cfg.get(key, default)
    if required:
        return getattr(cfg, key)
    return getattr(cfg, key, default) if not hasattr(cfg, "get") else cfg.get(key, default)


def _dict_select(dict_, inds):
    """det3d/datasets/pipelines/utils.py:3-8."""
    for k, v in dict_.items():
        if isinstance(v, dict):
            _dict_select(v, inds)
        else:
            dict_[k] = v[inds]


def filter_gt(res, pc_range):
    """Ground-truth boxes outside the (polar) range are dropped from every annotation array --
    det3d/datasets/pipelines/utils.py:11-27, the branch for voxel_shape != 'cuboid' (box centre radius in
    [rho_lo, rho_hi], azimuth in [phi_lo, phi_hi]; the reference multiplies the box diagonal by 0).  The cuboid
    branch (a numba polygon test on Cartesian ranges) is outside the polar front end."""
    gt_dict = res["lidar"]["annotations"]
    if len(gt_dict["gt_boxes"]) > 0:
        bv_range = np.asarray(pc_range)[[0, 1, 3, 4]]
        if res.get("voxel_shape", "cylinder") == "cuboid":
            raise NotImplementedError("filter_gt on Cartesian ranges (cuboid voxels) is outside the polar front end")
        boxes = gt_dict["gt_boxes"]
        gt_rho = np.linalg.norm(boxes[:, :2], axis=1)
        gt_diag = np.linalg.norm(boxes[:, 3:5], axis=1)
        gt_diag *= 0
        gt_az = np.arctan2(boxes[:, 1], boxes[:, 0])
        mask = ((gt_rho - gt_diag) >= bv_range[0]) & ((gt_rho + gt_diag) <= bv_range[2]) & (
            gt_az >= bv_range[1]) & (gt_az <= bv_range[3])
        _dict_select(gt_dict, mask)
        res["lidar"]["annotations"] = gt_dict


def rotation_points_single_angle(points, angle, axis=2):
    """det3d/core/bbox/box_np_ops.py:182-204, rotation about z (the only axis the streaming path uses)."""
    if axis not in (2, -1):
        raise ValueError("only the z axis is used by the polar front end")
    rot_sin, rot_cos = np.sin(angle), np.cos(angle)
    rot_mat_T = np.array([[rot_cos, -rot_sin, 0], [rot_sin, rot_cos, 0], [0, 0, 1]], dtype=points.dtype)
    return points @ rot_mat_T


def sector_annotations(cur_res, cur_pc_range, pc_range):
    """Training-time ground truth of one azimuth sector (voxelization.py:332-349): boxes outside the wedge are
    dropped, the rest are rotated into the first wedge (centres, heading and -- if present -- velocities)."""
    filter_gt(cur_res, cur_pc_range)
    gt_boxes = cur_res["lidar"]["annotations"]["gt_boxes"]
    if len(gt_boxes):
        angle = cur_pc_range[1] - pc_range[1]
        gt_boxes[:, :3] = rotation_points_single_angle(gt_boxes[:, :3], angle, axis=2)
        gt_boxes[:, -1] += angle
        if gt_boxes.shape[1] > 7:
            gt_boxes[:, 6:8] = rotation_points_single_angle(
                np.hstack([gt_boxes[:, 6:8], np.zeros((gt_boxes.shape[0], 1))]), angle, axis=2)[:, :2]
        cur_res["lidar"]["annotations"]["gt_boxes"] = gt_boxes


class Voxelization(object):
    def __init__(self, **kwargs):
        cfg = kwargs.get("cfg", None)
        self.dynamic = _get(cfg, "dynamic", False)
        self.range = _get(cfg, "range", required=True)
        self.voxel_size = _get(cfg, "voxel_size", required=True)
        self.max_points_in_voxel = _get(cfg, "max_points_in_voxel", required=True)
        mv = _get(cfg, "max_voxel_num", required=True)
        self.max_voxel_num = [mv, mv] if isinstance(mv, int) else mv
        self.voxel_generator = VoxelGenerator(voxel_size=self.voxel_size, point_cloud_range=self.range,
                                              max_num_points=self.max_points_in_voxel,
                                              max_voxels=self.max_voxel_num[0])
        self.return_density = _get(cfg, "return_density", False)
        self.double_flip = _get(cfg, "double_flip", False)
        self.super_tasks = kwargs.get("super_tasks", ["det"])
        self.nsectors = _get(cfg, "nsectors", 1)
        self.return_pc_grid_ind = False
        if "seg" in self.super_tasks:
            self.return_pc_grid_ind = True
            assert not self.double_flip, "currently not supporting double flip for segmentation"

    # voxelization.py:40-60
    def get_grid_ind(self, res, pc_grid_ind, grid_size):
        if res["mode"] in ["train", "debug_gt"]:
            # :42-54 -- drop unlabelled points, majority label per cell (AssignLabel.assign_voxel_labels,
            # preprocess.py:170-191) on the GPU: pv_seg_voxel_labels
            import torch
            import ctypes
            from ._lib import PvConfig
            dev = torch.device("cuda", torch.cuda.current_device())
            cfg = PvConfig()
            ctypes.memmove(ctypes.byref(cfg), ctypes.byref(self.voxel_generator._cfg), ctypes.sizeof(PvConfig))
            for j in range(3):      # a sector's reduced grid (:316-317): only the grid extent matters here
                cfg.grid[j] = int(grid_size[j])
            out_dtype = pc_grid_ind.dtype
            gi = torch.from_numpy(np.ascontiguousarray(pc_grid_ind, dtype=np.int32)).to(dev)
            lab = torch.from_numpy(np.ascontiguousarray(res["lidar"]["pc_label"]).reshape(-1).astype(np.int32)).to(dev)
            off = torch.tensor([0, gi.shape[0]], dtype=torch.int32, device=dev)
            labels, valid, _ = F.seg_voxel_labels(cfg, gi, lab, off, 1)
            labels, valid = F.to_numpy(labels, valid)
            res["lidar"]["voxels"].update({"labels": labels})                       # [1, nz, ny, nx] int64 (:52)
            pc_grid_ind = valid.astype(out_dtype)
        else:
            pc_grid_ind = pc_grid_ind[:res["lidar"]["n_key_points"]]
        res["lidar"]["voxels"].update({"valid_grid_ind": pc_grid_ind.copy()})
        return res

    def _pack(self, out, f):
        """Frame f of a generate_batch result -> the reference's per-sample dict (:79-87)."""
        vg = self.voxel_generator
        counts = out["num_voxels"].numpy()
        lo = int(counts[:f].sum())
        hi = lo + int(counts[f])
        voxels, coordinates, num_points = F.to_numpy(out["voxels"][lo:hi], out["coordinates"][lo:hi, 1:].contiguous(),
                                                     out["num_points"][lo:hi])
        return dict(voxels=voxels, coordinates=coordinates, num_points=num_points,
                    num_voxels=np.array([hi - lo], dtype=np.int64),
                    shape=vg.grid_size, range=vg.point_cloud_range, size=vg.voxel_size)

    def voxelize_hard(self, res, info):
        vg = self.voxel_generator
        if res["mode"] in ["train", "debug_gt"]:
            filter_gt(res, vg.point_cloud_range)                        # voxelization.py:67-68
        max_voxels = self.max_voxel_num[0] if res["mode"] in ["train", "debug_gt"] else self.max_voxel_num[1]
        double_flip = self.double_flip and (res["mode"] != "train")
        if not double_flip:
            voxels, coordinates, num_points, pc_grid_ind, density = vg.generate(
                res["lidar"]["points"], max_voxels=max_voxels, return_pc_grid_ind=self.return_pc_grid_ind,
                return_density=self.return_density)
            res["lidar"]["voxels"] = dict(voxels=voxels, coordinates=coordinates, num_points=num_points,
                                          num_voxels=np.array([voxels.shape[0]], dtype=np.int64),
                                          shape=vg.grid_size, range=vg.point_cloud_range, size=vg.voxel_size)
        else:
            # :98-144 -- the flipped copies use the generator's default max_voxels, like the reference
            main = vg.generate_batch([res["lidar"]["points"]], max_voxels=max_voxels,
                                     return_pc_grid_ind=self.return_pc_grid_ind, return_density=self.return_density)
            res["lidar"]["voxels"] = self._pack(main, 0)
            pc_grid_ind, density = F.to_numpy(main["pc_grid_ind"] if self.return_pc_grid_ind else None,
                                              main["n_points"][0] if self.return_density else None)
            keys = ["yflip", "xflip", "double_flip"]
            flips = vg.generate_batch([res["lidar"][k + "_points"] for k in keys])      # one launch sequence
            for f, k in enumerate(keys):
                res["lidar"][k + "_voxels"] = self._pack(flips, f)
        if "seg" in self.super_tasks:
            res = self.get_grid_ind(res, pc_grid_ind, vg.grid_size)
            if "part" in self.super_tasks:
                raise NotImplementedError("AssignLabel.assign_part_2d is outside the front end")
        if self.return_density:
            res["lidar"]["voxels"].update({"n_points": density})
        return res, info

    def voxelize_dynamic(self, res, info, **kwargs):
        import torch
        vg = self.voxel_generator
        if res["mode"] in ["train", "debug_gt"]:                        # voxelization.py:156-163
            if res.get("voxel_shape", "cylinder") != "cuboid":
                cur_pc_range = vg.point_cloud_range.copy()
                cur_pc_range[1] = -np.pi
                cur_pc_range[5] = np.pi                                 # (sic: index 5, as the reference writes it)
                filter_gt(res, cur_pc_range)
            else:
                filter_gt(res, vg.point_cloud_range)
        points = np.ascontiguousarray(res["lidar"]["points"], dtype=np.float32)
        n = points.shape[0]
        dev = torch.device("cuda", torch.cuda.current_device())
        off = torch.tensor([0, n], dtype=torch.int32, device=dev)
        # only the clamped grid index is needed here (:169-172): the binning-only kernel, any grid size
        gi = F.dynamic_grid_ind(vg._cfg, torch.from_numpy(points).to(dev), off, 1, False)
        pc_grid_ind = F.to_numpy(gi[:, 1:].contiguous())[0].astype(np.int64)           # (z, y, x), np.int of the reference
        res["lidar"]["voxels"] = dict(grid_ind=pc_grid_ind.copy(), shape=vg.grid_size, range=vg.point_cloud_range,
                                      size=vg.voxel_size)
        if ("seg" in self.super_tasks) and kwargs.get("seg", True):
            res = self.get_grid_ind(res, pc_grid_ind, vg.grid_size)
        return res, info

    def voxelize_streaming_polar(self, res, info, **kwargs):
        """voxelization.py:305-392: all sectors in one stable GPU partition; in training the per-sector ground
        truth (filter + rotation into the first wedge, :332-349) and point labels (:374-377) as the reference."""
        import copy
        import torch
        train = res["mode"] in ["train", "debug_gt"]
        vg = self.voxel_generator
        grid_size, pc_range, voxel_size = vg.grid_size, vg.point_cloud_range, vg.voxel_size
        nsectors = self.nsectors
        min_az, max_az = pc_range[1], pc_range[4]
        interval = (max_az - min_az) / nsectors
        cur_grid_size = grid_size.copy()
        cur_grid_size[1] //= nsectors
        ref_pc_range = pc_range.copy()
        ref_pc_range[4] = min_az + interval
        load_range = info["load_range"] if "load_range" in info else range(nsectors)
        points = np.ascontiguousarray(res["lidar"]["points"], dtype=np.float32)
        dev = torch.device("cuda", torch.cuda.current_device())
        out, gi, idx, counts = F.stream_sectors(vg._cfg, torch.from_numpy(points).to(dev), nsectors, max_az)
        counts, out, gi, idx = F.to_numpy(counts, out, gi, idx)
        offs = np.concatenate([[0], np.cumsum(counts)])
        gi, idx = gi.astype(np.int64), idx.astype(np.int64)
        lidar_rest = {k: v for k, v in res["lidar"].items() if k != "points"}
        sectors = []
        for i in load_range:
            lo, hi = int(offs[i]), int(offs[i + 1])
            cur_res = {k: copy.deepcopy(v) for k, v in res.items() if k != "lidar"}
            cur_res["lidar"] = copy.deepcopy(lidar_rest)
            if train:
                cur_pc_range = pc_range.copy()                          # :328-331
                cur_pc_range[1] = min_az + i * interval
                cur_pc_range[4] = min_az + (i + 1) * interval
                sector_annotations(cur_res, cur_pc_range, pc_range)
            cur_res["lidar"]["points"] = out[lo:hi].copy()
            pc_grid_ind = gi[lo:hi]
            cur_res["lidar"]["voxels"] = dict(grid_ind=pc_grid_ind.copy(), shape=cur_grid_size, range=ref_pc_range,
                                              size=voxel_size)
            if ("seg" in self.super_tasks) and kwargs.get("seg", True):
                points_index = idx[lo:hi]
                if train:                                               # :374-380
                    cur_res["lidar"]["pc_label"] = cur_res["lidar"]["pc_label"][points_index]
                if (not train) or res["mode"] == "debug_gt":
                    key_points_index = points_index[points_index < res["lidar"]["n_key_points"]]
                    cur_res["lidar"]["n_key_points"] = len(key_points_index)
                    cur_res["lidar"]["key_points_index"] = key_points_index
                cur_res = self.get_grid_ind(cur_res, pc_grid_ind, cur_grid_size)
                if ("part" in self.super_tasks) and train:
                    raise NotImplementedError("AssignLabel.assign_part_2d is outside the front end")
            sectors.append(cur_res)
        return {"sectors": sectors}, info

    def voxelize_streaming_by_sweep(self, res, info):
        """voxelization.py:393-460, cylinder branch (bidirectional padding): the later sweeps as they are
        and the earlier sweeps warped back by one pose, each through transform_points +
        voxelize_streaming_polar; the warp, the cylinder transform and the sector partition run on the GPU."""
        import copy
        import torch
        if res.get("voxel_shape", "cylinder") == "cuboid":
            raise NotImplementedError("Cartesian sweep streaming is outside the polar front end")
        train = res["mode"] in ["train", "debug_gt"]
        dev = torch.device("cuda", torch.cuda.current_device())
        npoints_sweep = np.cumsum(res["lidar"]["npoints_sweep"])
        nsweeps = len(npoints_sweep)
        npoints_sweep = np.insert(npoints_sweep, 0, 0)
        pivot = nsweeps - 1
        points = np.ascontiguousarray(res["lidar"]["points"], dtype=np.float32)
        rest = {k: v for k, v in res.items() if k != "lidar"}
        lidar_rest = {k: v for k, v in res["lidar"].items() if k != "points"}

        def run(cart_dev, seg, extra=None):
            cur = copy.deepcopy(rest)
            if extra:
                cur.update(extra)
            cur["lidar"] = copy.deepcopy(lidar_rest)
            polar = F.transform_points(cart_dev[:, :5].contiguous(), "cylinder")          # utils.py:34-44
            cur["lidar"]["points"] = polar.cpu().numpy()
            out, _ = self.voxelize_streaming_polar(cur, info, seg=seg)
            return out["sectors"]

        # later sweeps (:424-429)
        later = run(torch.from_numpy(points[:npoints_sweep[pivot]]).to(dev), True)
        # earlier sweeps, warped back to the previous pose (:431-449)
        prev = points[npoints_sweep[1]:]
        tm = np.linalg.inv(res["lidar"]["transform_matrices"][1])
        t0 = prev[0, -1] if prev.shape[0] else np.float32(0)
        warped = F.affine_points(torch.from_numpy(np.ascontiguousarray(prev)).to(dev), tm, t0)
        earlier = run(warped, False, extra={"transform_matrix": tm[:2, :2]})
        if ("seg" in self.super_tasks) and train:                       # :451-453 the earlier sweeps reuse the labels
            for i in range(len(earlier)):
                earlier[i]["lidar"]["voxels"]["labels"] = later[i]["lidar"]["voxels"]["labels"].copy()
        sweeps = earlier + later
        return {"sweeps": sweeps, "nsweeps": 2, "nsectors": len(earlier)}, info

    def __call__(self, res, info):
        if res["lidar"].get("transform_type") == "feature":
            return self.voxelize_streaming_by_sweep(res, info)
        if self.nsectors > 1:
            if res.get("voxel_shape", "cylinder") == "cuboid":
                raise NotImplementedError("Cartesian sector streaming is outside the polar front end")
            return self.voxelize_streaming_polar(res, info)
        if not self.dynamic:
            return self.voxelize_hard(res, info)
        return self.voxelize_dynamic(res, info)
