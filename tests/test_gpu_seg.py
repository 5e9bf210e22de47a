"""GPU parity of the segmentation row through the C ABI: pv_seg_voxel_labels
(Voxelization.get_grid_ind train branch + AssignLabel.assign_voxel_labels) and pv_seg_gather_points
(SegHead.predict) vs the reference's own outputs (tests/golden/seg.npz) and vs the oracle on
batches at full frame size.  All outputs are integers: bit-exact."""
import os

import numpy as np
import pytest

import oracle
from partner_b200 import synth
from test_oracle_golden import seg_pred_map
from util import densify

pytestmark = pytest.mark.gpu


def _cfg(grid):
    from partner_b200 import functional as F
    g = synth.GRIDS[grid]
    return F.make_config(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])[0], g


def _run(cfg, gi, lab, offsets):
    import torch
    from partner_b200 import functional as F
    dev = torch.device("cuda", 0)
    labels, valid, voff = F.seg_voxel_labels(cfg, torch.from_numpy(np.ascontiguousarray(gi, np.int32)).to(dev),
                                             torch.from_numpy(np.ascontiguousarray(lab, np.int32).reshape(-1)).to(dev),
                                             torch.from_numpy(np.asarray(offsets, np.int32)).to(dev), len(offsets) - 1)
    return labels, valid, voff


@pytest.mark.parametrize("case,grid", [("pillar", "NUSC-PILLAR"), ("cyl", "NUSC-CYL")])
def test_seg_labels_match_reference_golden(case, grid, golden_dir):
    import torch
    from partner_b200 import functional as F
    g = np.load(os.path.join(golden_dir, "seg.npz"))
    cfg, _ = _cfg(grid)
    gs = g[f"{case}_grid_size"]
    n = g[f"{case}_grid_ind"].shape[0]
    labels, valid, voff = _run(cfg, g[f"{case}_grid_ind"], g[f"{case}_label"], [0, n])
    want = densify(g[f"{case}_labels_nz_index"], g[f"{case}_labels_nz_value"], (1,) + tuple(gs[::-1]))
    assert labels.dtype == torch.int64 and np.array_equal(labels.cpu().numpy(), want)
    assert np.array_equal(valid.cpu().numpy(), g[f"{case}_valid_grid_ind"])
    assert voff.cpu().tolist() == [0, g[f"{case}_valid_grid_ind"].shape[0]]
    nx, ny, nz = (int(v) for v in gs)
    pred = torch.from_numpy(seg_pred_map(nz, ny, nx)).cuda()
    pred = pred if nz == 1 else pred[None]                   # [1, ny, nx] (2-D form) or [1, nz, ny, nx]
    got = F.seg_gather_points(pred.contiguous(), valid, voff)
    assert np.array_equal(got.cpu().numpy(), g[f"{case}_point_preds"])


@pytest.mark.parametrize("case", ["wrap_a", "wrap_b"])
def test_seg_labels_uint16_counter_wraps(case, golden_dir):
    from partner_b200 import functional as F
    g = np.load(os.path.join(golden_dir, "seg.npz"))
    cfg = F.make_config([1.0, 1.0, 8.0], [0, 0, -4, 8, 8, 4], 5, 100)[0]
    assert [int(v) for v in cfg.grid] == [8, 8, 1]
    labels, _, _ = _run(cfg, g[f"{case}_grid_ind"], g[f"{case}_label"], [0, g[f"{case}_grid_ind"].shape[0]])
    assert np.array_equal(labels.cpu().numpy(), g[f"{case}_labels"])


def test_seg_labels_batch_vs_oracle_full_size():
    """Batch of 3 full nuScenes frames (one empty frame in between) through the voxelizer's own
    pc_grid_ind; per frame the oracle is evaluated separately."""
    import torch
    from partner_b200 import functional as F
    cfg, g = _cfg("NUSC-PILLAR")
    frames = [oracle.transform_points(synth.nusc_frame(61)), np.zeros((0, 7), np.float32),
              oracle.transform_points(synth.nusc_frame(62, nsweeps=3))]
    sizes = [f.shape[0] for f in frames]
    off = np.zeros(len(frames) + 1, np.int32)
    np.cumsum(sizes, out=off[1:])
    pts = torch.from_numpy(np.concatenate(frames)).cuda()
    d_off = torch.from_numpy(off).cuda()
    vb = F.voxelize(cfg, pts, d_off, len(frames), max(sizes), False, want_grid_ind=True)
    rng = np.random.default_rng(5)
    gi = vb.pc_grid_ind.cpu().numpy()
    lab = ((gi[:, 1] // 29 + gi[:, 2] // 41) % 20).astype(np.int32)
    flip = rng.random(lab.shape[0]) < 0.3
    lab[flip] = rng.integers(0, 256, int(flip.sum()))
    lab[rng.random(lab.shape[0]) < 0.1] = -1
    labels, valid, voff = F.seg_voxel_labels(cfg, vb.pc_grid_ind, torch.from_numpy(lab).cuda(), d_off, len(frames))
    labels, valid, voff = labels.cpu().numpy(), valid.cpu().numpy(), voff.cpu().numpy()
    gs = oracle.grid_size(g["voxel_size"], g["range"])
    for b in range(len(frames)):
        lo, hi = off[b], off[b + 1]
        want_l, want_v = oracle.seg_voxel_labels(gi[lo:hi], lab[lo:hi], gs)
        assert np.array_equal(labels[b], want_l[0]), b
        assert np.array_equal(valid[voff[b]:voff[b + 1]], want_v), b
    assert voff[-1] == int((lab >= 0).sum())


def test_seg_labels_reject_bad_input():
    import torch
    from partner_b200 import functional as F
    cfg, _ = _cfg("NUSC-PILLAR")
    gi = torch.zeros((4, 3), dtype=torch.int32, device="cuda")
    off = torch.tensor([0, 4], dtype=torch.int32, device="cuda")
    with pytest.raises(ValueError):
        F.seg_voxel_labels(cfg, gi, torch.tensor([1, 2, 300, 4], dtype=torch.int32, device="cuda"), off, 1)
    bad = gi.clone()
    bad[2, 1] = 512                                           # y outside the 512 x 512 grid
    with pytest.raises(ValueError):
        F.seg_voxel_labels(cfg, bad, torch.tensor([1, 2, 3, 4], dtype=torch.int32, device="cuda"), off, 1)
    # all points unlabelled: empty result, all-zero map
    labels, valid, voff = F.seg_voxel_labels(cfg, gi, torch.full((4,), -1, dtype=torch.int32, device="cuda"), off, 1)
    assert valid.shape[0] == 0 and int(labels.abs().sum()) == 0 and voff.cpu().tolist() == [0, 0]


def test_voxelization_step_train_seg_labels():
    """The pipeline-step mirror in train mode: labels [1, nz, ny, nx] + valid_grid_ind like the reference (:42-58)."""
    from partner_b200 import Voxelization
    g = synth.GRIDS["NUSC-PILLAR"]
    cfg = dict(range=g["range"], voxel_size=g["voxel_size"], max_points_in_voxel=g["max_points"],
               max_voxel_num=[g["max_voxels"], g["max_voxels"]])
    step = Voxelization(cfg=cfg, super_tasks=["det", "seg"])
    polar = oracle.transform_points(synth.nusc_frame(63, nsweeps=2))
    rng = np.random.default_rng(3)
    label = rng.integers(-1, 17, (polar.shape[0], 1)).astype(np.int64)
    # training mode filters the ground truth first (voxelization.py:67-68), so the sample carries annotations:
    # three boxes, one of them beyond the 50.476 m range
    boxes = np.zeros((3, 9), np.float32)
    boxes[:, 0] = [10.0, 60.0, -20.0]
    boxes[:, 3:6] = 2.0
    res = {"mode": "train", "voxel_shape": "cylinder",
           "lidar": {"points": polar, "pc_label": label,
                     "annotations": {"gt_boxes": boxes, "gt_names": np.array(["car", "car", "bus"])}}}
    res, _ = step(res, {})
    assert res["lidar"]["annotations"]["gt_names"].tolist() == ["car", "bus"]          # utils.py:11-27
    assert np.array_equal(res["lidar"]["annotations"]["gt_boxes"][:, 0], [10.0, -20.0])
    vg = oracle.VoxelGenerator(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
    _, _, _, ind, _ = vg.generate(polar, return_pc_grid_ind=True)
    want_l, want_v = oracle.seg_voxel_labels(ind, label, vg.grid_size)
    assert np.array_equal(res["lidar"]["voxels"]["labels"], want_l)
    assert np.array_equal(res["lidar"]["voxels"]["valid_grid_ind"], want_v)
