"""GPU parity of the dynamic-voxelization row through the C ABI:
pv_dynamic_voxelize vs the reference's own outputs (tests/golden/dynamic.npz, produced by
tests/golden/make_golden_dynamic.py) and vs the oracle at full frame size.  Integer outputs
(grid_ind, unq, unq_inv, unq_cnt, counts) bit-exact; means / canvas within the fp32 gate."""
import os

import numpy as np
import pytest

import oracle
from partner_b200 import synth
from util import assert_close_fp32, densify

pytestmark = pytest.mark.gpu


def _cfg(grid="NUSC-PILLAR"):
    from partner_b200 import functional as F
    g = synth.GRIDS[grid]
    return F.make_config(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])[0], g


def test_drop_in_reader_modules_match_reference_golden(golden_dir):
    """DynamicVoxelEncoderV1 / DynamicPPScatter on the reference's own points + grid_ind."""
    import torch
    from partner_b200.readers import DynamicVoxelEncoderV1, DynamicPPScatter
    g = np.load(os.path.join(golden_dir, "dynamic.npz"))
    pts = torch.from_numpy(g["polar"]).cuda()
    gi = torch.from_numpy(g["grid_ind"].astype(np.int64)).cuda()
    feats, unq = DynamicVoxelEncoderV1(num_input_features=7)(dict(points=pts, grid_ind=gi))
    assert unq.dtype == torch.int64
    assert np.array_equal(unq.cpu().numpy(), g["unq"])
    assert_close_fp32(feats.cpu().numpy(), g["features"], "features")
    canvas = DynamicPPScatter()(feats, unq, len(g["sizes"]), [512, 512, 1])
    assert_close_fp32(canvas.cpu().numpy(), densify(g["canvas_idx"], g["canvas_val"], g["canvas_shape"]), "canvas")


@pytest.mark.parametrize("mode", ["grid_ind", "polar", "cartesian"])
def test_dynamic_voxelize_matches_reference_golden(mode, golden_dir):
    import torch
    from partner_b200 import functional as F
    g = np.load(os.path.join(golden_dir, "dynamic.npz"))
    cfg, _ = _cfg()
    sizes = g["sizes"]
    off = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)).cuda()
    B = len(sizes)
    if mode == "grid_ind":
        r = F.dynamic_voxelize(cfg, torch.from_numpy(g["polar"]).cuda(), None, B, 0, False,
                               grid_ind=torch.from_numpy(g["grid_ind"]).cuda(), canvas=True)
    else:
        src = g["polar"] if mode == "polar" else g["cart"]
        r = F.dynamic_voxelize(cfg, torch.from_numpy(src).cuda(), off, B, int(sizes.max()), mode == "cartesian",
                               want_grid_ind=True, canvas=True)
        # phi of the fused Cartesian path is the library's own atan2 (<= 4 ulp from numpy's): compare
        # its bins with the oracle evaluated on the same definition, everything else with the golden
        gi_ref = g["grid_ind"]
        if mode == "cartesian":
            polar = oracle.transform_points(g["cart"])
            o = np.concatenate([[0], np.cumsum(sizes)])
            gi_ref = np.concatenate([np.pad(oracle.dynamic_grid_ind(polar[o[b]:o[b + 1]], g["voxel_size"], g["range"]),
                                            ((0, 0), (1, 0)), constant_values=b) for b in range(B)])
        assert np.array_equal(r.grid_ind.cpu().numpy(), gi_ref)
    F.read_status(r)
    m = r.total()
    if mode != "cartesian":
        assert m == g["unq"].shape[0]
        assert np.array_equal(r.unq[:m].cpu().numpy(), g["unq"])
        assert np.array_equal(r.unq_inv.cpu().numpy(), g["unq_inv"])
        assert np.array_equal(r.unq_cnt[:m].cpu().numpy(), g["unq_cnt"])
        assert_close_fp32(r.mean_feats[:m].cpu().numpy(), g["features"], "features")
        assert_close_fp32(r.canvas.cpu().numpy(), densify(g["canvas_idx"], g["canvas_val"], g["canvas_shape"]), "canvas")
    else:
        mean, unq, inv, cnt = oracle.dynamic_mean(gi_ref, polar)
        assert np.array_equal(r.unq[:m].cpu().numpy(), unq)
        assert np.array_equal(r.unq_inv.cpu().numpy(), inv)
        assert np.array_equal(r.unq_cnt[:m].cpu().numpy(), cnt)
        assert_close_fp32(r.mean_feats[:m].cpu().numpy(), mean, "features")
    counts = r.voxel_counts.cpu().numpy()
    assert counts.sum() == m and counts[2] == 0          # frame 2 of the golden batch is empty


def test_dynamic_full_size_batch_vs_oracle():
    """BASELINE-sized nuScenes frames, fused Cartesian input, run twice on the same workspace."""
    import torch
    from partner_b200 import functional as F
    cfg, g = _cfg()
    frames = synth.make_batch("nusc", 2, 3)
    sizes = [f.shape[0] for f in frames]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    polars = [oracle.transform_points(f) for f in frames]
    gi = np.concatenate([np.pad(oracle.dynamic_grid_ind(p, g["voxel_size"], g["range"]), ((0, 0), (1, 0)),
                                constant_values=b) for b, p in enumerate(polars)])
    mean, unq, inv, cnt = oracle.dynamic_mean(gi, np.concatenate(polars))
    canvas, _ = oracle.scatter(mean, unq, len(frames), [512, 512, 1])
    pts = torch.from_numpy(np.concatenate(frames)).cuda()
    for _ in range(2):
        r = F.dynamic_voxelize(cfg, pts, torch.from_numpy(off).cuda(), len(frames), max(sizes), True,
                               want_grid_ind=True, canvas=True)
        F.read_status(r)
        m = r.total()
        assert np.array_equal(r.grid_ind.cpu().numpy(), gi)
        assert np.array_equal(r.unq[:m].cpu().numpy(), unq)
        assert np.array_equal(r.unq_inv.cpu().numpy(), inv)
        assert np.array_equal(r.unq_cnt[:m].cpu().numpy(), cnt)
        assert_close_fp32(r.mean_feats[:m].cpu().numpy(), mean, "features")
        assert_close_fp32(r.canvas.cpu().numpy(), canvas, "canvas")


def test_dynamic_rejects_bad_rows_and_large_grids():
    import torch
    from partner_b200 import functional as F
    cfg, _ = _cfg()
    pts = torch.zeros((3, 7), dtype=torch.float32).cuda()
    gi = torch.tensor([[0, 0, 1, 1], [0, 0, 600, 2], [1, 0, 3, 3]], dtype=torch.int32).cuda()   # y = 600 outside 512
    r = F.dynamic_voxelize(cfg, pts, None, 2, 0, False, grid_ind=gi)
    assert r.unq_inv.cpu().tolist() == [0, -1, 1]
    with pytest.raises(ValueError):
        F.read_status(r)
    big, _ = _cfg("WAYMO-PARTNER")                         # 1152 x 2048 x 40 = 94 M cells: beyond the 2^26-cell bitmap
    with pytest.raises(ValueError):
        F.dynamic_voxelize(big, pts, torch.tensor([0, 3], dtype=torch.int32).cuda(), 1, 3, False)


PFN_CASES = [("pfn_cuboid_64_128", "cuboid", (64, 128), dict(xyz_cluster=True, raz_cluster=True, xy_center=True, ra_center=True)),
             ("pfn_cyl_64", "cylinder", (64,), dict(xyz_cluster=True, raz_cluster=True, xy_center=True, ra_center=True)),
             ("pfn_cyl_raz_32_64", "cylinder", (32, 64), dict(raz_cluster=True, ra_center=True))]


@pytest.mark.parametrize("tag,shape,filters,flags", PFN_CASES)
def test_dynamic_pfnet_matches_reference_golden(tag, shape, filters, flags, golden_dir):
    """Drop-in DynamicPFNet with the reference's weights (state_dict keys of the reference) on the
    reference's points + grid_ind -> the reference module's own output."""
    import torch
    from partner_b200.readers import DynamicPFNet
    g = np.load(os.path.join(golden_dir, "dynamic.npz"))
    gcfg = synth.GRIDS["NUSC-PILLAR"]
    net = DynamicPFNet(num_input_features=7, num_filters=filters, voxel_shape=shape, voxel_size=gcfg["voxel_size"],
                       pc_range=gcfg["range"], **flags)
    sd = net.state_dict()
    for i in range(len(filters)):
        key = "pfn_layers.%d.linear.weight" % i
        assert key in sd and tuple(sd[key].shape) == g[f"{tag}_w{i}"].shape
        sd[key] = torch.from_numpy(g[f"{tag}_w{i}"])
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    feats, unq = net(dict(points=torch.from_numpy(g["polar"]).cuda(),
                          grid_ind=torch.from_numpy(g["grid_ind"].astype(np.int64)).cuda()))
    assert np.array_equal(unq.cpu().numpy(), g["unq"])
    assert_close_fp32(feats.cpu().numpy(), g[f"{tag}_out"], tag)


def test_dynamic_pfnet_full_size_vs_oracle():
    """Whole pipeline at frame size: fused Cartesian voxelization -> DynamicPFNet (polarstream reader config)."""
    import torch
    from partner_b200 import functional as F
    cfg, g = _cfg()
    frames = synth.make_batch("nusc", 3, 2)
    sizes = [f.shape[0] for f in frames]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    polars = [oracle.transform_points(f) for f in frames]
    polar = np.concatenate(polars)
    gi = np.concatenate([np.pad(oracle.dynamic_grid_ind(p, g["voxel_size"], g["range"]), ((0, 0), (1, 0)),
                                constant_values=b) for b, p in enumerate(polars)])
    mean, unq, inv, cnt = oracle.dynamic_mean(gi, polar)
    rng = np.random.default_rng(0)
    ws = [rng.normal(0, 0.2, (32, 16)).astype(np.float32), rng.normal(0, 0.1, (128, 64)).astype(np.float32)]
    flags = dict(xyz_cluster=True, raz_cluster=True, xy_center=True, ra_center=True)
    ref = oracle.dynamic_pfn(polar, inv, unq, ws, g["voxel_size"], g["range"], "cuboid", **flags)
    pts = torch.from_numpy(np.concatenate(frames)).cuda()
    r = F.dynamic_voxelize(cfg, pts, torch.from_numpy(off).cuda(), len(frames), max(sizes), True)
    m = r.total()
    assert np.array_equal(r.unq[:m].cpu().numpy(), unq)
    polar_dev = F.transform_points(pts)                    # the rows the reader sees (bit-exact vs the oracle)
    vx, vy = g["voxel_size"][0], g["voxel_size"][1]
    out = F.dynamic_pfn(polar_dev, r, m, [torch.from_numpy(w).cuda() for w in ws], vx, vy,
                        vx / 2 + g["range"][0], vy / 2 + g["range"][1], False, True, True, True, True)
    assert_close_fp32(out.cpu().numpy(), ref, "dynamic pfn")


def test_voxelization_pipeline_step_hard_double_flip_and_dynamic():
    """partner_b200.Voxelization: the reference's res['lidar'] keys for hard voxelization with
    double-flip TTA (voxelization.py:62-144) and for dynamic voxelization (:148-181)."""
    from partner_b200 import Voxelization
    g = synth.GRIDS["NUSC-PILLAR"]
    cart = synth.nusc_frame(31)[:60000]
    polar = oracle.transform_points(cart)

    def flipped(fx, fy):
        c = cart.copy()
        if fy:
            c[:, 1] = -c[:, 1]
        if fx:
            c[:, 0] = -c[:, 0]
        return oracle.transform_points(c)
    cfg = dict(range=g["range"], voxel_size=g["voxel_size"], max_points_in_voxel=20, max_voxel_num=[3000, 5000],
               return_density=True, double_flip=True)
    step = Voxelization(cfg=cfg)
    res = dict(mode="val", lidar=dict(points=polar, yflip_points=flipped(False, True), xflip_points=flipped(True, False),
                                      double_flip_points=flipped(True, True), transform_type="point"))
    res, _ = step(res, {})
    ref = oracle.VoxelGenerator(g["voxel_size"], g["range"], 20, 3000)
    v, c, n, _, den = ref.generate(polar, max_voxels=5000, return_density=True)
    got = res["lidar"]["voxels"]
    assert np.array_equal(got["voxels"], v) and np.array_equal(got["coordinates"], c) and np.array_equal(got["num_points"], n)
    assert np.array_equal(got["n_points"], den) and got["num_voxels"].tolist() == [v.shape[0]]
    for key, pts in (("yflip", res["lidar"]["yflip_points"]), ("xflip", res["lidar"]["xflip_points"]),
                     ("double_flip", res["lidar"]["double_flip_points"])):
        v, c, n, _, _ = ref.generate(pts)                     # flipped copies: the generator's own cap (3000)
        got = res["lidar"][key + "_voxels"]
        assert np.array_equal(got["voxels"], v), key
        assert np.array_equal(got["coordinates"], c) and np.array_equal(got["num_points"], n), key
        assert got["num_voxels"].tolist() == [v.shape[0]]
    dyn = Voxelization(cfg=dict(cfg, dynamic=True, double_flip=False), super_tasks=["det", "seg"])
    res = dict(mode="val", lidar=dict(points=polar, n_key_points=1000, transform_type="point"))
    res, _ = dyn(res, {})
    gi = oracle.dynamic_grid_ind(polar, g["voxel_size"], g["range"])
    assert res["lidar"]["voxels"]["grid_ind"].dtype == np.int64
    assert np.array_equal(res["lidar"]["voxels"]["grid_ind"], gi)
    assert np.array_equal(res["lidar"]["voxels"]["valid_grid_ind"], gi[:1000])
    with pytest.raises(NotImplementedError):
        Voxelization(cfg=dict(cfg, nsectors=4, double_flip=False))(
            dict(mode="val", voxel_shape="cuboid", lidar=dict(points=polar, transform_type="point")), {})


@pytest.mark.parametrize("nsec", [1, 4, 8])
def test_stream_sectors_matches_reference_golden(nsec, golden_dir):
    """pv_stream_sectors vs the reference's own per-sector statements (tests/golden/stream.npz)."""
    import torch
    from partner_b200 import functional as F
    g = np.load(os.path.join(golden_dir, "stream.npz"))
    cfg, _ = _cfg()
    out, gi, idx, counts = F.stream_sectors(cfg, torch.from_numpy(g["polar"]).cuda(), nsec, float(g["range"][4]))
    counts = counts.cpu().numpy()
    offs = np.concatenate([[0], np.cumsum(counts)])
    out, gi, idx = out.cpu().numpy(), gi.cpu().numpy(), idx.cpu().numpy()
    for i in range(nsec):
        lo, hi = offs[i], offs[i + 1]
        ref_pts = g[f"n{nsec}_s{i}_points"]
        assert np.array_equal(idx[lo:hi], g[f"n{nsec}_s{i}_index"])                 # selection + stable order
        assert np.array_equal(gi[lo:hi], g[f"n{nsec}_s{i}_grid_ind"])
        other = [k for k in range(ref_pts.shape[1]) if k not in (3, 4)]
        assert np.array_equal(out[lo:hi][:, other], ref_pts[:, other])                 # shifted azimuth bit-exact
        assert np.allclose(out[lo:hi][:, 3:5], ref_pts[:, 3:5], rtol=1e-6, atol=2e-5)  # cos / sin: a few ulp


def test_voxelization_streaming_polar_vs_oracle():
    """Full-size frame through Voxelization(nsectors=4) with the seg keys; oracle as the checker."""
    from partner_b200 import Voxelization
    g = synth.GRIDS["NUSC-PILLAR"]
    polar = oracle.transform_points(synth.nusc_frame(33))
    cfg = dict(range=g["range"], voxel_size=g["voxel_size"], max_points_in_voxel=20, max_voxel_num=[30000, 60000],
               dynamic=True, nsectors=4)
    step = Voxelization(cfg=cfg, super_tasks=["det", "seg"])
    res, _ = step(dict(mode="val", voxel_shape="cylinder",
                       lidar=dict(points=polar, n_key_points=30000, transform_type="point")), {})
    secs = oracle.stream_polar(polar, g["voxel_size"], g["range"], 4)
    assert len(res["sectors"]) == 4
    for cur, (pts, gi, idx) in zip(res["sectors"], secs):
        got = cur["lidar"]["points"]
        other = [k for k in range(pts.shape[1]) if k not in (3, 4)]
        assert np.array_equal(got[:, other], pts[:, other])
        assert np.allclose(got[:, 3:5], pts[:, 3:5], rtol=1e-6, atol=2e-5)
        assert np.array_equal(cur["lidar"]["voxels"]["grid_ind"], gi)
        assert list(cur["lidar"]["voxels"]["shape"]) == [512, 128, 1]
        key = idx[idx < 30000]
        assert np.array_equal(cur["lidar"]["key_points_index"], key) and cur["lidar"]["n_key_points"] == len(key)
        assert np.array_equal(cur["lidar"]["voxels"]["valid_grid_ind"], gi[:len(key)])


def test_voxelization_sweep_streaming_bidirectional_vs_oracle():
    """transform_type == 'feature', cylinder branch of voxelize_streaming_by_sweep (voxelization.py:393-460):
    later sweeps + earlier sweeps warped back by one pose, each cut into azimuth sectors."""
    from partner_b200 import Voxelization
    g = synth.GRIDS["NUSC-PILLAR"]
    pts = synth.nusc_frame(35, nsweeps=4)
    counts = [int((pts[:, 4] == np.float32(0.05 * s)).sum()) for s in range(4)]
    assert sum(counts) == pts.shape[0]
    th = 0.03
    tm1 = np.array([[np.cos(th), -np.sin(th), 0, 0.4], [np.sin(th), np.cos(th), 0, -0.1], [0, 0, 1, 0.02], [0, 0, 0, 1]])
    mats = [np.eye(4), tm1, tm1 @ tm1, tm1 @ tm1 @ tm1]
    cfg = dict(range=g["range"], voxel_size=g["voxel_size"], max_points_in_voxel=20, max_voxel_num=[30000, 60000],
               dynamic=True, nsectors=4)
    step = Voxelization(cfg=cfg)
    res, _ = step(dict(mode="val", voxel_shape="cylinder",
                       lidar=dict(points=pts, npoints_sweep=counts, transform_matrices=mats, transform_type="feature")), {})
    assert res["nsweeps"] == 2 and res["nsectors"] == 4 and len(res["sweeps"]) == 8
    cs = np.insert(np.cumsum(counts), 0, 0)
    later = oracle.stream_polar(oracle.transform_points(np.ascontiguousarray(pts[:cs[3], :5])), g["voxel_size"], g["range"], 4)
    p = pts[cs[1]:].copy()
    p[:, -1] -= p[0, -1]
    tm = np.linalg.inv(mats[1])
    p[:, :3] = (np.hstack((p[:, :3], np.ones((p.shape[0], 1)))) @ tm.T)[:, :3]
    earlier = oracle.stream_polar(oracle.transform_points(np.ascontiguousarray(p[:, :5])), g["voxel_size"], g["range"], 4)
    assert np.allclose(res["sweeps"][0]["transform_matrix"], tm[:2, :2])
    for cur, (rp, gi, idx) in zip(res["sweeps"], earlier + later):
        got = cur["lidar"]["points"]
        assert got.shape == rp.shape
        assert np.allclose(got, rp, rtol=1e-6, atol=2e-5)          # float64 warp / cos / sin: last-bit differences
        assert (cur["lidar"]["voxels"]["grid_ind"] != gi).sum() <= 2   # a warped coordinate within an ulp of a bin edge


@pytest.mark.parametrize("tag,vs", [("NUSC-CYL 1024x1024x40", [0.049, 0.00615, 0.2]),
                                    ("seg cylinder 640x640x40", [0.0784, 0.00984, 0.2])])
def test_dynamic_voxelize_on_3d_cylinder_grids(tag, vs):
    """The reference's dynamic=True configs on 3-D cylinder grids (voxelnet_det_cylinder_singlehead.py:8-18,68-74:
    1024 x 1024 x 40; voxelnet_seg_cylinder.py: 640 x 640 x 40) -- tens of millions of cells per frame, so the
    rows live in the hash map while the torch.unique order still comes from the cell-order bitmap.  Two full
    nuScenes frames (fused Cartesian input) + the drop-in DynamicVoxelEncoderV1 on a caller-made grid_ind,
    twice on the same workspace; integers bit-exact vs the oracle."""
    import torch
    from partner_b200 import DynamicVoxelEncoderV1
    from partner_b200 import functional as F
    rng = [0.3, -3.1488, -5.0, 50.476, 3.1488, 3.0]
    cfg, _, _, grid = F.make_config(vs, rng, 30, 120000)
    assert int(grid[0]) * int(grid[1]) * int(grid[2]) > (1 << 20)
    frames = synth.make_batch("nusc", 2, 2)
    sizes = [f.shape[0] for f in frames]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    polars = [oracle.transform_points(f) for f in frames]
    gi = np.concatenate([np.pad(oracle.dynamic_grid_ind(p, vs, rng), ((0, 0), (1, 0)), constant_values=b)
                         for b, p in enumerate(polars)])
    mean, unq, inv, cnt = oracle.dynamic_mean(gi, np.concatenate(polars))
    pts = torch.from_numpy(np.concatenate(frames)).cuda()
    for _ in range(2):
        r = F.dynamic_voxelize(cfg, pts, torch.from_numpy(off).cuda(), len(frames), max(sizes), True, want_grid_ind=True)
        F.read_status(r)
        m = r.total()
        assert m == unq.shape[0], tag
        assert np.array_equal(r.grid_ind.cpu().numpy(), gi)
        assert np.array_equal(r.unq[:m].cpu().numpy(), unq)
        assert np.array_equal(r.unq_inv.cpu().numpy(), inv)
        assert np.array_equal(r.unq_cnt[:m].cpu().numpy(), cnt)
        assert_close_fp32(r.mean_feats[:m].cpu().numpy(), mean, "features")
    # binning only (Voxelization.voxelize_dynamic): any grid, no map
    gi_dev = F.dynamic_grid_ind(cfg, pts, torch.from_numpy(off).cuda(), len(frames), True)
    assert np.array_equal(gi_dev.cpu().numpy(), gi)
    # drop-in reader on the reference's dict (grid_ind made by the caller), grid size given or inferred
    polar_dev = torch.from_numpy(np.concatenate(polars)).cuda()
    for gs in ((int(grid[0]), int(grid[1]), int(grid[2])), None):
        enc = DynamicVoxelEncoderV1(num_input_features=7, grid_size=gs, batch_size=2 if gs else None)
        feats, u = enc(dict(points=polar_dev, grid_ind=torch.from_numpy(gi).cuda().long()))
        assert np.array_equal(u.cpu().numpy(), unq)
        assert_close_fp32(feats.cpu().numpy(), mean, "DynamicVoxelEncoderV1 features")


def test_dynamic_grid_ind_on_the_waymo_grid():
    """Binning only works where no map fits at all: the 1152 x 2048 x 40 Waymo grid (94 M cells)."""
    import torch
    from partner_b200 import functional as F
    g = synth.GRIDS["WAYMO-PARTNER"]
    cfg, _, _, _ = F.make_config(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
    f = synth.waymo_frame(77, nsweeps=1, time_column=True)
    polar = oracle.transform_points(f)
    ref = np.pad(oracle.dynamic_grid_ind(polar, g["voxel_size"], g["range"]), ((0, 0), (1, 0)))
    off = torch.tensor([0, f.shape[0]], dtype=torch.int32).cuda()
    assert np.array_equal(F.dynamic_grid_ind(cfg, torch.from_numpy(f).cuda(), off, 1, True).cpu().numpy(), ref)
    assert np.array_equal(F.dynamic_grid_ind(cfg, torch.from_numpy(polar).cuda(), off, 1, False).cpu().numpy(), ref)


def test_voxelization_streaming_polar_training_mode():
    """mode == 'train' through Voxelization(nsectors=4): per-sector ground truth (filtered + rotated into the
    first wedge, voxelization.py:332-349), per-sector point labels (:374-377) and the voxel labels of
    get_grid_ind / assign_voxel_labels, checked against the oracle composition; and the sweep-streaming variant
    hands the labels of the later sweeps to the earlier ones (:451-453)."""
    import copy
    from partner_b200 import Voxelization
    from partner_b200.voxelization import sector_annotations
    g = synth.GRIDS["NUSC-PILLAR"]
    cart = synth.nusc_frame(36, nsweeps=3)
    polar = oracle.transform_points(cart)
    rng = np.random.default_rng(5)
    labels = rng.integers(-1, 17, (polar.shape[0], 1)).astype(np.int64)          # -1 = unlabelled (dropped)
    nb = 40
    rho, az = rng.uniform(0, 60, nb), rng.uniform(-np.pi, np.pi, nb)
    boxes = np.zeros((nb, 9), np.float32)
    boxes[:, 0], boxes[:, 1], boxes[:, 3:6], boxes[:, 8] = rho * np.cos(az), rho * np.sin(az), 2.0, rng.uniform(-3, 3, nb)
    boxes[:, 6:8] = rng.normal(0, 2, (nb, 2))
    ann = dict(gt_boxes=boxes, gt_names=np.array(["car"] * nb))
    cfg = dict(range=g["range"], voxel_size=g["voxel_size"], max_points_in_voxel=20, max_voxel_num=[30000, 60000],
               dynamic=True, nsectors=4)
    step = Voxelization(cfg=cfg, super_tasks=["det", "seg"])
    res_in = dict(mode="train", voxel_shape="cylinder",
                  lidar=dict(points=polar, pc_label=labels, annotations=copy.deepcopy(ann), transform_type="point"))
    res, _ = step(copy.deepcopy(res_in), {})
    secs = oracle.stream_polar(polar, g["voxel_size"], g["range"], 4)
    pc_range = step.voxel_generator.point_cloud_range
    interval = (pc_range[4] - pc_range[1]) / 4
    total_boxes = 0
    for i, (cur, (pts, gi, idx)) in enumerate(zip(res["sectors"], secs)):
        assert np.array_equal(cur["lidar"]["voxels"]["grid_ind"], gi)
        assert np.array_equal(cur["lidar"]["pc_label"], labels[idx])
        vl, valid = oracle.seg_voxel_labels(gi, labels[idx].reshape(-1), [512, 128, 1])
        assert np.array_equal(cur["lidar"]["voxels"]["labels"].reshape(vl.shape), vl)
        assert np.array_equal(cur["lidar"]["voxels"]["valid_grid_ind"], valid)
        ref = {"mode": "train", "voxel_shape": "cylinder", "lidar": {"annotations": copy.deepcopy(ann)}}
        cr = pc_range.copy()
        cr[1], cr[4] = pc_range[1] + i * interval, pc_range[1] + (i + 1) * interval
        sector_annotations(ref, cr, pc_range)
        assert np.array_equal(cur["lidar"]["annotations"]["gt_boxes"], ref["lidar"]["annotations"]["gt_boxes"])
        total_boxes += cur["lidar"]["annotations"]["gt_boxes"].shape[0]
    assert 0 < total_boxes <= nb
    # sweep streaming in training mode
    counts = [int((cart[:, 4] == np.float32(0.05 * s)).sum()) for s in range(3)]
    th = 0.02
    tm1 = np.array([[np.cos(th), -np.sin(th), 0, 0.3], [np.sin(th), np.cos(th), 0, 0.1], [0, 0, 1, 0], [0, 0, 0, 1]])
    sw, _ = step(dict(mode="train", voxel_shape="cylinder",
                      lidar=dict(points=cart, pc_label=labels, annotations=copy.deepcopy(ann), npoints_sweep=counts,
                                 transform_matrices=[np.eye(4), tm1, tm1 @ tm1], transform_type="feature")), {})
    assert sw["nsweeps"] == 2 and len(sw["sweeps"]) == 8
    for i in range(4):
        assert np.array_equal(sw["sweeps"][i]["lidar"]["voxels"]["labels"], sw["sweeps"][4 + i]["lidar"]["voxels"]["labels"])
        assert "valid_grid_ind" in sw["sweeps"][4 + i]["lidar"]["voxels"]
