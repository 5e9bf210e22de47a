"""Pins oracle/ (the C restatement) against vectors produced by the reference itself.

Golden files come from tests/golden/make_golden.py, which runs the reference's numba
voxelizer / numpy transform / torch reader modules in the build container.
"""
import glob
import os

import numpy as np
import pytest

import oracle
from util import assert_close_fp32, densify, ulp_diff

CASES = sorted(os.path.basename(p)[6:-4] for p in
               glob.glob(os.path.join(os.path.dirname(__file__), "golden", "voxel_*.npz")))


def test_golden_present():
    assert len(CASES) >= 6


@pytest.mark.parametrize("case", CASES)
def test_points_to_voxel_bit_exact(case, golden_dir):
    g = np.load(os.path.join(golden_dir, f"voxel_{case}.npz"))
    vg = oracle.VoxelGenerator(g["voxel_size"], g["range"], int(g["max_points"]), int(g["max_voxels"]))
    assert np.array_equal(vg.grid_size, g["grid_size"])
    want_ind = "pc_grid_ind" in g.files
    want_den = "den_idx" in g.files
    voxels, coors, num, ind, den = vg.generate(g["polar"], return_pc_grid_ind=want_ind,
                                               return_density=want_den)
    assert np.array_equal(coors, g["coors"])
    assert np.array_equal(num, g["num_points"])
    assert np.array_equal(voxels, g["voxels"])           # pure gather -> bit exact
    assert coors.dtype == np.int32 and num.dtype == np.int32 and voxels.dtype == np.float32
    if want_ind:
        assert np.array_equal(ind, g["pc_grid_ind"])
    else:
        assert ind is None
    if want_den:
        assert np.array_equal(den, densify(g["den_idx"], g["den_val"], g["den_shape"]))
    else:
        assert den is None


@pytest.mark.parametrize("case", CASES)
def test_transform_then_voxelize_matches_reference_bins(case, golden_dir):
    """rho is bit-exact; phi is the front end's own fixed-op-sequence atan2 (<= 4 ulp of numpy)."""
    g = np.load(os.path.join(golden_dir, f"voxel_{case}.npz"))
    polar = oracle.transform_points(g["cart"], "cylinder")
    ref = g["polar"]
    assert np.array_equal(polar[:, 0], ref[:, 0], equal_nan=True)
    assert np.array_equal(polar[:, 2:], ref[:, 2:], equal_nan=True)
    assert ulp_diff(polar[:, 1], ref[:, 1]).max() <= 4


def test_transform_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "transform.npz"))
    for shape in ("cylinder", "cuboid"):
        out = oracle.transform_points(g["cart"], shape)
        ref = g[shape]
        phi_col = 1 if shape == "cylinder" else ref.shape[1] - 1
        other = [k for k in range(ref.shape[1]) if k != phi_col]
        assert np.array_equal(out[:, other], ref[:, other])
        assert ulp_diff(out[:, phi_col], ref[:, phi_col]).max() <= 4


def test_atan2_special_values():
    f = oracle.lib().po_atan2f
    assert f(0.0, 0.0) == 0.0
    assert f(0.0, -0.0) == np.float32(np.pi)
    assert f(-0.0, -1.0) == -np.float32(np.pi)
    assert f(1.0, 0.0) == np.float32(np.pi / 2)
    assert f(np.inf, np.inf) == np.float32(np.pi / 4)
    assert f(np.inf, -np.inf) == np.float32(3 * np.pi / 4)
    assert np.isnan(f(np.nan, 1.0)) and np.isnan(f(1.0, np.nan))


def _layers(g, tag):
    layers = []
    i = 0
    while f"{tag}_w{i}" in g.files:
        layers.append(dict(weight=g[f"{tag}_w{i}"], mean=g[f"{tag}_mean{i}"], var=g[f"{tag}_var{i}"],
                           gamma=g[f"{tag}_gamma{i}"], beta=g[f"{tag}_beta{i}"]))
        i += 1
    return layers


def test_vfe_mean_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "readers.npz"))
    assert_close_fp32(oracle.vfe_mean(g["voxels"], g["num_points"]), g["vfe_mean"], "vfe_mean")


@pytest.mark.parametrize("tag,dist", [("pfn64_128", False), ("pfn64", False), ("pfn32_32_64_dist", True)])
def test_pfn_golden(tag, dist, golden_dir):
    g = np.load(os.path.join(golden_dir, "readers.npz"))
    out = oracle.pfn_forward(g["voxels"], g["num_points"], g["coors"], _layers(g, tag),
                             g["voxel_size"], g["pc_range"], with_distance=dist,
                             eps=float(g[f"{tag}_eps"]))
    assert_close_fp32(out, g[f"{tag}_out"], tag)


def test_scatter_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "readers.npz"))
    canvas, idx = oracle.scatter(g["vfe_mean"], g["coors"], 2, [512, 512, 1])
    assert np.array_equal(idx, g["bev_index"])
    assert np.array_equal(canvas, densify(g["canvas_idx"], g["canvas_val"], g["canvas_shape"]))


# ---- dynamic voxelization: golden from tests/golden/make_golden_dynamic.py ----
def test_dynamic_grid_ind_and_mean_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "dynamic.npz"))
    sizes = g["sizes"]
    off = np.concatenate([[0], np.cumsum(sizes)])
    gi = []
    for b in range(len(sizes)):
        zyx = oracle.dynamic_grid_ind(g["polar"][off[b]:off[b + 1]], g["voxel_size"], g["range"])
        gi.append(np.pad(zyx, ((0, 0), (1, 0)), constant_values=b))
    gi = np.concatenate(gi)
    assert np.array_equal(gi, g["grid_ind"])                         # bit-exact incl. the clamped border points
    mean, unq, inv, cnt = oracle.dynamic_mean(gi, g["polar"])
    assert np.array_equal(unq, g["unq"])
    assert np.array_equal(inv, g["unq_inv"])
    assert np.array_equal(cnt, g["unq_cnt"])
    assert_close_fp32(mean, g["features"], "scatter_mean")
    canvas, _ = oracle.scatter(mean, unq, len(sizes), [512, 512, 1])
    assert_close_fp32(canvas, densify(g["canvas_idx"], g["canvas_val"], g["canvas_shape"]), "canvas")


@pytest.mark.parametrize("tag,shape,flags", [
    ("pfn_cuboid_64_128", "cuboid", dict(xyz_cluster=True, raz_cluster=True, xy_center=True, ra_center=True)),
    ("pfn_cyl_64", "cylinder", dict(xyz_cluster=True, raz_cluster=True, xy_center=True, ra_center=True)),
    ("pfn_cyl_raz_32_64", "cylinder", dict(raz_cluster=True, ra_center=True))])
def test_dynamic_pfn_golden(tag, shape, flags, golden_dir):
    g = np.load(os.path.join(golden_dir, "dynamic.npz"))
    ws = []
    while f"{tag}_w{len(ws)}" in g.files:
        ws.append(g[f"{tag}_w{len(ws)}"])
    out = oracle.dynamic_pfn(g["polar"], g["unq_inv"], g["unq"], ws, g["voxel_size"], g["range"], shape, **flags)
    assert_close_fp32(out, g[f"{tag}_out"], tag)


# ---- azimuth-sector streaming: golden from tests/golden/make_golden_stream.py ----
@pytest.mark.parametrize("nsec", [1, 4, 8])
def test_stream_polar_golden(nsec, golden_dir):
    g = np.load(os.path.join(golden_dir, "stream.npz"))
    secs = oracle.stream_polar(g["polar"], g["voxel_size"], g["range"], nsec)
    assert len(secs) == nsec
    for i, (pts, gi, idx) in enumerate(secs):
        ref_pts = g[f"n{nsec}_s{i}_points"]
        assert np.array_equal(idx, g[f"n{nsec}_s{i}_index"])                  # which points, in which order
        assert np.array_equal(gi, g[f"n{nsec}_s{i}_grid_ind"])
        other = [k for k in range(pts.shape[1]) if k not in (3, 4)]
        assert np.array_equal(pts[:, other], ref_pts[:, other])                  # incl. the shifted azimuth: bit-exact
        # x, y = rho * cos / sin(phi): numpy's float32 cos / sin are vendor SIMD routines (not reproducible)
        assert np.allclose(pts[:, 3:5], ref_pts[:, 3:5], rtol=1e-6, atol=2e-5)


# ---- segmentation voxel labels, tests/golden/make_golden_seg.py ----
def seg_pred_map(nz, ny, nx):
    """The formula-defined prediction map of make_golden_seg.pred_map."""
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    return ((z * 7 + y * 3 + x * 5 + (x * y) // 11) % 16 + 1).astype(np.int64)


@pytest.mark.parametrize("case", ["pillar", "cyl"])
def test_seg_voxel_labels_match_reference(case, golden_dir):
    g = np.load(os.path.join(golden_dir, "seg.npz"))
    gs = g[f"{case}_grid_size"]
    labels, valid = oracle.seg_voxel_labels(g[f"{case}_grid_ind"], g[f"{case}_label"], gs)
    want = densify(g[f"{case}_labels_nz_index"], g[f"{case}_labels_nz_value"], (1,) + tuple(gs[::-1]))
    assert labels.dtype == np.int64 and np.array_equal(labels, want)
    assert np.array_equal(valid, g[f"{case}_valid_grid_ind"])
    nx, ny, nz = (int(v) for v in gs)
    pred = seg_pred_map(nz, ny, nx)
    got = oracle.seg_gather_points(pred[0] if nz == 1 else pred, valid)
    assert np.array_equal(got, g[f"{case}_point_preds"])


@pytest.mark.parametrize("case", ["wrap_a", "wrap_b"])
def test_seg_voxel_labels_uint16_counter_wraps(case, golden_dir):
    g = np.load(os.path.join(golden_dir, "seg.npz"))
    labels, _ = oracle.seg_voxel_labels(g[f"{case}_grid_ind"], g[f"{case}_label"], [8, 8, 1])
    assert np.array_equal(labels, g[f"{case}_labels"])
    assert int(labels[0, 0, 3, 4]) == (9 if case == "wrap_a" else 2)


def test_seg_voxel_labels_oracle_vs_numpy_vote():
    """Independent check of the C restatement: per-cell majority vote written with numpy (ties -> the
    smallest label, as np.argmax over the reference's 256-entry counter)."""
    rng = np.random.default_rng(11)
    gs = np.array([9, 7, 3])                                  # nx, ny, nz
    n = 5000
    gi = np.stack([rng.integers(0, gs[2], n), rng.integers(0, gs[1], n), rng.integers(0, gs[0], n)], 1).astype(np.int32)
    lab = rng.integers(-1, 6, n).astype(np.int32)             # few labels -> many ties
    labels, valid = oracle.seg_voxel_labels(gi, lab, gs)
    keep = lab >= 0
    assert np.array_equal(valid, gi[keep])
    want = np.zeros((1, gs[2], gs[1], gs[0]), np.int64)
    cells = (gi[keep, 0] * gs[1] + gi[keep, 1]) * gs[0] + gi[keep, 2]
    counts = np.zeros((gs.prod(), 256), np.int64)
    np.add.at(counts, (cells, lab[keep]), 1)
    want.reshape(-1)[:] = np.where(counts.sum(1) > 0, counts.argmax(1), 0)
    assert np.array_equal(labels, want)
