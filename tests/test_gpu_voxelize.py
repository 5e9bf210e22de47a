"""GPU parity: CUDA voxelizer (through the C ABI) vs the oracle and the reference golden vectors.

Integer outputs, voxel order, per-voxel point lists and the padded voxels tensor are compared
bit-exactly (np.array_equal); mean features / canvas within rtol 1e-5, atol 1e-5*max|ref|.
"""
import glob
import os

import numpy as np
import pytest

import oracle
from partner_b200 import synth
from util import assert_close_fp32, densify

pytestmark = pytest.mark.gpu

CASES = sorted(os.path.basename(p)[6:-4] for p in
               glob.glob(os.path.join(os.path.dirname(__file__), "golden", "voxel_*.npz")))


def _vg(grid, max_points=None, max_voxels=None):
    from partner_b200 import VoxelGenerator
    g = synth.GRIDS[grid]
    return (VoxelGenerator(g["voxel_size"], g["range"], max_points or g["max_points"],
                           max_voxels or g["max_voxels"]),
            oracle.VoxelGenerator(g["voxel_size"], g["range"], max_points or g["max_points"],
                                  max_voxels or g["max_voxels"]))


def _compare_single(gpu_out, ref_out):
    names = ("voxels", "coors", "num_points", "pc_grid_ind", "density")
    for name, a, b in zip(names, gpu_out, ref_out):
        if b is None:
            assert a is None, name
            continue
        assert a.dtype == b.dtype, (name, a.dtype, b.dtype)
        assert a.shape == b.shape, (name, a.shape, b.shape)
        assert np.array_equal(a, b), "%s differs (%d mismatches)" % (name, int((a != b).sum()))


@pytest.mark.parametrize("case", CASES)
def test_generate_matches_reference_golden(case, golden_dir):
    """Drop-in VoxelGenerator.generate on the polar points the reference saw -> bit-exact."""
    from partner_b200 import VoxelGenerator
    g = np.load(os.path.join(golden_dir, f"voxel_{case}.npz"))
    vg = VoxelGenerator(g["voxel_size"], g["range"], int(g["max_points"]), int(g["max_voxels"]))
    assert np.array_equal(vg.grid_size, g["grid_size"])
    want_ind, want_den = "pc_grid_ind" in g.files, "den_idx" in g.files
    out = vg.generate(g["polar"], return_pc_grid_ind=want_ind, return_density=want_den)
    ref = (g["voxels"], g["coors"], g["num_points"], g["pc_grid_ind"] if want_ind else None,
           densify(g["den_idx"], g["den_val"], g["den_shape"]) if want_den else None)
    _compare_single(out, ref)


@pytest.mark.parametrize("grid,kind,kw", [
    ("NUSC-PILLAR", "nusc", {}),
    ("NUSC-CYL", "nusc", {}),
    ("WAYMO-PARTNER", "waymo", dict(nsweeps=1)),
    ("WAYMO-PARTNER", "waymo", dict(nsweeps=3)),
])
def test_full_size_frame_vs_oracle(grid, kind, kw):
    """BASELINE.json-sized frames (cap V binds on NUSC-PILLAR and Waymo 3-sweep)."""
    gpu, ref = _vg(grid)
    cart = synth.make_batch(kind, 2, 1, **kw)[0]
    polar = oracle.transform_points(cart)
    want_den = grid == "NUSC-PILLAR"          # dense density maps of the 3-D grids are 100s of MB
    out = gpu.generate(polar, return_pc_grid_ind=True, return_density=want_den)
    exp = ref.generate(polar, return_pc_grid_ind=True, return_density=want_den)
    _compare_single(out, exp)


def test_shuffled_points_and_small_caps():
    gpu, ref = _vg("NUSC-PILLAR", max_points=3, max_voxels=5000)
    cart = synth.make_batch("nusc", 3, 1, shuffle=True)[0][:120000]
    polar = oracle.transform_points(cart)
    _compare_single(gpu.generate(polar, return_pc_grid_ind=True, return_density=True),
                    ref.generate(polar, return_pc_grid_ind=True, return_density=True))


def test_max_voxels_override_and_t1():
    gpu, ref = _vg("NUSC-PILLAR", max_points=1, max_voxels=60000)
    polar = oracle.transform_points(synth.nusc_frame(11)[:50000])
    _compare_single(gpu.generate(polar, max_voxels=777), ref.generate(polar, max_voxels=777))


def test_heavy_cells_many_points_per_voxel():
    """Thousands of points in a handful of cells: exercises the T-smallest selection."""
    rng = np.random.default_rng(5)
    n = 40000
    polar = np.zeros((n, 7), np.float32)
    polar[:, 0] = rng.choice([1.0, 1.05, 7.3, 20.0], n) + rng.uniform(0, 0.01, n)
    polar[:, 1] = rng.choice([-1.0, 0.5], n)
    polar[:, 3:] = rng.normal(size=(n, 4))
    gpu, ref = _vg("NUSC-PILLAR", max_points=20)
    _compare_single(gpu.generate(polar, return_density=True), ref.generate(polar, return_density=True))


def test_edge_inputs():
    gpu, ref = _vg("NUSC-PILLAR")
    empty = np.zeros((0, 7), np.float32)
    v, c, n, ind, den = gpu.generate(empty, return_pc_grid_ind=True)
    assert v.shape == (0, 20, 7) and c.shape == (0, 3) and n.shape == (0,) and ind.shape == (0, 3)
    one = np.array([[10.0, 0.1, 0.0, 1, 2, 3, 4]], np.float32)
    _compare_single(gpu.generate(one, return_pc_grid_ind=True), ref.generate(one, return_pc_grid_ind=True))
    far = np.array([[1e9, 0.1, 0.0, 1, 2, 3, 4], [10.0, 9.0, 0.0, 0, 0, 0, 0], [np.inf, 0, 0, 0, 0, 0, 0],
                    [10.0, 0.0, -np.inf, 0, 0, 0, 0]], np.float32)
    _compare_single(gpu.generate(far, return_pc_grid_ind=True), ref.generate(far, return_pc_grid_ind=True))
    # NaN is undefined behaviour in the reference; both oracle and kernels drop it (bin clamped to 0)
    nan = np.array([[np.nan, 0.1, 0.0, 1, 2, 3, 4], [10.0, np.nan, 0.0, 0, 0, 0, 0],
                    [10.0, 0.1, np.nan, 0, 0, 0, 0], [10.0, 0.1, 0.0, 5, 5, 5, 5]], np.float32)
    _compare_single(gpu.generate(nan, return_pc_grid_ind=True), ref.generate(nan, return_pc_grid_ind=True))


def test_rejects_bad_inputs():
    gpu, _ = _vg("NUSC-PILLAR")
    with pytest.raises(ValueError):
        gpu.generate(np.zeros((4, 7), np.float64))
    with pytest.raises(ValueError):
        gpu.generate(np.zeros((4, 2), np.float32))
    from partner_b200 import VoxelGenerator
    with pytest.raises(ValueError):
        VoxelGenerator([0.1, 0.1, 0.1], [0, 0, 0, 1, 1, 1], 5, 10).generate(np.zeros((1, 40), np.float32))


def _oracle_batch(ref, polars, **kw):
    outs = [ref.generate(p, **kw) for p in polars]
    vox, coor, num, nv = oracle.collate([(o[0], o[1], o[2]) for o in outs])
    return vox, coor, num, nv, outs


def test_batch_collated_like_collate_kitti():
    """Several frames incl. an empty one: [SM,4] (b,z,y,x) coordinates, per-frame counts."""
    import torch
    gpu, ref = _vg("NUSC-PILLAR", max_voxels=9000)
    frames = [synth.nusc_frame(21)[:70000], np.zeros((0, 5), np.float32), synth.nusc_frame(22)[:30000],
              synth.nusc_frame(23)[:500]]
    polars = [oracle.transform_points(f) for f in frames]
    vox, coor, num, nv, outs = _oracle_batch(ref, polars, return_pc_grid_ind=True, return_density=True)
    got = gpu.generate_batch(polars, return_pc_grid_ind=True, return_density=True, return_mean=True)
    assert np.array_equal(got["num_voxels"].numpy(), nv)
    assert got["num_voxels"].dtype == torch.int64
    assert np.array_equal(got["coordinates"].cpu().numpy(), coor)
    assert np.array_equal(got["num_points"].cpu().numpy(), num)
    assert np.array_equal(got["voxels"].cpu().numpy(), vox)
    assert np.array_equal(got["pc_grid_ind"].cpu().numpy(), np.concatenate([o[3] for o in outs]))
    assert np.array_equal(got["n_points"].cpu().numpy(), np.stack([o[4] for o in outs]))
    assert_close_fp32(got["mean_features"].cpu().numpy(), oracle.vfe_mean(vox, num), "mean_features")


@pytest.mark.parametrize("grid,kind,kw,nf", [
    ("NUSC-PILLAR", "nusc", {}, 3),
    ("WAYMO-PARTNER", "waymo", dict(nsweeps=1, time_column=True), 2),
    ("NUSC-CYL", "nusc", {}, 2),
])
def test_fused_cartesian_path_bit_exact(grid, kind, kw, nf):
    """Cartesian input, transform fused in the kernels: every integer output and the gathered
    voxels equal oracle.transform_points + oracle voxelizer, because phi is the same fixed
    IEEE op sequence on both sides."""
    gpu, ref = _vg(grid)
    frames = synth.make_batch(kind, 4, nf, **kw)
    polars = [oracle.transform_points(f) for f in frames]
    vox, coor, num, nv, outs = _oracle_batch(ref, polars, return_pc_grid_ind=True)
    got = gpu.generate_batch(frames, cartesian=True, return_pc_grid_ind=True, return_mean=True)
    assert np.array_equal(got["num_voxels"].numpy(), nv)
    assert np.array_equal(got["coordinates"].cpu().numpy(), coor)
    assert np.array_equal(got["num_points"].cpu().numpy(), num)
    assert np.array_equal(got["pc_grid_ind"].cpu().numpy(), np.concatenate([o[3] for o in outs]))
    assert np.array_equal(got["voxels"].cpu().numpy(), vox)
    assert_close_fp32(got["mean_features"].cpu().numpy(), oracle.vfe_mean(vox, num), "mean_features")


def test_transform_points_kernel(golden_dir):
    import torch
    from partner_b200 import transform_points
    from util import ulp_diff
    g = np.load(os.path.join(golden_dir, "transform.npz"))
    cart = torch.from_numpy(g["cart"]).cuda()
    for shape in ("cylinder", "cuboid"):
        out = transform_points(cart, shape).cpu().numpy()
        assert np.array_equal(out, oracle.transform_points(g["cart"], shape))      # bit-exact vs oracle
        ref = g[shape]                                                             # numpy reference
        phi = 1 if shape == "cylinder" else ref.shape[1] - 1
        other = [k for k in range(ref.shape[1]) if k != phi]
        assert np.array_equal(out[:, other], ref[:, other])
        assert ulp_diff(out[:, phi], ref[:, phi]).max() <= 4


def test_fused_front_end_mean_canvas():
    """PolarFrontEnd: Cartesian batch -> coors/num_points/mean features/canvas in one pass."""
    from partner_b200 import PolarFrontEnd
    g = synth.GRIDS["NUSC-PILLAR"]
    fe = PolarFrontEnd(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
    ref = oracle.VoxelGenerator(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
    frames = synth.make_batch("nusc", 2, 3)
    polars = [oracle.transform_points(f) for f in frames]
    vox, coor, num, nv, _ = _oracle_batch(ref, polars)
    feats = oracle.vfe_mean(vox, num)
    canvas, bev = oracle.scatter(feats, coor, len(frames), [512, 512, 1])
    got = fe(frames)
    assert np.array_equal(got["num_voxels"], nv)
    assert np.array_equal(got["coordinates"], coor)
    assert np.array_equal(got["num_points"], num)
    assert_close_fp32(got["features"], feats, "features")
    assert_close_fp32(got["canvas"], canvas, "canvas")
    # BEV index map: the canvas is non-zero exactly where the oracle scattered a voxel
    nz_ref = canvas[:, 0] != 0          # channel 0 = mean rho > 0 for every voxel
    assert np.array_equal(got["canvas"][:, 0] != 0, nz_ref)
    # the canvas holds exactly the feature rows of the same call
    c2, _ = oracle.scatter(got["features"], coor, len(frames), [512, 512, 1])
    assert np.array_equal(got["canvas"], c2)
    # run to run: integer outputs are bit-reproducible; the means are sums of the same addends in
    # an unspecified order (fp32 reductions at L2), so they agree to rounding
    again = fe(frames)
    for k in ("num_voxels", "coordinates", "num_points"):
        assert np.array_equal(got[k], again[k]), k
    assert_close_fp32(again["features"], got["features"], "features, second run")
    assert_close_fp32(again["canvas"], got["canvas"], "canvas, second run")


def test_cuda_graph_replay_matches_eager():
    import torch
    from partner_b200 import PolarFrontEnd
    g = synth.GRIDS["NUSC-PILLAR"]
    fe = PolarFrontEnd(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
    frames = synth.make_batch("nusc", 5, 2)
    sizes = [f.shape[0] for f in frames]
    pts = torch.from_numpy(np.concatenate(frames)).cuda()
    off = torch.tensor([0, sizes[0], sizes[0] + sizes[1]], dtype=torch.int32).cuda()
    eager = fe.forward_device(pts, off, 2, max(sizes))
    torch.cuda.synchronize()
    e = {k: getattr(eager, k).clone() for k in ("coors", "num_points", "voxel_counts", "mean_feats", "canvas")}
    out = fe.capture(pts, off, 2, max(sizes))
    for _ in range(3):
        fe.replay()
    torch.cuda.synchronize()
    m = int(e["voxel_counts"].sum())
    assert torch.equal(out.voxel_counts, e["voxel_counts"])
    assert torch.equal(out.coors[:m], e["coors"][:m])
    assert torch.equal(out.num_points[:m], e["num_points"][:m])
    assert_close_fp32(out.mean_feats[:m].cpu().numpy(), e["mean_feats"][:m].cpu().numpy(), "mean_feats")
    assert_close_fp32(out.canvas.cpu().numpy(), e["canvas"].cpu().numpy(), "canvas")


def test_hash_mode_pillar_canvas():
    """A pillar grid too large for the direct map (2048x2048 cells) -> hash map + canvas lookup."""
    from partner_b200 import PolarFrontEnd
    rng_ = [0.3, -3.1488, -5.0, 50.476, 3.1488, 3.0]
    vs = [0.0245, 0.003075, 8.0]
    fe = PolarFrontEnd(vs, rng_, 8, 40000)
    assert tuple(fe.grid_size) == (2048, 2048, 1)
    ref = oracle.VoxelGenerator(vs, rng_, 8, 40000)
    frames = [synth.nusc_frame(31)[:100000], synth.nusc_frame(32)[:60000]]
    polars = [oracle.transform_points(f) for f in frames]
    vox, coor, num, nv, _ = _oracle_batch(ref, polars)
    feats = oracle.vfe_mean(vox, num)
    canvas, _ = oracle.scatter(feats, coor, 2, [2048, 2048, 1])
    got = fe(frames)
    assert np.array_equal(got["num_voxels"], nv)
    assert np.array_equal(got["coordinates"], coor)
    assert np.array_equal(got["num_points"], num)
    assert_close_fp32(got["features"], feats, "features")
    assert_close_fp32(got["canvas"], canvas, "canvas")


# ---------------------------------------------------------------------------------------------
# list-free pipeline (pv_voxelize with voxels == NULL, fused.cu): same integer contract, means
# within the float gate
# ---------------------------------------------------------------------------------------------
def _list_free(grid, frames_polar, max_points=None, max_voxels=None, want_density=False, cartesian=False,
               frames_in=None):
    """Runs pv_voxelize without the padded voxels tensor; returns numpy outputs."""
    import torch
    from partner_b200 import functional as F
    g = synth.GRIDS[grid]
    T, V = max_points or g["max_points"], max_voxels or g["max_voxels"]
    cfg, _, _, gs = F.make_config(g["voxel_size"], g["range"], T, V)
    src = frames_in if frames_in is not None else frames_polar
    sizes = [f.shape[0] for f in src]
    off = np.zeros(len(src) + 1, np.int32)
    np.cumsum(sizes, out=off[1:])
    c = src[0].shape[1]
    pts = torch.from_numpy(np.concatenate(src) if sum(sizes) else np.zeros((0, c), np.float32)).cuda()
    vb = F.voxelize(cfg, pts, torch.from_numpy(off).cuda(), len(src), max(sizes + [1]), cartesian,
                    want_voxels=False, want_mean=True, want_grid_ind=True, want_density=want_density)
    F.read_status(vb)
    m = vb.total()
    return dict(coors=vb.coors[:m].cpu().numpy(), num=vb.num_points[:m].cpu().numpy(),
                nv=vb.voxel_counts.cpu().numpy(), mean=vb.mean_feats[:m].cpu().numpy(),
                ind=vb.pc_grid_ind.cpu().numpy(),
                den=vb.density.cpu().numpy() if want_density else None)


@pytest.fixture(autouse=True, params=["auto", "list-free"])
def pipeline(request):
    """Every test runs under the production pipeline choice and with the list-free pipeline forced
    (hash-map grids default to the list-based one): pv_config.pipeline of every configuration the
    test builds, through the host-side default of functional.make_config."""
    from partner_b200 import functional as F
    F.set_default_pipeline(2 if request.param == "list-free" else 0)
    yield request.param
    F.set_default_pipeline(0)


def _check_list_free(grid, polars, max_points=None, max_voxels=None, want_density=False, **kw):
    g = synth.GRIDS[grid]
    ref = oracle.VoxelGenerator(g["voxel_size"], g["range"], max_points or g["max_points"],
                                max_voxels or g["max_voxels"])
    vox, coor, num, nv, outs = _oracle_batch(ref, polars, return_pc_grid_ind=True, return_density=want_density)
    got = _list_free(grid, polars, max_points, max_voxels, want_density, **kw)
    assert np.array_equal(got["nv"], nv)
    assert np.array_equal(got["coors"], coor)
    assert np.array_equal(got["num"], num)
    assert np.array_equal(got["ind"], np.concatenate([o[3] for o in outs]))
    if want_density:
        assert np.array_equal(got["den"], np.stack([o[4] for o in outs]))
    assert_close_fp32(got["mean"], oracle.vfe_mean(vox, num), "mean_features")


@pytest.mark.parametrize("grid,kind,kw,cap", [
    ("NUSC-PILLAR", "nusc", {}, None),
    ("NUSC-PILLAR", "nusc", {}, 9000),
    ("NUSC-CYL", "nusc", {}, None),
    ("WAYMO-PARTNER", "waymo", dict(nsweeps=1), None),
    ("WAYMO-PARTNER", "waymo", dict(nsweeps=3), None),
])
def test_list_free_full_size_batches(grid, kind, kw, cap):
    """Two BASELINE-sized frames + an empty one, polar input, on every grid (direct map and hash)."""
    cart = synth.make_batch(kind, 2, 2, **kw)
    polars = [oracle.transform_points(cart[0]), np.zeros((0, cart[0].shape[1] + 2), np.float32),
              oracle.transform_points(cart[1])]
    _check_list_free(grid, polars, max_voxels=cap, want_density=(grid == "NUSC-PILLAR"))


def test_list_free_cartesian_input():
    """Fused cylinder transform: integer outputs bit-exact vs the oracle on the oracle's polar points."""
    cart = synth.make_batch("waymo", 4, 2, nsweeps=1)
    polars = [oracle.transform_points(f) for f in cart]
    _check_list_free("WAYMO-PARTNER", polars, cartesian=True, frames_in=cart)
    cart = synth.make_batch("nusc", 2, 2)
    polars = [oracle.transform_points(f) for f in cart]
    _check_list_free("NUSC-PILLAR", polars, cartesian=True, frames_in=cart, want_density=True)


def test_list_free_heavy_cells_shuffled_and_small_caps():
    """T-smallest selection for cells far above T; T = 1; shuffled order; both caps binding."""
    rng = np.random.default_rng(5)
    n = 40000
    polar = np.zeros((n, 7), np.float32)
    polar[:, 0] = rng.choice([1.0, 1.05, 7.3, 20.0], n) + rng.uniform(0, 0.01, n)
    polar[:, 1] = rng.choice([-1.0, 0.5], n)
    polar[:, 3:] = rng.normal(size=(n, 4))
    _check_list_free("NUSC-PILLAR", [polar], max_points=20, want_density=True)
    _check_list_free("NUSC-PILLAR", [polar, polar[::-1].copy()], max_points=1)
    cart = synth.make_batch("nusc", 3, 1, shuffle=True)[0][:120000]
    _check_list_free("NUSC-PILLAR", [oracle.transform_points(cart)], max_points=3, max_voxels=5000, want_density=True)
    _check_list_free("WAYMO-PARTNER", [oracle.transform_points(synth.waymo_frame(9, nsweeps=3))], max_points=2)


def test_list_free_edge_inputs():
    one = np.array([[10.0, 0.1, 0.0, 1, 2, 3, 4]], np.float32)
    far = np.array([[1e9, 0.1, 0.0, 1, 2, 3, 4], [10.0, 9.0, 0.0, 0, 0, 0, 0], [np.inf, 0, 0, 0, 0, 0, 0],
                    [10.0, 0.0, -np.inf, 0, 0, 0, 0]], np.float32)
    nan = np.array([[np.nan, 0.1, 0.0, 1, 2, 3, 4], [10.0, np.nan, 0.0, 0, 0, 0, 0],
                    [10.0, 0.1, np.nan, 0, 0, 0, 0], [10.0, 0.1, 0.0, 5, 5, 5, 5]], np.float32)
    empty = np.zeros((0, 7), np.float32)
    _check_list_free("NUSC-PILLAR", [one, far, empty, nan, one], want_density=True)
    _check_list_free("NUSC-CYL", [empty])
    # odd channel counts: 3 (xyz only), 9 and 14 channels go through the generic insert kernel
    rng = np.random.default_rng(3)
    for c in (3, 9, 14):
        p = np.zeros((5000, c), np.float32)
        p[:, 0] = rng.uniform(0.2, 52.0, 5000)
        p[:, 1] = rng.uniform(-3.2, 3.2, 5000)
        p[:, 2] = rng.uniform(-6.0, 4.0, 5000)
        p[:, 3:] = rng.normal(size=(5000, c - 3))
        _check_list_free("NUSC-PILLAR", [p, p[:777]], max_points=4)
        _check_list_free("WAYMO-PARTNER", [p], max_points=4)
