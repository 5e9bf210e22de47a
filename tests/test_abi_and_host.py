"""CPU-side checks: the C-ABI library loads and exports every symbol include/*.h declares, the
host-only entry points behave, and the drop-in classes mirror the reference's construction-time
arithmetic.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle
from partner_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "polar_voxel_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pv_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.SO_PATH):
        pytest.skip("library not built yet (run __graft_entry__.build())")
    lib = ctypes.CDLL(_lib.SO_PATH)
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert set(names) == set(_lib.EXPORTS), "ctypes binding table and header disagree"
    # ... and nothing else: every unmangled pv_* export of the library is declared in the header
    import shutil
    import subprocess
    if shutil.which("nm"):
        out = subprocess.run(["nm", "-D", "--defined-only", _lib.SO_PATH], capture_output=True, text=True).stdout
        exported = {l.split()[-1] for l in out.splitlines() if " T pv_" in l}
        assert exported == set(names), exported ^ set(names)


def test_host_only_entry_points():
    lib = _lib.load()
    assert lib.pv_version() >= 200
    assert lib.pv_error_string(0) == b"ok"
    assert b"frame_capacity" in lib.pv_error_string(-5)
    from partner_b200.functional import make_config
    cfg, vs, rng, grid = make_config([0.098, 0.0123, 8], [0.3, -3.1488, -5, 50.476, 3.1488, 3], 20, 60000)
    assert lib.pv_workspace_bytes(cfg, 2_400_000, 8, 300_000, 7) > 0
    bad = _lib.PvConfig()
    assert lib.pv_workspace_bytes(bad, 1000, 1, 1000, 7) == 0            # zero grid
    cfg.max_points = 0
    assert lib.pv_workspace_bytes(cfg, 1000, 1, 1000, 7) == 0
    assert lib.pv_scatter_workspace_bytes(2, 512, 512) == 2 * 512 * 512 * 4
    with pytest.raises(ValueError):
        _lib.check(-1, "x")
    with pytest.raises(RuntimeError):
        _lib.check(-4, "x")


@pytest.mark.parametrize("grid", sorted(synth.GRIDS))
def test_voxel_generator_constructor_arithmetic_matches_reference(grid):
    """voxel_generator.py:6-17: f32 rounding of size/range, grid = round((hi - lo) / vs)."""
    from partner_b200 import VoxelGenerator
    g = synth.GRIDS[grid]
    vg = VoxelGenerator(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
    ref = oracle.VoxelGenerator(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
    assert np.array_equal(vg.grid_size, ref.grid_size) and vg.grid_size.dtype == np.int64
    assert np.array_equal(vg.voxel_size, ref.voxel_size) and vg.voxel_size.dtype == np.float32
    assert np.array_equal(vg.point_cloud_range, ref.point_cloud_range)
    assert vg.max_num_points_per_voxel == g["max_points"]
    expect = {"NUSC-PILLAR": (512, 512, 1), "NUSC-CYL": (1024, 1024, 40), "WAYMO-PARTNER": (1152, 2048, 40)}[grid]
    assert tuple(vg.grid_size) == expect


def test_reader_state_dict_keys_match_reference():
    """Checkpoints saved by the reference load unchanged."""
    from partner_b200 import PillarFeatureNet
    net = PillarFeatureNet(7, (64, 128), False, (0.098, 0.0123, 8), (0.3, -3.1488, -5, 50.476, 3.1488, 3))
    keys = set(net.state_dict())
    for i, (u, k) in enumerate(((32, 12), (128, 64))):
        assert tuple(net.pfn_layers[i].linear.weight.shape) == (u, k)
        for name in ("linear.weight", "norm.weight", "norm.bias", "norm.running_mean", "norm.running_var",
                     "norm.num_batches_tracked"):
            assert f"pfn_layers.{i}.{name}" in keys
    assert net.pfn_layers[0].norm.eps == 1e-3 and net.pfn_layers[0].norm.momentum == 0.01
    assert abs(net.x_offset - (0.098 / 2 + 0.3)) < 1e-12


def test_cpu_tensors_are_rejected_not_silently_processed():
    import torch
    from partner_b200 import functional as F
    with pytest.raises(ValueError):
        F.vfe_mean(torch.zeros(2, 3, 4), torch.ones(2, dtype=torch.int32))
    with pytest.raises(ValueError):
        F.transform_points(torch.zeros(4, 5))


def test_synthetic_generator_reproduces_survey_statistics():
    f = synth.nusc_frame(0)
    assert f.dtype == np.float32 and f.shape[1] == 5 and 280_000 < f.shape[0] < 300_000
    w = synth.waymo_frame(0, nsweeps=1)
    assert w.shape == (169_600, 6)
    assert synth.waymo_frame(0, nsweeps=1, time_column=False).shape == (169_600, 5)


def test_new_entry_points_validate_arguments_on_the_host():
    """Argument checks of the dynamic / streaming entry points return before any CUDA call."""
    lib = _lib.load()
    from partner_b200.functional import make_config
    cfg, _, _, _ = make_config([0.098, 0.0123, 8], [0.3, -3.1488, -5, 50.476, 3.1488, 3], 20, 60000)
    N = ctypes.c_void_p(0)
    assert lib.pv_profile_pipeline(cfg) == 2                                  # direct map -> list-free
    forced, _, _, _ = make_config([0.098, 0.0123, 8], [0.3, -3.1488, -5, 50.476, 3.1488, 3], 20, 60000, pipeline=1)
    assert lib.pv_profile_pipeline(forced) == 1                               # pv_config.pipeline, no process state
    forced.pipeline = 7
    assert lib.pv_profile_pipeline(forced) == -1                              # bad configuration
    big, _, _, _ = make_config([0.065, 0.00307, 0.15], [0.3, -3.14368, -2.0, 75.18, 3.14368, 4.0], 5, 150000)
    assert lib.pv_profile_pipeline(big) == 1                                  # hash map -> list-based
    assert lib.pv_dynamic_voxelize(cfg, N, N, N, 1, 10, 5, 1, 100, 100, N, 0, N, N, N, N, N, N, N, N) == -2
    assert lib.pv_stream_workspace_bytes(1000, 0) == 0 and lib.pv_stream_workspace_bytes(1000, 65) == 0
    assert lib.pv_stream_workspace_bytes(1000, 4) > 0
    assert lib.pv_stream_sectors(cfg, N, 10, 4, 4, 3.1488, N, 0, N, N, N, N, N) == -2      # c < 5
    assert lib.pv_dynamic_pfn_workspace_bytes(1000, 100) > 0 and lib.pv_dynamic_pfn_workspace_bytes(-1, 1) == 0
    layers = (_lib.PvPfnLayer * 3)()
    assert lib.pv_dynamic_pfn(N, N, N, N, N, 10, 5, 7, 0, 15, 0.1, 0.1, 0.0, 0.0, layers, 3, N, 0, N, N) == -6   # 3 layers
    assert lib.pv_dynamic_pfn(N, N, N, N, N, 10, 0, 7, 0, 15, 0.1, 0.1, 0.0, 0.0, layers, 2, N, 0, N, N) == 0    # no voxels


def test_dynamic_reader_state_dict_keys_match_reference():
    """pillar_encoder.py:262-335: DynamicPFNet keeps the PFNLayer parameter names (norm included, unused)."""
    from partner_b200 import DynamicPFNet
    net = DynamicPFNet(num_input_features=7, num_filters=[64, 128], xyz_cluster=True, raz_cluster=True, xy_center=True,
                       ra_center=True, voxel_size=[0.098, 0.0123, 8], pc_range=[0.3, -3.1488, -5.0, 50.476, 3.1488, 3.0])
    assert tuple(net.pfn_layers[0].linear.weight.shape) == (32, 16)
    assert tuple(net.pfn_layers[1].linear.weight.shape) == (128, 64)
    keys = set(net.state_dict())
    for i in range(2):
        for name in ("linear.weight", "norm.weight", "norm.bias", "norm.running_mean", "norm.running_var"):
            assert f"pfn_layers.{i}.{name}" in keys
    assert net.voxel_shape == "cuboid"                 # the reference's default, which its configs rely on


@pytest.mark.parametrize("nsec", [1, 4, 8])
def test_sector_ground_truth_matches_reference_golden(nsec, golden_dir):
    """Training-time sector streaming, annotation side (host numpy): filter_gt + rotation into the first wedge
    vs the reference's own statements (voxelization.py:332-349, utils.py:11-27; tests/golden/make_golden_stream_train.py)."""
    import copy
    import os
    from partner_b200.voxelization import sector_annotations
    g = np.load(os.path.join(golden_dir, "stream_train.npz"))
    pc_range = g["pc_range"]
    interval = (pc_range[4] - pc_range[1]) / nsec
    n = g["gt_boxes"].shape[0]
    for i in range(nsec):
        cur = pc_range.copy()
        cur[1] = pc_range[1] + i * interval
        cur[4] = pc_range[1] + (i + 1) * interval
        res = {"mode": "train", "voxel_shape": "cylinder",
               "lidar": {"annotations": {"gt_boxes": g["gt_boxes"].copy(), "gt_names": g["gt_names"].copy(),
                                         "gt_index": np.arange(n)}}}
        cur_res = copy.deepcopy(res)
        sector_annotations(cur_res, cur, pc_range)
        ann = cur_res["lidar"]["annotations"]
        assert np.array_equal(ann["gt_index"], g[f"n{nsec}_s{i}_index"])
        assert np.array_equal(ann["gt_boxes"], g[f"n{nsec}_s{i}_boxes"])          # same numpy statements: bit for bit
        assert ann["gt_names"].shape[0] == ann["gt_boxes"].shape[0]


def test_out_buffers_are_validated():
    """`out=` reuses a previous call's tensors: buffers that do not fit the current call are refused on the host
    instead of being overrun on the device (functional._check_out)."""
    import torch
    from partner_b200 import functional as F
    dev = torch.device("cpu")
    r = F.VoxelBatch()
    r.coors = torch.empty((100, 4), dtype=torch.int32)
    r.num_points = torch.empty((100,), dtype=torch.int32)
    r.voxel_counts = torch.empty((2,), dtype=torch.int32)
    r.mean_feats = torch.empty((100, 7), dtype=torch.float32)
    r.canvas = None
    F._check_out(r, dev, 100, 2, {"mean_feats": ((100, 7), torch.float32)})
    F._check_out(r, dev, 60, 2, {"mean_feats": ((60, 7), torch.float32)})          # larger buffers are fine
    with pytest.raises(ValueError):
        F._check_out(r, dev, 101, 2, {})                                          # more rows than the buffers hold
    with pytest.raises(ValueError):
        F._check_out(r, dev, 100, 3, {})                                          # another batch size
    with pytest.raises(ValueError):
        F._check_out(r, dev, 100, 2, {"mean_feats": ((100, 8), torch.float32)})   # another channel count
    with pytest.raises(ValueError):
        F._check_out(r, dev, 100, 2, {"canvas": ((2, 7, 8, 8), torch.float32)})   # an output the old call did not produce
