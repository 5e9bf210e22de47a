"""GPU parity of the reader / scatter drop-ins vs the reference's own torch modules (golden
vectors) and the oracle.  Float gate: rtol 1e-5, atol 1e-5 * max|ref|."""
import os

import numpy as np
import pytest

import oracle
from util import assert_close_fp32, densify

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "readers.npz"))


def _cuda(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_vfe_v3_matches_reference(g):
    from partner_b200 import VoxelFeatureExtractorV3
    net = VoxelFeatureExtractorV3(num_input_features=7)
    out = net(_cuda(g["voxels"]), _cuda(g["num_points"])).cpu().numpy()
    assert_close_fp32(out, g["vfe_mean"], "vfe vs reference")
    assert_close_fp32(out, oracle.vfe_mean(g["voxels"], g["num_points"]), "vfe vs oracle")


def _load_pfn(g, tag, filters, dist):
    import torch
    from partner_b200 import PillarFeatureNet
    net = PillarFeatureNet(7, filters, dist, tuple(g["voxel_size"]), tuple(g["pc_range"]))
    sd = {}
    for i in range(len(filters)):
        sd[f"pfn_layers.{i}.linear.weight"] = torch.from_numpy(g[f"{tag}_w{i}"])
        sd[f"pfn_layers.{i}.norm.running_mean"] = torch.from_numpy(g[f"{tag}_mean{i}"])
        sd[f"pfn_layers.{i}.norm.running_var"] = torch.from_numpy(g[f"{tag}_var{i}"])
        sd[f"pfn_layers.{i}.norm.weight"] = torch.from_numpy(g[f"{tag}_gamma{i}"])
        sd[f"pfn_layers.{i}.norm.bias"] = torch.from_numpy(g[f"{tag}_beta{i}"])
        sd[f"pfn_layers.{i}.norm.num_batches_tracked"] = torch.tensor(0)
    net.load_state_dict(sd, strict=True)        # reference checkpoint keys load unchanged
    return net.cuda().eval()


@pytest.mark.parametrize("tag,filters,dist", [("pfn64_128", (64, 128), False), ("pfn64", (64,), False),
                                              ("pfn32_32_64_dist", (32, 32, 64), True)])
def test_pillar_feature_net_matches_reference(g, tag, filters, dist):
    net = _load_pfn(g, tag, filters, dist)
    out = net(_cuda(g["voxels"]), _cuda(g["num_points"]), _cuda(g["coors"])).cpu().numpy()
    assert_close_fp32(out, g[f"{tag}_out"], tag + " vs reference")


@pytest.mark.parametrize("tag,filters", [("pfn64_128", (64, 128))])
def test_pillar_feature_net_tensor_core_path_matches_reference(g, tag, filters, monkeypatch):
    """Second layer on tcgen05 (3xTF32, accumulators in TMEM): same 1e-5 gate as the fp32 kernel."""
    monkeypatch.setenv("PV_PFN_TC", "1")
    net = _load_pfn(g, tag, filters, False)
    out = net(_cuda(g["voxels"]), _cuda(g["num_points"]), _cuda(g["coors"])).cpu().numpy()
    assert_close_fp32(out, g[f"{tag}_out"], tag + " (tcgen05) vs reference")
    monkeypatch.setenv("PV_PFN_TC", "0")
    ref = net(_cuda(g["voxels"]), _cuda(g["num_points"]), _cuda(g["coors"])).cpu().numpy()
    assert_close_fp32(out, ref, "tcgen05 vs fp32 kernel")


def test_pfn_training_mode_is_refused(g):
    net = _load_pfn(g, "pfn64", (64,), False).train()
    with pytest.raises(RuntimeError):
        net(_cuda(g["voxels"][:4]), _cuda(g["num_points"][:4]), _cuda(g["coors"][:4]))


def test_pfn_single_voxel_squeeze(g):
    net = _load_pfn(g, "pfn64", (64,), False)
    out = net(_cuda(g["voxels"][:1]), _cuda(g["num_points"][:1]), _cuda(g["coors"][:1]))
    assert tuple(out.shape) == (64,)            # pillar_encoder.py:169 squeeze quirk


def test_scatter_matches_reference(g):
    from partner_b200 import PointPillarsScatter
    from partner_b200 import functional as F
    sc = PointPillarsScatter(num_input_features=7)
    canvas = sc(_cuda(g["vfe_mean"]), _cuda(g["coors"]), 2, [512, 512, 1]).cpu().numpy()
    assert canvas.shape == (2, 7, 512, 512)
    assert np.array_equal(canvas, densify(g["canvas_idx"], g["canvas_val"], g["canvas_shape"]))
    _, bev = F.scatter(_cuda(g["vfe_mean"]), _cuda(g["coors"]), 2, 512, 512, want_bev_index=True)
    assert np.array_equal(bev.cpu().numpy(), g["bev_index"])


def test_scatter_wide_features_and_odd_canvas():
    """C = 128 (PFN output width) and a canvas whose cell count is not a multiple of 4."""
    from partner_b200 import functional as F
    rng = np.random.default_rng(0)
    for (ny, nx, c, b) in ((512, 512, 128, 2), (37, 23, 5, 3)):
        cells = rng.permutation(ny * nx)[: min(3000, ny * nx // 2)]
        coors = np.stack([rng.integers(0, b, cells.size), np.zeros_like(cells), cells // nx, cells % nx],
                         axis=1).astype(np.int32)
        feats = rng.normal(size=(cells.size, c)).astype(np.float32)
        ref, _ = oracle.scatter(feats, coors, b, [nx, ny, 1])
        got = F.scatter(_cuda(feats), _cuda(coors), b, ny, nx).cpu().numpy()
        assert np.array_equal(got, ref)


def test_registry_builds_reference_config_dicts():
    from partner_b200 import BACKBONES, READERS, build_from_cfg
    reader = build_from_cfg(dict(type="PillarFeatureNet", num_filters=[64, 64], num_input_features=5,
                                 with_distance=False, voxel_size=(0.32, 0.32, 6.0),
                                 pc_range=(-74.88, -74.88, -2, 74.88, 74.88, 4.0)), READERS)
    assert [tuple(l.linear.weight.shape) for l in reader.pfn_layers] == [(32, 10), (64, 64)]
    bb = build_from_cfg(dict(type="PointPillarsScatter", ds_factor=1, num_input_features=64), BACKBONES)
    assert bb.nchannels == 64
