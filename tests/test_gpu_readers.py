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


@pytest.mark.parametrize("filters,dist,t,c", [((64, 128), False, 20, 7), ((64, 64), False, 20, 7), ((64, 128), True, 20, 7),
                                              ((64, 96), False, 5, 8), ((64, 128), False, 32, 5), ((64, 32), True, 1, 9)])
def test_pfn_tensor_core_kernel_vs_oracle(filters, dist, t, c):
    """The default two-layer path (pfn_fused.cu: second layer on tcgen05, 3xTF32, accumulators in TMEM) on
    random padded tensors: every fill level (full voxels, single points, padded-slot quirk), M not a
    multiple of anything, same 1e-5 gate as the fp32 kernels."""
    import torch
    from partner_b200 import functional as F
    rng = np.random.default_rng(hash((filters, dist, t, c)) % 2 ** 31)
    m = 20011
    num = rng.integers(1, t + 1, m).astype(np.int32)
    num[::7] = t
    vox = rng.normal(0, 3, (m, t, c)).astype(np.float32) * (np.arange(t)[None, :, None] < num[:, None, None])
    coors = np.stack([np.zeros(m), np.zeros(m), rng.integers(0, 512, m), rng.integers(0, 512, m)], 1).astype(np.int32)
    vs, rg = [0.098, 0.0123, 8.0], [0.3, -3.1488, -5.0, 50.476, 3.1488, 3.0]
    width, layers, dev_layers = c + 5 + (1 if dist else 0), [], []
    for i, fo in enumerate(filters):
        u = fo if i == len(filters) - 1 else fo // 2
        L = dict(weight=rng.normal(0, 0.3, (u, width)).astype(np.float32), mean=rng.normal(0, 1, u).astype(np.float32),
                 var=rng.uniform(0.5, 2, u).astype(np.float32), gamma=rng.normal(0, 1, u).astype(np.float32),
                 beta=rng.normal(0, 1, u).astype(np.float32))
        layers.append(L)
        dev_layers.append(tuple(torch.from_numpy(L[k]).cuda() for k in ("weight", "mean", "var", "gamma", "beta")))
        width = 2 * u
    ref = oracle.pfn_forward(vox, num, coors, layers, vs, rg, with_distance=dist, eps=1e-3)
    out = F.pfn_forward(_cuda(vox), _cuda(num), _cuda(coors), dev_layers, vs[0], vs[1], vs[0] / 2 + rg[0], vs[1] / 2 + rg[1],
                        dist, 1e-3)
    assert_close_fp32(out.cpu().numpy(), ref, "tcgen05 PFN %s dist=%s t=%d c=%d" % (filters, dist, t, c))


@pytest.mark.parametrize("tag,filters,dist", [("t64_128", (64, 128), False), ("t32_32_64_dist", (32, 32, 64), True)])
def test_pfn_training_mode_matches_reference_autograd(tag, filters, dist, golden_dir):
    """PillarFeatureNet.train(): forward with batch statistics over all M * T rows (padded slots included),
    running statistics after the step, and the gradients of every parameter, against the REFERENCE module
    run in .train() mode under torch.autograd (tests/golden/make_golden_train.py)."""
    import torch
    from partner_b200 import PillarFeatureNet
    t = np.load(os.path.join(golden_dir, "pfn_train.npz"))
    net = PillarFeatureNet(7, filters, dist, tuple(t["voxel_size"]), tuple(t["pc_range"]))
    sd = {}
    for i in range(len(filters)):
        sd[f"pfn_layers.{i}.linear.weight"] = torch.from_numpy(t[f"{tag}_w{i}"])
        sd[f"pfn_layers.{i}.norm.running_mean"] = torch.from_numpy(t[f"{tag}_mean{i}"])
        sd[f"pfn_layers.{i}.norm.running_var"] = torch.from_numpy(t[f"{tag}_var{i}"])
        sd[f"pfn_layers.{i}.norm.weight"] = torch.from_numpy(t[f"{tag}_gamma{i}"])
        sd[f"pfn_layers.{i}.norm.bias"] = torch.from_numpy(t[f"{tag}_beta{i}"])
        sd[f"pfn_layers.{i}.norm.num_batches_tracked"] = torch.tensor(0)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    out = net(_cuda(t["voxels"]), _cuda(t["num_points"]), _cuda(t["coors"]))
    assert_close_fp32(out.detach().cpu().numpy(), t[f"{tag}_out"], tag + " forward (batch statistics)")
    (out * _cuda(t[f"{tag}_G"])).sum().backward()

    def close(a, b, what, tol=2e-4):
        # gradients are sums over 30 000 rows in a different order than torch's: relative to the tensor's scale
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)
        assert err < tol, "%s: max error %.3g of the largest entry" % (what, err)

    for i, l in enumerate(net.pfn_layers):
        close(l.norm.running_mean.cpu().numpy(), t[f"{tag}_mean{i}_after"], f"running_mean {i}", 1e-5)
        close(l.norm.running_var.cpu().numpy(), t[f"{tag}_var{i}_after"], f"running_var {i}", 1e-5)
        assert int(l.norm.num_batches_tracked) == 1
        close(l.linear.weight.grad.cpu().numpy(), t[f"{tag}_dw{i}"], f"d weight {i}")
        close(l.norm.weight.grad.cpu().numpy(), t[f"{tag}_dgamma{i}"], f"d gamma {i}")
        close(l.norm.bias.grad.cpu().numpy(), t[f"{tag}_dbeta{i}"], f"d beta {i}")
    # eval mode afterwards uses the updated running statistics through the inference kernels
    net.eval()
    ev = net(_cuda(t["voxels"]), _cuda(t["num_points"]), _cuda(t["coors"]))
    assert tuple(ev.shape) == tuple(out.shape) and bool(torch.isfinite(ev).all())


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


def _random_pfn_state(net, seed=0):
    """SURVEY.md section 8d weights: Linear default init, BN running stats / affine drawn at random
    (eval mode), so the padded-slot quirk (relu(shift) != 0 takes part in the max) is exercised."""
    import torch
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for layer in net.pfn_layers:
            u = layer.norm.num_features
            layer.norm.running_mean.copy_(torch.randn(u, generator=gen))
            layer.norm.running_var.copy_(torch.rand(u, generator=gen) * 1.5 + 0.5)
            layer.norm.weight.copy_(torch.randn(u, generator=gen))
            layer.norm.bias.copy_(torch.randn(u, generator=gen))
    return net


def test_static_pfn_full_size_vs_oracle():
    """BASELINE config 3 at full size: two nuScenes 10-sweep frames on the NUSC-PILLAR grid (max_voxels
    binding, T = 20, heavy and non-full voxels mixed) through VoxelGenerator.generate ->
    PillarFeatureNet(7, [64, 128]) -> PointPillarsScatter(128) vs oracle.pfn_forward / oracle.scatter."""
    import torch
    from partner_b200 import PillarFeatureNet, PointPillarsScatter, VoxelGenerator, synth
    g = synth.GRIDS["NUSC-PILLAR"]
    gen = VoxelGenerator(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
    ref_gen = oracle.VoxelGenerator(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
    frames = [oracle.transform_points(synth.nusc_frame(3100 + k)) for k in range(2)]
    vox, coor, num = [], [], []
    for b, f in enumerate(frames):
        v, c, n = gen.generate(f)[:3]
        rv, rc, rn = ref_gen.generate(f)[:3]
        assert np.array_equal(c, rc) and np.array_equal(n, rn) and np.array_equal(v, rv)
        vox.append(v)
        num.append(n)
        coor.append(np.pad(c, ((0, 0), (1, 0)), constant_values=b))      # collate_kitti batch column
    vox, coor, num = np.concatenate(vox), np.concatenate(coor).astype(np.int32), np.concatenate(num)
    assert vox.shape[0] == 2 * g["max_voxels"] and (num == g["max_points"]).any() and (num < g["max_points"]).any()
    torch.manual_seed(0)
    net = _random_pfn_state(PillarFeatureNet(7, (64, 128), False, tuple(g["voxel_size"]), tuple(g["range"]))).cuda().eval()
    layers = [dict(weight=L.linear.weight.detach().cpu().numpy(), mean=L.norm.running_mean.cpu().numpy(),
                   var=L.norm.running_var.cpu().numpy(), gamma=L.norm.weight.detach().cpu().numpy(),
                   beta=L.norm.bias.detach().cpu().numpy()) for L in net.pfn_layers]
    ref = oracle.pfn_forward(vox, num, coor, layers, g["voxel_size"], g["range"], with_distance=False, eps=1e-3)
    dc = _cuda(coor)
    out = net(_cuda(vox), _cuda(num), dc)
    assert_close_fp32(out.cpu().numpy(), ref, "static PFN [64, 128], full size")
    canvas = PointPillarsScatter(128)(out, dc, 2, [512, 512, 1])
    rc, _ = oracle.scatter(ref, coor, 2, [512, 512, 1])
    got = canvas.cpu().numpy()
    assert got.shape == rc.shape == (2, 128, 512, 512)
    assert_close_fp32(got, rc, "PFN canvas, full size")


def test_fused_pillar_front_end_vs_oracle():
    """pv_forward_pfn_canvas (raw Cartesian sweeps -> voxelize -> PFN [64, 128] -> scatter, no [M, T, C]
    tensor, second layer on tcgen05) on three frames -- one full nuScenes 10-sweep frame with max_voxels
    binding, one sparse, one empty -- vs the oracle chain transform_points -> VoxelGenerator.generate ->
    pfn_forward -> scatter.  Integers bit-exact, features / canvas within the 1e-5 gate."""
    import torch
    from partner_b200 import PillarFeatureNet, PillarFrontEnd, synth
    g = synth.GRIDS["NUSC-PILLAR"]
    frames = [synth.nusc_frame(3200), synth.nusc_frame(3201, nsweeps=1), np.zeros((0, 5), np.float32)]
    torch.manual_seed(1)
    net = _random_pfn_state(PillarFeatureNet(7, (64, 128), False, tuple(g["voxel_size"]), tuple(g["range"])), seed=1).cuda().eval()
    fe = PillarFrontEnd(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"], net, cartesian=True)
    got = fe(frames)
    ref_gen = oracle.VoxelGenerator(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
    outs = [ref_gen.generate(oracle.transform_points(f))[:3] for f in frames]
    vox, coor, num, nv = oracle.collate(outs)
    assert np.array_equal(got["num_voxels"], nv)
    assert np.array_equal(got["coordinates"], coor)
    assert np.array_equal(got["num_points"], num)
    layers = [dict(weight=L.linear.weight.detach().cpu().numpy(), mean=L.norm.running_mean.cpu().numpy(),
                   var=L.norm.running_var.cpu().numpy(), gamma=L.norm.weight.detach().cpu().numpy(),
                   beta=L.norm.bias.detach().cpu().numpy()) for L in net.pfn_layers]
    ref = oracle.pfn_forward(vox, num, coor, layers, g["voxel_size"], g["range"], with_distance=False, eps=1e-3)
    assert_close_fp32(got["features"], ref, "fused PFN features")
    rc, _ = oracle.scatter(ref, coor, len(frames), [512, 512, 1])
    assert_close_fp32(got["canvas"], rc, "fused PFN canvas")
    again = fe(frames)                      # the point lists and the map were restored: a second call is identical
    assert np.array_equal(again["coordinates"], coor) and np.array_equal(again["num_points"], num)
    assert_close_fp32(again["features"], ref, "fused PFN features, second call")


@pytest.mark.parametrize("name,cartesian,c_in,with_distance,filters", [
    ("cartesian5_distance_64", True, 5, True, (64, 64)),     # C = 7, decorated width 13: four float4 planes per row
    ("polar3_96", False, 3, False, (64, 96)),                # C = 3, decorated width 8: two planes
    ("polar4_32", False, 4, False, (64, 32)),                # C = 4, decorated width 9: three planes, one TMEM quarter
])
def test_fused_pillar_front_end_row_widths(name, cartesian, c_in, with_distance, filters):
    """pv_forward_pfn_canvas on the other decorated-row widths (the tensor-core kernel is instantiated per
    number of float4 planes), with and without the distance channel, Cartesian and polar input, several
    last-layer widths; two 1-sweep frames + one 2-sweep frame.  Same gates as the full-size test."""
    import torch
    from partner_b200 import PillarFeatureNet, PillarFrontEnd, synth
    g = synth.GRIDS["NUSC-PILLAR"]
    cart = [synth.nusc_frame(3300, nsweeps=1), synth.nusc_frame(3301, nsweeps=2), synth.nusc_frame(3302, nsweeps=1)[:777]]
    if cartesian:
        frames = [np.ascontiguousarray(f[:, :c_in]) for f in cart]
        polar = [oracle.transform_points(f) for f in frames]
    else:
        polar = [np.ascontiguousarray(oracle.transform_points(f)[:, :c_in]) for f in cart]
        frames = polar
    C = polar[0].shape[1]
    net = _random_pfn_state(PillarFeatureNet(C, filters, with_distance, tuple(g["voxel_size"]), tuple(g["range"])), seed=7).cuda().eval()
    fe = PillarFrontEnd(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"], net, cartesian=cartesian)
    got = fe(frames)
    ref_gen = oracle.VoxelGenerator(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
    vox, coor, num, nv = oracle.collate([ref_gen.generate(p)[:3] for p in polar])
    assert np.array_equal(got["num_voxels"], nv)
    assert np.array_equal(got["coordinates"], coor)
    assert np.array_equal(got["num_points"], num)
    layers = [dict(weight=L.linear.weight.detach().cpu().numpy(), mean=L.norm.running_mean.cpu().numpy(),
                   var=L.norm.running_var.cpu().numpy(), gamma=L.norm.weight.detach().cpu().numpy(),
                   beta=L.norm.bias.detach().cpu().numpy()) for L in net.pfn_layers]
    ref = oracle.pfn_forward(vox, num, coor, layers, g["voxel_size"], g["range"], with_distance=with_distance, eps=1e-3)
    assert_close_fp32(got["features"], ref, "fused PFN features " + name)
    rc, _ = oracle.scatter(ref, coor, len(frames), [512, 512, 1])
    assert_close_fp32(got["canvas"], rc, "fused PFN canvas " + name)


def test_static_pfn_slices_agree_with_one_slice_calls():
    """The drop-in PillarFeatureNet on the padded tensor runs in slices of 524 288 voxels (pfn_fused.cu, P2_SLICE).
    Size-independent property at 655 360 voxels (two slices, the second one partial): the rows of the big call are
    bit-identical to the rows of calls on sub-ranges that fit one slice -- every voxel is evaluated on its own --,
    and the first 120 000 rows match the oracle."""
    import torch
    from partner_b200 import PillarFeatureNet, synth
    g = synth.GRIDS["NUSC-PILLAR"]
    ref_gen = oracle.VoxelGenerator(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
    vox, coor, num = [], [], []
    for b in range(2):
        v, c, n = ref_gen.generate(oracle.transform_points(synth.nusc_frame(3400 + b)))[:3]
        vox.append(v); num.append(n); coor.append(np.pad(c, ((0, 0), (1, 0)), constant_values=b))
    vox, coor, num = np.concatenate(vox), np.concatenate(coor).astype(np.int32), np.concatenate(num)
    m0 = vox.shape[0]
    reps = -(-655360 // m0)
    sel = (np.arange(reps * m0) * 7919 % m0)[:655360]            # a permuted tiling: slice boundaries fall inside frames
    V, N, Cc = _cuda(vox[sel]), _cuda(num[sel]), _cuda(coor[sel])
    net = _random_pfn_state(PillarFeatureNet(7, (64, 128), False, tuple(g["voxel_size"]), tuple(g["range"])), seed=5).cuda().eval()
    big = net(V, N, Cc)
    assert big.shape == (655360, 128)
    for lo, hi in ((0, 300000), (300000, 655360), (524288 - 1000, 524288 + 1000)):
        part = net(V[lo:hi].contiguous(), N[lo:hi].contiguous(), Cc[lo:hi].contiguous())
        assert torch.equal(part, big[lo:hi]), "rows [%d, %d) differ between the sliced and the one-slice call" % (lo, hi)
    layers = [dict(weight=L.linear.weight.detach().cpu().numpy(), mean=L.norm.running_mean.cpu().numpy(),
                   var=L.norm.running_var.cpu().numpy(), gamma=L.norm.weight.detach().cpu().numpy(),
                   beta=L.norm.bias.detach().cpu().numpy()) for L in net.pfn_layers]
    k = 20000
    ref = oracle.pfn_forward(vox[sel[:k]], num[sel[:k]], coor[sel[:k]], layers, g["voxel_size"], g["range"], with_distance=False, eps=1e-3)
    assert_close_fp32(big[:k].cpu().numpy(), ref, "static PFN, sliced call")


def test_fused_pillar_front_end_without_points():
    """pv_forward_pfn_canvas on batches that contain no voxel at all (no points; only out-of-range points): the
    tensor-core kernel's tile protocol ends in its first round, nothing is written but the zero canvas."""
    from partner_b200 import PillarFeatureNet, PillarFrontEnd, synth
    g = synth.GRIDS["NUSC-PILLAR"]
    net = _random_pfn_state(PillarFeatureNet(7, (64, 128), False, tuple(g["voxel_size"]), tuple(g["range"])), seed=3).cuda().eval()
    fe = PillarFrontEnd(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"], net, cartesian=True)
    far = np.zeros((500, 5), np.float32)
    far[:, 0] = 500.0                                          # rho beyond the grid
    for frames in ([np.zeros((0, 5), np.float32)], [np.zeros((0, 5), np.float32), far, np.zeros((0, 5), np.float32)]):
        got = fe(frames)
        assert got["features"].shape == (0, 128) and got["coordinates"].shape == (0, 4)
        assert np.array_equal(got["num_voxels"], np.zeros(len(frames), np.int64))
        assert got["canvas"].shape == (len(frames), 128, 512, 512) and not got["canvas"].any()
    one = fe([synth.nusc_frame(3500, nsweeps=1)[:300]])       # and the next call on the same workspaces is a normal one
    assert one["features"].shape[0] == int(one["num_voxels"].sum()) > 0 and np.isfinite(one["features"]).all()
