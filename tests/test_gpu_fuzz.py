"""Randomised differential parity (seeded): random grids, caps, channel counts, batch splits and point
distributions through every pipeline of the voxelizer, against the oracle.  Exercises the code
paths the fixed-size tests do not: generic channel counts (row widths of 1-5 vector reductions),
dense 3-D direct maps, hash maps, frames that are empty / tiny / not multiples of the tile, input
rows that start at unaligned addresses (no TMA bulk copy), caps that bind and caps that do not."""
import numpy as np
import pytest

import oracle
from util import assert_close_fp32

pytestmark = pytest.mark.gpu


def _random_case(seed):
    rng = np.random.default_rng(seed)
    dense = rng.random() < 0.6
    if dense:
        grid = [int(rng.integers(4, 200)), int(rng.integers(4, 200)), int(rng.choice([1, 1, 2, 5]))]
    else:
        grid = [int(rng.integers(300, 1500)), int(rng.integers(300, 1500)), int(rng.integers(1, 40))]
        while grid[0] * grid[1] * grid[2] <= (1 << 20):
            grid[0] *= 2
    lo = np.array([rng.uniform(0.1, 1.0), -3.1488, rng.uniform(-5, -1)], np.float32)
    vs = np.array([rng.uniform(0.05, 0.5), 2 * 3.1488 / grid[1], rng.uniform(0.2, 8.0)], np.float32)
    hi = lo + vs * np.array(grid, np.float32)
    c = int(rng.choice([3, 4, 5, 7, 8, 9, 12, 14]))
    T = int(rng.choice([1, 2, 5, 20, 40]))
    B = int(rng.integers(1, 5))
    sizes = [int(rng.choice([0, 1, 3, 1023, 1024, 1025, int(rng.integers(100, 30000))])) for _ in range(B)]
    frames = []
    for n in sizes:
        p = np.zeros((n, c), np.float32)
        spread = rng.choice([0.02, 0.3, 1.2])              # tight clusters (heavy cells) ... beyond the range
        p[:, 0] = lo[0] + rng.random(n) * (hi[0] - lo[0]) * spread * rng.choice([1.0, 1.0, 1.1])
        p[:, 1] = rng.uniform(-3.3, 3.3, n) if rng.random() < 0.5 else rng.choice(np.linspace(-3, 3, 7), n) + rng.normal(0, 0.01, n)
        p[:, 2] = rng.uniform(lo[2] - 1, hi[2] + 1, n)
        p[:, 3:] = rng.normal(0, 10, (n, c - 3))
        if n > 10:                                          # consecutive duplicates (runs merged inside a thread)
            k = int(rng.integers(0, n - 5))
            p[k + 1:k + 4] = p[k]
        frames.append(p.astype(np.float32))
    total = sum(sizes)
    V = int(rng.choice([max(1, total // 50), max(1, total // 3), total + 10]))
    return dict(grid=grid, range=[float(lo[0]), float(lo[1]), float(lo[2]), float(hi[0]), float(hi[1]), float(hi[2])],
                voxel_size=[float(v) for v in vs], c=c, T=T, V=V, frames=frames, dense=dense)


@pytest.mark.parametrize("seed", range(96))
def test_random_configs_all_pipelines_vs_oracle(seed):
    import torch
    from partner_b200 import functional as F, _lib
    case = _random_case(seed)
    ref = oracle.VoxelGenerator(case["voxel_size"], case["range"], case["T"], case["V"])
    if not np.array_equal(ref.grid_size, case["grid"]):
        pytest.skip("rounding moved the grid size")       # grid = round((hi - lo) / vs) of f32 values
    cfg, _, _, _ = F.make_config(case["voxel_size"], case["range"], case["T"], case["V"])
    frames = case["frames"]
    # every third seed: Cartesian rows through the fused cylinder transform (oracle sees its own polar rows)
    cart = seed % 3 == 0 and case["c"] >= 5
    if cart:
        dev_frames = []
        for k, f in enumerate(frames):
            c = np.concatenate([f[:, 0:1] * np.cos(f[:, 1:2]), f[:, 0:1] * np.sin(f[:, 1:2]), f[:, 2:3], f[:, 5:]], axis=1).astype(np.float32)
            dev_frames.append(c)
            frames[k] = oracle.transform_points(c) if c.shape[0] else np.zeros((0, c.shape[1] + 2), np.float32)
    else:
        dev_frames = frames
    want_den = int(np.prod(case["grid"])) <= (1 << 22)
    outs = [ref.generate(f, return_pc_grid_ind=True, return_density=want_den) for f in frames]
    vox, coor, num, nv = oracle.collate([(o[0], o[1], o[2]) for o in outs])
    mean = oracle.vfe_mean(vox, num) if vox.shape[0] else np.zeros((0, case["c"]), np.float32)
    sizes = [f.shape[0] for f in frames]
    off = np.zeros(len(frames) + 1, np.int32)
    np.cumsum(sizes, out=off[1:])
    n = int(off[-1])
    allp = np.concatenate(frames) if n else np.zeros((0, case["c"]), np.float32)
    devp = np.concatenate(dev_frames) if n else np.zeros((0, dev_frames[0].shape[1]), np.float32)
    c_dev = devp.shape[1]
    # odd seeds: the rows start 4 bytes past a 16-byte boundary (plain-load staging instead of TMA)
    pad = 1 if seed % 2 else 0
    buf = torch.zeros(n * c_dev + pad + 4, dtype=torch.float32, device="cuda")
    pts = buf[pad:pad + n * c_dev].view(n, c_dev)
    pts.copy_(torch.from_numpy(devp))
    d_off = torch.from_numpy(off).cuda()
    for pipeline, want_voxels in ((2, False), (0, False), (0, True)):
        cfg.pipeline = pipeline                 # per-call configuration, no process state
        vb = F.voxelize(cfg, pts, d_off, len(frames), max(sizes + [1]), cart, want_voxels=want_voxels,
                        want_mean=True, want_grid_ind=True, want_density=want_den)
        F.read_status(vb)
        m = vb.total()
        tag = "pipeline %d voxels %s" % (pipeline, want_voxels)
        assert np.array_equal(vb.voxel_counts.cpu().numpy(), nv), tag
        assert np.array_equal(vb.coors[:m].cpu().numpy(), coor), tag
        assert np.array_equal(vb.num_points[:m].cpu().numpy(), num), tag
        assert np.array_equal(vb.pc_grid_ind.cpu().numpy(), np.concatenate([o[3] for o in outs])), tag
        if want_den:
            assert np.array_equal(vb.density.cpu().numpy(), np.stack([o[4] for o in outs])), tag
        if want_voxels:
            assert np.array_equal(vb.voxels[:m].cpu().numpy(), vox), tag
        assert_close_fp32(vb.mean_feats[:m].cpu().numpy(), mean, "mean, " + tag)
    if case["dense"] and case["grid"][1] * case["grid"][2] <= 65535:
        gi = np.concatenate([np.pad(oracle.dynamic_grid_ind(f, case["voxel_size"], case["range"]), ((0, 0), (1, 0)),
                                    constant_values=b) for b, f in enumerate(frames)]) if n else np.zeros((0, 4), np.int32)
        dmean, unq, inv, cnt = oracle.dynamic_mean(gi, allp)
        r = F.dynamic_voxelize(cfg, pts, d_off, len(frames), max(sizes + [1]), cart, want_grid_ind=True)
        F.read_status(r)
        m = r.total()
        assert np.array_equal(r.grid_ind.cpu().numpy(), gi)
        assert np.array_equal(r.unq[:m].cpu().numpy(), unq)
        assert np.array_equal(r.unq_inv.cpu().numpy(), inv)
        assert np.array_equal(r.unq_cnt[:m].cpu().numpy(), cnt)
        assert_close_fp32(r.mean_feats[:m].cpu().numpy(), dmean, "dynamic mean")


@pytest.mark.parametrize("seed", range(24))
def test_random_dynamic_pfnet_vs_oracle(seed):
    """Random decoration flags / voxel shape / layer widths / channel counts / voxel sizes (1 point up
    to hundreds per voxel, voxels spanning several 64-point chunks) through pv_dynamic_pfn."""
    import torch
    from partner_b200 import functional as F
    rng = np.random.default_rng(1000 + seed)
    nx, ny = int(rng.integers(4, 60)), int(rng.integers(4, 60))
    vs = [float(rng.uniform(0.1, 0.5)), 2 * 3.1488 / ny, 8.0]
    rg = [0.3, -3.1488, -5.0, 0.3 + vs[0] * nx, -3.1488 + vs[1] * ny, 3.0]
    cfg, _, _, gs = F.make_config(vs, rg, 20, 1000)
    if tuple(gs) != (nx, ny, 1):
        pytest.skip("rounding moved the grid size")
    c = int(rng.choice([5, 6, 7, 9]))
    B = int(rng.integers(1, 4))
    n = int(rng.choice([1, 70, 3000, 20000]))
    pts = np.zeros((n, c), np.float32)
    concentrate = rng.random() < 0.5                        # a few very full voxels
    pts[:, 0] = rg[0] + rng.random(n) * (rg[3] - rg[0]) * (0.05 if concentrate else 1.0)
    pts[:, 1] = rng.uniform(-3.1, 3.1, n) * (0.05 if concentrate else 1.0)
    pts[:, 2] = rng.uniform(-5, 3, n)
    pts[:, 3] = pts[:, 0] * np.cos(pts[:, 1])
    pts[:, 4] = pts[:, 0] * np.sin(pts[:, 1])
    pts[:, 5:] = rng.normal(0, 1, (n, c - 5))
    pts = pts.astype(np.float32)
    cuts = np.sort(rng.integers(0, n + 1, B - 1))
    off = np.concatenate([[0], cuts, [n]]).astype(np.int32)
    flags = dict(xyz_cluster=bool(rng.integers(2)), raz_cluster=bool(rng.integers(2)), xy_center=bool(rng.integers(2)),
                 ra_center=bool(rng.integers(2)))
    shape = str(rng.choice(["cuboid", "cylinder"]))
    c0 = c + (3 if flags["xyz_cluster"] else 0) + (2 if flags["xy_center"] else 0) + \
        ((2 if flags["xyz_cluster"] else 3) if flags["raz_cluster"] else 0) + (2 if flags["ra_center"] else 0)
    if rng.random() < 0.4:
        units = [int(rng.choice([16, 64, 128]))]
        ws = [rng.normal(0, 0.3, (units[0], c0)).astype(np.float32)]
    else:
        u1, u2 = int(rng.choice([8, 32, 64])), int(rng.choice([32, 64, 128]))
        ws = [rng.normal(0, 0.3, (u1, c0)).astype(np.float32), rng.normal(0, 0.2, (u2, 2 * u1)).astype(np.float32)]
    gi = np.concatenate([np.pad(oracle.dynamic_grid_ind(pts[off[b]:off[b + 1]], vs, rg), ((0, 0), (1, 0)), constant_values=b)
                         for b in range(B)])
    mean, unq, inv, cnt = oracle.dynamic_mean(gi, pts)
    ref = oracle.dynamic_pfn(pts, inv, unq, ws, vs, rg, shape, **flags)
    d = torch.from_numpy(pts).cuda()
    r = F.dynamic_voxelize(cfg, d, torch.from_numpy(off).cuda(), B, n, False)
    m = r.total()
    assert np.array_equal(r.unq[:m].cpu().numpy(), unq)
    out = F.dynamic_pfn(d, r, m, [torch.from_numpy(w).cuda() for w in ws], vs[0], vs[1], vs[0] / 2 + rg[0], vs[1] / 2 + rg[1],
                        shape != "cuboid", flags["xyz_cluster"], flags["raz_cluster"], flags["xy_center"], flags["ra_center"])
    assert_close_fp32(out.cpu().numpy(), ref, "dynamic pfn %s %s" % (shape, flags))


@pytest.mark.parametrize("seed", range(24))
def test_random_static_readers_vs_oracle(seed):
    """Random padded voxel tensors (M, T, C, fill levels incl. full and single-point voxels) and random
    PFN stacks (1-3 layers, widths, with_distance) through pv_vfe_mean / pv_pfn_forward / pv_scatter."""
    import torch
    from partner_b200 import functional as F
    rng = np.random.default_rng(2000 + seed)
    m = int(rng.choice([1, 2, 257, 5000, 40000]))
    t = int(rng.choice([1, 5, 20, 32]))
    c = int(rng.choice([4, 5, 7, 9]))
    nx, ny, B = int(rng.integers(8, 200)), int(rng.integers(8, 200)), int(rng.integers(1, 4))
    num = rng.integers(1, t + 1, m).astype(np.int32)
    if rng.random() < 0.3:
        num[:] = t                                          # every voxel full: no padded slot anywhere
    vox = rng.normal(0, 3, (m, t, c)).astype(np.float32)
    vox *= (np.arange(t)[None, :, None] < num[:, None, None])                 # zero padding, as points_to_voxel leaves it
    cells = rng.choice(B * ny * nx, size=min(m, B * ny * nx), replace=False)
    m = len(cells)
    vox, num = vox[:m], num[:m]
    coors = np.stack([cells // (ny * nx), np.zeros(m, np.int64), (cells // nx) % ny, cells % nx], 1).astype(np.int32)
    dv, dn, dc = torch.from_numpy(vox).cuda(), torch.from_numpy(num).cuda(), torch.from_numpy(coors).cuda()
    assert_close_fp32(F.vfe_mean(dv, dn).cpu().numpy(), oracle.vfe_mean(vox, num), "vfe_mean")
    dist = bool(rng.integers(2))
    nl = int(rng.integers(1, 4))
    filters = [int(rng.choice([16, 32, 64, 128])) for _ in range(nl)]
    vs, rg = [0.1, 0.02, 8.0], [0.3, -3.1, -5.0, 0.3 + 0.1 * nx, -3.1 + 0.02 * ny, 3.0]
    width, layers, dev_layers = c + 5 + (1 if dist else 0), [], []
    for i, fo in enumerate(filters):
        u = fo if i == nl - 1 else fo // 2
        L = dict(weight=rng.normal(0, 0.3, (u, width)).astype(np.float32), mean=rng.normal(0, 1, u).astype(np.float32),
                 var=rng.uniform(0.5, 2, u).astype(np.float32), gamma=rng.normal(0, 1, u).astype(np.float32),
                 beta=rng.normal(0, 1, u).astype(np.float32))
        layers.append(L)
        dev_layers.append(tuple(torch.from_numpy(L[k]).cuda() for k in ("weight", "mean", "var", "gamma", "beta")))
        width = u if i == nl - 1 else 2 * u
    ref = oracle.pfn_forward(vox, num, coors, layers, vs, rg, with_distance=dist, eps=1e-3)
    out = F.pfn_forward(dv, dn, dc, dev_layers, vs[0], vs[1], vs[0] / 2 + rg[0], vs[1] / 2 + rg[1], dist, 1e-3)
    assert_close_fp32(out.cpu().numpy(), ref, "pfn %s dist=%s m=%d t=%d c=%d" % (filters, dist, m, t, c))
    canvas, bev = F.scatter(out, dc, B, ny, nx, want_bev_index=True)
    rc, rb = oracle.scatter(ref, coors, B, [nx, ny, 1])
    assert np.array_equal(bev.cpu().numpy(), rb)
    assert_close_fp32(canvas.cpu().numpy(), rc, "canvas")


@pytest.mark.parametrize("seed", range(32))
def test_random_fused_pillar_front_end_vs_oracle(seed):
    """pv_forward_pfn_canvas (gather pre-pass + tensor-core kernel + scatter) on random batches: 1-6 frames of
    0 ... 40 000 points (empty and single-point frames, batches where only one producer set finds work, batches
    that end in partial tiles), random caps (T down to 1, V binding or not), polar 3- to 7-channel or Cartesian
    4- / 5-channel input, with and without the distance channel, output widths 32-128.  Integers bit-exact,
    features / canvas within the 1e-5 gate; exercises the exit paths of the kernel's tile protocol."""
    import torch
    from partner_b200 import PillarFeatureNet, PillarFrontEnd
    rng = np.random.default_rng(1000 + seed)
    nx, ny = int(rng.choice([64, 128, 256])), int(rng.choice([64, 128, 512]))
    lo = np.array([0.3, -3.1488, -5.0], np.float32)
    vs = [float(rng.uniform(0.08, 0.4)), 2 * 3.1488 / ny, 8.0]
    rg = [float(lo[0]), float(lo[1]), -5.0, float(lo[0] + vs[0] * nx), float(lo[1] + vs[1] * ny), 3.0]
    T = int(rng.choice([1, 3, 20, 32]))
    V = int(rng.choice([50, 2000, 60000]))
    cartesian = bool(rng.random() < 0.5)
    c_in = int(rng.choice([4, 5])) if cartesian else int(rng.choice([3, 4, 5, 7]))
    with_distance = bool(rng.random() < 0.4)
    units = int(rng.choice([32, 64, 96, 128]))
    B = int(rng.integers(1, 7))
    sizes = [int(rng.choice([0, 1, 40, 700, int(rng.integers(1000, 40000))])) for _ in range(B)]
    frames, polar = [], []
    for n in sizes:
        rho = lo[0] + rng.random(n) * vs[0] * nx * float(rng.choice([0.05, 0.5, 1.05]))
        phi = rng.uniform(-3.2, 3.2, n)
        z = rng.uniform(-5.5, 3.5, n)
        extra = rng.normal(0, 5, (n, 4))
        if cartesian:
            f = np.column_stack([rho * np.cos(phi), rho * np.sin(phi), z, extra])[:, :c_in].astype(np.float32)
            frames.append(np.ascontiguousarray(f)); polar.append(oracle.transform_points(frames[-1]))
        else:
            f = np.column_stack([rho, phi, z, extra])[:, :c_in].astype(np.float32)
            frames.append(np.ascontiguousarray(f)); polar.append(frames[-1])
    C = polar[0].shape[1]
    if C + 5 + (1 if with_distance else 0) > 16:
        with_distance = False
    torch.manual_seed(seed)
    net = PillarFeatureNet(C, (64, units), with_distance, tuple(vs), tuple(rg)).cuda().eval()
    g = torch.Generator().manual_seed(seed)
    for L in net.pfn_layers:
        u = L.norm.num_features
        L.norm.running_mean.copy_(torch.randn(u, generator=g)); L.norm.running_var.copy_(torch.rand(u, generator=g) * 1.5 + 0.5)
        L.norm.weight.data.copy_(torch.randn(u, generator=g)); L.norm.bias.data.copy_(torch.randn(u, generator=g))
    fe = PillarFrontEnd(vs, rg, T, V, net, cartesian=cartesian)
    got = fe(frames)
    ref_gen = oracle.VoxelGenerator(vs, rg, T, V)
    vox, coor, num, nv = oracle.collate([ref_gen.generate(p)[:3] for p in polar])
    assert np.array_equal(got["num_voxels"], nv)
    assert np.array_equal(got["coordinates"], coor)
    assert np.array_equal(got["num_points"], num)
    if coor.shape[0] == 0:
        assert got["features"].shape[0] == 0 and not got["canvas"].any()
        return
    layers = [dict(weight=L.linear.weight.detach().cpu().numpy(), mean=L.norm.running_mean.cpu().numpy(),
                   var=L.norm.running_var.cpu().numpy(), gamma=L.norm.weight.detach().cpu().numpy(),
                   beta=L.norm.bias.detach().cpu().numpy()) for L in net.pfn_layers]
    ref = oracle.pfn_forward(vox, num, coor, layers, vs, rg, with_distance=with_distance, eps=1e-3)
    assert_close_fp32(got["features"], ref, "fused PFN features, random case %d" % seed)
    rc, _ = oracle.scatter(ref, coor, len(frames), [nx, ny, 1])
    assert_close_fp32(got["canvas"], rc, "fused PFN canvas, random case %d" % seed)
