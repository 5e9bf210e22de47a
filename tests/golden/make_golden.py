"""Generate the golden vectors in tests/golden/*.npz FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, numba, torch CPU):

    python tests/golden/make_golden.py

It executes the reference's own functions -- numba ``points_to_voxel`` through
``VoxelGenerator.generate``, ``transform_points``, and the torch modules
``VoxelFeatureExtractorV3`` / ``PillarFeatureNet`` / ``PointPillarsScatter`` --
on small seeded synthetic clouds and freezes inputs + outputs.  The committed
.npz files are what pins ``oracle/`` (tests/test_oracle_golden.py); nothing at
test / bench time reads /root/reference.

Loader recipe: the voxelizer imports as a namespace
package; ``voxel_generator.py`` and the reader modules are loaded by file path
with stub parent packages, because their package __init__ files import
dependencies that are not installed (terminaltables, torch_scatter, spconv).
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from partner_b200 import synth  # noqa: E402  (input generator only)


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    sys.path.insert(0, REF)
    from det3d.ops.point_cloud.point_cloud_ops import points_to_voxel  # noqa: F401
    vg = _load("ref_voxel_generator", f"{REF}/det3d/core/input/voxel_generator.py")
    # transform_points: exec the function source alone (its module's imports fail)
    src = open(f"{REF}/det3d/datasets/pipelines/utils.py").read().splitlines()
    start = next(i for i, l in enumerate(src) if l.startswith("def transform_points"))
    end = next(i for i in range(start + 1, len(src)) if src[i].startswith("def "))
    ns = {"np": np}
    exec("\n".join(src[start:end]), ns)
    # reader modules: stub parents + torch_scatter
    import torch
    ts = types.ModuleType("torch_scatter")
    ts.scatter_mean = ts.scatter_max = None
    sys.modules["torch_scatter"] = ts
    for pkg in ("det3d.models", "det3d.models.readers", "det3d.models.utils", "det3d.utils"):
        m = types.ModuleType(pkg)
        m.__path__ = []
        sys.modules[pkg] = m

    class Registry:
        def __init__(self, name):
            self.name = name
            self.module_dict = {}

        def register_module(self, cls):
            self.module_dict[cls.__name__] = cls
            return cls
    sys.modules["det3d.utils"].Registry = Registry
    _load("det3d.models.registry", f"{REF}/det3d/models/registry.py")
    misc = _load("det3d.models.utils.misc", f"{REF}/det3d/models/utils/misc.py")
    sys.modules["det3d.utils.dist"] = types.ModuleType("det3d.utils.dist")
    sys.modules["det3d.utils.dist"].__path__ = []
    _load("det3d.utils.dist.dist_common", f"{REF}/det3d/utils/dist/dist_common.py")
    norm = _load("det3d.models.utils.norm", f"{REF}/det3d/models/utils/norm.py")
    sys.modules["det3d.models.utils"].get_paddings_indicator = misc.get_paddings_indicator
    sys.modules["det3d.models.utils"].build_norm_layer = norm.build_norm_layer
    pe = _load("det3d.models.readers.pillar_encoder", f"{REF}/det3d/models/readers/pillar_encoder.py")
    ve = _load("det3d.models.readers.voxel_encoder", f"{REF}/det3d/models/readers/voxel_encoder.py")
    return dict(VoxelGenerator=vg.VoxelGenerator, transform_points=ns["transform_points"],
                pe=pe, ve=ve, torch=torch)


def sparse_density(d):
    flat = d.reshape(-1)
    nz = np.flatnonzero(flat)
    return nz.astype(np.int64), flat[nz].astype(np.int32), np.array(d.shape, np.int64)


def voxel_case(ref, name, cart, grid, max_points, max_voxels, ind=True, den=True):
    polar = ref["transform_points"](cart, "cylinder").astype(np.float32)
    g = synth.GRIDS[grid]
    vg = ref["VoxelGenerator"](g["voxel_size"], g["range"], max_points, max_voxels)
    voxels, coors, num, pind, pden = vg.generate(polar, return_pc_grid_ind=ind, return_density=den)
    out = dict(cart=cart, polar=polar, voxel_size=vg.voxel_size, range=vg.point_cloud_range,
               grid_size=vg.grid_size, max_points=np.int64(max_points), max_voxels=np.int64(max_voxels),
               voxels=voxels, coors=coors, num_points=num)
    if ind:
        out["pc_grid_ind"] = pind
    if den:
        out["den_idx"], out["den_val"], out["den_shape"] = sparse_density(pden)
    np.savez_compressed(os.path.join(HERE, f"voxel_{name}.npz"), **out)
    print(f"voxel_{name}: N={cart.shape[0]} M={voxels.shape[0]} maxnum={num.max() if len(num) else 0}")
    return polar, voxels, coors, num


def main():
    ref = load_reference()
    torch = ref["torch"]
    nusc = synth.nusc_frame(7)
    way = synth.waymo_frame(8, nsweeps=1, time_column=True)
    # contiguous slices keep the beam/azimuth order (adjacent points share cells)
    a = nusc[:9000]
    b = nusc[100000:106000]
    w = way[40000:48000]
    voxel_case(ref, "nusc_pillar", a, "NUSC-PILLAR", 20, 60000)
    voxel_case(ref, "nusc_pillar_cap", a, "NUSC-PILLAR", 4, 1500)        # both caps bind
    voxel_case(ref, "nusc_cyl", b, "NUSC-CYL", 30, 180000)
    voxel_case(ref, "waymo", w, "WAYMO-PARTNER", 5, 150000, den=True)
    rng = np.random.default_rng(3)
    sh = nusc[rng.permutation(nusc.shape[0])[:8000]]                     # shuffled order
    voxel_case(ref, "nusc_shuffled", sh, "NUSC-PILLAR", 20, 3000, ind=False, den=False)
    # edge values: exactly on bin edges / range bounds, +-inf, far points
    g = synth.GRIDS["NUSC-PILLAR"]
    vs = np.array(g["voxel_size"], np.float32)
    lo = np.array(g["range"][:3], np.float32)
    hi = np.array(g["range"][3:], np.float32)
    k = np.arange(0, 513, dtype=np.float32)
    rho = (lo[0] + k * vs[0]).astype(np.float32)
    phi = (lo[1] + k * vs[1]).astype(np.float32)
    pts = []
    for r_, p_ in zip(rho, phi):
        for dz in (lo[2], hi[2], np.float32(0.0), np.nextafter(hi[2], np.float32(-9))):
            pts.append([r_ * np.cos(p_), r_ * np.sin(p_), dz, 1.0, 0.0])
    pts += [[np.inf, 0, 0, 0, 0], [-np.inf, 1, 0, 0, 0], [1e20, 1e20, 0, 0, 0], [0, 0, 0, 0, 0],
            [-0.0, 0.0, 0, 0, 0], [0.0, -0.0, 0, 0, 0], [-1, 0.0, 0, 0, 0], [-1, -0.0, 0, 0, 0]]
    edge = np.array(pts, np.float32)
    # the reference's NaN handling is UB; keep NaN out of the golden set. inf*0 -> nan in
    # transform (inf**2 fine, arctan2 fine) so only rho=inf appears.
    voxel_case(ref, "edges", edge, "NUSC-PILLAR", 3, 60000)

    # ---- transform_points golden (rho bit-exact, phi <= 4 ulp) ----
    t_in = np.concatenate([nusc[::40], way[::40, :5]], axis=0)
    np.savez_compressed(os.path.join(HERE, "transform.npz"), cart=t_in,
                        cylinder=ref["transform_points"](t_in, "cylinder").astype(np.float32),
                        cuboid=ref["transform_points"](t_in, "cuboid").astype(np.float32))

    # ---- readers / scatter golden ----
    frames = []
    gcfg = synth.GRIDS["NUSC-PILLAR"]
    vg = ref["VoxelGenerator"](gcfg["voxel_size"], gcfg["range"], 20, 60000)
    for f, sl in enumerate((slice(0, 5000), slice(150000, 154000))):
        polar = ref["transform_points"](nusc[sl], "cylinder").astype(np.float32)
        v, c, n, _, _ = vg.generate(polar)
        frames.append((v, c, n))
    vox = np.concatenate([f[0] for f in frames])
    coor = np.concatenate([np.pad(f[1], ((0, 0), (1, 0)), constant_values=i) for i, f in enumerate(frames)])
    num = np.concatenate([f[2] for f in frames])
    tv, tc, tn = torch.from_numpy(vox), torch.from_numpy(coor.astype(np.int32)), torch.from_numpy(num)
    out = dict(voxels=vox, coors=coor.astype(np.int32), num_points=num,
               voxel_size=np.array(gcfg["voxel_size"], np.float64), pc_range=np.array(gcfg["range"], np.float64))
    with torch.no_grad():
        vfe = ref["ve"].VoxelFeatureExtractorV3(num_input_features=7)
        out["vfe_mean"] = vfe(tv, tn).numpy()
        for tag, filters, wd in (("pfn64_128", (64, 128), False), ("pfn64", (64,), False),
                                 ("pfn32_32_64_dist", (32, 32, 64), True)):
            torch.manual_seed(0)
            net = ref["pe"].PillarFeatureNet(7, filters, wd, gcfg["voxel_size"], gcfg["range"])
            for l in net.pfn_layers:
                u = l.norm.num_features
                l.norm.running_mean.copy_(torch.randn(u))
                l.norm.running_var.copy_(torch.rand(u) * 1.5 + 0.5)
                l.norm.weight.copy_(torch.randn(u))
                l.norm.bias.copy_(torch.randn(u))
            net.eval()
            feats = net(tv, tn, tc)
            out[f"{tag}_out"] = feats.numpy()
            for i, l in enumerate(net.pfn_layers):
                out[f"{tag}_w{i}"] = l.linear.weight.numpy()
                out[f"{tag}_mean{i}"] = l.norm.running_mean.numpy()
                out[f"{tag}_var{i}"] = l.norm.running_var.numpy()
                out[f"{tag}_gamma{i}"] = l.norm.weight.numpy()
                out[f"{tag}_beta{i}"] = l.norm.bias.numpy()
            out[f"{tag}_eps"] = np.float64(net.pfn_layers[0].norm.eps)
        sc = ref["pe"].PointPillarsScatter(num_input_features=7)
        canvas = sc(torch.from_numpy(out["vfe_mean"]), tc, 2, [512, 512, 1])
        # canvas is mostly zero: store sparse
        cv = canvas.numpy()
        nzi = np.flatnonzero(cv.reshape(-1))
        out["canvas_shape"] = np.array(cv.shape, np.int64)
        out["canvas_idx"] = nzi.astype(np.int64)
        out["canvas_val"] = cv.reshape(-1)[nzi]
        out["bev_index"] = (tc[:, 2].long() * 512 + tc[:, 3].long()).numpy()
    np.savez_compressed(os.path.join(HERE, "readers.npz"), **out)
    print("readers: M=%d" % vox.shape[0])


if __name__ == "__main__":
    main()
