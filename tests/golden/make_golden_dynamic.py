"""Golden vectors for the dynamic-voxelization rows, produced FROM THE REFERENCE ITSELF.

    python tests/golden/make_golden_dynamic.py         (build container only: needs /root/reference)

Runs, on seeded synthetic clouds,
  * the grid-index expression of ``Voxelization.voxelize_dynamic``
    (det3d/datasets/pipelines/voxelization.py:169-172; the three source lines are exec'ed as they
    stand, with ``np.int`` -- removed from numpy >= 1.24 -- mapped to ``int``),
  * the batch-index padding of ``collate_kitti`` (torchie/parallel/collate.py:157-164),
  * ``DynamicVoxelEncoderV1.forward`` (models/readers/voxel_encoder.py:38-44) and
    ``DynamicPPScatter.forward`` (models/readers/pillar_encoder.py:413-432), unmodified.
``torch_scatter`` is a third-party module the reference neither vendors nor pins; the stub below
implements ``scatter_mean`` with its documented semantics (sum / count), which is the only thing
DynamicVoxelEncoderV1 takes from it.  Output: tests/golden/dynamic.npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import make_golden as mg  # noqa: E402
from partner_b200 import synth  # noqa: E402


class _NP:
    """numpy with the removed ``np.int`` alias restored (the reference was written for numpy < 1.24)."""
    int = int

    def __getattr__(self, name):
        return getattr(np, name)


def reference_grid_ind(points, pc_range, voxel_size, grid_size):
    src = open(f"{mg.REF}/det3d/datasets/pipelines/voxelization.py").read().splitlines()
    start = next(i for i, l in enumerate(src) if "pc_grid_ind = (np.floor(" in l and i > 140)
    stmt = "\n".join(l.strip() for l in src[start:start + 3])
    ns = {"np": _NP(), "points": points, "pc_range": pc_range, "voxel_size": voxel_size, "grid_size": grid_size}
    exec(stmt, ns)
    return ns["pc_grid_ind"]


def main():
    ref = mg.load_reference()
    torch = ref["torch"]
    import torch_scatter

    def scatter_mean(src, index, dim=0):
        m = int(index.max()) + 1
        out = torch.zeros((m,) + tuple(src.shape[1:]), dtype=src.dtype)
        out.index_add_(0, index, src)
        cnt = torch.zeros(m, dtype=src.dtype).index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        return out / cnt.clamp(min=1).view(-1, *([1] * (src.dim() - 1)))
    torch_scatter.scatter_mean = scatter_mean

    def scatter_max(src, index, dim=0):
        m = int(index.max()) + 1
        out = torch.full((m,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype)
        out = out.scatter_reduce(0, index.view(-1, *([1] * (src.dim() - 1))).expand_as(src), src, "amax", include_self=True)
        return out, None
    torch_scatter.scatter_max = scatter_max

    g = synth.GRIDS["NUSC-PILLAR"]
    vg = ref["VoxelGenerator"](g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
    nusc = synth.nusc_frame(21)
    frames = [nusc[:7000], nusc[120000:125000], np.zeros((0, 5), np.float32), nusc[250000:253000]]
    # out-of-range / edge points are CLAMPED into border cells by the dynamic path
    frames[1] = np.concatenate([frames[1], np.array([[200.0, 3.0, 9.0, 1.0, 0.0], [0.05, 0.0, -9.0, 2.0, 0.0],
                                                     [-60.0, -1e-3, 0.0, 3.0, 0.0], [3.0e4, 1.0, 0.0, 4.0, 0.0]], np.float32)])
    polars = [ref["transform_points"](f, "cylinder").astype(np.float32) for f in frames]
    ginds = [reference_grid_ind(p, vg.point_cloud_range, vg.voxel_size, vg.grid_size) for p in polars]
    gi4 = np.concatenate([np.pad(gi, ((0, 0), (1, 0)), mode="constant", constant_values=i)
                          for i, gi in enumerate(ginds)], axis=0)          # collate.py:157-164
    pts = np.concatenate(polars, axis=0)
    with torch.no_grad():
        enc = ref["ve"].DynamicVoxelEncoderV1(num_input_features=7)
        feats, unq = enc(dict(points=torch.from_numpy(pts), grid_ind=torch.from_numpy(gi4.astype(np.int64))))
        _, inv, cnt = torch.unique(torch.from_numpy(gi4.astype(np.int64)), return_inverse=True, return_counts=True, dim=0)
        canvas = ref["pe"].DynamicPPScatter()(feats, unq, len(frames), [512, 512, 1]).numpy()
    # ---- DynamicPFNet (pillar_encoder.py:262-411): the polarstream config (all four decorations,
    # [64, 128] filters, reader left at voxel_shape='cuboid') and a cylinder-shaped single-layer one
    pfn = {}
    gl = torch.from_numpy(gi4.astype(np.int64))
    with torch.no_grad():
        for tag, shape, filters, flags in (("pfn_cuboid_64_128", "cuboid", (64, 128), dict(xyz_cluster=True, raz_cluster=True, xy_center=True, ra_center=True)),
                                           ("pfn_cyl_64", "cylinder", (64,), dict(xyz_cluster=True, raz_cluster=True, xy_center=True, ra_center=True)),
                                           ("pfn_cyl_raz_32_64", "cylinder", (32, 64), dict(raz_cluster=True, ra_center=True))):
            torch.manual_seed(1)
            net = ref["pe"].DynamicPFNet(num_input_features=7, num_filters=filters, voxel_shape=shape,
                                         voxel_size=g["voxel_size"], pc_range=g["range"], **flags)
            net.eval()
            o, u = net(dict(points=torch.from_numpy(pts), grid_ind=gl))
            pfn[f"{tag}_out"] = o.numpy()
            for i, l in enumerate(net.pfn_layers):
                pfn[f"{tag}_w{i}"] = l.linear.weight.numpy()
            assert np.array_equal(u.numpy(), unq.numpy())
    nzi = np.flatnonzero(canvas.reshape(-1))
    out = dict(voxel_size=vg.voxel_size, range=vg.point_cloud_range, grid_size=vg.grid_size,
               sizes=np.array([f.shape[0] for f in frames], np.int64), cart=np.concatenate(frames), polar=pts,
               grid_ind=gi4.astype(np.int32), features=feats.numpy(), unq=unq.numpy().astype(np.int32),
               unq_inv=inv.numpy(), unq_cnt=cnt.numpy(), canvas_shape=np.array(canvas.shape, np.int64),
               canvas_idx=nzi.astype(np.int64), canvas_val=canvas.reshape(-1)[nzi], **pfn)
    np.savez_compressed(os.path.join(HERE, "dynamic.npz"), **out)
    print("dynamic: N=%d M=%d frames=%d" % (pts.shape[0], feats.shape[0], len(frames)))


if __name__ == "__main__":
    main()
