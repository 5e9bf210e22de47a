"""Golden vectors for azimuth-sector streaming, produced by the REFERENCE'S OWN SOURCE LINES.

    python tests/golden/make_golden_stream.py          (build container only: needs /root/reference)

``Voxelization.voxelize_streaming_polar`` (det3d/datasets/pipelines/voxelization.py:305-371) cannot
be called as a method here (its module imports numba-era numpy aliases and det3d packages that do
not import), so the statements of its per-sector loop body that touch the points -- the index
selection (:350-358), the azimuth shift and x / y recomputation (:360-362) and the grid index
(:366-368) -- are exec'ed verbatim from the source file with the variables the method defines
before them.  ``np.int`` (removed from numpy >= 1.24) is mapped to ``int``.  Output: stream.npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import make_golden as mg  # noqa: E402
from make_golden_dynamic import _NP  # noqa: E402
from partner_b200 import synth  # noqa: E402


def reference_sector(src, points, pc_range, voxel_size, grid_size, nsectors, i):
    """Runs the reference's own statements for sector i; returns (points, pc_grid_ind, points_index)."""
    class Self:
        pass
    self = Self()
    self.nsectors = nsectors
    npx = _NP()
    min_az, max_az = pc_range[1], pc_range[4]                       # :314
    interval = (max_az - min_az) / nsectors                          # :315
    cur_grid_size = grid_size.copy()                                 # :316-317
    cur_grid_size[1] //= nsectors
    cur_pc_range = pc_range.copy()                                   # :328-330
    cur_pc_range[1] = min_az + i * interval
    cur_pc_range[4] = min_az + (i + 1) * interval
    ns = dict(np=npx, self=self, i=i, points=points.copy(), cur_pc_range=cur_pc_range, pc_range=pc_range,
              voxel_size=voxel_size, cur_grid_size=cur_grid_size)
    start = next(k for k, l in enumerate(src) if l.strip() == "if i == 0:" and k > 340 and "points[:, 1] < cur_pc_range[4]" in src[k + 1])
    end = next(k for k in range(start, len(src)) if src[k].strip().startswith("np.int)[:, ::-1]"))
    body = [l for l in src[start:end + 1] if 'cur_res["lidar"]["points"] = points' not in l]
    indent = len(src[start]) - len(src[start].lstrip())
    exec("\n".join(l[indent:] if l.strip() else l for l in body), ns)
    return ns["points"], ns["pc_grid_ind"], ns["points_index"], cur_grid_size


def main():
    ref = mg.load_reference()
    src = open(f"{mg.REF}/det3d/datasets/pipelines/voxelization.py").read().splitlines()
    g = synth.GRIDS["NUSC-PILLAR"]
    vg = ref["VoxelGenerator"](g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
    cart = synth.nusc_frame(41)[::13][:20000]
    polar = ref["transform_points"](cart, "cylinder").astype(np.float32)
    # values exactly on the sector boundaries and outside the azimuth range
    extra = polar[:16].copy()
    lo, hi = vg.point_cloud_range[1], vg.point_cloud_range[4]
    iv = (hi - lo) / 4
    extra[:8, 1] = [lo, lo + iv, lo + 2 * iv, lo + 3 * iv, hi, np.nextafter(lo + iv, np.float32(-9)), -3.2, 3.2]
    polar = np.concatenate([polar, extra])
    out = dict(polar=polar, voxel_size=vg.voxel_size, range=vg.point_cloud_range, grid_size=vg.grid_size)
    for nsec in (1, 4, 8):
        for i in range(nsec):
            pts, gi, idx, cgs = reference_sector(src, polar, vg.point_cloud_range, vg.voxel_size, vg.grid_size, nsec, i)
            out[f"n{nsec}_s{i}_points"] = pts
            out[f"n{nsec}_s{i}_grid_ind"] = gi.astype(np.int32)
            out[f"n{nsec}_s{i}_index"] = idx.astype(np.int32)
        out[f"n{nsec}_grid"] = cgs
        print("nsectors", nsec, [int(out[f"n{nsec}_s{i}_index"].shape[0]) for i in range(nsec)])
    np.savez_compressed(os.path.join(HERE, "stream.npz"), **out)


if __name__ == "__main__":
    main()
