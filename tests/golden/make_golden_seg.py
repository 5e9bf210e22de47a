"""Golden vectors for the segmentation voxel labels, produced by the REFERENCE'S OWN SOURCE.

    python tests/golden/make_golden_seg.py          (build container only: needs /root/reference, numba)

``Voxelization.get_grid_ind`` (det3d/datasets/pipelines/voxelization.py:40-60) and
``AssignLabel.assign_voxel_labels`` (det3d/datasets/pipelines/preprocess.py:170-191) live in modules
whose imports fail here (det3d.core, np.long), so
  * the numba function is exec'ed from its own source lines (decorators included) and JIT-compiled,
  * the method body of ``get_grid_ind`` is exec'ed verbatim as a plain function, with ``AssignLabel``
    bound to a holder of that numba function and ``np.long`` (removed from numpy >= 1.24) mapped to
    ``np.int64``.
``pc_grid_ind`` comes from the reference's own ``VoxelGenerator.generate(return_pc_grid_ind=True)``.
SegHead.predict's per-point lookup (det3d/models/seg_heads/seg_head.py:184-191) is the two indexing
expressions of that method, evaluated with numpy on a formula-defined prediction map (pred_map).  Output: seg.npz.
"""
import os
import sys
import textwrap

import numba
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import make_golden as mg  # noqa: E402
from partner_b200 import synth  # noqa: E402


class _NP:
    """numpy with the removed ``np.long`` alias restored."""
    long = np.int64

    def __getattr__(self, name):
        return getattr(np, name)


def load_assign_voxel_labels():
    src = open(f"{mg.REF}/det3d/datasets/pipelines/preprocess.py").read().splitlines()
    start = next(i for i, l in enumerate(src) if l.strip().startswith("def assign_voxel_labels"))
    end = next(i for i in range(start + 1, len(src)) if src[i].strip().startswith("def "))
    body = textwrap.dedent("\n".join(src[start:end]))
    ns = {"np": np, "numba": numba}
    exec("@numba.njit(cache=False, parallel=False)\n" + body, ns)        # the decorators of :168-169
    return ns["assign_voxel_labels"]


def load_get_grid_ind(assign):
    src = open(f"{mg.REF}/det3d/datasets/pipelines/voxelization.py").read().splitlines()
    start = next(i for i, l in enumerate(src) if l.strip().startswith("def get_grid_ind"))
    end = next(i for i in range(start + 1, len(src)) if src[i].strip().startswith("def "))
    body = textwrap.dedent("\n".join(src[start:end]))

    class AssignLabel:
        assign_voxel_labels = staticmethod(assign)
    ns = {"np": _NP(), "AssignLabel": AssignLabel}
    exec(body, ns)
    return ns["get_grid_ind"]


def pred_map(nz, ny, nx):
    """Deterministic stand-in for argmax(seg_preds) + 1: int64 [nz, ny, nx] (tests rebuild it)."""
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    return ((z * 7 + y * 3 + x * 5 + (x * y) // 11) % 16 + 1).astype(np.int64)


def main():
    ref = mg.load_reference()
    assign = load_assign_voxel_labels()
    get_grid_ind = load_get_grid_ind(assign)
    out = {}
    rng = np.random.default_rng(7)
    cases = {"pillar": ("NUSC-PILLAR", 30000), "cyl": ("NUSC-CYL", 20000)}
    for name, (tag, npts) in cases.items():
        g = synth.GRIDS[tag]
        vg = ref["VoxelGenerator"](g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
        cart = synth.nusc_frame(51 if name == "pillar" else 52)[::9][:npts]
        polar = ref["transform_points"](cart, "cylinder").astype(np.float32)
        _, _, _, pc_grid_ind, _ = vg.generate(polar, return_pc_grid_ind=True)
        n = polar.shape[0]
        # labels 0..16 with a spatial pattern (so cells have a real majority), 15 % unlabelled (-1), ties on purpose
        label = ((pc_grid_ind[:, 1] // 37 + pc_grid_ind[:, 2] // 53) % 17).astype(np.int64)
        noise = rng.random(n)
        label[noise < 0.25] = rng.integers(0, 17, int((noise < 0.25).sum()))
        label[rng.random(n) < 0.15] = -1
        label = label.reshape(-1, 1)
        res = {"mode": "train", "lidar": {"pc_label": label, "voxels": {}}}
        res = get_grid_ind(None, res, pc_grid_ind, vg.grid_size)
        out[f"{name}_grid_ind"] = pc_grid_ind.astype(np.int32)
        out[f"{name}_label"] = label.astype(np.int32)
        out[f"{name}_grid_size"] = vg.grid_size
        lab = res["lidar"]["voxels"]["labels"]
        assert lab.shape == (1,) + tuple(int(v) for v in vg.grid_size[::-1]) and lab.dtype == np.int64
        nzc = np.nonzero(lab.reshape(-1))[0]                  # the dense map is mostly zeros: store it sparse
        out[f"{name}_labels_nz_index"] = nzc.astype(np.int64)
        out[f"{name}_labels_nz_value"] = lab.reshape(-1)[nzc]
        out[f"{name}_valid_grid_ind"] = res["lidar"]["voxels"]["valid_grid_ind"].astype(np.int32)
        # evaluation branch (:56): the first n_key_points rows
        res_e = {"mode": "val", "lidar": {"n_key_points": n // 3, "voxels": {}}}
        res_e = get_grid_ind(None, res_e, pc_grid_ind, vg.grid_size)
        out[f"{name}_eval_valid_grid_ind"] = res_e["lidar"]["voxels"]["valid_grid_ind"].astype(np.int32)
        # SegHead.predict lookups (seg_head.py:186-191)
        v = res["lidar"]["voxels"]["valid_grid_ind"]
        nx, ny, nz = (int(t) for t in vg.grid_size)
        pred = pred_map(nz, ny, nx)                           # a formula, so the map itself is not stored
        if nz == 1:
            pred3 = pred[0]                                   # ndim == 3 (:188): [0, y, x] after unsqueeze(1)
            out[f"{name}_point_preds"] = pred3[None][0, v[:, 1], v[:, 2]]
        else:
            out[f"{name}_point_preds"] = pred[v[:, 0], v[:, 1], v[:, 2]]
        print(name, "points", n, "valid", v.shape[0], "labelled cells", nzc.shape[0])
    # a cell with > 65535 points of one label: the uint16 counter wraps (preprocess.py:178)
    gi = np.zeros((70000 + 5, 3), np.int32)
    gi[:, 1] = 3
    gi[:, 2] = 4
    lab = np.full((70000 + 5, 1), 9, np.int64)
    lab[70000:] = 2                                           # 5 points of label 2 beat 70000 mod 65536 = 4464? no: 4464 > 5
    gi2 = gi.copy()
    lab2 = lab.copy()
    lab2[65536 + 3:70000] = 2                                 # label 9: 65539 -> wraps to 3; label 2: 4466
    for tag, (g_, l_) in {"wrap_a": (gi, lab), "wrap_b": (gi2, lab2)}.items():
        res = {"mode": "train", "lidar": {"pc_label": l_, "voxels": {}}}
        res = get_grid_ind(None, res, g_, np.array([8, 8, 1]))
        out[f"{tag}_grid_ind"] = g_
        out[f"{tag}_label"] = l_.astype(np.int32)
        out[f"{tag}_labels"] = res["lidar"]["voxels"]["labels"]
        print(tag, "label of the cell:", int(res["lidar"]["voxels"]["labels"][0, 0, 3, 4]))
    np.savez_compressed(os.path.join(HERE, "seg.npz"), **out)


if __name__ == "__main__":
    main()
