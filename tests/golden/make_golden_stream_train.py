"""Golden vectors for the TRAINING-time ground truth of azimuth-sector streaming, produced by the reference's
own source lines.

    python tests/golden/make_golden_stream_train.py        (build container only: needs /root/reference)

Exec'ed verbatim: ``filter_gt`` + ``_dict_select`` (det3d/datasets/pipelines/utils.py:3-27),
``rotation_points_single_angle`` (det3d/core/bbox/box_np_ops.py:182-204) and the per-sector ground-truth block
of ``Voxelization.voxelize_streaming_polar`` (det3d/datasets/pipelines/voxelization.py:332-349), on synthetic
annotation dicts (boxes with velocities, some exactly on wedge boundaries / range limits).  Output:
stream_train.npz -- per sector the surviving box indices and the rotated boxes.
"""
import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load():
    ns = {"np": np, "prep": None}
    src = open(f"{REF}/det3d/datasets/pipelines/utils.py").read().splitlines()
    a = next(i for i, l in enumerate(src) if l.startswith("def _dict_select"))
    b = next(i for i, l in enumerate(src) if l.startswith("def drop_arrays_by_name"))
    exec("\n".join(src[a:b]), ns)
    box = {"np": np}
    bsrc = open(f"{REF}/det3d/core/bbox/box_np_ops.py").read().splitlines()
    a = next(i for i, l in enumerate(bsrc) if l.startswith("def rotation_points_single_angle"))
    b = next(i for i in range(a + 1, len(bsrc)) if bsrc[i].startswith("def "))
    exec("\n".join(bsrc[a:b]), box)

    class BoxOps:
        rotation_points_single_angle = staticmethod(box["rotation_points_single_angle"])
    ns["box_np_ops"] = BoxOps
    vsrc = open(f"{REF}/det3d/datasets/pipelines/voxelization.py").read().splitlines()
    a = next(i for i, l in enumerate(vsrc) if i > 320 and l.strip() == 'if res["mode"] in ["train", "debug_gt"]:' and "filter_gt(cur_res, cur_pc_range)" in vsrc[i + 1])
    b = next(i for i in range(a, len(vsrc)) if vsrc[i].strip() == "cur_res[\"lidar\"][\"annotations\"]['gt_boxes'] = gt_boxes")
    indent = len(vsrc[a]) - len(vsrc[a].lstrip())
    block = "\n".join(l[indent:] if l.strip() else l for l in vsrc[a:b + 1])
    return ns, block


def main():
    ns, block = load()
    rng = np.random.default_rng(21)
    pc_range = np.array([0.3, -3.1488, -5.0, 50.476, 3.1488, 3.0], np.float32)
    n = 120
    rho = rng.uniform(0.0, 60.0, n)
    az = rng.uniform(-np.pi, np.pi, n)
    az[:6] = [-3.1488, -1.5744, 0.0, 1.5744, 3.1488, np.pi]            # on wedge boundaries / range limits
    rho[6:9] = [0.3, 50.476, 0.2999]
    boxes = np.zeros((n, 9), np.float32)
    boxes[:, 0], boxes[:, 1] = rho * np.cos(az), rho * np.sin(az)
    boxes[:, 2] = rng.uniform(-3, 1, n)
    boxes[:, 3:6] = rng.uniform(0.5, 5.0, (n, 3))
    boxes[:, 6:8] = rng.normal(0, 3, (n, 2))
    boxes[:, 8] = rng.uniform(-np.pi, np.pi, n)
    names = np.array(["car", "truck", "pedestrian"])[rng.integers(0, 3, n)]
    out = dict(gt_boxes=boxes, gt_names=names, pc_range=pc_range)
    for nsec in (1, 4, 8):
        min_az, max_az = pc_range[1], pc_range[4]
        interval = (max_az - min_az) / nsec
        for i in range(nsec):
            cur_pc_range = pc_range.copy()
            cur_pc_range[1] = min_az + i * interval
            cur_pc_range[4] = min_az + (i + 1) * interval
            res = {"mode": "train", "voxel_shape": "cylinder",
                   "lidar": {"annotations": {"gt_boxes": boxes.copy(), "gt_names": names.copy(),
                                             "gt_index": np.arange(n)}}}
            env = dict(ns)
            env.update(res=res, cur_res=copy.deepcopy(res), cur_pc_range=cur_pc_range, pc_range=pc_range)
            exec(block, env)
            ann = env["cur_res"]["lidar"]["annotations"]
            out[f"n{nsec}_s{i}_index"] = ann["gt_index"].astype(np.int64)
            out[f"n{nsec}_s{i}_boxes"] = ann["gt_boxes"]
        print(nsec, [int(out[f"n{nsec}_s{i}_index"].shape[0]) for i in range(nsec)])
    np.savez_compressed(os.path.join(HERE, "stream_train.npz"), **out)


if __name__ == "__main__":
    main()
