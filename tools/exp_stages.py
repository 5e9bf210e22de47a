#!/usr/bin/env python
"""Experiment aid: per-stage CUDA-event times of the fused front end with ROTATING input sets (cold
inputs, like bench.py), optionally with ONE workspace shared by all sets (L2-residency experiment).
Usage: tools/exp_stages.py [--lib path.so] [--share-ws] [--sets 4] [--reps 20] [--workload name]"""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=None)
ap.add_argument("--share-ws", action="store_true")
ap.add_argument("--sets", type=int, default=4)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--workload", default="nusc_pillar_mean_canvas_b8")
args = ap.parse_args()

import torch  # noqa: E402
from partner_b200 import PolarFrontEnd, synth, _lib  # noqa: E402
if args.lib:
    _lib.SO_PATH = os.path.abspath(args.lib)
from partner_b200 import functional as F  # noqa: E402
from partner_b200._lib import ptr, current_stream  # noqa: E402
import bench  # noqa: E402

grid, kind, kw, per_gpu, cfg_id, has_canvas = bench.WORKLOADS[args.workload]
g = synth.GRIDS[grid]
dev = torch.device("cuda", 0)
lib = _lib.load()
sets = []
for s in range(args.sets):
    frames = synth.make_batch(kind, cfg_id, per_gpu, first_frame=s * per_gpu, **kw)
    sizes = [f.shape[0] for f in frames]
    off = np.zeros(per_gpu + 1, np.int32)
    np.cumsum(sizes, out=off[1:])
    sets.append(dict(pts=torch.from_numpy(np.concatenate(frames)).to(dev), off=torch.from_numpy(off).to(dev),
                     n=int(off[-1]), cap=max(sizes)))
cap_all = max(s["cap"] for s in sets)
for k, s in enumerate(sets):
    fe = PolarFrontEnd(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"], cartesian=True, device=dev,
                       workspace_tag=0 if args.share_ws else k)
    s["fe"] = fe
    s["out"] = fe.forward_device(s["pts"], s["off"], per_gpu, cap_all)
torch.cuda.synchronize()
names = bench.STAGES_BY_PIPELINE[lib.pv_profile_pipeline(sets[0]["fe"].cfg)]
tot = np.zeros(len(names))
ms = (ctypes.c_float * len(names))()
for rep in range(args.reps + 2):
    for s in sets:
        o = s["out"]
        F.check(lib.pv_profile_mean_canvas(s["fe"].cfg, ptr(s["pts"]), ptr(s["off"]), per_gpu, s["n"], s["pts"].shape[1], 1,
                                           o.n_cap, o.f_cap, ptr(o.ws), o.ws.numel(), ptr(o.coors), ptr(o.num_points),
                                           ptr(o.voxel_counts), ptr(o.mean_feats), ptr(o.canvas) if o.canvas is not None else ptr(None),
                                           current_stream(dev), 1, ms), "profile")
        if rep >= 2:
            tot += np.array(list(ms))
tot /= args.reps * len(sets)
print("%-28s %s share_ws=%d: " % (os.path.basename(args.lib or "default"), args.workload, args.share_ws) +
      "  ".join("%s %.1f" % (n, v * 1e3) for n, v in zip(names, tot)) + "  | sum %.1f us" % (tot.sum() * 1e3))
