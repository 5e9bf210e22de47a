#!/usr/bin/env python
"""Per-stage CUDA-event times of the fused front end on the bench workload (experiment aid).
Usage: tools/stage_times.py [workload] [reps]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from partner_b200 import PolarFrontEnd, synth, _lib  # noqa: E402
from partner_b200 import functional as F  # noqa: E402
from partner_b200._lib import ptr, current_stream  # noqa: E402
import bench  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "nusc_pillar_mean_canvas_b8"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
grid, kind, kw, per_gpu, cfg_id, has_canvas = bench.WORKLOADS[wl]
g = synth.GRIDS[grid]
dev = torch.device("cuda", 0)
fe = PolarFrontEnd(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"], cartesian=True, device=dev)
frames = synth.make_batch(kind, cfg_id, per_gpu, **kw)
sizes = [f.shape[0] for f in frames]
off = np.zeros(per_gpu + 1, np.int32)
np.cumsum(sizes, out=off[1:])
pts = torch.from_numpy(np.concatenate(frames)).to(dev)
d_off = torch.from_numpy(off).to(dev)
out = fe.forward_device(pts, d_off, per_gpu, max(sizes))
torch.cuda.synchronize()
lib = _lib.load()
names = bench.STAGES_BY_PIPELINE[lib.pv_profile_pipeline(fe.cfg)]
ms = (ctypes.c_float * len(names))()
for _ in range(2):
    F.check(lib.pv_profile_mean_canvas(fe.cfg, ptr(pts), ptr(d_off), per_gpu, int(off[-1]), pts.shape[1], 1,
                                       out.n_cap, out.f_cap, ptr(out.ws), out.ws.numel(), ptr(out.coors),
                                       ptr(out.num_points), ptr(out.voxel_counts), ptr(out.mean_feats),
                                       ptr(out.canvas), current_stream(dev), reps, ms), "profile")
print("PV_PIPELINE=%s %s: " % (os.environ.get("PV_PIPELINE", "0"), wl) +
      "  ".join("%s %.1f" % (n, v * 1e3) for n, v in zip(names, ms)) + "  | sum %.1f us" % (sum(ms) * 1e3))
