#!/usr/bin/env python
"""Throughput of the dynamic-voxelization path (pv_dynamic_voxelize: bins + unique + scatter_mean +
DynamicPPScatter canvas) on the headline batch: nuScenes 10-sweep frames, NUSC-PILLAR grid, B = 8.
CUDA-graph replay, CUDA events, 4 rotating input sets.  Usage: tools/bench_dynamic.py [steps]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from partner_b200 import synth  # noqa: E402
from partner_b200 import functional as F  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
g = synth.GRIDS["NUSC-PILLAR"]
cfg = F.make_config(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])[0]
dev = torch.device("cuda", 0)
B, sets = 8, []
for s in range(4):
    frames = synth.make_batch("nusc", 2, B, first_frame=s * B)
    sizes = [f.shape[0] for f in frames]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    sets.append((torch.from_numpy(np.concatenate(frames)).to(dev), torch.from_numpy(off).to(dev), max(sizes), int(off[-1])))
cap = max(s[2] for s in sets)
graphs = []
for k, (pts, off, _, n) in enumerate(sets):
    for inverse in (False,):
        F.dynamic_voxelize(cfg, pts, off, B, cap, True, want_inverse=inverse, canvas=True, ws_tag=k)   # warm-up
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            out = F.dynamic_voxelize(cfg, pts, off, B, cap, True, want_inverse=inverse, canvas=True, ws_tag=k)
        graphs.append((gr, out))
torch.cuda.synchronize()
for k in range(8):
    graphs[k % 4][0].replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(steps):
    graphs[k % 4][0].replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
npts = np.mean([s[3] for s in sets])
m = np.mean([int(o.voxel_counts.sum()) for _, o in graphs])
print(json.dumps({"workload": "nusc_pillar_dynamic_mean_canvas_b8", "ms_per_step": ms, "Mpoints_per_s": npts / ms / 1e3,
                  "points": npts, "voxels": m, "launch": "cuda-graph replay, single stream"}))

# ---- DynamicPFNet (polarstream reader config: 16-d decoration, Linear 16->32, 64->128) on the same batch ----
rng = np.random.default_rng(0)
ws = [torch.from_numpy(rng.normal(0, 0.2, (32, 16)).astype(np.float32)).to(dev),
      torch.from_numpy(rng.normal(0, 0.1, (128, 64)).astype(np.float32)).to(dev)]
vx, vy = g["voxel_size"][0], g["voxel_size"][1]
pts, off, _, n = sets[0]
polar = F.transform_points(pts)
r = F.dynamic_voxelize(cfg, pts, off, B, cap, True, ws_tag=9)
m = r.total()
run = lambda: F.dynamic_pfn(polar, r, m, ws, vx, vy, vx / 2 + g["range"][0], vy / 2 + g["range"][1], False, True, True, True, True)
for _ in range(3):
    run()
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
flop = 2.0 * n * (16 * 32 * 2 + 64 * 128)          # layer 1 runs twice (pass A and B)
print(json.dumps({"workload": "nusc_dynamic_pfnet_b8", "ms_per_step": ms, "Mpoints_per_s": n / ms / 1e3, "voxels": m,
                  "fp32_TFLOPs": flop / ms / 1e9}))
