#!/usr/bin/env python
"""Config 3 of BASELINE.json: nuScenes polar pillars -> PFN (linear + BN + ReLU + max) -> scatter to the
512x512 polar BEV canvas, batch 16, through the drop-in modules.  Prints per-stage CUDA-event times."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partner_b200 import PillarFeatureNet, PointPillarsScatter, synth  # noqa: E402
from partner_b200 import functional as F  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--filters", default="64,128")
args = ap.parse_args()
dev = torch.device("cuda", 0)
g = synth.GRIDS["NUSC-PILLAR"]
frames = synth.make_batch("nusc", 3, args.batch)
sizes = [f.shape[0] for f in frames]
off = np.zeros(args.batch + 1, np.int32)
np.cumsum(sizes, out=off[1:])
pts = torch.from_numpy(np.concatenate(frames)).to(dev)
d_off = torch.from_numpy(off).to(dev)
cfg, _, _, grid = F.make_config(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
filters = tuple(int(x) for x in args.filters.split(","))
torch.manual_seed(0)
net = PillarFeatureNet(7, filters, False, g["voxel_size"], g["range"]).to(dev).eval()
for l in net.pfn_layers:
    u = l.norm.num_features
    l.norm.running_mean.copy_(torch.randn(u))
    l.norm.running_var.copy_(torch.rand(u) * 1.5 + 0.5)
    l.norm.weight.data.copy_(torch.randn(u))
    l.norm.bias.data.copy_(torch.randn(u))
scat = PointPillarsScatter(num_input_features=filters[-1])

ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
tot = np.zeros(3)
for it in range(args.iters + 2):
    ev[0].record()
    vb = F.voxelize(cfg, pts, d_off, args.batch, max(sizes), True, want_voxels=True)
    ev[1].record()
    m = int(vb.voxel_counts.sum().item()) if it == 0 else m
    feats = net(vb.voxels[:m], vb.num_points[:m], vb.coors[:m])
    ev[2].record()
    canvas = scat(feats, vb.coors[:m], args.batch, [int(grid[0]), int(grid[1]), 1])
    ev[3].record()
    torch.cuda.synchronize()
    if it >= 2:
        tot += [ev[k].elapsed_time(ev[k + 1]) for k in range(3)]
    del canvas
tot /= args.iters
n = int(off[-1])
K = int(vb.num_points[:m].sum().item())
nonfull = int((vb.num_points[:m] < g["max_points"]).sum().item())
macs = sum(a * b for a, b in zip([12] + [f for f in filters[:-1]], [f // 2 for f in filters[:-1]] + [filters[-1]]))
print(json.dumps({"workload": "nusc_pillar_pfn_canvas_b%d" % args.batch, "points": n, "voxels": m, "kept_points": K,
                  "useful_rows": K + nonfull, "ms": {"voxelize_with_voxels_tensor": tot[0], "pfn": tot[1], "scatter": tot[2]},
                  "ms_total": float(tot.sum()), "Mpoints_per_s": n / tot.sum() / 1e3, "frames_per_s": args.batch / tot.sum() * 1e3,
                  "pfn_useful_tflops": 2.0 * (K + nonfull) * macs / (tot[1] * 1e-3) / 1e12,
                  "canvas_GBps": 4.0 * filters[-1] * 512 * 512 * args.batch / (tot[2] * 1e-3) / 1e9}))
