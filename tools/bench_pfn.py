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
from partner_b200 import _lib  # noqa: E402
if os.environ.get("PV_LIB"):                     # development aid: time another build of the library
    _lib.SO_PATH = os.path.abspath(os.environ["PV_LIB"])
from partner_b200 import PillarFeatureNet, PointPillarsScatter, synth  # noqa: E402
from partner_b200 import functional as F  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--filters", default="64,128")
ap.add_argument("--streams", type=int, default=1, help="replay the fused front end round-robin on this many streams (independent batches overlap)")
ap.add_argument("--fused-only", action="store_true", help="skip the three-call path (for ncu launch lists of the fused front end)")
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
for it in range(0 if args.fused_only else args.iters + 2):
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
if args.fused_only:
    tot[:] = 1.0
# fused front end: pv_forward_pfn_canvas (no [M, T, C] tensor), CUDA-graph replay on rotating input sets
from partner_b200 import PillarFrontEnd  # noqa: E402
n_sets = 3
fused_sets = []
for k in range(n_sets):
    fr = synth.make_batch("nusc", 3, args.batch, first_frame=k * args.batch)
    sz = [f.shape[0] for f in fr]
    of = np.zeros(args.batch + 1, np.int32)
    np.cumsum(sz, out=of[1:])
    fused_sets.append((torch.from_numpy(np.concatenate(fr)).to(dev), torch.from_numpy(of).to(dev), max(sz)))
cap_all = max(s_[2] for s_ in fused_sets)
fes, graphs, outs = [], [], []
for k, (p_, o_, _) in enumerate(fused_sets):
    fe = PillarFrontEnd(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"], net, cartesian=True, device=dev,
                        workspace_tag=k)
    out = fe.forward_device(p_, o_, args.batch, cap_all)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fe.forward_device(p_, o_, args.batch, cap_all, out=out)
    fes.append(fe); graphs.append(gr); outs.append(out)
for gr in graphs:
    gr.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 4 * args.iters
streams = [torch.cuda.Stream(dev) for _ in range(args.streams)] if args.streams > 1 else None
e0.record()
if streams is None:
    for it in range(reps):
        graphs[it % n_sets].replay()
else:
    cur = torch.cuda.current_stream(dev)
    for st in streams:
        st.wait_stream(cur)
    for it in range(reps):
        with torch.cuda.stream(streams[it % args.streams]):
            graphs[it % n_sets].replay()
    for st in streams:
        cur.wait_stream(st)
e1.record()
torch.cuda.synchronize()
fused_ms = e0.elapsed_time(e1) / reps
fused_pts = float(np.mean([int(s_[1][-1].item()) for s_ in fused_sets]))
fused_m = float(np.mean([int(o.voxel_counts.sum().item()) for o in outs]))
n = int(off[-1])
if args.fused_only:
    m = int(outs[0].voxel_counts.sum().item())
    K = int(outs[0].num_points[:m].sum().item())
    nonfull = int((outs[0].num_points[:m] < g["max_points"]).sum().item())
else:
    K = int(vb.num_points[:m].sum().item())
    nonfull = int((vb.num_points[:m] < g["max_points"]).sum().item())
macs = sum(a * b for a, b in zip([12] + [f for f in filters[:-1]], [f // 2 for f in filters[:-1]] + [filters[-1]]))
print(json.dumps({"workload": "nusc_pillar_pfn_canvas_b%d" % args.batch, "points": n, "voxels": m, "kept_points": K,
                  "useful_rows": K + nonfull, "ms": {"voxelize_with_voxels_tensor": tot[0], "pfn": tot[1], "scatter": tot[2]},
                  "ms_total": float(tot.sum()), "Mpoints_per_s": n / tot.sum() / 1e3, "frames_per_s": args.batch / tot.sum() * 1e3,
                  "pfn_useful_tflops": 2.0 * (K + nonfull) * macs / (tot[1] * 1e-3) / 1e12,
                  "canvas_GBps": 4.0 * filters[-1] * 512 * 512 * args.batch / (tot[2] * 1e-3) / 1e9,
                  "fused": {"ms_per_step": fused_ms, "Mpoints_per_s": fused_pts / fused_ms / 1e3,
                            "frames_per_s": args.batch / fused_ms * 1e3,
                            "algorithmic_MB": (fused_pts * 20 + fused_m * (20 + 4 * filters[-1]) + 4.0 * filters[-1] * 512 * 512 * args.batch) / 1e6,
                            "hbm_frac_of_6551": (fused_pts * 20 + fused_m * (20 + 4 * filters[-1]) + 4.0 * filters[-1] * 512 * 512 * args.batch)
                                                / (fused_ms * 1e-3) / 6551e9,
                            "note": "pv_forward_pfn_canvas, CUDA-graph replay, 3 rotating input sets, single stream"}}))
