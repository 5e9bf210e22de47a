import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from partner_b200 import PillarFeatureNet, PillarFrontEnd, synth
g = synth.GRIDS["NUSC-PILLAR"]
frames = [synth.nusc_frame(3300, nsweeps=1)[:20000], synth.nusc_frame(3301, nsweeps=1)[:777]]
net = PillarFeatureNet(7, (64, 128), False, tuple(g["voxel_size"]), tuple(g["range"])).cuda().eval()
fe = PillarFrontEnd(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"], net, cartesian=True)
out = fe(frames)
print("ok", out["features"].shape)
