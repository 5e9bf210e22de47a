#!/usr/bin/env python
"""Development aid: time line of the roles of k_pfn_fused (block 0) from a -DP2_TRACE build.
   SRC=pfn_fused tools/build_variant.sh trace -DP2_TRACE ; PV_LIB=build/lib_trace.so python tools/pfn_trace.py"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partner_b200 import _lib  # noqa: E402

_lib.SO_PATH = os.path.abspath(os.environ["PV_LIB"])
from partner_b200 import PillarFrontEnd, synth  # noqa: E402
from partner_b200.readers import PillarFeatureNet  # noqa: E402

dev = torch.device("cuda", 0)
g = synth.GRIDS["NUSC-PILLAR"]
B = 16
frames = synth.make_batch("nusc", 3, B)
sizes = [f.shape[0] for f in frames]
off = np.zeros(B + 1, np.int32)
np.cumsum(sizes, out=off[1:])
pts = torch.from_numpy(np.concatenate(frames)).to(dev)
d_off = torch.from_numpy(off).to(dev)
torch.manual_seed(0)
net = PillarFeatureNet(7, (64, 128), False, g["voxel_size"], g["range"]).to(dev).eval()
fe = PillarFrontEnd(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"], net, cartesian=True, device=dev)
for _ in range(3):
    out = fe.forward_device(pts, d_off, B, max(sizes))
torch.cuda.synchronize()
lib = _lib.load()
lib.pv_debug_p2_trace.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
buf = np.zeros((32, 64, 2, 3), np.int64)
print("rc", lib.pv_debug_p2_trace(buf.ctypes.data, buf.nbytes))
nprod = int(os.environ.get("NPROD", "16"))
NEPI = 12
t0 = buf[buf > 0].min()
mhz = 1.9
def us(x): return (x - t0) / mhz / 1e3
R = range(8, 16)
print("producer warp 0 (set 0): round: loads issued / stage free / full arrived  [us]")
for w in (0, nprod // 2):
    for r in R:
        print("  w%d r%d  %.2f %.2f %.2f   compute %.2f" % (w, r, us(buf[w, r, 0, 0]), us(buf[w, r, 0, 1]), us(buf[w, r, 0, 2]),
                                                     (buf[w, r, 0, 2] - buf[w, r, 0, 1]) / mhz / 1e3))
print("issuer: round s: full seen / free seen / committed")
for r in R:
    for s in (0, 1):
        print("  r%d s%d  %.2f %.2f %.2f" % (r, s, us(buf[nprod + NEPI, r, s, 0]), us(buf[nprod + NEPI, r, s, 1]), us(buf[nprod + NEPI, r, s, 2])))
# per-producer compute time statistics
comp = (buf[:nprod, 8:48, 0, 2] - buf[:nprod, 8:48, 0, 1]) / mhz / 1e3
wait = (buf[:nprod, 8:48, 0, 1] - buf[:nprod, 8:48, 0, 0]) / mhz / 1e3
print("producer compute us: mean %.2f  (per warp %s)" % (comp.mean(), np.round(comp.mean(1), 2)))
print("producer wait-for-stage us: mean %.2f" % wait.mean())
pre = (buf[:nprod, 9:48, 0, 0] - buf[:nprod, 8:47, 0, 2]) / mhz / 1e3
print("producer fetch + layer 0 (before the wait) us: mean %.2f" % pre.mean())
ew = []
for w in range(NEPI):                       # team = w >> 2 serves tiles = team (mod 3); tile = 2 * round + s
    for r in range(8, 48):
        for s_ in (0, 1):
            if (2 * r + s_) % 3 == (w >> 2) and buf[nprod + w, r, s_, 1] > 0:
                ew.append((buf[nprod + w, r, s_, 1] - buf[nprod + w, r, s_, 0]) / mhz / 1e3)
print("epilogue work per tile and warp us: mean %.2f" % np.mean(ew))
iss = buf[nprod + NEPI, 8:48, :, :]
print("issuer: wait full->free %.2f  issue %.2f   period per tile %.2f" % (((iss[:, :, 1] - iss[:, :, 0]) / mhz / 1e3).mean(), ((iss[:, :, 2] - iss[:, :, 1]) / mhz / 1e3).mean(),
      (iss[-1, 1, 2] - iss[0, 0, 2]) / mhz / 1e3 / (2 * 40 - 1)))

