"""Development aid: one small pv_pfn_forward call on the tensor-core path, watchdog word printed."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from partner_b200 import functional as F, _lib
if os.environ.get('PV_LIB'):
    _lib.SO_PATH = os.path.abspath(os.environ['PV_LIB'])
m = int(sys.argv[1]) if len(sys.argv) > 1 else 300
t, c = 20, 7
rng = np.random.default_rng(0)
num = rng.integers(1, t + 1, m).astype(np.int32)
vox = rng.normal(0, 3, (m, t, c)).astype(np.float32) * (np.arange(t)[None, :, None] < num[:, None, None])
coors = np.stack([np.zeros(m), np.zeros(m), rng.integers(0, 512, m), rng.integers(0, 512, m)], 1).astype(np.int32)
vs, rg = [0.098, 0.0123, 8.0], [0.3, -3.1488, -5.0, 50.476, 3.1488, 3.0]
layers, dev = [], []
width = 12
for u in (32, 128):
    L = dict(weight=rng.normal(0, 0.3, (u, width)).astype(np.float32), mean=rng.normal(0, 1, u).astype(np.float32),
             var=rng.uniform(0.5, 2, u).astype(np.float32), gamma=rng.normal(0, 1, u).astype(np.float32),
             beta=rng.normal(0, 1, u).astype(np.float32))
    layers.append(L); dev.append(tuple(torch.from_numpy(L[k]).cuda() for k in ("weight", "mean", "var", "gamma", "beta")))
    width = 2 * u
out = F.pfn_forward(torch.from_numpy(vox).cuda(), torch.from_numpy(num).cuda(), torch.from_numpy(coors).cuda(), dev,
                    vs[0], vs[1], vs[0] / 2 + rg[0], vs[1] / 2 + rg[1], False, 1e-3)
torch.cuda.synchronize()
ws = [w for k, w in F._workspaces.items() if k[-1] == "pfn"][0]
words = ws[:256].view(torch.int32).cpu().numpy()
print("counter %d  status %d  watchdog 0x%x block %d" % (int(words[0]), int(words[1]), int(words[2]) & 0xffffffff, int(words[3])), " ".join("w%d:%x" % (k, int(words[4 + k]) & 0xffffffff) for k in range(29)))
ref = oracle.pfn_forward(vox, num, coors, layers, vs, rg, with_distance=False, eps=1e-3)
o = out.cpu().numpy()
err = np.abs(o - ref)
print("max abs err %g (max |ref| %g), rows with err > 1e-4: %d / %d" % (err.max(), np.abs(ref).max(), (err.max(1) > 1e-5 * np.abs(ref).max()).sum(), m))
bad = np.where(err.max(1) > 1e-5 * np.abs(ref).max())[0][:10]
print("first bad rows", bad, "num", num[bad])
out2 = F.pfn_forward(torch.from_numpy(vox).cuda(), torch.from_numpy(num).cuda(), torch.from_numpy(coors).cuda(), dev,
                     vs[0], vs[1], vs[0] / 2 + rg[0], vs[1] / 2 + rg[1], False, 1e-3).cpu().numpy()
print("second run: rows differing from first run %d, bad rows %d" % ((np.abs(out2 - o).max(1) > 0).sum(), (np.abs(out2 - ref).max(1) > 1e-4).sum()))
badrows = np.where(err.max(1) > 1e-5 * np.abs(ref).max())[0]
print("bad rows per 64-chunk (first 20 chunks):", np.bincount(badrows // 64, minlength=20)[:20])
r = badrows[0]; print("row", r, "out", o[r, :6], "ref", ref[r, :6])
