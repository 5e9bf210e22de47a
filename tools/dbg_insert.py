import sys, ctypes, numpy as np, torch
sys.path.insert(0, '/root/repo')
import oracle
from partner_b200 import synth, _lib
from partner_b200 import functional as F
lib = _lib.load()
g = synth.GRIDS["NUSC-PILLAR"]
cfg, vs, rng, gs = F.make_config(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
P = ctypes.c_void_p
lib.pv_debug_insert.argtypes = [ctypes.POINTER(_lib.PvConfig), P, P, ctypes.c_int32, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                ctypes.c_int64, ctypes.c_int64, P, ctypes.c_int64, P, P, ctypes.c_int32]
def run(frames, generic, npts=None):
    if npts: frames = [f[:npts] for f in frames]
    sizes = [f.shape[0] for f in frames]
    off = np.zeros(len(frames) + 1, np.int32); np.cumsum(sizes, out=off[1:])
    allp = np.concatenate(frames)
    pts = torch.from_numpy(allp).cuda(); d_off = torch.from_numpy(off).cuda()
    n = int(off[-1]); B = len(frames); n_cap = F._bucket(n); f_cap = min(n_cap, F._bucket(max(sizes)))
    F.drop_voxel_workspaces()
    ws = F.voxel_workspace(cfg, n_cap, B, f_cap, 7, pts.device)
    torch.cuda.synchronize()
    slots = B * 262144
    first = np.empty(slots, np.uint32); acc = np.empty((slots, 8), np.float32)
    rc = lib.pv_debug_insert(cfg, pts.data_ptr(), d_off.data_ptr(), B, n, 5, 1, n_cap, f_cap, ws.data_ptr(), slots,
                             first.ctypes.data, acc.ctypes.data, 1 if generic else 0)
    assert rc == 0, rc
    # numpy reference
    cnt = np.zeros(slots); fmin = np.full(slots, 0xFFFFFFFF, np.uint64); isum = np.zeros(slots)
    for b, f in enumerate(frames):
        pol = oracle.transform_points(f)
        c = np.floor((pol[:, :3] - rng[:3]) / vs).astype(np.int64)
        ok = ((c >= 0) & (c < gs)).all(1)
        s = b * 262144 + (c[:, 1] * 512 + c[:, 0])[ok]
        idx = (np.arange(f.shape[0]) + off[b])[ok]
        np.add.at(cnt, s, 1); np.minimum.at(fmin, s, idx.astype(np.uint64)); np.add.at(isum, s, f[ok, 3].astype(np.float64))
    badc = np.nonzero(acc[:, 7] != cnt)[0]
    badf = np.nonzero(first.astype(np.uint64) != fmin)[0]
    print("generic" if generic else "special", "frames", len(frames), "n", n, ": count mismatches", len(badc), "first mismatches", len(badf),
          "| intensity-sum max rel err", np.abs(acc[:, 5] - isum).max() / max(1, isum.max()))
    for s in badc[:6]:
        print("   slot", s, "b", s // 262144, "y", (s % 262144) // 512, "x", s % 512, "true", cnt[s], "got row", acc[s])
    del ws; F.drop_voxel_workspaces()
frames = synth.make_batch("nusc", 2, 3)
run(frames[:1], False, 1024)
run(frames[:1], False, 4096)
run(frames[:1], False)
run(frames[:1], True)
run(frames, False)
run(frames, True)
