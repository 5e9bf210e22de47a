import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import oracle
from partner_b200 import synth, PolarFrontEnd
g = synth.GRIDS["NUSC-PILLAR"]
fe = PolarFrontEnd(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
ref = oracle.VoxelGenerator(g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
frames = synth.make_batch("nusc", 2, 3)
polars = [oracle.transform_points(f) for f in frames]
outs = [ref.generate(p, return_density=True) for p in polars]
vox, coor, num, nv = oracle.collate([(o[0], o[1], o[2]) for o in outs])
got = fe(frames)
bad = np.nonzero(got["num_points"] != num)[0]
print("mismatches", len(bad), "of", len(num))
den = np.stack([o[4] for o in outs])
for i in bad[:30]:
    b, z, y, x = coor[i]
    print(i, "coor", coor[i], "ref num", num[i], "got", got["num_points"][i], "true count", den[b, z, y, x])
d = got["num_points"][bad].astype(int) - num[bad].astype(int)
print("diff hist", np.unique(d, return_counts=True))
print("frames of bad", np.unique(coor[bad, 0], return_counts=True))
import torch
from partner_b200 import functional as F
sizes = [f.shape[0] for f in frames]
off = np.zeros(len(frames) + 1, np.int32); np.cumsum(sizes, out=off[1:])
pts = torch.from_numpy(np.concatenate(frames)).cuda()
vb = F.voxelize(fe.cfg, pts, torch.from_numpy(off).cuda(), len(frames), max(sizes), True, want_mean=True, want_density=True)
m = vb.total()
gnum = vb.num_points[:m].cpu().numpy(); gden = vb.density.cpu().numpy(); gmean = vb.mean_feats[:m].cpu().numpy()
rmean = oracle.vfe_mean(vox, num)
bad = np.nonzero(gnum != num)[0]
print("voxelize path mismatches", len(bad))
np.set_printoptions(precision=4, suppress=True, linewidth=200)
for i in bad[:12]:
    b, z, y, x = coor[i]
    print(i, "true", den[b, z, y, x], "got dens", gden[b, z, y, x], "got num", gnum[i])
    print("   ref mean", rmean[i]); print("   got mean", gmean[i], " got mean*num", gmean[i] * gnum[i])
