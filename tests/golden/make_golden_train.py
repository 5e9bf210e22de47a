"""Golden vectors for the TRAINING-mode PillarFeatureNet, generated FROM THE REFERENCE MODULE ITSELF.

Run in the build container only (needs /root/reference and torch CPU):

    python tests/golden/make_golden_train.py

The reference's own ``PillarFeatureNet`` (det3d/models/readers/pillar_encoder.py:75-169, PFNLayer :19-61) is put
in ``.train()`` mode and run forward + backward with torch.autograd on the voxels of readers.npz: BatchNorm1d
with batch statistics over all M * T rows (padded slots included), running statistics updated with momentum
0.01, loss = sum(out * G) for a fixed random G.  Frozen: the output, the updated running statistics and the
gradients of every parameter -> tests/golden/pfn_train.npz (two configurations).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference  # noqa: E402


def main():
    ref = load_reference()
    torch = ref["torch"]
    g = np.load(os.path.join(HERE, "readers.npz"))
    m = 1500
    vox, num, coor = g["voxels"][:m], g["num_points"][:m], g["coors"][:m]
    out = dict(voxels=vox, num_points=num, coors=coor, voxel_size=g["voxel_size"], pc_range=g["pc_range"])
    tv, tn, tc = torch.from_numpy(vox), torch.from_numpy(num), torch.from_numpy(coor)
    for tag, filters, wd in (("t64_128", (64, 128), False), ("t32_32_64_dist", (32, 32, 64), True)):
        torch.manual_seed(5)
        net = ref["pe"].PillarFeatureNet(7, filters, wd, list(g["voxel_size"]), list(g["pc_range"]))
        with torch.no_grad():
            for l in net.pfn_layers:
                u = l.norm.num_features
                l.norm.running_mean.copy_(torch.randn(u))
                l.norm.running_var.copy_(torch.rand(u) * 1.5 + 0.5)
                l.norm.weight.copy_(torch.randn(u))
                l.norm.bias.copy_(torch.randn(u))
        for i, l in enumerate(net.pfn_layers):
            out[f"{tag}_w{i}"] = l.linear.weight.detach().numpy().copy()
            out[f"{tag}_mean{i}"] = l.norm.running_mean.numpy().copy()
            out[f"{tag}_var{i}"] = l.norm.running_var.numpy().copy()
            out[f"{tag}_gamma{i}"] = l.norm.weight.detach().numpy().copy()
            out[f"{tag}_beta{i}"] = l.norm.bias.detach().numpy().copy()
        net.train()
        y = net(tv, tn, tc)
        G = torch.from_numpy(np.random.default_rng(9).normal(0, 1, tuple(y.shape)).astype(np.float32))
        (y * G).sum().backward()
        out[f"{tag}_out"] = y.detach().numpy()
        out[f"{tag}_G"] = G.numpy()
        for i, l in enumerate(net.pfn_layers):
            out[f"{tag}_mean{i}_after"] = l.norm.running_mean.numpy().copy()
            out[f"{tag}_var{i}_after"] = l.norm.running_var.numpy().copy()
            out[f"{tag}_dw{i}"] = l.linear.weight.grad.numpy().copy()
            out[f"{tag}_dgamma{i}"] = l.norm.weight.grad.numpy().copy()
            out[f"{tag}_dbeta{i}"] = l.norm.bias.grad.numpy().copy()
        out[f"{tag}_eps"] = np.float64(net.pfn_layers[0].norm.eps)
        out[f"{tag}_momentum"] = np.float64(net.pfn_layers[0].norm.momentum)
        print(tag, "out", tuple(y.shape), "eps", net.pfn_layers[0].norm.eps, "momentum", net.pfn_layers[0].norm.momentum)
    np.savez_compressed(os.path.join(HERE, "pfn_train.npz"), **out)


if __name__ == "__main__":
    main()
