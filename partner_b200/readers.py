"""Drop-in reader / scatter modules (inference) backed by the sm_100a kernels.

Same class names, constructor kwargs, forward signatures and state_dict keys as
det3d/models/readers/voxel_encoder.py:7-22 and det3d/models/readers/pillar_encoder.py:19-225, so
reference configs (``type="PillarFeatureNet"`` ...) and checkpoints
(``reader.pfn_layers.{i}.linear.weight``, ``...norm.{weight,bias,running_mean,running_var}``)
load unchanged.  Forward runs the CUDA kernels; there is no eager-PyTorch or CPU fallback.
PillarFeatureNet also trains: in ``.train()`` mode forward() uses batch statistics over all padded
rows and is differentiable with respect to its parameters (pv_pfn_train_forward / _backward).
"""
import torch
from torch import nn

from . import functional as F
from .registry import BACKBONES, READERS


@READERS.register_module
class VoxelFeatureExtractorV3(nn.Module):
    def __init__(self, num_input_features=4, norm_cfg=None, name="VoxelFeatureExtractorV3"):
        super(VoxelFeatureExtractorV3, self).__init__()
        self.name = name
        self.num_input_features = num_input_features

    def forward(self, features, num_voxels, coors=None):
        assert self.num_input_features == features.shape[-1]
        return F.vfe_mean(features, _as_i32(num_voxels))


def _as_i32(t):
    # the reference collates num_points as int32 (point_cloud_ops.py:185); accept int64 too
    return t if t.dtype == torch.int32 else t.to(torch.int32)


class PFNLayer(nn.Module):
    """Parameter container matching pillar_encoder.py:19-47 (Linear no-bias + BatchNorm1d)."""

    def __init__(self, in_channels, out_channels, norm_cfg=None, last_layer=False):
        super().__init__()
        self.name = "PFNLayer"
        self.last_vfe = last_layer
        if not self.last_vfe:
            out_channels = out_channels // 2
        self.units = out_channels
        if norm_cfg is None:
            norm_cfg = dict(type="BN1d", eps=1e-3, momentum=0.01)
        self.norm_cfg = norm_cfg
        cfg = dict(norm_cfg)
        kind = cfg.pop("type")
        if kind != "BN1d":
            raise ValueError("PFNLayer kernels implement BN1d only, got %s" % kind)
        cfg.pop("requires_grad", None)
        cfg.setdefault("eps", 1e-5)          # det3d/models/utils/norm.py:113
        self.linear = nn.Linear(in_channels, self.units, bias=False)
        self.norm = nn.BatchNorm1d(self.units, **cfg)


@READERS.register_module
class PillarFeatureNet(nn.Module):
    def __init__(self, num_input_features=4, num_filters=(64,), with_distance=False,
                 voxel_size=(0.2, 0.2, 4), pc_range=(0, -40, -3, 70.4, 40, 1), norm_cfg=None):
        super().__init__()
        self.name = "PillarFeatureNet"
        assert len(num_filters) > 0
        self.num_input = num_input_features
        num_input_features += 5
        if with_distance:
            num_input_features += 1
        self._with_distance = with_distance
        num_filters = [num_input_features] + list(num_filters)
        layers = []
        for i in range(len(num_filters) - 1):
            layers.append(PFNLayer(num_filters[i], num_filters[i + 1], norm_cfg=norm_cfg,
                                   last_layer=(i >= len(num_filters) - 2)))
        self.pfn_layers = nn.ModuleList(layers)
        # pillar_encoder.py:123-126 -- Python doubles, used as f32 scalars by the kernels
        self.vx = voxel_size[0]
        self.vy = voxel_size[1]
        self.x_offset = self.vx / 2 + pc_range[0]
        self.y_offset = self.vy / 2 + pc_range[1]

    def forward(self, features, num_voxels, coors):
        if self.training:
            return self._forward_train(features, num_voxels, coors)
        eps = {l.norm.eps for l in self.pfn_layers}
        if len(eps) != 1:
            raise ValueError("all PFN layers must share one BatchNorm eps")
        layers = [(l.linear.weight.detach().contiguous(), l.norm.running_mean, l.norm.running_var,
                   l.norm.weight.detach(), l.norm.bias.detach()) for l in self.pfn_layers]
        out = F.pfn_forward(features, _as_i32(num_voxels), coors if coors.dtype == torch.int32 else coors.int(),
                            layers, self.vx, self.vy, self.x_offset, self.y_offset,
                            self._with_distance, eps.pop())
        return out.squeeze()                      # pillar_encoder.py:169 (M == 1 collapses)


    def _forward_train(self, features, num_voxels, coors):
        """Training mode (pillar_encoder.py:49-61 with the norm in training mode): batch statistics over all
        M * T rows, running statistics updated in place, gradients for linear.weight / norm.weight / norm.bias
        through ``pv_pfn_train_backward`` (torch.autograd.Function, no eager fallback)."""
        eps = {l.norm.eps for l in self.pfn_layers}
        mom = {l.norm.momentum for l in self.pfn_layers}
        if len(eps) != 1 or len(mom) != 1:
            raise ValueError("all PFN layers must share one BatchNorm eps / momentum")
        params = []
        for l in self.pfn_layers:
            params += [l.linear.weight, l.norm.weight, l.norm.bias]
        out = _PfnTrain.apply(self, features.contiguous(), _as_i32(num_voxels),
                              coors if coors.dtype == torch.int32 else coors.int(), eps.pop(), mom.pop(), *params)
        with torch.no_grad():
            for l in self.pfn_layers:
                l.norm.num_batches_tracked += 1
        return out.squeeze()


class _PfnTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, features, num_voxels, coors, eps, momentum, *params):
        layers = [(params[3 * i].detach().contiguous(), l.norm.running_mean, l.norm.running_var,
                   params[3 * i + 1].detach().contiguous(), params[3 * i + 2].detach().contiguous())
                  for i, l in enumerate(net.pfn_layers)]
        out, saved = F.pfn_train_forward(features, num_voxels, coors, layers, net.vx, net.vy, net.x_offset, net.y_offset,
                                         net._with_distance, eps, momentum)
        ctx.saved, ctx.layers = saved, layers
        return out

    @staticmethod
    def backward(ctx, d_out):
        dw, dg, db = F.pfn_train_backward(d_out.contiguous(), ctx.saved, ctx.layers)
        grads = []
        for a, b, c in zip(dw, dg, db):
            grads += [a, b, c]
        ctx.saved = None                       # release the activations
        return (None, None, None, None, None, None, *grads)


@BACKBONES.register_module
class PointPillarsScatter(nn.Module):
    def __init__(self, num_input_features=64, norm_cfg=None, name="PointPillarsScatter", **kwargs):
        super().__init__()
        self.name = "PointPillarsScatter"
        self.nchannels = num_input_features

    def forward(self, voxel_features, coords, batch_size, input_shape):
        self.nx = int(input_shape[0])
        self.ny = int(input_shape[1])
        if voxel_features.dim() == 1:             # the reader's .squeeze() on a single voxel
            voxel_features = voxel_features.view(1, -1)
        return F.scatter(voxel_features.contiguous(), coords if coords.dtype == torch.int32 else coords.int(),
                         int(batch_size), self.ny, self.nx)


@READERS.register_module
class DynamicVoxelEncoderV1(nn.Module):
    """det3d/models/readers/voxel_encoder.py:26-44: torch.unique(grid_ind, dim=0) + scatter_mean.

    ``forward(data)`` takes the reference's dict (``data["points"]`` [N, C] f32, ``data["grid_ind"]``
    [N, 4] (b, z, y, x)) and returns ``(features [M, C], unq [M, 4])`` with voxels in sorted
    (b, z, y, x) order.  ``grid_size`` = (nx, ny, nz) and ``batch_size`` may be given to avoid the
    one host read that otherwise sizes the direct map from ``grid_ind.max()``."""

    def __init__(self, num_input_features=7, out_channels=16, name="DynamicVoxelEncoderV1", grid_size=None,
                 batch_size=None):
        super().__init__()
        self.name = name
        self.num_input_features = num_input_features
        self.point_density = False
        self.grid_size = grid_size
        self.batch_size = batch_size

    def forward(self, data):
        features = data["points"]
        grid_ind = data["grid_ind"]
        r = _dynamic_from_grid_ind(features, grid_ind, self.grid_size, self.batch_size, want_inverse=False)
        m = r.total()
        return r.mean_feats[:m], r.unq[:m].to(grid_ind.dtype)


def _dynamic_from_grid_ind(features, grid_ind, grid_size, batch_size, want_inverse):
    gi = grid_ind if grid_ind.dtype == torch.int32 else grid_ind.to(torch.int32)
    gi = gi.contiguous()
    if grid_size is None or batch_size is None:
        mx = gi.max(dim=0).values.tolist() if gi.shape[0] else [0, 0, 0, 0]
        if batch_size is None:
            batch_size = mx[0] + 1
        if grid_size is None:
            grid_size = (mx[3] + 1, mx[2] + 1, mx[1] + 1)
    nx, ny, nz = (int(v) for v in grid_size)
    cfg, _, _, _ = F.make_config([1.0, 1.0, 1.0], [0.0, 0.0, 0.0, float(nx), float(ny), float(nz)], 1, 1)
    r = F.dynamic_voxelize(cfg, features.contiguous(), None, int(batch_size), 0, False, grid_ind=gi,
                           want_inverse=want_inverse)
    F.read_status(r)      # a row outside grid_size / batch_size raises instead of silently dropping out of the means
    return r              # (torch.unique keeps every row); the readers synchronise on the voxel count anyway


@READERS.register_module
class DynamicPFNet(nn.Module):
    """det3d/models/readers/pillar_encoder.py:262-411 (same constructor, forward(data) and state_dict keys;
    the BatchNorm of each PFNLayer exists as in the reference but the dynamic forward never applies it)."""

    def __init__(self, num_input_features=4, num_filters=(64,), voxel_shape="cuboid", xyz_cluster=False,
                 raz_cluster=False, xy_center=False, ra_center=False, voxel_size=(0.2, 0.2, 4),
                 pc_range=(0, -40, -3, 70.4, 40, 1), norm_cfg=None, grid_size=None, batch_size=None):
        super().__init__()
        self.name = "DynamicPFNet"
        assert len(num_filters) > 0
        self.num_input = num_input_features
        self.voxel_shape = voxel_shape
        self.xyz_cluster, self.raz_cluster = xyz_cluster, raz_cluster
        self.xy_center, self.ra_center = xy_center, ra_center
        if xyz_cluster:
            num_input_features += 3
        if xy_center:
            num_input_features += 2
        if raz_cluster:
            num_input_features += 2 if xyz_cluster else 3
        if ra_center:
            num_input_features += 2
        num_filters = [num_input_features] + list(num_filters)
        self.pfn_layers = nn.ModuleList(
            [PFNLayer(num_filters[i], num_filters[i + 1], norm_cfg=norm_cfg, last_layer=(i >= len(num_filters) - 2))
             for i in range(len(num_filters) - 1)])
        self.vx = voxel_size[0]
        self.vy = voxel_size[1]
        self.x_offset = self.vx / 2 + pc_range[0]
        self.y_offset = self.vy / 2 + pc_range[1]
        self.grid_size, self.batch_size = grid_size, batch_size

    def forward(self, data):
        points = data["points"].contiguous()
        grid_ind = data["grid_ind"]
        r = _dynamic_from_grid_ind(points, grid_ind, self.grid_size, self.batch_size, want_inverse=True)
        m = r.total()
        weights = [l.linear.weight.detach().contiguous() for l in self.pfn_layers]
        feats = F.dynamic_pfn(points, r, m, weights, self.vx, self.vy, self.x_offset, self.y_offset,
                              self.voxel_shape != "cuboid", self.xyz_cluster, self.raz_cluster, self.xy_center,
                              self.ra_center)
        return feats, r.unq[:m].to(torch.int64)


@BACKBONES.register_module
class DynamicPPScatter(nn.Module):
    """det3d/models/readers/pillar_encoder.py:413-432."""

    def __init__(self, **kwargs):
        super().__init__()
        self.name = "DynamicPPScatter"

    def forward(self, voxel_features, unq, batch_size, grid_size):
        nx, ny = int(grid_size[0]), int(grid_size[1])
        return F.scatter(voxel_features.contiguous(), unq if unq.dtype == torch.int32 else unq.to(torch.int32),
                         int(batch_size), ny, nx)
