"""partner_b200 -- B200-native (sm_100a) polar front end of fudan-zvg/PARTNER.

Drop-in classes (same names / signatures as the reference):
    VoxelGenerator                         det3d/core/input/voxel_generator.py
    VoxelFeatureExtractorV3                det3d/models/readers/voxel_encoder.py
    PillarFeatureNet, PointPillarsScatter  det3d/models/readers/pillar_encoder.py
    DynamicVoxelEncoderV1                  det3d/models/readers/voxel_encoder.py
    DynamicPFNet, DynamicPPScatter         det3d/models/readers/pillar_encoder.py
    Voxelization                           det3d/datasets/pipelines/voxelization.py (hard + double flip, dynamic)
    transform_points                       det3d/datasets/pipelines/utils.py
and the fused batched paths ``PolarFrontEnd`` (mean VFE + canvas) and ``PillarFrontEnd`` (PFN + canvas).  All compute runs in hand-written CUDA kernels
reached through the C ABI of include/polar_voxel_b200.h; importing the package does not need a
GPU, calling anything does.
"""
from . import synth  # noqa: F401
from ._lib import build, load  # noqa: F401
from .registry import BACKBONES, READERS, build_from_cfg  # noqa: F401
from . import readers as _readers  # noqa: F401  (populates READERS / BACKBONES)


def __getattr__(name):
    # torch-dependent modules are imported lazily so `import partner_b200` stays cheap
    if name in ("VoxelGenerator",):
        from .voxel_generator import VoxelGenerator
        return VoxelGenerator
    if name == "Voxelization":
        from .voxelization import Voxelization
        return Voxelization
    if name in ("VoxelFeatureExtractorV3", "PillarFeatureNet", "PointPillarsScatter", "PFNLayer",
                "DynamicVoxelEncoderV1", "DynamicPFNet", "DynamicPPScatter"):
        from . import readers
        return getattr(readers, name)
    if name in ("PolarFrontEnd", "PillarFrontEnd", "shard_range"):
        from . import frontend
        return getattr(frontend, name)
    if name == "transform_points":
        from .functional import transform_points
        return transform_points
    raise AttributeError(name)
