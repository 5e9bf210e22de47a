"""Staging + loading of the handful of UNMODIFIED reference files the measurements use as comparators.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (like everything under oracle/): nothing under partner_b200/
imports this.  The files are copied verbatim from the reference checkout into the git-ignored
``baseline/_ref/`` (it travels to the GPU box with the snapshot, it never enters the history) by
``stage()``, which ``__graft_entry__.build()`` calls when /root/reference is present.  File list =
SURVEY.md section 8c.  ``bench.py`` then times
  * the reference's own numba ``points_to_voxel`` + ``transform_points`` on the box's host cores
    (``--impl reference`` / ``cpu_baseline``, kind "reference"), and
  * its eager-PyTorch readers (VoxelFeatureExtractorV3 / PillarFeatureNet / PointPillarsScatter) on
    the same B200 (``ref_eager_gpu``),
and falls back to the C port of oracle/ when the staged files (or numba) are missing.
"""
import importlib.util
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference"
REF_DST = os.path.join(ROOT, "baseline", "_ref")
FILES = [
    "det3d/__init__.py",
    "det3d/ops/point_cloud/__init__.py",
    "det3d/ops/point_cloud/point_cloud_ops.py",
    "det3d/core/input/voxel_generator.py",
    "det3d/datasets/pipelines/utils.py",
    "det3d/models/registry.py",
    "det3d/models/utils/misc.py",
    "det3d/models/utils/norm.py",
    "det3d/utils/dist/dist_common.py",
    "det3d/models/readers/pillar_encoder.py",
    "det3d/models/readers/voxel_encoder.py",
]


def stage():
    """Copy the reference files into baseline/_ref (no-op without /root/reference)."""
    if not os.path.isdir(REF_SRC):
        return False
    for rel in FILES:
        dst = os.path.join(REF_DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_SRC, rel), dst)
    return True


def available():
    return all(os.path.exists(os.path.join(REF_DST, rel)) for rel in FILES)


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_voxelizer = None


def load_voxelizer():
    """(transform_points, VoxelGenerator class) of the reference: numba points_to_voxel behind it."""
    global _voxelizer
    if _voxelizer is not None:
        return _voxelizer
    import numpy as np
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    from det3d.ops.point_cloud.point_cloud_ops import points_to_voxel  # noqa: F401  (namespace package)
    vg = _load("ref_voxel_generator", os.path.join(REF_DST, "det3d/core/input/voxel_generator.py"))
    src = open(os.path.join(REF_DST, "det3d/datasets/pipelines/utils.py")).read().splitlines()
    start = next(i for i, l in enumerate(src) if l.startswith("def transform_points"))
    end = next(i for i in range(start + 1, len(src)) if src[i].startswith("def "))
    ns = {"np": np}
    exec("\n".join(src[start:end]), ns)          # the function alone: its module's other imports are not installed
    _voxelizer = (ns["transform_points"], vg.VoxelGenerator)
    return _voxelizer


_readers = None


def load_readers():
    """(VoxelFeatureExtractorV3, PillarFeatureNet, PointPillarsScatter) of the reference, unmodified
    (stub parent packages + a torch_scatter stub: only the Dynamic* classes would call it)."""
    global _readers
    if _readers is not None:
        return _readers
    ts = types.ModuleType("torch_scatter")
    ts.scatter_mean = ts.scatter_max = None
    sys.modules.setdefault("torch_scatter", ts)
    for pkg in ("det3d", "det3d.models", "det3d.models.readers", "det3d.models.utils", "det3d.utils", "det3d.utils.dist"):
        if pkg not in sys.modules or not hasattr(sys.modules[pkg], "__path__"):
            m = sys.modules.get(pkg) or types.ModuleType(pkg)
            m.__path__ = getattr(m, "__path__", [])
            sys.modules[pkg] = m

    class Registry:
        def __init__(self, name):
            self.name = name
            self.module_dict = {}

        def register_module(self, cls):
            self.module_dict[cls.__name__] = cls
            return cls
    sys.modules["det3d.utils"].Registry = Registry
    p = lambda rel: os.path.join(REF_DST, rel)      # noqa: E731
    _load("det3d.models.registry", p("det3d/models/registry.py"))
    misc = _load("det3d.models.utils.misc", p("det3d/models/utils/misc.py"))
    _load("det3d.utils.dist.dist_common", p("det3d/utils/dist/dist_common.py"))
    norm = _load("det3d.models.utils.norm", p("det3d/models/utils/norm.py"))
    sys.modules["det3d.models.utils"].get_paddings_indicator = misc.get_paddings_indicator
    sys.modules["det3d.models.utils"].build_norm_layer = norm.build_norm_layer
    pe = _load("det3d.models.readers.pillar_encoder", p("det3d/models/readers/pillar_encoder.py"))
    ve = _load("det3d.models.readers.voxel_encoder", p("det3d/models/readers/voxel_encoder.py"))
    _readers = (ve.VoxelFeatureExtractorV3, pe.PillarFeatureNet, pe.PointPillarsScatter)
    return _readers
