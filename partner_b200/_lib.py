"""Build + ctypes binding of the C-ABI library declared in include/polar_voxel_b200.h.

The library is the product path: there is no CPU or PyTorch fallback.  If the shared object
is missing, ``load()`` raises; callers never catch that to substitute another implementation.
"""
import ctypes
import glob
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
SO_PATH = os.path.join(_PKG, "libpolar_voxel_b200.so")
_SOURCES = sorted(glob.glob(os.path.join(_PKG, "csrc", "*.cu")))
_HEADERS = sorted(glob.glob(os.path.join(_PKG, "csrc", "*.cuh"))) + [
    os.path.join(_ROOT, "include", "polar_voxel_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only
    "-lineinfo",
    "-fmad=false",            # no implicit contraction: bins must match numba's separate f32 ops
    "-Xcompiler", "-fPIC", "-shared",
]

PV_OK = 0
PV_MAX_CHANNELS = 16
PV_MAX_PFN_LAYERS = 4
_VALUE_ERRORS = (-1, -2, -6)   # bad config / bad argument / unsupported -> ValueError


class PvConfig(ctypes.Structure):
    _fields_ = [("lo", ctypes.c_float * 3), ("vs", ctypes.c_float * 3),
                ("grid", ctypes.c_int32 * 3), ("max_points", ctypes.c_int32),
                ("max_voxels", ctypes.c_int32), ("pipeline", ctypes.c_int32)]


class PvPfnLayer(ctypes.Structure):
    _fields_ = [("weight", ctypes.c_void_p), ("bn_mean", ctypes.c_void_p),
                ("bn_var", ctypes.c_void_p), ("bn_gamma", ctypes.c_void_p),
                ("bn_beta", ctypes.c_void_p), ("in_channels", ctypes.c_int32),
                ("units", ctypes.c_int32)]


def needs_build():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.getmtime(f) > t for f in _SOURCES + _HEADERS)


def build(force=False, verbose=False):
    """nvcc-compile every kernel for sm_100a into the in-tree shared object (one object per
    translation unit, compiled in parallel, objects kept under build/ and reused when unchanged)."""
    if not force and not needs_build():
        return SO_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = os.path.join(_ROOT, "build", "obj")
    os.makedirs(objdir, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in _HEADERS)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj
        cmd = [nvcc] + flags + ["-I", os.path.join(_ROOT, "include"), "-c", "-o", obj, src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(_SOURCES))) as ex:
        objs = list(ex.map(compile_one, _SOURCES))
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO_PATH] + objs)
    return SO_PATH


_lib = None
P = ctypes.c_void_p
I32 = ctypes.c_int32
I64 = ctypes.c_int64
F32 = ctypes.c_float
SZ = ctypes.c_size_t

EXPORTS = {
    "pv_version": (ctypes.c_int, []),
    "pv_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "pv_profile_pipeline": (ctypes.c_int, [ctypes.POINTER(PvConfig)]),
    "pv_workspace_bytes": (SZ, [ctypes.POINTER(PvConfig), I64, I32, I64, I32]),
    "pv_transform_points": (ctypes.c_int, [P, I64, I32, I32, P, P]),
    "pv_workspace_init": (ctypes.c_int, [ctypes.POINTER(PvConfig), I64, I32, I64, I32, P, SZ, P]),
    "pv_voxelize": (ctypes.c_int, [ctypes.POINTER(PvConfig), P, P, I32, I64, I32, I32, I64, I64, P, SZ,
                                   P, P, P, P, P, P, P, P]),
    "pv_forward_mean_canvas": (ctypes.c_int, [ctypes.POINTER(PvConfig), P, P, I32, I64, I32, I32, I64, I64,
                                              P, SZ, P, P, P, P, P, P]),
    "pv_pfn_canvas_workspace_bytes": (SZ, [I32, I32, I32, I64, I32]),
    "pv_forward_pfn_canvas": (ctypes.c_int, [ctypes.POINTER(PvConfig), P, P, I32, I64, I32, I32, I64, I64, P, SZ, P, SZ,
                                             ctypes.POINTER(PvPfnLayer), I32, I32, F32, F32, F32, F32, F32,
                                             P, P, P, P, P, P]),
    "pv_profile_mean_canvas": (ctypes.c_int, [ctypes.POINTER(PvConfig), P, P, I32, I64, I32, I32, I64, I64,
                                              P, SZ, P, P, P, P, P, P, I32, ctypes.POINTER(F32)]),
    "pv_dynamic_voxelize": (ctypes.c_int, [ctypes.POINTER(PvConfig), P, P, P, I32, I64, I32, I32, I64, I64, P, SZ,
                                           P, P, P, P, P, P, P, P]),
    "pv_dynamic_grid_ind": (ctypes.c_int, [ctypes.POINTER(PvConfig), P, P, I32, I64, I32, I32, P, P]),
    "pv_dynamic_pfn_workspace_bytes": (SZ, [I64, I64]),
    "pv_dynamic_pfn": (ctypes.c_int, [P, P, P, P, P, I64, I64, I32, I32, I32, F32, F32, F32, F32,
                                      ctypes.POINTER(PvPfnLayer), I32, P, SZ, P, P]),
    "pv_stream_workspace_bytes": (SZ, [I64, I32]),
    "pv_stream_sectors": (ctypes.c_int, [ctypes.POINTER(PvConfig), P, I64, I32, I32, F32, P, SZ, P, P, P, P, P]),
    "pv_affine_points": (ctypes.c_int, [P, I64, I32, ctypes.POINTER(ctypes.c_double), F32, P, P]),
    "pv_seg_workspace_bytes": (SZ, [ctypes.POINTER(PvConfig), I64, I32]),
    "pv_seg_voxel_labels": (ctypes.c_int, [ctypes.POINTER(PvConfig), P, P, P, I32, I64, P, SZ, P, P, P, P, P]),
    "pv_seg_gather_points": (ctypes.c_int, [P, I32, I32, I32, P, P, I32, I64, P, P, P]),
    "pv_read_status": (ctypes.c_int, [P, P]),
    "pv_vfe_mean": (ctypes.c_int, [P, P, I64, I32, I32, P, P]),
    "pv_pfn_forward": (ctypes.c_int, [P, P, P, I64, I32, I32, I32, F32, F32, F32, F32,
                                      ctypes.POINTER(PvPfnLayer), I32, F32, P, SZ, P, P]),
    "pv_pfn_workspace_bytes": (SZ, [I64, I32]),
    "pv_pfn_train_workspace_bytes": (SZ, [I64, I32, I32, I32, ctypes.POINTER(PvPfnLayer), I32]),
    "pv_pfn_train_forward": (ctypes.c_int, [P, P, P, I64, I32, I32, I32, F32, F32, F32, F32, ctypes.POINTER(PvPfnLayer), I32,
                                            F32, F32, P, SZ, P, P]),
    "pv_pfn_train_backward": (ctypes.c_int, [P, I64, I32, I32, I32, ctypes.POINTER(PvPfnLayer), I32, P, SZ,
                                             ctypes.POINTER(P), ctypes.POINTER(P), ctypes.POINTER(P), P]),
    "pv_tc_gemm_tf32x3": (ctypes.c_int, [P, P, I32, I32, I32, P, I32, P]),
    "pv_scatter_workspace_bytes": (SZ, [I32, I32, I32]),
    "pv_scatter": (ctypes.c_int, [P, P, I64, I32, I32, I32, I32, P, SZ, P, P, P]),
}


def load():
    """dlopen the library and bind every symbol of the header; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            "partner_b200: %s is not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % SO_PATH)
    lib = ctypes.CDLL(SO_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc == PV_OK:
        return
    msg = load().pv_error_string(rc).decode()
    text = "polar_voxel_b200: %s%s (code %d)" % (what + ": " if what else "", msg, rc)
    if rc in _VALUE_ERRORS:
        raise ValueError(text)
    raise RuntimeError(text)


def ptr(t):
    """Device pointer of a torch tensor (or NULL for None)."""
    return ctypes.c_void_p(0) if t is None else ctypes.c_void_p(t.data_ptr())


def current_stream(device=None):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
