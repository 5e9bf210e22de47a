"""Shared helpers for the parity tests."""
import numpy as np


def ulp_diff(a, b):
    """Distance in float32 ulps between two float32 arrays (sign-aware)."""
    ai = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    bi = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


def assert_close_fp32(a, b, what="", scale="auto"):
    """The float gate of SURVEY.md section 8d: rtol 1e-5, atol 1e-5 * max|b|.

    The atol floor is relative to the magnitude of what is being compared, so for tensors whose
    columns live on different scales (mean features: intensity <= 255 next to a time lag <= 0.5) the
    maximum is taken PER COLUMN ([M, C] rows, C <= 16) or PER CHANNEL ([B, C, H, W] canvases,
    C <= 16); a tensor-wide maximum would let the large column hide errors in the small ones.
    Wide tensors (PFN outputs, C >= 32 homogeneous post-ReLU units) keep the tensor-wide floor the
    survey wrote the gate for.  ``scale="global"`` forces the tensor-wide floor."""
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if b.size == 0:
        return
    mag = np.where(np.isfinite(b), np.abs(b), 0.0)
    if scale == "auto" and b.ndim == 2 and b.shape[1] <= 16:
        atol = 1e-5 * mag.max(axis=0, keepdims=True)
    elif scale == "auto" and b.ndim == 4 and b.shape[1] <= 16:
        atol = 1e-5 * mag.max(axis=(0, 2, 3), keepdims=True)
    else:
        atol = np.float64(1e-5 * float(mag.max()))
    bad = ~(np.abs(a - b) <= atol + 1e-5 * np.abs(b))
    bad &= ~(a == b)                                          # equal infinities pass, NaN never does
    assert not bad.any(), "%s: %d / %d outside rtol=1e-5 atol=%s (max abs err %g)" % (
        what, int(bad.sum()), b.size, np.array2string(np.asarray(atol).ravel(), precision=3),
        float(np.nanmax(np.abs(a - b)[bad])))


def densify(idx, val, shape):
    out = np.zeros(int(np.prod(shape)), dtype=val.dtype)
    out[idx] = val
    return out.reshape([int(s) for s in shape])
