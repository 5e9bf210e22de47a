"""Shared helpers for the parity tests."""
import numpy as np


def ulp_diff(a, b):
    """Distance in float32 ulps between two float32 arrays (sign-aware)."""
    ai = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    bi = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


def assert_close_fp32(a, b, what=""):
    """The float gate of SURVEY.md section 8d: rtol 1e-5, atol 1e-5 * max|b|."""
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if b.size == 0:
        return
    atol = 1e-5 * float(np.abs(b).max())
    bad = ~np.isclose(a, b, rtol=1e-5, atol=atol)
    assert not bad.any(), "%s: %d / %d outside rtol=1e-5 atol=%g (max abs err %g)" % (
        what, int(bad.sum()), b.size, atol, float(np.abs(a - b).max()))


def densify(idx, val, shape):
    out = np.zeros(int(np.prod(shape)), dtype=val.dtype)
    out[idx] = val
    return out.reshape([int(s) for s in shape])
