"""tcgen05 building block of the PFN linear layers: fp32 GEMM through TF32 tensor cores with the
3xTF32 split must match an fp64 reference at fp32 accuracy (tolerance: 1e-5 * max|ref|, the
float gate of SURVEY.md section 8d; measured ~1e-6)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n,k", [(128, 32, 16), (128, 128, 32), (300, 128, 64), (1000, 64, 8), (5000, 128, 64)])
def test_tc_gemm_tf32x3_matches_fp64(m, n, k):
    import torch
    from partner_b200 import _lib
    from partner_b200._lib import check, current_stream, ptr
    torch.manual_seed(m + n + k)
    a = torch.randn(m, k, device="cuda") * 3
    b = torch.randn(n, k, device="cuda")
    d = torch.empty(m, n, device="cuda")
    check(_lib.load().pv_tc_gemm_tf32x3(ptr(a), ptr(b), m, n, k, ptr(d), 0, current_stream()), "pv_tc_gemm_tf32x3")
    ref = a.double() @ b.double().t()
    err = (d.double() - ref).abs().max().item()
    assert err <= 1e-5 * ref.abs().max().item(), err
    # plain TF32 (no split) would be ~1e-3: make sure the split is really in effect
    assert err <= 5e-6 * ref.abs().max().item()


def test_tc_gemm_rejects_unsupported_shapes():
    import torch
    from partner_b200 import _lib
    from partner_b200._lib import current_stream, ptr
    a = torch.zeros(128, 12, device="cuda")
    b = torch.zeros(32, 12, device="cuda")
    d = torch.zeros(128, 32, device="cuda")
    assert _lib.load().pv_tc_gemm_tf32x3(ptr(a), ptr(b), 128, 32, 12, ptr(d), 0, current_stream()) == -6
