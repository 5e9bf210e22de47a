import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from partner_b200 import _lib
from partner_b200._lib import ptr, current_stream
lib = _lib.load()
torch.manual_seed(0)
for (m, n, k) in ((128, 32, 16), (128, 128, 32), (300, 128, 64), (1000, 64, 8)):
    a = torch.randn(m, k, device="cuda") * 3
    b = torch.randn(n, k, device="cuda")
    ref = (a.double() @ b.double().t())
    for variant in (0,):
        d = torch.zeros(m, n, device="cuda")
        rc = lib.pv_tc_gemm_tf32x3(ptr(a), ptr(b), m, n, k, ptr(d), variant, current_stream())
        torch.cuda.synchronize()
        err = (d.double() - ref).abs().max().item()
        print((m, n, k), "variant", variant, "rc", rc, "max abs err %.3e" % err, "rel %.3e" % (err / ref.abs().max().item()), flush=True)
