"""The sa1-sized GEMM launches of the S3DIS step, one launch each (for `ncu --set full`): python scripts/ncu_gemm_cases.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from repsurf_b200 import tc  # noqa: E402
from scripts.ab_gemm import make_opnd  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
R = 2621440


def rows(K, Nn, kind, epi):
    A, _ = make_opnd(kind, R, K, g)
    W = torch.randn(Nn, K, generator=g, device=dev) / K ** 0.5
    Wp, _, _ = tc.prep_weight(W)
    Y = torch.empty(R, Nn, device=dev)
    if epi == "stats":
        st = torch.zeros(2 * Nn, dtype=torch.float64, device=dev)
        return lambda: tc.gemm_rows(R, Nn, A, Wp, Y=Y, bias=torch.zeros(Nn, device=dev), stats=st)
    dual = epi == "maskdual"
    w = Nn * (2 if dual else 1)
    Yl = torch.randn(R, w, generator=g, device=dev)
    sc, sh, mu, inv = [torch.rand(w, generator=g, device=dev) + 0.5 for _ in range(4)]
    st = torch.zeros((3 if dual else 2) * Nn, dtype=torch.float64, device=dev)
    return lambda: tc.gemm_rows(R, Nn, A, Wp, Y=Y, stats=st, mask=(Yl, sc, sh, mu, inv, dual))


def wgrad(M, Nn, gk, xk):
    G, _ = make_opnd(gk, R, M, g)
    X, _ = make_opnd(xk, R, Nn, g)
    dW = torch.zeros(M, Nn, device=dev)
    return lambda: tc.gemm_wgrad(R, G, X, dW)


cases = [rows(32, 32, "aff", "maskdual"), rows(20, 64, "raw", "stats"), rows(64, 32, "raw", "mask"), rows(32, 64, "bn", "stats"),
         rows(32, 32, "dual", "stats"), wgrad(64, 20, "affwrap", "raw"), wgrad(32, 32, "aff", "dual"), wgrad(64, 32, "raw", "bn")]
for f in cases:       # warm-up (not captured: ncu -s skips these launches)
    f()
torch.cuda.synchronize()
for f in cases:
    f()
torch.cuda.synchronize()
