"""Python face of the tcgen05 shared-MLP kernels (csrc/mlp_tc.cu) — thin, allocation + launch only."""
import torch

from . import _native as N


def prep_weight(W, transposed=False):
    """W [N,K] (or [K,N] when transposed) fp32 -> pre-split hi/lo tf32 operand buffer for linear_forward."""
    W = W.detach()
    if transposed:
        K, Nn = W.shape
    else:
        Nn, K = W.shape
    W = W.contiguous()
    buf = torch.empty(int(N.lib().rsb_linear_tc_weight_floats(Nn, K)), dtype=torch.float32, device=W.device)
    N.call("rsb_linear_tc_prep_weight", Nn, K, W, W.shape[1], 1 if transposed else 0, buf)
    return buf, Nn, K


def linear_forward(X, W, bias=None, mode=0, sc=None, sh=None, want_stats=False, prepped=None):
    """Y = act(X) @ W^T + bias on the tensor cores (3xTF32).  Returns (Y [rows,N], stats fp64 [2N] | None)."""
    Wp, Nn, K = prepped if prepped is not None else prep_weight(W)
    rows = X.shape[0]
    assert X.is_contiguous() and X.shape[1] == K * (2 if mode == 2 else 1)
    Y = torch.empty(rows, Nn, dtype=torch.float32, device=X.device)
    stats = torch.zeros(2 * Nn, dtype=torch.float64, device=X.device) if want_stats else None
    N.call("rsb_linear_tc_forward", rows, K, Nn, X, X.shape[1], Wp, bias, mode, sc, sh, Y, stats)
    return Y, stats
