"""GPU suite: the tcgen05 3xTF32 GEMM (csrc/mlp_tc.cu) against an fp64 reference of the same op.
Tolerance: 1e-5 relative to the output's magnitude (north star: MLP features within 1e-5 rel fp32)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
cuda = torch.device("cuda")


def _run(rows, K, N, mode, bias=True, stats=True, seed=0):
    from repsurf_b200 import tc
    g = torch.Generator().manual_seed(seed)
    ldx = K * (2 if mode == 2 else 1)
    X = torch.randn(rows, ldx, generator=g).to(cuda)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda)
    b = torch.randn(N, generator=g).to(cuda) if bias else None
    sc = (torch.rand(ldx, generator=g) + 0.5).to(cuda)
    sh = (torch.randn(ldx, generator=g) * 0.3).to(cuda)
    Y, st = tc.linear_forward(X, W, b, mode, sc if mode else None, sh if mode else None, want_stats=stats)
    Xd, Wd = X.double(), W.double()
    if mode == 0:
        A = Xd
    elif mode == 1:
        A = torch.relu(Xd * sc.double() + sh.double())
    else:
        A = torch.relu(Xd[:, :K] * sc[:K].double() + sh[:K].double() + Xd[:, K:] * sc[K:].double() + sh[K:].double())
    ref = A @ Wd.t() + (b.double() if bias else 0)
    scale = ref.abs().max().item()
    err = (Y.double() - ref).abs().max().item() / scale
    assert err < 1e-5, (rows, K, N, mode, err)
    if stats:
        s1, s2 = ref.sum(0), (ref * ref).sum(0)
        assert ((st[:N] - s1).abs().max() / s1.abs().max().clamp_min(1.0)).item() < 1e-5
        assert ((st[N:] - s2).abs().max() / s2.abs().max()).item() < 1e-5
    return err


@pytest.mark.parametrize("rows,K,N", [(128, 32, 32), (1000, 3, 32), (4096, 19, 64), (777, 64, 64), (5000, 74, 128),
                                      (2048, 138, 256), (300, 266, 512), (130, 512, 1024), (1, 8, 16), (129, 33, 10)])
def test_linear_tc_plain(rows, K, N):
    _run(rows, K, N, 0)


@pytest.mark.parametrize("rows,K,N,mode", [(3000, 32, 64, 1), (3000, 32, 32, 2), (500, 128, 256, 1), (999, 64, 128, 2)])
def test_linear_tc_prologue_modes(rows, K, N, mode):
    _run(rows, K, N, mode)


def test_linear_tc_many_tiles_persistent():
    # more row tiles than SMs: exercises the persistent loop, both accumulator buffers and stage phases
    _run(148 * 128 * 3 + 77, 40, 48, 1, seed=3)
