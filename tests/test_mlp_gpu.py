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


# ------------------------------------------------------------------------------------------------------------
# fused shared MLP (forward + backward on tcgen05) vs the same block in torch fp64
# ------------------------------------------------------------------------------------------------------------
class _Block(torch.nn.Module):
    def __init__(self, pos_c, feat_c, mlp, conv):
        super().__init__()
        import torch.nn as nn
        Conv, BN = (nn.Conv2d, nn.BatchNorm2d) if conv == 2 else (nn.Conv1d, nn.BatchNorm1d)
        self.mlp_l0, self.mlp_f0 = Conv(pos_c, mlp[0], 1), Conv(feat_c, mlp[0], 1)
        self.bn_l0, self.bn_f0 = BN(mlp[0]), BN(mlp[0])
        self.mlp_convs = nn.ModuleList(Conv(i, o, 1) for i, o in zip(mlp[:-1], mlp[1:]))
        self.mlp_bns = nn.ModuleList(BN(o) for o in mlp[1:])
        for m in self.modules():
            if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
                torch.nn.init.uniform_(m.weight, 0.5, 1.5)
                torch.nn.init.normal_(m.bias, 0, 0.2)


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12)).item()


@pytest.mark.parametrize("train", [True, False])
@pytest.mark.parametrize("G,ns,pos_c,feat_c,mlp,conv", [(300, 32, 3, 16, [32, 32, 64], 1), (64, 64, 6, 138, [128, 128, 256], 2),
                                                         (40, 32, 3, 266, [256, 256, 512], 1), (1000, 16, 6, 10, [64, 64, 128], 2),
                                                         (37, 24, 3, 17, [32, 32, 64], 1)])   # ragged: rows % 32 != 0
def test_fused_sa_mlp_matches_fp64_torch(G, ns, pos_c, feat_c, mlp, conv, train):
    """Fused shared MLP + max-pool (tc.sa_mlp_fused) against the same nn modules evaluated in fp64, in training mode
    (batch statistics, running-buffer side effects) and in eval mode (running statistics; backward = frozen BatchNorm).
    The routing of the checker (max-pool arg-max, ReLU masks) is pinned to the kernel's own decisions: both are only
    piecewise differentiable and an fp32-vs-fp64 flip of one near-zero element moves a whole gradient row.  The free
    fp64 evaluation is compared on the forward values; gradients, with the routing fixed, at 1e-4."""
    import copy
    from repsurf_b200 import tc
    from tests.torch_ref import kernel_relu_masks, sa_mlp_rows
    torch.manual_seed(G + ns)
    blk = _Block(pos_c, feat_c, mlp, conv).to(cuda)
    with torch.no_grad():
        for m in blk.modules():
            if hasattr(m, "running_mean"):
                m.running_mean.normal_(0, 0.2)
                m.running_var.uniform_(0.5, 2.0)
    blk.train(train)
    ref = copy.deepcopy(blk).double()
    X = torch.randn(G * ns, pos_c + feat_c, device=cuda)
    Xa = X.clone().requires_grad_(True)
    Xb = X.double().requires_grad_(True)
    out = tc.sa_mlp_fused(Xa, pos_c, blk, ns)
    arg = out.grad_fn.saved[-1]                          # [G, C'] sample index the kernel pooled
    free = sa_mlp_rows(Xb.detach(), pos_c, copy.deepcopy(ref), ns)
    assert _rel(out, free) < 1e-5                        # the free max-pool agrees on the values
    want = sa_mlp_rows(Xb, pos_c, ref, ns, arg=arg, masks=kernel_relu_masks(out.grad_fn))
    assert _rel(out, want) < 1e-5
    go = torch.randn_like(out)
    out.backward(go)
    want.backward(go.double())
    # running statistics: updated in training mode, untouched in eval mode
    for a, b in zip(blk.buffers(), ref.buffers()):
        assert _rel(a.double(), b) < 1e-5 or a.dtype == torch.int64 and int(a) == int(b)
    # gradients: feature columns of X (position columns carry none on the RepSurf path), weights, BN affine
    assert _rel(Xa.grad[:, pos_c:], Xb.grad[:, pos_c:]) < 1e-4
    assert float(Xa.grad[:, :pos_c].abs().max()) == 0.0
    for (n, p), (_, q) in zip(blk.named_parameters(), ref.named_parameters()):
        if train and n.endswith("bias") and "bn" not in n:
            # a bias in front of a batch-statistics BatchNorm has an exactly zero gradient
            assert float(p.grad.abs().max()) == 0.0 and float(q.grad.abs().max()) < 1e-6 * max(1.0, float(go.abs().sum()))
        else:
            assert _rel(p.grad, q.grad) < 1e-4, n
    with pytest.raises(RuntimeError):                    # the saved activations were consumed by the first backward
        out.backward(go)


def test_fused_sa_mlp_propagates_nan():
    """A NaN activation must reach the output as NaN (torch.relu / torch.max semantics), not be clamped to zero."""
    from repsurf_b200 import tc
    torch.manual_seed(3)
    blk = _Block(3, 13, [32, 32, 64], 1).to(cuda).eval()
    X = torch.randn(40 * 16, 16, device=cuda)
    X[5 * 16 + 3, 7] = float("nan")
    out = tc.sa_mlp_fused(X, 3, blk, 16)
    assert torch.isnan(out[5]).all() and not torch.isnan(out[6:]).any() and not torch.isnan(out[:5]).any()


@pytest.mark.parametrize("train", [True, False])
@pytest.mark.parametrize("R,K,Nn,relu", [(5000, 128, 128, True), (777, 512, 256, False), (3000, 64, 256, True), (100, 256, 128, False),
                                         (4096, 10, 10, True), (999, 9, 9, True)])   # 10 / 9 channels: zero-padded to 12
def test_linear_bn_layer_matches_fp64_torch(R, K, Nn, relu, train):
    import copy
    import torch.nn as nn
    from repsurf_b200 import tc
    torch.manual_seed(R)
    lin, bn = nn.Linear(K, Nn).to(cuda), nn.BatchNorm1d(Nn).to(cuda)
    nn.init.uniform_(bn.weight, 0.5, 1.5)
    nn.init.normal_(bn.bias, 0, 0.2)
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.2)
        bn.running_var.uniform_(0.5, 2.0)
    bn.train(train)
    lin2, bn2 = copy.deepcopy(lin).double(), copy.deepcopy(bn).double()
    Kp = (K + 3) // 4 * 4
    x = torch.randn(R, K, device=cuda)
    xp = torch.zeros(R, Kp, device=cuda)
    xp[:, :K] = x
    xa, xb = xp.clone().requires_grad_(True), x.double().requires_grad_(True)
    out = tc.linear_bn(xa, lin, bn, relu)
    want = bn2(lin2(xb))
    if relu:
        # ReLU decisions pinned to the kernel's (z = fma(Y, sc, sh) in fp32 from its stored pre-BatchNorm output)
        _X, _W, Ysv, sc, sh = out.grad_fn.saved[:5]
        mask = ((Ysv.double() * sc.double() + sh.double()).float() > 0)[:, :Nn]
        assert _rel(out[:, :Nn].detach(), torch.relu(want).detach()) < 1e-5
        want = want * mask.double()
    assert out.shape[1] == (Nn + 3) // 4 * 4 and float(out[:, Nn:].detach().abs().sum()) == 0.0
    assert _rel(out[:, :Nn], want) < 1e-5
    go = torch.randn_like(want).float()
    gop = torch.zeros_like(out)
    gop[:, :Nn] = go
    out.backward(gop)
    want.backward(go.double())
    assert _rel(xa.grad[:, :K], xb.grad) < 2e-4
    assert _rel(lin.weight.grad, lin2.weight.grad) < 2e-4
    assert _rel(bn.weight.grad, bn2.weight.grad) < 2e-4 and _rel(bn.bias.grad, bn2.bias.grad) < 2e-4
    if not train:
        assert _rel(lin.bias.grad, lin2.bias.grad) < 2e-4
    assert _rel(bn.running_mean.double(), bn2.running_mean) < 1e-5 and _rel(bn.running_var.double(), bn2.running_var) < 1e-5
    assert int(bn.num_batches_tracked) == int(bn2.num_batches_tracked)


def test_plain_linear_matches_fp64_torch():
    import copy
    import torch.nn as nn
    from repsurf_b200 import tc
    torch.manual_seed(5)
    lin = nn.Linear(128, 13).to(cuda)
    lin2 = copy.deepcopy(lin).double()
    x = torch.randn(4000, 128, device=cuda)
    xa, xb = x.clone().requires_grad_(True), x.double().requires_grad_(True)
    out, want = tc.linear(xa, lin), lin2(xb)
    assert _rel(out, want) < 1e-5
    go = torch.randn_like(out)
    out.backward(go)
    want.backward(go.double())
    assert _rel(xa.grad, xb.grad) < 1e-5 and _rel(lin.weight.grad, lin2.weight.grad) < 1e-5 and _rel(lin.bias.grad, lin2.bias.grad) < 1e-5


# ------------------------------------------------------------------------------------------- fused umbrella MLP
@pytest.mark.parametrize("n,g,train", [(5000, 9, True), (777, 8, True), (3000, 5, True), (2000, 9, False)])
def test_umbrella_mlp_fused_matches_fp64_modules(n, g, train):
    """csrc/umbrella_mlp.cu (Conv1d+BN+ReLU+Conv1d+sum over triangles, recomputing) against the same nn modules
    evaluated in fp64 (segmentation/modules/repsurface_utils.py:297-302, :322-327): output, all six parameter
    gradients and the BatchNorm running buffers."""
    import copy
    import torch.nn as nn
    from repsurf_b200 import tc
    gen = torch.Generator().manual_seed(5)
    feat = torch.randn(n, g, 10, generator=gen).to(cuda)
    mlps = nn.Sequential(nn.Conv1d(10, 10, 1), nn.BatchNorm1d(10), nn.ReLU(True), nn.Conv1d(10, 10, 1)).to(cuda)
    with torch.no_grad():
        mlps[1].weight.uniform_(0.5, 1.5)
        mlps[1].bias.normal_(0, 0.3)
        mlps[1].running_mean.normal_(0, 0.1)
        mlps[1].running_var.uniform_(0.5, 2.0)
    ref = copy.deepcopy(mlps).double()
    mlps.train(train)
    ref.train(train)
    # the ReLU gradient is discontinuous at z == 0: points with a pre-activation within 1e-4 of it get a zero output
    # gradient, so that an fp32 and an fp64 evaluation cannot disagree on a mask that matters (one flipped element
    # moves dW1 by ~1/sqrt(rows) of its norm)
    w = torch.randn(n, 10, generator=gen).to(cuda)
    with torch.no_grad():
        y = ref[0](feat.double().view(n * g, 10).t().unsqueeze(0))
        z = nn.functional.batch_norm(y, None if train else ref[1].running_mean, None if train else ref[1].running_var,
                                     ref[1].weight, ref[1].bias, train, 0.0, ref[1].eps).squeeze(0).t()
        w[(z.abs().amin(1) < 1e-4).view(n, g).any(1)] = 0
    out = tc.umbrella_mlp_fused(feat, mlps[0], mlps[1], mlps[3])
    (out * w).sum().backward()
    want = ref(feat.double().view(n * g, 10).t().unsqueeze(0)).squeeze(0).t().reshape(n, g, 10).sum(1)
    (want * w.double()).sum().backward()
    scale = want.abs().max().item()
    assert ((out.double() - want).abs().max() / scale).item() < 1e-5
    errs = {}
    gmax = max(q.grad.abs().max().item() for q in ref.parameters())
    for (name, p), (_, q) in zip(mlps.named_parameters(), ref.named_parameters()):
        # conv1.bias has an analytically zero gradient under train-mode BatchNorm: absolute criterion there
        den = q.grad.abs().max().item() if (name != "0.bias" or not train) else gmax
        errs[name] = ((p.grad.double() - q.grad).abs().max() / den).item()
    assert max(errs.values()) < 1e-4, errs
    if train:
        assert torch.allclose(mlps[1].running_mean.double(), ref[1].running_mean, atol=1e-5)
        assert torch.allclose(mlps[1].running_var.double(), ref[1].running_var, rtol=1e-5, atol=1e-6)
    assert int(mlps[1].num_batches_tracked) == int(ref[1].num_batches_tracked)


@pytest.mark.parametrize("n,M,ns,cn,cf,mlp", [(5000, 700, 32, 10, 6, [32, 32, 64]), (3000, 500, 32, 10, 64, [64, 64, 128]),
                                              (900, 37, 24, 10, 256, [256, 256, 512]), (4000, 1000, 16, 10, 0, [32, 64])])
def test_gather_fused_first_layer_equals_materialised_rows(n, M, ns, cn, cf, mlp):
    """First shared-MLP layer reading its rows through TMA gather4 from the per-point table (RSB_OPND_GATHER, forward and
    weight gradient) and scattering the input gradient from the GEMM epilogue, against the same block fed with the row
    matrix group_rows_fwd materialises: forward bit-identical (same subtraction, same GEMM), gradients to atomics' order."""
    import copy
    from repsurf_b200.mlp import gather_rows, group_rows, sa_mlp
    g = torch.Generator().manual_seed(n + M)
    xyz = torch.rand(n, 3, generator=g).to(cuda)
    new_xyz = xyz[torch.randperm(n, generator=g)[:M].to(cuda)].contiguous()
    idx = torch.randint(0, n, (M, ns), generator=g, dtype=torch.int32).to(cuda)
    normal = torch.randn(n, cn, generator=g).to(cuda)
    feat = torch.randn(n, cf, generator=g).to(cuda) if cf else None
    blk = _Block(3, cn + cf, mlp, 1).to(cuda).train()
    ref = copy.deepcopy(blk)
    na, fa = normal.clone().requires_grad_(True), (feat.clone().requires_grad_(True) if cf else None)
    nb, fb = normal.clone().requires_grad_(True), (feat.clone().requires_grad_(True) if cf else None)
    out = sa_mlp(gather_rows(xyz, new_xyz, idx, na, fa, ns), 3, blk, ns)
    rows, layout = group_rows(xyz, new_xyz, idx, nb, fb, ns, False)
    want = sa_mlp(rows, 3, ref, ns, layout)
    assert torch.equal(out, want)
    go = torch.randn_like(out)
    out.backward(go)
    want.backward(go)
    assert _rel(na.grad, nb.grad) < 1e-5
    if cf:
        assert _rel(fa.grad, fb.grad) < 1e-5
    for (name, p), (_, q) in zip(blk.named_parameters(), ref.named_parameters()):
        assert _rel(p.grad, q.grad) < 2e-5 or float(q.grad.abs().max()) == 0.0, name
    for a, b in zip(blk.buffers(), ref.buffers()):
        assert torch.equal(a, b)
