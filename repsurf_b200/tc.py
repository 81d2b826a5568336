"""Python face of the tcgen05 shared-MLP kernels (csrc/mlp_tc.cu, csrc/mlp_aux.cu): descriptor structs,
launch helpers, and the fused shared-MLP autograd function used by the SurfaceAbstractionCD modules.

Nothing here computes on the host or through torch GEMM/BatchNorm kernels: torch only allocates."""
import ctypes

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _native as N

OPND_RAW, OPND_BN_RELU, OPND_DUAL, OPND_AFFINE2, OPND_POOLED, OPND_GATHER = 0, 1, 2, 3, 4, 5
EPI_BIAS_STATS, EPI_RELU_MASK = 0, 1


class Opnd(ctypes.Structure):
    _fields_ = [("U", ctypes.c_void_p), ("V", ctypes.c_void_p), ("a", ctypes.c_void_p), ("b", ctypes.c_void_p),
                ("d", ctypes.c_void_p), ("arg", ctypes.c_void_p), ("ldu", ctypes.c_int), ("ldv", ctypes.c_int),
                ("K", ctypes.c_int), ("k0", ctypes.c_int), ("ku", ctypes.c_int), ("kind", ctypes.c_int),
                ("ns", ctypes.c_int)]


class Epi(ctypes.Structure):
    _fields_ = [("Y", ctypes.c_void_p), ("ldy", ctypes.c_int), ("bias", ctypes.c_void_p), ("stats", ctypes.c_void_p),
                ("Yl", ctypes.c_void_p), ("ldl", ctypes.c_int), ("sc", ctypes.c_void_p), ("sh", ctypes.c_void_p),
                ("mu", ctypes.c_void_p), ("inv", ctypes.c_void_p), ("kind", ctypes.c_int), ("dual", ctypes.c_int),
                ("scatter", ctypes.c_void_p)]


def _dp(t, off=0):
    return None if t is None else t.data_ptr() + 4 * off


def opnd(kind, U, K, a=None, d=None, V=None, b=None, arg=None, ku=None, k0=0, ns=1):
    o = Opnd()
    o.U, o.V, o.a, o.b, o.d, o.arg = _dp(U), _dp(V), _dp(a), _dp(b), _dp(d), (None if arg is None else arg.data_ptr())
    o.ldu = U.stride(0)
    o.ldv = 0 if V is None else V.stride(0)
    o.K, o.k0, o.ku, o.kind, o.ns = K, k0, (K if ku is None else ku), kind, ns
    o._keep = (U, V, a, b, d, arg)
    return o


def prep_weight(W, transposed=False):
    """W [N,K] (or stored [K,N] when transposed) fp32 -> pre-split hi/lo tf32 operand buffer."""
    W = W.detach().contiguous()
    if transposed:
        K, Nn = W.shape
    else:
        Nn, K = W.shape
    buf = torch.empty(int(N.lib().rsb_linear_tc_weight_floats(Nn, K)), dtype=torch.float32, device=W.device)
    N.call("rsb_linear_tc_prep_weight", Nn, K, W, W.shape[1], 1 if transposed else 0, buf)
    return buf, Nn, K


def gemm_rows(rows, Nn, A, Wp, Y=None, ldy=None, y_off=0, bias=None, stats=None, mask=None, scatter=None):
    """Y[:, y_off:y_off+N] = A @ W^T (+bias) ; mask = (Yl, sc, sh, mu, inv, dual) selects the dgrad epilogue;
    scatter = int32 [rows]: Y[scatter[r], y_off:] += row r instead (grouping backward fused into the GEMM)."""
    e = Epi()
    e.scatter = None if scatter is None else scatter.data_ptr()
    e.Y = _dp(Y, y_off)
    e.ldy = (Y.stride(0) if ldy is None else ldy) if Y is not None else 0
    e.bias, e.stats = _dp(bias), (None if stats is None else stats.data_ptr())
    if mask is None:
        e.kind, e.dual = EPI_BIAS_STATS, 0
    else:
        Yl, sc, sh, mu, inv, dual = mask
        e.Yl, e.ldl, e.sc, e.sh, e.mu, e.inv = _dp(Yl), Yl.stride(0), _dp(sc), _dp(sh), _dp(mu), _dp(inv)
        e.kind, e.dual = EPI_RELU_MASK, int(dual)
    N.call("rsb_gemm_rows", rows, Nn, A, Wp, e)


def gemm_wgrad(rows, G, X, dW):
    N.call("rsb_gemm_wgrad", rows, G, X, dW, dW.stride(0))


def linear_forward(X, W, bias=None, mode=0, sc=None, sh=None, want_stats=False, prepped=None):
    """Y = act(X) @ W^T + bias on the tensor cores (3xTF32).  Returns (Y [rows,N], stats fp64 [2N] | None)."""
    Wp, Nn, K = prepped if prepped is not None else prep_weight(W)
    rows = X.shape[0]
    assert X.is_contiguous() and X.shape[1] == K * (2 if mode == 2 else 1)
    Y = torch.empty(rows, Nn, dtype=torch.float32, device=X.device)
    stats = torch.zeros(2 * Nn, dtype=torch.float64, device=X.device) if want_stats else None
    N.call("rsb_linear_tc_forward", rows, K, Nn, X, X.shape[1], Wp, bias, mode, sc, sh, Y, stats)
    return Y, stats


def _bn_finalize(C, rows, stats, gamma, beta, eps, momentum, rm, rv, dev):
    out = torch.empty(4, C, dtype=torch.float32, device=dev)
    N.call("rsb_bn_finalize", C, rows, stats, gamma, beta, float(eps), float(momentum), rm, rv, out[0], out[1], out[2], out[3])
    return out[0], out[1], out[2], out[3]


def _bn_eval_coef(bn_list, pad_to=None):
    """Running-statistics BatchNorm(s) -> (sc, sh, mu, inv) rows of one [4, sum C] tensor (rsb_bn_eval_coef, one launch per
    module).  pad_to: total width with trailing identity channels (sc = 1/sqrt(1+eps), sh = mu = 0)."""
    dev = bn_list[0].running_mean.device
    widths = [b.running_mean.shape[0] for b in bn_list]
    total = sum(widths) if pad_to is None else pad_to
    out = torch.zeros(4, total, dtype=torch.float32, device=dev) if total > sum(widths) else \
        torch.empty(4, total, dtype=torch.float32, device=dev)
    c0 = 0
    for b, w in zip(bn_list, widths):
        N.call("rsb_bn_eval_coef", w, None if b.weight is None else b.weight.detach(), None if b.bias is None else b.bias.detach(),
               b.running_mean, b.running_var, float(b.eps), out[0, c0:c0 + w], out[1, c0:c0 + w], out[2, c0:c0 + w], out[3, c0:c0 + w])
        c0 += w
    return out[0], out[1], out[2], out[3]


def bn_uses_batch_stats(bn):
    return bn.training or bn.running_mean is None


class _FusedSAMLP(Function):
    """relu(bn_l(X_pos W_l^T) + bn_f(X_feat W_f^T)) -> [relu(bn(. W_i^T))]* -> max over nsample, fused.

    forward(X [R, Cin], meta, *params) with params = W_l, b_l, W_f, b_f, g_l, be_l, g_f, be_f, then (W, b, g, be) per
    further layer.  Running statistics are updated by the caller from the returned batch statistics."""

    @staticmethod
    def forward(ctx, X, meta, *params):
        ns, P, eps, n_extra = meta["ns"], meta["pos_channel"], meta["eps"], meta["n_extra"]
        frozen = meta["frozen"]          # BatchNorm layers normalise with their running statistics (eval / frozen)
        bns = meta["bns"]
        dev = X.device
        gat = meta.get("gather")            # (idx int32 [R], centres [G, 3]): X is the per-point table, rows are gathered by TMA
        Cin = X.shape[1]
        R = gat[0].numel() if gat is not None else X.shape[0]
        G = R // ns
        W_l, b_l, W_f, b_f, g_l, be_l, g_f, be_f = params[:8]
        C0 = W_l.shape[0]
        W_l2, W_f2 = W_l.reshape(C0, -1), W_f.reshape(C0, -1)
        P4, F = meta["feat_col"], meta["feat_channels"]      # feature columns live at [P4, P4 + F)
        # block-diagonal first layer: one GEMM produces [y_l | y_f]
        Wbd = torch.zeros(2 * C0, Cin, device=dev)
        Wbd[:C0, :P] = W_l2
        Wbd[C0:, P4:P4 + F] = W_f2
        bias0 = torch.cat([b_l, b_f]).detach()
        Wp0, _, _ = prep_weight(Wbd)
        Y0 = torch.empty(R, 2 * C0, device=dev)
        # batch statistics of every layer: slices of ONE zeroed fp64 buffer (one fill launch per block)
        widths = [4 * C0] + [2 * params[8 + 4 * i].shape[0] for i in range(n_extra)]
        if frozen:
            stats_of = [None] * (1 + n_extra)
        else:
            stbuf = torch.zeros(sum(widths), dtype=torch.float64, device=dev)
            stats_of = list(torch.split(stbuf, widths))
        st = stats_of[0]
        A0 = opnd(OPND_RAW, X, Cin) if gat is None else opnd(OPND_GATHER, X, Cin, V=gat[1], arg=gat[0], ku=X.shape[0], ns=ns)
        gemm_rows(R, 2 * C0, A0, Wp0, Y=Y0, bias=bias0, stats=st)
        if frozen:
            coefs = [_bn_eval_coef(bns[:2])]
        else:
            coefs = [_bn_finalize(2 * C0, R, st, torch.cat([g_l, g_f]).detach(), torch.cat([be_l, be_f]).detach(), eps, 0.0,
                                  None, None, dev)]
        batch_stats = [] if frozen else [st]
        Ys = [Y0]
        Ws = []
        prev = opnd(OPND_DUAL, Y0, C0, a=coefs[0][0], d=coefs[0][1], ku=C0)
        Cprev = C0
        for i in range(n_extra):
            W, b, g, be = params[8 + 4 * i: 12 + 4 * i]
            Ci = W.shape[0]
            W2 = W.reshape(Ci, -1).detach().contiguous()
            Wp, _, _ = prep_weight(W2)
            Yi = torch.empty(R, Ci, device=dev)
            sti = stats_of[i + 1]
            gemm_rows(R, Ci, prev, Wp, Y=Yi, bias=b.detach(), stats=sti)
            co = _bn_eval_coef([bns[2 + i]]) if frozen else _bn_finalize(Ci, R, sti, g.detach(), be.detach(), eps, 0.0, None, None, dev)
            coefs.append(co)
            if not frozen:
                batch_stats.append(sti)
            Ys.append(Yi)
            Ws.append(W2)
            prev = opnd(OPND_BN_RELU, Yi, Ci, a=co[0], d=co[1])
            Cprev = Ci
        out = torch.empty(G, Cprev, device=dev)
        arg = torch.empty(G, Cprev, dtype=torch.int32, device=dev)
        N.call("rsb_pool_forward", G, ns, Cprev, Ys[-1], Ys[-1].stride(0), coefs[-1][0], coefs[-1][1], out, arg)
        ctx.meta = meta
        ctx.saved = (X, Wbd, Ws, Ys, coefs, arg)
        ctx.consumed = False
        ctx.mark_non_differentiable(*batch_stats)
        return (out, *batch_stats)

    @staticmethod
    @once_differentiable
    def backward(ctx, dOut, *_unused):
        meta = ctx.meta
        ns, P, n_extra = meta["ns"], meta["pos_channel"], meta["n_extra"]
        if ctx.consumed:
            # the stored pre-BatchNorm activations are overwritten in place by dL/dY (rsb_pool_bn_backward_dense): a second
            # backward through the same graph (retain_graph=True) would read gradients where it expects activations
            raise RuntimeError("repsurf_b200 fused shared MLP: backward called twice on the same forward (its saved activations "
                               "are consumed by the first backward); run the forward again instead of retain_graph=True")
        ctx.consumed = True
        fz = 2 if meta["frozen"] else 0
        X, Wbd, Ws, Ys, coefs, arg = ctx.saved
        dev = X.device
        gat = meta.get("gather")
        Cin = X.shape[1]
        R = gat[0].numel() if gat is not None else X.shape[0]
        G = R // ns
        C0 = Wbd.shape[0] // 2
        L = n_extra  # index of the last layer (0 = block-diagonal first layer)
        CL = Ys[L].shape[1]
        sc, sh, mu, inv = coefs[L]
        grads = {}
        # zeroed accumulators of the whole backward: slices of one fp64 buffer (BatchNorm-backward statistics) and of
        # one fp32 buffer (weight gradients) - two fill launches per block instead of two per layer
        st_w = [2 * CL] + [(3 if l - 1 == 0 else 2) * Ws[l - 1].shape[1] for l in range(L, 0, -1)]
        st_parts = list(torch.split(torch.zeros(sum(st_w), dtype=torch.float64, device=dev), st_w))
        dw_n = [Ws[l - 1].numel() for l in range(L, 0, -1)] + [2 * C0 * Cin]
        dw_parts = list(torch.split(torch.zeros(sum(dw_n), device=dev), dw_n))
        # ---- max-pool backward -> BatchNorm-backward coefficients of the last layer
        st = st_parts.pop(0)
        dm = torch.empty(G, CL, device=dev)
        N.call("rsb_pool_backward_stats", G, ns, CL, dOut.contiguous(), arg, Ys[L], Ys[L].stride(0), sc, sh, mu, inv, dm, st)
        co = torch.empty(5, CL, device=dev)
        N.call("rsb_bn_backward_coef", CL, R, st, fz, sc, mu, inv, co[0], co[1], co[2], co[3], co[4])
        grads[("g", L)], grads[("be", L)] = co[3], co[4]
        scales = {L: sc}
        if L > 0 and CL % 4 == 0:
            # densify dL/dY_L in place of the stored Y_L (one streaming pass); the GEMMs then read a plain matrix
            N.call("rsb_pool_bn_backward_dense", G, ns, CL, dm, arg, Ys[L], Ys[L].stride(0), co[0], co[1], co[2])
            Gop = opnd(OPND_RAW, Ys[L], CL)
        else:
            Gop = opnd(OPND_POOLED, dm, CL, a=co[0], b=co[1], d=co[2], V=Ys[L], arg=arg, ns=ns) if L > 0 else None
        if L == 0:
            raise RuntimeError("shared MLP needs at least two layers")
        for l in range(L, 0, -1):
            W = Ws[l - 1]                                   # [C_l, C_{l-1}]
            Cl, Cp = W.shape
            # forward operand of layer l (= activated output of layer l-1)
            if l - 1 == 0:
                Aprev = opnd(OPND_DUAL, Ys[0], C0, a=coefs[0][0], d=coefs[0][1], ku=C0)
            else:
                Aprev = opnd(OPND_BN_RELU, Ys[l - 1], Cp, a=coefs[l - 1][0], d=coefs[l - 1][1])
            dW = dw_parts.pop(0).view(Cl, Cp)
            gemm_wgrad(R, Gop, Aprev, dW)
            grads[("W", l)] = dW
            # dgrad through W_l, ReLU mask + BatchNorm-backward statistics of layer l-1
            WpT, _, _ = prep_weight(W, transposed=True)      # operand [N=C_{l-1}, K=C_l]
            dual = (l - 1 == 0)
            dZ = torch.empty(R, Cp, device=dev)
            stp = st_parts.pop(0)
            scp, shp, mup, invp = coefs[l - 1]
            # dgrad with the ReLU mask of layer l-1 and its BatchNorm-backward statistics fused into the epilogue
            gemm_rows(R, Cp, Gop, WpT, Y=dZ, stats=stp, mask=(Ys[l - 1], scp, shp, mup, invp, dual))
            width = 2 * Cp if dual else Cp
            cop = torch.empty(5, width, device=dev)
            N.call("rsb_bn_backward_coef", Cp, R, stp, (1 if dual else 0) | fz, scp, mup, invp, cop[0], cop[1], cop[2], cop[3], cop[4])
            grads[("g", l - 1)], grads[("be", l - 1)] = cop[3], cop[4]
            scales[l - 1] = scp
            Gop = opnd(OPND_AFFINE2, dZ, width, a=cop[0], b=cop[1], d=cop[2], V=Ys[l - 1], ku=Cp)
            dZ0, cop0 = dZ, cop
        # ---- first layer: weight gradient of the block-diagonal GEMM, input gradient of the feature columns
        dWbd = dw_parts.pop(0).view(2 * C0, Cin)
        X0 = opnd(OPND_RAW, X, Cin) if gat is None else opnd(OPND_GATHER, X, Cin, V=gat[1], arg=gat[0], ku=X.shape[0], ns=ns)
        gemm_wgrad(R, Gop, X0, dWbd)
        dX = None
        P4, F = meta["feat_col"], meta["feat_channels"]
        if ctx.needs_input_grad[0]:
            dX = torch.zeros(X.shape[0], Cin, device=dev)
            Wf = Wbd[C0:, P4:P4 + F].contiguous()            # [C0, F]
            WpT, _, _ = prep_weight(Wf, transposed=True)     # operand [N=F, K=C0]
            Gf = opnd(OPND_AFFINE2, dZ0, C0, a=cop0[0], b=cop0[1], d=cop0[2], V=Ys[0], ku=C0, k0=C0)
            # gathered rows: the gradient of row r goes to the table row it was gathered from (scatter fused into the GEMM)
            gemm_rows(R, F, Gf, WpT, Y=dX, ldy=Cin, y_off=P4, scatter=None if gat is None else gat[0])
        # ---- assemble parameter gradients in the order of *params
        W_l_shape, W_f_shape = meta["W_l_shape"], meta["W_f_shape"]
        # biases in front of a train-mode BatchNorm have an exactly zero gradient: slices of one zero buffer
        zbuf = torch.zeros(2 * C0 + sum(grads[("W", i + 1)].shape[0] for i in range(n_extra)), device=dev)
        zpos = [0]

        def zero(n):
            z = zbuf[zpos[0]:zpos[0] + n]
            zpos[0] += n
            return z

        def dbias(layer, lo, hi):
            # frozen BatchNorm: dL/db = sum_r dY = sc * sum_r dZ  (per-channel vectors)
            return (scales[layer][lo:hi] * grads[("be", layer)][lo:hi]) if fz else zero(hi - lo)
        out = [dX, None,
               dWbd[:C0, :P].reshape(W_l_shape), dbias(0, 0, C0), dWbd[C0:, P4:P4 + F].reshape(W_f_shape), dbias(0, C0, 2 * C0),
               grads[("g", 0)][:C0], grads[("be", 0)][:C0], grads[("g", 0)][C0:], grads[("be", 0)][C0:]]
        for i in range(n_extra):
            Ci = grads[("W", i + 1)].shape[0]
            out += [grads[("W", i + 1)].reshape(meta["W_shapes"][i]), dbias(i + 1, 0, Ci),
                    grads[("g", i + 1)], grads[("be", i + 1)]]
        return tuple(out)


def _update_running(bn, R, sums, sumsq):
    """Side effects of a train-mode nn.BatchNorm forward on its buffers, from fp64 batch sums (views into the statistics
    vector of the GEMM epilogue): one kernel (rsb_bn_update_running)."""
    if bn.track_running_stats and bn.running_mean is not None:
        m = bn.momentum if bn.momentum is not None else 0.1
        N.call("rsb_bn_update_running", sums.shape[0], R, sums, sumsq, float(m), bn.running_mean, bn.running_var,
               bn.num_batches_tracked)


def sa_mlp_fused(rows, pos_channel, mod, nsample, layout=None):
    """Shared MLP + max-pool of a SurfaceAbstractionCD level on the tensor cores.  rows [G*nsample, C] -> [G, mlp[-1]].
    layout = (first feature column, feature channels) of the packed row matrix (csrc/group.cu group_rows_fwd).
    BatchNorm layers in training mode use batch statistics and update their running buffers; in eval mode (or frozen:
    bn.eval() inside a training module) they normalise with the running statistics, as nn.BatchNorm does."""
    bns = [mod.bn_l0, mod.bn_f0] + list(mod.mlp_bns)
    modes = {bn_uses_batch_stats(b) for b in bns}
    if len(modes) != 1:
        raise RuntimeError("repsurf_b200 fused shared MLP: the BatchNorm layers of one SurfaceAbstractionCD block must all be in "
                           "the same mode (all training or all eval)")
    frozen = not modes.pop()
    params = [mod.mlp_l0.weight, mod.mlp_l0.bias, mod.mlp_f0.weight, mod.mlp_f0.bias,
              mod.bn_l0.weight, mod.bn_l0.bias, mod.bn_f0.weight, mod.bn_f0.bias]
    for lin, bn in zip(mod.mlp_convs, mod.mlp_bns):
        params += [lin.weight, lin.bias, bn.weight, bn.bias]
    gather = None
    if hasattr(rows, "table"):              # mlp.GatheredRows: the row matrix exists only as (table, neighbour index, centres)
        gather, layout, n_rows = (rows.idx, rows.centres), rows.layout, rows.idx.numel()
        rows = rows.table
    else:
        n_rows = rows.shape[0]
    feat_col, feat_channels = layout if layout is not None else (pos_channel, rows.shape[1] - pos_channel)
    meta = dict(ns=nsample, pos_channel=pos_channel, feat_col=feat_col, feat_channels=feat_channels, gather=gather,
                eps=mod.bn_l0.eps, n_extra=len(mod.mlp_convs), frozen=frozen, bns=bns,
                W_l_shape=tuple(mod.mlp_l0.weight.shape), W_f_shape=tuple(mod.mlp_f0.weight.shape),
                W_shapes=[tuple(l.weight.shape) for l in mod.mlp_convs])
    res = _FusedSAMLP.apply(rows if rows.stride(1) == 1 else rows.contiguous(), meta, *params)
    out, stats = res[0], res[1:]
    if frozen:
        return out
    # running statistics (same side effects as the BatchNorm modules): batch mean, UNBIASED batch variance
    R = n_rows
    with torch.no_grad():
        C0 = mod.mlp_l0.weight.shape[0]
        st0 = stats[0]                                       # [sum l | sum f | sumsq l | sumsq f]
        _update_running(mod.bn_l0, R, st0[:C0], st0[2 * C0:3 * C0])
        _update_running(mod.bn_f0, R, st0[C0:2 * C0], st0[3 * C0:])
        for bn, sti in zip(mod.mlp_bns, stats[1:]):
            C = sti.shape[0] // 2
            _update_running(bn, R, sti[:C], sti[C:])
    return out


# ------------------------------------------------------------------------------------------------------------
# single layers on the tensor cores: Linear (+ train-mode BatchNorm (+ ReLU)) over rows, hand-written backward.
# Used by SurfaceFeaturePropagationCD, the umbrella MLP and the segmentation head.
# ------------------------------------------------------------------------------------------------------------
class _LinearBN(Function):
    """out = [relu](bn(X W^T + b)) over rows.  `frozen` = (sc, sh, mu, inv) of a running-statistics BatchNorm, or None for
    batch statistics (returned as the second output for the running-buffer update)."""

    @staticmethod
    def forward(ctx, X, W, bias, gamma, beta, relu, eps, frozen):
        dev = X.device
        X = X.contiguous()
        R, K = X.shape
        Nn = W.shape[0]
        W2 = W.reshape(Nn, -1).detach().contiguous()
        Wp, _, _ = prep_weight(W2)
        Y = torch.empty(R, Nn, device=dev)
        st = None if frozen is not None else torch.zeros(2 * Nn, dtype=torch.float64, device=dev)
        gemm_rows(R, Nn, opnd(OPND_RAW, X, K), Wp, Y=Y, bias=None if bias is None else bias.detach(), stats=st)
        if frozen is not None:
            sc, sh, mu, inv = frozen
        else:
            sc, sh, mu, inv = _bn_finalize(Nn, R, st, gamma.detach(), beta.detach(), eps, 0.0, None, None, dev)
        out = torch.empty(R, Nn, device=dev)
        N.call("rsb_bn_apply", R, Nn, Y, Nn, sc, sh, 1 if relu else 0, out, Nn)
        ctx.relu = relu
        ctx.w_shape = tuple(W.shape)
        ctx.has_bias = bias is not None
        ctx.frozen = frozen is not None
        ctx.saved = (X, W2, Y, sc, sh, mu, inv)
        if st is None:
            st = torch.empty(0, dtype=torch.float64, device=dev)
        ctx.mark_non_differentiable(st)
        return out, st

    @staticmethod
    @once_differentiable
    def backward(ctx, dOut, _st):
        X, W2, Y, sc, sh, mu, inv = ctx.saved
        dev = X.device
        R, K = X.shape
        Nn = W2.shape[0]
        dOut = dOut.contiguous()
        dZ = torch.empty_like(dOut)          # masked gradient written out of place (no clone of the incoming gradient)
        st = torch.zeros(2 * Nn, dtype=torch.float64, device=dev)
        if ctx.relu:
            msc, msh = sc, sh
        else:   # no activation: mask always on (z = 0*y + 1 > 0), statistics still needed
            msc, msh = torch.zeros(Nn, device=dev), torch.ones(Nn, device=dev)
        N.call("rsb_bn_relu_backward", R, Nn, dOut, Nn, dZ, Nn, Y, Nn, msc, msh, mu, inv, 0, st)
        co = torch.empty(5, Nn, device=dev)
        N.call("rsb_bn_backward_coef", Nn, R, st, 2 if ctx.frozen else 0, sc, mu, inv, co[0], co[1], co[2], co[3], co[4])
        G = opnd(OPND_AFFINE2, dZ, Nn, a=co[0], b=co[1], d=co[2], V=Y, ku=Nn)
        dW = torch.zeros(Nn, K, device=dev)
        gemm_wgrad(R, G, opnd(OPND_RAW, X, K), dW)
        dX = None
        if ctx.needs_input_grad[0]:
            WpT, _, _ = prep_weight(W2, transposed=True)
            dX = torch.empty(R, K, device=dev)
            gemm_rows(R, K, G, WpT, Y=dX)
        db = None
        if ctx.has_bias:   # a bias in front of a batch-statistics BatchNorm has zero gradient; frozen: sc * sum dZ
            db = sc * co[4] if ctx.frozen else torch.zeros(Nn, device=dev)
        return dX, dW.reshape(ctx.w_shape), db, co[3], co[4], None, None, None


def _pad4(n):
    return (n + 3) // 4 * 4


class _Linear(Function):
    """Y = X W^T + b.  Output and incoming gradient live in row-padded buffers (pitch a multiple of 4 floats) so that
    narrow layers (13 classes, 10 umbrella channels) stay on the TMA-fed kernels; the caller sees a [R, N] view."""

    @staticmethod
    def forward(ctx, X, W, bias):
        X = X.contiguous()
        R, K = X.shape
        Nn = W.shape[0]
        W2 = W.detach().contiguous()
        Wp, _, _ = prep_weight(W2)
        Ybuf = torch.empty(R, _pad4(Nn), device=X.device)
        gemm_rows(R, Nn, opnd(OPND_RAW, X, K), Wp, Y=Ybuf, bias=None if bias is None else bias.detach())
        ctx.saved = (X, W2)
        ctx.has_bias = bias is not None
        return Ybuf[:, :Nn]

    @staticmethod
    @once_differentiable
    def backward(ctx, dY):
        X, W2 = ctx.saved
        R, K = X.shape
        Nn = W2.shape[0]
        if dY.stride(1) != 1 or dY.stride(0) % 4 or dY.data_ptr() % 16:
            buf = torch.empty(R, _pad4(Nn), device=X.device)
            buf[:, :Nn].copy_(dY)
            dY = buf[:, :Nn]
        dW = torch.zeros(Nn, K, device=X.device)
        gemm_wgrad(R, opnd(OPND_RAW, dY, Nn), opnd(OPND_RAW, X, K), dW)
        dX = None
        if ctx.needs_input_grad[0]:
            WpT, _, _ = prep_weight(W2, transposed=True)
            dX = torch.empty(R, K, device=X.device)
            gemm_rows(R, K, opnd(OPND_RAW, dY, Nn), WpT, Y=dX)
        return dX, dW, (dY.sum(0) if ctx.has_bias else None)


def _pad_vec(v, n, value=0.0):
    return v if v is None or v.shape[0] == n else torch.nn.functional.pad(v, (0, n - v.shape[0]), value=value)


def linear_bn(x, lin, bn, relu):
    """[relu](bn(lin(x))) over rows [R, K'] on the tensor cores, K' >= in_features (extra input columns must be zero).
    Channel counts that are not multiples of 4 (the 10-channel umbrella layers) run zero-padded to the next multiple:
    the result then has pad4(out_features) columns, the padding columns exactly zero.  Train-mode BatchNorm uses batch
    statistics and updates the running buffers; eval-mode (or frozen) BatchNorm uses the running statistics."""
    if not x.is_cuda:
        raise RuntimeError("repsurf_b200 has no CPU path")
    W = lin.weight.reshape(lin.weight.shape[0], -1)
    C, K = W.shape
    Cp, Kp = _pad4(C), x.shape[1]
    assert Kp >= K and Kp % 4 == 0, "input rows must be padded to a multiple of 4 channels"
    if Cp != C or Kp != K:
        W = torch.nn.functional.pad(W, (0, Kp - K, 0, Cp - C))
    bias = _pad_vec(lin.bias, Cp)
    gamma = _pad_vec(bn.weight if bn.weight is not None else torch.ones(C, device=x.device), Cp, 1.0)
    beta = _pad_vec(bn.bias if bn.bias is not None else torch.zeros(C, device=x.device), Cp)
    batch = bn_uses_batch_stats(bn)
    frozen = None if batch else _bn_eval_coef([bn], pad_to=Cp)
    out, st = _LinearBN.apply(x, W, bias, gamma, beta, relu, bn.eps, frozen)
    if batch and bn.training:
        with torch.no_grad():
            _update_running(bn, x.shape[0], st[:C], st[Cp:Cp + C])
    return out


def linear(x, lin):
    """lin(x) over rows [R, K'], K' >= in_features with zero extra columns; -> [R, out_features] view."""
    if not x.is_cuda:
        raise RuntimeError("repsurf_b200 has no CPU path")
    W = lin.weight.reshape(lin.weight.shape[0], -1)
    if x.shape[1] != W.shape[1]:
        W = torch.nn.functional.pad(W, (0, x.shape[1] - W.shape[1]))
    return _Linear.apply(x, W, lin.bias)


class _UmbrellaMLP(Function):
    """Conv1d(10,10) + BN + ReLU + Conv1d(10,10) + sum over the g triangles, csrc/umbrella_mlp.cu (recomputing: the
    only tensors in HBM are the descriptor rows and the [n, 10] result)."""

    @staticmethod
    def forward(ctx, feat, W1, b1, gamma, beta, W2, b2, coef, train):
        n, g, cin = feat.shape
        C = W1.shape[0]
        W1c, W2c = W1.reshape(C, cin).contiguous(), W2.reshape(C, C).contiguous()
        out = torch.empty(n, C, device=feat.device)
        sc, sh, mu, inv = coef
        N.call("rsb_umbrella_mlp_forward", n * g, g, cin, C, feat, W1c, b1, W2c, b2, sc, sh, out)
        ctx.save_for_backward(feat, W1c, b1, W2c, sc, sh, mu, inv)
        ctx.g, ctx.train = g, train
        return out

    @staticmethod
    def backward(ctx, dOut):
        feat, W1c, b1, W2c, sc, sh, mu, inv = ctx.saved_tensors
        n, g, cin = feat.shape
        C = W1c.shape[0]
        acc = torch.zeros(C * C + 3 * C + C * cin + C, dtype=torch.float64, device=feat.device)
        acc1, acc2 = acc[:C * C + 3 * C], acc[C * C + 3 * C:]
        N.call("rsb_umbrella_mlp_backward", n * g, g, cin, C, 1 if ctx.train else 0, feat, dOut.contiguous(), W1c, b1, W2c, sc, sh, mu, inv, acc1, acc2)
        a = acc.float()
        dW2, db2 = a[:C * C].view(C, C, 1), a[C * C:C * C + C]
        dbeta, dgamma = a[C * C + C:C * C + 2 * C], a[C * C + 2 * C:C * C + 3 * C]
        o = C * C + 3 * C
        dW1, db1 = a[o:o + C * cin].view(C, cin, 1), a[o + C * cin:]
        return None, dW1, db1, dgamma, dbeta, dW2, db2, None, None


def umbrella_mlp_fused(feat, conv1, bn1, conv2):
    """feat [n, g, 10] (no gradient) -> [n, 10]; the segmentation umbrella MLP and its 'sum' aggregation
    (segmentation/modules/repsurface_utils.py:297-302, :322-327).  Train mode uses batch statistics and updates the
    running buffers exactly like nn.BatchNorm1d; eval mode uses the running statistics."""
    n, g, cin = feat.shape
    C = conv1.weight.shape[0]
    feat = feat.contiguous()
    batch_stats = bn1.training or bn1.running_mean is None
    if batch_stats:
        st = torch.zeros(2 * C, dtype=torch.float64, device=feat.device)
        N.call("rsb_umbrella_mlp_stats", n * g, cin, C, feat, conv1.weight.detach().reshape(C, cin).contiguous(),
               conv1.bias.detach(), st)
        track = bn1.training and bn1.track_running_stats and bn1.running_mean is not None
        coef = _bn_finalize(C, n * g, st, bn1.weight.detach(), bn1.bias.detach(), bn1.eps,
                            (bn1.momentum if bn1.momentum is not None else 0.1) if track else 0.0,
                            bn1.running_mean if track else None, bn1.running_var if track else None, feat.device)
        if track:
            bn1.num_batches_tracked.add_(1)
    else:
        coef = _bn_eval_coef([bn1])
    return _UmbrellaMLP.apply(feat, conv1.weight, conv1.bias, bn1.weight, bn1.bias, conv2.weight, conv2.bias, coef,
                              batch_stats)
