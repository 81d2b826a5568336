"""Shared-MLP evaluation on ROW MATRICES [rows, C] (rows = every (centre, sample) pair of a level).

The reference evaluates the per-group MLP as 1x1 Conv2d/Conv1d + BatchNorm2d/1d over [B,C,ns,m] / [M,C,ns]
(classification/modules/repsurface_utils.py:233-244, segmentation/modules/repsurface_utils.py:217-228), which
sends cuDNN down its worst paths on a B200 (grouped-direct wgrad, 1C11 batch-norm: 77 % of a step, see
profiles/r01_torch_profile_seg_v1.txt).  A 1x1 convolution over that layout IS a GEMM over rows, and
BatchNorm over (B, ns, m) per channel IS BatchNorm over rows, so everything here works on [rows, C] with the
parameters of the reference's layer objects (weights [out,in,1(,1)] viewed as [out,in]).

This module is the seam where the fused sm_100a kernels plug in (gather -> GEMM -> BN statistics epilogue).
"""
import torch
import torch.nn.functional as F


def linear_rows(x, layer):
    w = layer.weight
    return F.linear(x, w.view(w.shape[0], -1), layer.bias)


def bn_rows(x, bn):
    """Train/eval BatchNorm over rows with the module's buffers (same side effects as calling the module)."""
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    use_batch = bn.training or bn.running_mean is None
    return F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, use_batch,
                        0.0 if bn.momentum is None else bn.momentum, bn.eps)


def pack_rows(pos, feats):
    """Row matrix of a grouped level with 16-byte aligned column blocks: [pos | 0-pad to 4 | feats... | 0-pad to 4].
    pos [R, P]; feats: list of [R, c_i].  Returns (rows [R, P4 + F4], P4, F).  Aligned blocks let the tensor-core
    kernels use 128-bit asynchronous loads / stores for the operand and for the feature gradient."""
    R, P = pos.shape
    F_ = sum(f.shape[1] for f in feats)
    P4, F4 = (P + 3) // 4 * 4, (F_ + 3) // 4 * 4
    parts = [pos]
    if P4 > P:
        parts.append(pos.new_zeros(R, P4 - P))
    parts += feats
    if F4 > F_:
        parts.append(pos.new_zeros(R, F4 - F_))
    return torch.cat(parts, dim=-1), P4, F_


class _GroupRows(torch.autograd.Function):
    """Fused builder of the packed row matrix of a grouped level (csrc/group.cu group_rows_*)."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, idx, normal, feature, ns, polar):
        from . import _native as N
        rows = idx.numel()
        P = 6 if polar else 3
        P4 = (P + 3) // 4 * 4
        Cn = normal.shape[1]
        Cf = feature.shape[1] if feature is not None else 0
        ld = P4 + (Cn + Cf + 3) // 4 * 4
        out = torch.empty(rows, ld, device=xyz.device)
        normal = normal.contiguous()
        feature = feature.contiguous() if feature is not None else None
        N.call("rsb_group_rows_forward", rows, ns, 1 if polar else 0, P4, Cn, Cf, ld, xyz.contiguous(), new_xyz.contiguous(),
               idx, normal, feature, out)
        ctx.save_for_backward(idx)
        ctx.dims = (rows, P4, Cn, Cf, ld, normal.shape[0], feature is not None)
        return out

    @staticmethod
    def backward(ctx, drows):
        from . import _native as N
        idx, = ctx.saved_tensors
        rows, P4, Cn, Cf, ld, n, has_f = ctx.dims
        dn = torch.zeros(n, Cn, device=drows.device) if ctx.needs_input_grad[3] else None
        df = torch.zeros(n, Cf, device=drows.device) if (has_f and ctx.needs_input_grad[4]) else None
        if dn is not None or df is not None:
            N.call("rsb_group_rows_backward", rows, P4, Cn, Cf, ld, drows.contiguous(), idx, dn, df)
        return None, None, None, dn, df, None, None


def group_rows(xyz, new_xyz, idx, normal, feature, nsample, polar):
    """rows [M*ns, ld], layout (P4, Cn+Cf) — one kernel instead of three gathers + sub (+ polar) + cat + pad.
    xyz [n,3], new_xyz [M,3], idx [M,ns] GLOBAL row ids, normal [n,Cn], feature [n,Cf] | None."""
    rows = _GroupRows.apply(xyz, new_xyz, idx.reshape(-1).contiguous(), normal, feature, nsample, polar)
    P4 = 8 if polar else 4
    return rows, (P4, normal.shape[1] + (feature.shape[1] if feature is not None else 0))


def sa_mlp(rows, pos_channel, mod, nsample, layout=None):
    """Shared MLP + max-pool of a SurfaceAbstractionCD level.  Training mode runs the fused tcgen05 path
    (repsurf_b200.tc: 3xTF32 GEMMs with BatchNorm/ReLU/pool folded into operand loads and epilogues, hand-written
    backward); eval mode (running statistics, no autograd through BatchNorm statistics) uses the row-matrix
    composition below."""
    if mod.training and rows.is_cuda and len(mod.mlp_convs) >= 1:
        from . import tc
        return tc.sa_mlp_fused(rows, pos_channel, mod, nsample, layout)
    if layout is not None:   # strip the alignment padding for the plain composition
        P4, F_ = layout
        rows = torch.cat([rows[:, :pos_channel], rows[:, P4:P4 + F_]], dim=-1)
    return sa_mlp_rows(rows, pos_channel, mod, nsample)


def sa_mlp_rows(rows, pos_channel, mod, nsample):
    """Channel-de-differentiated shared MLP + max-pool.  rows [G*nsample, C] -> [G, mlp[-1]].
    mod provides mlp_l0/mlp_f0/bn_l0/bn_f0/mlp_convs/mlp_bns (the reference's attribute names)."""
    x = F.relu(bn_rows(linear_rows(rows[:, :pos_channel], mod.mlp_l0), mod.bn_l0)
               + bn_rows(linear_rows(rows[:, pos_channel:], mod.mlp_f0), mod.bn_f0))
    for lin, bn in zip(mod.mlp_convs, mod.mlp_bns):
        x = F.relu(bn_rows(linear_rows(x, lin), bn))
    return x.view(-1, nsample, x.shape[-1]).max(dim=1)[0]
