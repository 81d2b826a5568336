"""Row-matrix formulation of the per-group shared MLP (rows = every (centre, sample) pair of a level).

The reference evaluates the per-group MLP as 1x1 Conv2d/Conv1d + BatchNorm2d/1d over [B,C,ns,m] / [M,C,ns]
(classification/modules/repsurface_utils.py:233-244, segmentation/modules/repsurface_utils.py:217-228), which
sends cuDNN down its worst paths on a B200 (grouped-direct wgrad, 1C11 batch-norm: 77 % of a step, see
profiles/r01_torch_profile_seg_v1.txt).  A 1x1 convolution over that layout IS a GEMM over rows, and
BatchNorm over (B, ns, m) per channel IS BatchNorm over rows, so everything here works on [rows, C] with the
parameters of the reference's layer objects (weights [out,in,1(,1)] viewed as [out,in]).

This module is the seam between the module layer and the sm_100a kernels: the fused row builder (csrc/group.cu) and
the tcgen05 shared MLP (repsurf_b200.tc).  There is no torch / cuDNN / CPU evaluation path.
"""
import torch


class _GroupRows(torch.autograd.Function):
    """Fused builder of the packed row matrix of a grouped level (csrc/group.cu group_rows_*)."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, idx, normal, feature, ns, polar):
        from . import _native as N
        rows = idx.numel()
        P = 6 if polar else 3
        P4 = (P + 3) // 4 * 4
        Cn = normal.shape[1] if normal is not None else 0
        Cf = feature.shape[1] if feature is not None else 0
        ld = P4 + (Cn + Cf + 3) // 4 * 4
        out = torch.empty(rows, ld, device=xyz.device)
        normal = normal.contiguous() if normal is not None else None
        feature = feature.contiguous() if feature is not None else None
        N.call("rsb_group_rows_forward", rows, ns, 1 if polar else 0, P4, Cn, Cf, ld, xyz.contiguous(), new_xyz.contiguous(),
               idx, normal, feature, out)
        ctx.save_for_backward(idx)
        ctx.dims = (rows, P4, Cn, Cf, ld, xyz.shape[0], feature is not None)
        return out

    @staticmethod
    def backward(ctx, drows):
        from . import _native as N
        idx, = ctx.saved_tensors
        rows, P4, Cn, Cf, ld, n, has_f = ctx.dims
        dn = torch.zeros(n, Cn, device=drows.device) if (Cn and ctx.needs_input_grad[3]) else None
        df = torch.zeros(n, Cf, device=drows.device) if (has_f and ctx.needs_input_grad[4]) else None
        if dn is not None or df is not None:
            N.call("rsb_group_rows_backward", rows, P4, Cn, Cf, ld, drows.contiguous(), idx, dn, df)
        return None, None, None, dn, df, None, None


def group_rows(xyz, new_xyz, idx, normal, feature, nsample, polar):
    """rows [M*ns, ld], layout (P4, Cn+Cf) — one kernel instead of three gathers + sub (+ polar) + cat + pad.
    xyz [n,3], new_xyz [M,3], idx [M,ns] GLOBAL row ids, normal [n,Cn] | None, feature [n,Cf] | None."""
    rows = _GroupRows.apply(xyz, new_xyz, idx.reshape(-1).contiguous(), normal, feature, nsample, polar)
    P4 = 8 if polar else 4
    return rows, (P4, (normal.shape[1] if normal is not None else 0) + (feature.shape[1] if feature is not None else 0))


class _PointTable(torch.autograd.Function):
    """[xyz | 0 | normal | feature | 0-pad] per POINT (csrc/group.cu point_table_kernel): the table the first shared-MLP GEMM
    gathers its rows from.  Backward: the GEMM's scatter epilogue has already summed the row gradients per point."""

    @staticmethod
    def forward(ctx, xyz, normal, feature):
        from . import _native as N
        n = xyz.shape[0]
        Cn = normal.shape[1] if normal is not None else 0
        Cf = feature.shape[1] if feature is not None else 0
        ld = 4 + (Cn + Cf + 3) // 4 * 4
        out = torch.empty(n, ld, device=xyz.device)
        N.call("rsb_point_table", n, Cn, Cf, ld, xyz.contiguous(), None if normal is None else normal.contiguous(),
               None if feature is None else feature.contiguous(), out)
        ctx.dims = (Cn, Cf)
        return out

    @staticmethod
    def backward(ctx, dT):
        Cn, Cf = ctx.dims
        dn = dT[:, 4:4 + Cn].contiguous() if (Cn and ctx.needs_input_grad[1]) else None
        df = dT[:, 4 + Cn:4 + Cn + Cf].contiguous() if (Cf and ctx.needs_input_grad[2]) else None
        return None, dn, df


class GatheredRows:
    """The row matrix of a grouped level, NOT materialised: rows[r] = table[idx[r]] with the first three columns made
    relative to centres[r // nsample].  The first-layer GEMMs (forward, weight gradient) gather it tile by tile with TMA
    and the input-gradient GEMM scatters into the table's gradient (repsurf_b200.tc, RSB_OPND_GATHER)."""

    def __init__(self, table, idx, centres, nsample, layout):
        self.table, self.idx, self.centres, self.nsample, self.layout = table, idx, centres, nsample, layout


def gather_rows(xyz, new_xyz, idx, normal, feature, nsample):
    """Fused form of group_rows for levels without the polar columns: returns a GatheredRows (layout: position columns
    0..2 (+ zero pad), features from column 4)."""
    table = _PointTable.apply(xyz, normal, feature)
    F_ = (normal.shape[1] if normal is not None else 0) + (feature.shape[1] if feature is not None else 0)
    return GatheredRows(table, idx.reshape(-1).contiguous(), new_xyz.contiguous(), nsample, (4, F_))


def sa_mlp(rows, pos_channel, mod, nsample, layout=None):
    """Shared MLP + max-pool of a SurfaceAbstractionCD level: the fused tcgen05 path of repsurf_b200.tc (3xTF32 GEMMs with
    BatchNorm / ReLU / pool folded into operand loads and epilogues, hand-written backward) in training AND eval mode."""
    if not (rows.table if isinstance(rows, GatheredRows) else rows).is_cuda:
        raise RuntimeError("repsurf_b200 has no CPU path")
    if len(mod.mlp_convs) < 1:
        raise RuntimeError("the fused shared MLP needs at least two layers (mlp = [c0, c1, ...])")
    from . import tc
    return tc.sa_mlp_fused(rows, pos_channel, mod, nsample, layout)
