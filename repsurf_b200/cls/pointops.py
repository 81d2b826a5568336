"""Dense-layout operator API — same names, argument meaning and error behaviour as the reference's
classification/modules/pointops/functions/pointops.py (cited per op), backed by the sm_100a
kernels of librepsurf_b200.so through the C-ABI.  No CPU path: CPU tensors raise.

Differences that are deliberate and invisible to callers:
  * index outputs are written completely by the kernels (no `.zero_()` pre-fill pass),
  * everything runs on torch's CURRENT stream (the reference uses the legacy default stream for
    most ops), so the ops compose with CUDA graphs / side streams,
  * the FPS scratch buffer of the reference (`temp`) is not allocated: min-distances live in registers.
"""
from typing import Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from .. import _native as N


def _chk(t, name):
    assert t.is_contiguous(), f"{name} must be contiguous"  # same assert as the reference wrappers


class FurthestSampling(Function):
    """ref: pointops.py:35-54.  xyz (b,n,3) fp32, m -> idx (b,m) int32, idx[:,0]=0."""

    @staticmethod
    def forward(ctx, xyz, m):
        _chk(xyz, "xyz")
        b, n, _ = xyz.size()
        idx = torch.empty(b, m, dtype=torch.int32, device=xyz.device)
        N.call("rsb_furthestsampling_dense", b, n, m, xyz, None, idx, None)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthestsampling = FurthestSampling.apply


def furthestsampling_with_xyz(xyz, m):
    """Fused FPS + gather of the sampled coordinates: returns (idx (b,m) int32, new_xyz (b,m,3)).
    Coordinates never carry gradient on the RepSurf path, so this is forward-only."""
    _chk(xyz, "xyz")
    b, n, _ = xyz.size()
    idx = torch.empty(b, m, dtype=torch.int32, device=xyz.device)
    new_xyz = torch.empty(b, m, 3, dtype=torch.float32, device=xyz.device)
    N.call("rsb_furthestsampling_dense", b, n, m, xyz.detach(), None, idx, new_xyz)
    return idx, new_xyz


class Gathering(Function):
    """ref: pointops.py:57-83.  features (b,c,n), idx (b,m) -> (b,c,m)."""

    @staticmethod
    def forward(ctx, features, idx):
        _chk(features, "features")
        _chk(idx, "idx")
        b, c, n = features.size()
        m = idx.size(1)
        out = torch.empty(b, c, m, dtype=torch.float32, device=features.device)
        N.call("rsb_gathering_forward", b, c, n, m, features, idx, out)
        ctx.for_backwards = (idx, c, n)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, c, n = ctx.for_backwards
        b, m = idx.size()
        grad = torch.zeros(b, c, n, dtype=torch.float32, device=grad_out.device)
        N.call("rsb_gathering_backward", b, c, n, m, grad_out.contiguous(), idx, grad)
        return grad, None


gathering = Gathering.apply


class NearestNeighbor(Function):
    """ref: pointops.py:86-109.  unknown (b,n,3), known (b,m,3) -> (sqrt(dist2) (b,n,3), idx (b,n,3))."""

    @staticmethod
    def forward(ctx, unknown, known) -> Tuple[torch.Tensor, torch.Tensor]:
        _chk(unknown, "unknown")
        _chk(known, "known")
        b, n, _ = unknown.size()
        m = known.size(1)
        dist2 = torch.empty(b, n, 3, dtype=torch.float32, device=unknown.device)
        idx = torch.empty(b, n, 3, dtype=torch.int32, device=unknown.device)
        N.call("rsb_nearestneighbor", b, n, m, unknown, known, dist2, idx)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


nearestneighbor = NearestNeighbor.apply


class Interpolation(Function):
    """ref: pointops.py:112-147.  features (b,c,m), idx/weight (b,n,3) -> (b,c,n)."""

    @staticmethod
    def forward(ctx, features, idx, weight):
        features = features.contiguous()
        _chk(idx, "idx")
        _chk(weight, "weight")
        b, c, m = features.size()
        n = idx.size(1)
        ctx.interpolation_for_backward = (idx, weight, m)
        out = torch.empty(b, c, n, dtype=torch.float32, device=features.device)
        N.call("rsb_interpolation_forward", b, c, m, n, features, idx, weight, out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.interpolation_for_backward
        b, c, n = grad_out.size()
        grad = torch.zeros(b, c, m, dtype=torch.float32, device=grad_out.device)
        N.call("rsb_interpolation_backward", b, c, n, m, grad_out.contiguous(), idx, weight, grad)
        return grad, None, None


interpolation = Interpolation.apply


class Grouping(Function):
    """ref: pointops.py:150-180.  features (b,c,n), idx (b,m,nsample) -> (b,c,m,nsample)."""

    @staticmethod
    def forward(ctx, features, idx):
        _chk(features, "features")
        _chk(idx, "idx")
        b, c, n = features.size()
        _, m, nsample = idx.size()
        out = torch.empty(b, c, m, nsample, dtype=torch.float32, device=features.device)
        N.call("rsb_grouping_forward", b, c, n, m, nsample, features, idx, out)
        ctx.for_backwards = (idx, n)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, n = ctx.for_backwards
        b, c, m, nsample = grad_out.size()
        grad = torch.zeros(b, c, n, dtype=torch.float32, device=grad_out.device)
        N.call("rsb_grouping_backward", b, c, n, m, nsample, grad_out.contiguous(), idx, grad)
        return grad, None


grouping = Grouping.apply


class GroupingInt(Function):
    """ref: pointops.py:183-203.  int64 features (b,c,n) -> (b,c,m,nsample) int64."""

    @staticmethod
    def forward(ctx, features, idx):
        _chk(features, "features")
        _chk(idx, "idx")
        b, c, n = features.size()
        _, m, nsample = idx.size()
        out = torch.empty(b, c, m, nsample, dtype=torch.int64, device=features.device)
        N.call("rsb_grouping_int_forward", b, c, n, m, nsample, features, idx, out)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, a=None):
        return None, None


grouping_int = GroupingInt.apply


class BallQuery(Function):
    """ref: pointops.py:206-229.  -> idx (b,m,nsample) int32."""

    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        _chk(xyz, "xyz")
        _chk(new_xyz, "new_xyz")
        b, n, _ = xyz.size()
        m = new_xyz.size(1)
        idx = torch.empty(b, m, nsample, dtype=torch.int32, device=xyz.device)
        N.call("rsb_ballquery", b, n, m, float(radius), int(nsample), new_xyz, xyz, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ballquery = BallQuery.apply


def pairwise_distances(x, y=None):
    """ref: pointops.py:232-249 (pure torch helper kept for API completeness)."""
    x_norm = (x ** 2).sum(1).view(-1, 1)
    if y is None:
        y = x
    y_norm = (y ** 2).sum(1).view(1, -1)
    return torch.clamp(x_norm + y_norm - 2.0 * torch.mm(x, y.t()), min=0.0)


# clouds with at least this many points go through the exact uniform-grid search (None-like: set very large to disable)
KNN_GRID_MIN_POINTS = 4096


class KNNQuery(Function):
    """ref: pointops.py:294-323.  -> idx (b,m,nsample) int32 sorted by (d2, index)."""

    @staticmethod
    def forward(ctx, nsample, xyz, new_xyz=None):
        if new_xyz is None:
            new_xyz = xyz
        xyz = xyz.contiguous()
        new_xyz = new_xyz.contiguous()
        b, m, _ = new_xyz.size()
        n = xyz.size(1)
        idx = torch.empty(b, m, nsample, dtype=torch.int32, device=xyz.device)
        if n >= KNN_GRID_MIN_POINTS:
            # same indices through the exact uniform-grid search (csrc/knn_grid.cu) instead of the all-pairs scan
            nbytes = int(N.lib().rsb_knn_grid_workspace_bytes(b * n, b))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=xyz.device)
            N.call("rsb_knnquery_grid", 0, 0, b, n, m, b * n, b * m, int(nsample), xyz, new_xyz, None, None, idx, None, 0, ws, nbytes)
        else:
            N.call("rsb_knnquery_dense", b, n, m, int(nsample), xyz, new_xyz, idx, None)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None


knnquery = KNNQuery.apply
# the reference's sort-based variant (pointops.py:252-291) returns the same indices up to ties
knnquery_naive = KNNQuery.apply


class KNNQuery_Heap(Function):
    """ref: pointops.py:326-354.  heap-order semantics."""

    @staticmethod
    def forward(ctx, nsample, xyz, new_xyz=None):
        if new_xyz is None:
            new_xyz = xyz
        _chk(xyz, "xyz")
        _chk(new_xyz, "new_xyz")
        b, m, _ = new_xyz.size()
        n = xyz.size(1)
        idx = torch.empty(b, m, nsample, dtype=torch.int32, device=xyz.device)
        dist2 = torch.empty(b, m, nsample, dtype=torch.float32, device=xyz.device)
        if n >= KNN_GRID_MIN_POINTS:
            nbytes = int(N.lib().rsb_knn_grid_workspace_bytes(b * n, b))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=xyz.device)
            N.call("rsb_knnquery_grid", 0, 1, b, n, m, b * n, b * m, int(nsample), xyz, new_xyz, None, None, idx, dist2, 0, ws, nbytes)
        else:
            N.call("rsb_knnquery_heap_dense", b, n, m, int(nsample), xyz, new_xyz, idx, dist2)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None


knnquery_heap = KNNQuery_Heap.apply


class QueryAndGroup(nn.Module):
    """ref: pointops.py:357-408 (ball query or heap kNN, then grouping of xyz (+features))."""

    def __init__(self, radius=None, nsample=32, use_xyz=True, return_idx=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz, self.return_idx = radius, nsample, use_xyz, return_idx

    def forward(self, xyz, new_xyz=None, features=None, idx=None):
        if new_xyz is None:
            new_xyz = xyz
        if idx is None:
            if self.radius is not None:
                idx = ballquery(self.radius, self.nsample, xyz, new_xyz)
            else:
                idx = knnquery_heap(self.nsample, xyz, new_xyz)
        grouped_xyz = grouping(xyz.transpose(1, 2).contiguous(), idx)  # (b,3,m,ns)
        diff = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is not None:
            gf = grouping(features, idx)
            new_features = torch.cat([diff, gf], dim=1) if self.use_xyz else gf
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = diff
        if self.return_idx:
            return new_features, grouped_xyz, idx.long()
        return new_features, grouped_xyz
