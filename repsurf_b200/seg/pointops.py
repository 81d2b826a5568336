"""Packed-layout operator API — same names and argument meaning as the reference's
segmentation/modules/pointops/functions/pointops.py (cited per op), backed by librepsurf_b200.so.

Packed layout: coord [N,3], feat [N,C], offset [B] int32 cumulative ends (device tensor).

Host synchronisation.  The reference reads `offset` element by element on the host
(`offset[i].item()`, tensor `max()` in Python loops: pointops.py:39-42, 61-93).  Here the host
copy of an offset tensor is taken ONCE (`host_offsets`) and remembered for tensors created by
this package (`register_offsets`), so a forward pass performs no device->host round trips after
the first look at the input offsets.
"""
import weakref

import torch
from torch.autograd import Function

from .. import _native as N

# kNN: clouds with at least this many points (on average) go through the grid search; None = always all-pairs
KNN_GRID_MIN_POINTS = 2048

# largest segment (in priority positions) the FPS kernel keeps in registers: 16 CTAs x 512 threads x 16 points; beyond
# it the kernel streams the running minima through a scratch buffer the caller provides
_FPS_REGISTER_CAPACITY = 16 * 512 * 16

# ---------------------------------------------------------------------------------------------
# host copies of offset tensors
# ---------------------------------------------------------------------------------------------
_HOST = {}


def register_offsets(t, values):
    """Remember the host values of an offset tensor (keyed by identity AND in-place version: a later copy_() into the same
    buffer invalidates the mirror, so static input buffers of a training loop are re-read, not trusted)."""
    key = id(t)
    _HOST[key] = (weakref.ref(t, lambda _r, k=key: _HOST.pop(k, None)), tuple(int(v) for v in values), t._version)
    return t


def host_offsets(t):
    ent = _HOST.get(id(t))
    if ent is not None and ent[0]() is t and ent[2] == t._version:
        return ent[1]
    vals = tuple(t.tolist())  # one D2H sync, only for offsets that came from outside (or were modified in place)
    register_offsets(t, vals)
    return vals


_CONST = {}


def const_tensor(values, dtype, device):
    """Device copy of a small host list, cached by value and never written to.  torch.tensor(..., device=cuda)
    synchronises the stream on every call (pageable host-to-device copy); a training loop over a fixed batch layout
    would drain the launch queue a dozen times per step.  First use stages through pinned memory, asynchronously."""
    key = (tuple(values), dtype, str(device))
    t = _CONST.get(key)
    if t is None:
        if len(_CONST) > 1024:
            _CONST.clear()
        t = torch.tensor(list(values), dtype=dtype)
        if torch.device(device).type == "cuda":
            t = t.pin_memory().to(device, non_blocking=True)
            # the constant is shared by every later call on ANY stream: finish the upload once, here (first use only)
            torch.cuda.current_stream(t.device).synchronize()
        _CONST[key] = t
    return t


def make_offsets(values, device):
    """Offset tensor for module outputs: a cached device constant (shared between calls with the same values; callers must
    not write to it - an in-place edit bumps its version and is caught by host_offsets, but would still corrupt the other
    holders of the constant)."""
    return register_offsets(const_tensor(values, torch.int32, device), values)


def _sizes(off):
    return [b - a for a, b in zip((0,) + tuple(off[:-1]), off)]


# ---------------------------------------------------------------------------------------------
class FurthestSampling(Function):
    """ref: pointops.py:31-49.  xyz (n,3), offset (b), new_offset (b) -> idx (m) int32, GLOBAL row ids."""

    @staticmethod
    def forward(ctx, xyz, offset, new_offset):
        assert xyz.is_contiguous()
        off, noff = host_offsets(offset), host_offsets(new_offset)
        b = len(off)
        idx = torch.empty(noff[-1], dtype=torch.int32, device=xyz.device)
        n_max = max(_sizes(off))
        tmp = torch.empty(xyz.shape[0], device=xyz.device) if n_max + 1024 > _FPS_REGISTER_CAPACITY else None
        N.call("rsb_furthestsampling_packed", b, n_max, None, xyz, offset, new_offset, tmp, idx, None)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None


furthestsampling = FurthestSampling.apply


class SectorizedFurthestSampling(Function):
    """ref: pointops.py:52-111.  Azimuth-sectorized FPS; returns int64 global row ids, sector-major per cloud.

    The reference does the sector split in host Python (per cloud and per sector `torch.where`, `.item()`); here it is
    csrc/sector.cu (angles + per-cloud range, classification against the fp32 linspace edges, stable counting sort)
    followed by ONE packed FPS launch over all (cloud, sector) segments and a map-back kernel, with no host
    synchronisation: the sector sizes stay on the device and the FPS kernel derives the reference's tie rule from the
    device-side maximum (rsb_furthestsampling_packed_bounded)."""

    @staticmethod
    def forward(ctx, xyz, offset, new_offset, num_sectors, min_points=10000):
        assert xyz.is_contiguous()
        dev = xyz.device
        off, noff = host_offsets(offset), host_offsets(new_offset)
        sizes, new_sizes = _sizes(off), _sizes(noff)
        b = len(off)
        n = xyz.shape[0]
        nsec = [1 if s < min_points else num_sectors for s in sizes]          # host, from host offsets
        quotas, seg_first = [], []
        for i in range(b):
            q = [new_sizes[i] // nsec[i]] * nsec[i]
            q[-1] += new_sizes[i] % nsec[i]
            seg_first.append(len(quotas))
            quotas += q
        nseg = len(quotas)
        acc_q, run = [], 0
        for q in quotas:
            run += q
            acc_q.append(run)
        new_sector_offset = const_tensor(acc_q, torch.int32, dev)
        nbytes = int(N.lib().rsb_sector_split_workspace_bytes(n, b, nseg))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        order = torch.empty(n, dtype=torch.int32, device=dev)
        sector_xyz = torch.empty(n, 3, dtype=torch.float32, device=dev)
        sector_offset = torch.empty(nseg, dtype=torch.int32, device=dev)
        count_max = torch.empty(1, dtype=torch.int32, device=dev)
        N.call("rsb_sector_split", b, n, int(num_sectors), nseg, xyz, offset, const_tensor(nsec, torch.int32, dev),
               const_tensor(seg_first, torch.int32, dev), ws, nbytes, order, sector_xyz, sector_offset, count_max)
        idx = torch.empty(noff[-1], dtype=torch.int32, device=dev)
        # The sector sizes exist on the device only.  No read-back: sectors of up to 1.1x the mean size take a launch
        # planned for that size, anything larger a second launch planned for the whole cloud (it exits at once when the
        # sectors are balanced) - rsb_furthestsampling_packed_bounded.  The reference's tie rule depends on the largest
        # sector only through min(2^floor(log2 n_max), 1024): known on the host whenever some sector must hold >= 1024
        # points (pigeonhole), else taken from the device-side maximum.
        mean_sector = max(-(-sz // k) for sz, k in zip(sizes, nsec))
        n_expect = min(max(sizes), mean_sector + mean_sector // 10 + 1)
        n_max_dev = None if mean_sector >= 1024 else count_max
        tmp = torch.empty(n, device=dev) if max(sizes) + 1024 > _FPS_REGISTER_CAPACITY else None
        N.call("rsb_furthestsampling_packed_bounded", nseg, n_expect, max(sizes), n_max_dev, sector_xyz, sector_offset,
               new_sector_offset, tmp, idx, None)
        out = torch.empty(noff[-1], dtype=torch.int64, device=dev)
        N.call("rsb_sector_map_back", noff[-1], order, idx, out)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None, None


sectorized_fps = SectorizedFurthestSampling.apply


class KNNQuery(Function):
    """ref: pointops.py:114-130.  -> (idx (m,nsample) int32 global ids, sqrt(dist2) (m,nsample))."""

    @staticmethod
    def forward(ctx, nsample, xyz, new_xyz, offset, new_offset):
        if new_xyz is None:
            new_xyz = xyz
        assert xyz.is_contiguous() and new_xyz.is_contiguous()
        m = new_xyz.shape[0]
        n = xyz.shape[0]
        b = offset.shape[0]
        idx = torch.empty(m, nsample, dtype=torch.int32, device=xyz.device)
        dist = torch.empty(m, nsample, dtype=torch.float32, device=xyz.device)
        if KNN_GRID_MIN_POINTS is not None and n >= KNN_GRID_MIN_POINTS * b:
            # same indices / distances as the all-pairs kernel, through a uniform grid (csrc/knn_grid.cu)
            nbytes = int(N.lib().rsb_knn_grid_workspace_bytes(n, b))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=xyz.device)
            N.call("rsb_knnquery_grid", 1, 1, b, 0, 0, n, m, int(nsample), xyz, new_xyz, offset, new_offset, idx, dist, 1,
                   ws, nbytes)
        else:
            N.call("rsb_knnquery_packed", b, m, int(nsample), xyz, new_xyz, offset, new_offset, idx, dist, 1)
        ctx.mark_non_differentiable(idx, dist)
        return idx, dist

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None, None, None


knnquery = KNNQuery.apply


class Grouping(Function):
    """ref: pointops.py:133-162.  input (n,c), idx (m,nsample) -> (m,nsample,c)."""

    @staticmethod
    def forward(ctx, input, idx):
        assert input.is_contiguous() and idx.is_contiguous()
        m, nsample, n, c = idx.shape[0], idx.shape[1], input.shape[0], input.shape[1]
        out = torch.empty(m, nsample, c, dtype=torch.float32, device=input.device)
        N.call("rsb_grouping_packed_forward", m, nsample, c, input, idx, out)
        ctx.n = n
        ctx.save_for_backward(idx)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        idx, = ctx.saved_tensors
        m, nsample, c = grad_output.shape
        grad = torch.zeros(ctx.n, c, dtype=torch.float32, device=grad_output.device)
        N.call("rsb_grouping_packed_backward", m, nsample, c, grad_output.contiguous(), idx, grad)
        return grad, None


grouping = Grouping.apply


def queryandgroup(nsample, xyz, new_xyz, feat, idx, offset, new_offset, use_xyz=True):
    """ref: pointops.py:165-186.  -> (m, nsample, 3+c) (or (m,nsample,c))."""
    assert xyz.is_contiguous() and feat.is_contiguous()
    if new_xyz is None:
        new_xyz = xyz
    if idx is None:
        idx, _ = knnquery(nsample, xyz, new_xyz, offset, new_offset)
    grouped_xyz = grouping(xyz, idx) - new_xyz.unsqueeze(1)
    grouped_feat = grouping(feat, idx)
    return torch.cat((grouped_xyz, grouped_feat), -1) if use_xyz else grouped_feat


class Subtraction(Function):
    """ref: pointops.py:189-218.  input1 (n,c), input2 (n,c), idx (n,nsample) -> (n,nsample,c) = input1[n] - input2[idx]."""

    @staticmethod
    def forward(ctx, input1, input2, idx):
        assert input1.is_contiguous() and input2.is_contiguous()
        n, c = input1.shape
        nsample = idx.shape[-1]
        output = torch.empty(n, nsample, c, dtype=torch.float32, device=input1.device)
        N.call("rsb_subtraction_forward", n, nsample, c, input1, input2, idx.contiguous(), output)
        ctx.save_for_backward(idx)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        idx, = ctx.saved_tensors
        n, nsample, c = grad_output.shape
        grad_input1 = torch.empty(n, c, dtype=torch.float32, device=grad_output.device)
        grad_input2 = torch.zeros(n, c, dtype=torch.float32, device=grad_output.device)
        N.call("rsb_subtraction_backward", n, nsample, c, idx.contiguous(), grad_output.contiguous(), grad_input1, grad_input2)
        return grad_input1, grad_input2, None


subtraction = Subtraction.apply


class Aggregation(Function):
    """ref: pointops.py:221-253.  input (n,c), position (n,nsample,c), weight (n,nsample,c'), idx (n,nsample) -> (n,c):
    sum over the samples of (input[idx] + position) * weight[..., c % c']."""

    @staticmethod
    def forward(ctx, input, position, weight, idx):
        assert input.is_contiguous() and position.is_contiguous() and weight.is_contiguous()
        n, nsample, c = position.shape
        w_c = weight.shape[-1]
        output = torch.empty(n, c, dtype=torch.float32, device=input.device)
        N.call("rsb_aggregation_forward", n, nsample, c, w_c, input, position, weight, idx.contiguous(), output)
        ctx.save_for_backward(input, position, weight, idx)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input, position, weight, idx = ctx.saved_tensors
        n, nsample, c = position.shape
        w_c = weight.shape[-1]
        dev = grad_output.device
        grad_input = torch.zeros(n, c, dtype=torch.float32, device=dev)
        grad_position = torch.empty(n, nsample, c, dtype=torch.float32, device=dev)
        grad_weight = torch.empty(n, nsample, w_c, dtype=torch.float32, device=dev)
        N.call("rsb_aggregation_backward", n, nsample, c, w_c, input, position, weight, idx.contiguous(), grad_output.contiguous(),
               grad_input, grad_position, grad_weight)
        return grad_input, grad_position, grad_weight, None


aggregation = Aggregation.apply


def _idw(dist):
    """inverse-distance weights of pointops.py:262-265 / 283-285 (un-squared distance, eps 1e-8)."""
    r = 1.0 / (dist + 1e-8)
    return r / torch.sum(r, dim=1, keepdim=True)


class _InterpApply(Function):
    @staticmethod
    def forward(ctx, input, idx, weight):
        n, k = idx.shape
        m, c = input.shape
        out = torch.zeros(n, c, dtype=torch.float32, device=input.device)
        N.call("rsb_interpolation_packed_forward", n, c, k, input.contiguous(), idx, weight, out)
        ctx.m = m
        ctx.save_for_backward(idx, weight)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        idx, weight = ctx.saved_tensors
        n, c = grad_output.shape
        grad = torch.zeros(ctx.m, c, dtype=torch.float32, device=grad_output.device)
        N.call("rsb_interpolation_packed_backward", n, c, idx.shape[1], grad_output.contiguous(), idx, weight, grad)
        return grad, None, None


def interpolation(xyz, new_xyz, feat, offset, new_offset, k=3):
    """ref: pointops.py:256-270.  xyz (m,3), new_xyz (n,3), feat (m,c) -> (n,c)."""
    assert xyz.is_contiguous() and new_xyz.is_contiguous() and feat.is_contiguous()
    idx, dist = knnquery(k, xyz, new_xyz, offset, new_offset)
    return _InterpApply.apply(feat, idx, _idw(dist).contiguous())


def interpolation2(xyz, new_xyz, input, offset, new_offset, k=3):
    """ref: pointops.py:273-307 (the autograd.Function variant of `interpolation`: k nearest, inverse-distance weights,
    gradient to `input` only)."""
    assert xyz.is_contiguous() and new_xyz.is_contiguous() and input.is_contiguous()
    idx, dist = knnquery(k, xyz, new_xyz, offset, new_offset)
    return _InterpApply.apply(input, idx, _idw(dist).contiguous())
