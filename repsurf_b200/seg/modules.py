"""Packed-layout RepSurf-U modules with the reference's constructor / forward signatures and
state_dict keys (segmentation/modules/repsurface_utils.py:176-329), running on the sm_100a operator
layer (`repsurf_b200.seg.pointops`).  CUDA only.

Drop-in: `from repsurf_b200.seg.modules import UmbrellaSurfaceConstructor, SurfaceAbstractionCD,
SurfaceFeaturePropagationCD` replaces `from modules.repsurface_utils import ...` in
segmentation/models/repsurf/repsurf_umb_ssg.py:8.

Differences from the reference implementation (same results):
  * no per-cloud `.item()` round trips: offsets keep a host mirror (pointops.host_offsets),
    strided offsets are computed on the host once and uploaded asynchronously;
  * sectorized FPS runs fully on the device (pointops.sectorized_fps);
  * FPS emits the sampled coordinates; kNN emits sqrt distances; gathers use the grouping kernel.
"""
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointops as P
from .. import _native as N
from .. import tc
from ..mlp import gather_rows, group_rows, sa_mlp


def strided_offsets(offset, stride):
    """new_n_i = n_i // stride, cumulative (segmentation/modules/repsurface_utils.py:17-22), host arithmetic."""
    off = P.host_offsets(offset)
    acc, out, prev = 0, [], 0
    for o in off:
        acc += (o - prev) // stride
        out.append(acc)
        prev = o
    return P.make_offsets(out, offset.device)


# ---------------------------------------------------------------------------------------------------------------
# Geometry plan: sampling and neighbour search of ALL levels depend on the coordinates only (never on features), so a
# caller that knows the level structure can start them ahead, on side streams, while the main stream runs the umbrella
# constructor and the shared MLPs: the FPS launches occupy a few SMs for milliseconds (one CTA cluster per segment) and
# hide completely behind the GEMMs.  The modules pick the results up by the identity of their `center` tensor; without
# an active plan they compute everything themselves, in order, on the current stream (same results either way).
# ---------------------------------------------------------------------------------------------------------------
_ACTIVE_PLAN = None
# first shared-MLP layer reads its rows through a TMA gather instead of a materialised row matrix (levels without polar columns)
FUSE_GATHER = True
# GeometryPlan runs sampling / neighbour search on side streams; False = everything in order on the caller's stream (per-kernel
# timing without interference from concurrent kernels: bench.py's roofline pass)
USE_SIDE_STREAMS = True
# while the cluster FPS of the NEXT level holds SMs on the side stream, the forward GEMMs of a level launch one persistent CTA per
# FREE SM instead of per SM (rsb_tc_set_sm_budget: no second wave behind the CTAs that found no SM).  RSB_SA_SM_BUDGET=0: off
SM_BUDGET = os.environ.get("RSB_SA_SM_BUDGET", "1") != "0"
# ... only for FPS launches of at least this many CTAs: a handful of single-CTA segments is over before the level's GEMMs start
SM_BUDGET_MIN_CTAS = int(os.environ.get("RSB_SA_SM_BUDGET_MIN_CTAS", "16"))
_SIDE_STREAMS = {}


def _fps_ctas(sizes):
    """SMs the packed FPS launch over clouds of these sizes occupies: one cluster per cloud, cluster size by the largest cloud's
    padded position count (csrc/fps.cu fps_plan; every CTA of the kernel owns an SM)."""
    n_max = -(-max(sizes) // 1024) * 1024
    cs = 1 if n_max <= 8192 else 4 if n_max <= 16384 else 8 if n_max <= 65536 else 16
    return cs * len(sizes)


def _side_streams(device):
    key = str(device)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = (torch.cuda.Stream(device=device), torch.cuda.Stream(device=device))
    return _SIDE_STREAMS[key]


def reserve_allocator_headroom(device, nbytes=4 << 30):
    """Give torch's caching allocator one large free block on the main stream and on each side stream of the geometry plan.
    Blocks are cached PER STREAM, and which side-stream blocks are reusable at a given moment depends on event timing, so a
    training loop keeps hitting first-time cudaMalloc calls (device-wide stalls of tens of milliseconds for GB-sized blocks) for
    many steps after the shapes have stopped changing; a big cached block is split instead.  Optional; results are unaffected."""
    dev = torch.device(device)
    streams = (torch.cuda.current_stream(dev),) + _side_streams(dev)
    for st in streams:
        with torch.cuda.stream(st):
            x = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            small = [torch.empty(512 << 10, dtype=torch.uint8, device=dev) for _ in range(64)]     # the <1 MB pool is separate
            del x, small
    torch.cuda.synchronize(dev)


def _sample(stride, center, offset, num_sector, training):
    """FPS of one level -> (fps_idx int64, new_center, new_offset); ref: segmentation/modules/repsurface_utils.py:24-33."""
    new_offset = strided_offsets(offset, stride)
    if num_sector > 1 and training:
        fps_idx = P.sectorized_fps(center, offset, new_offset, num_sector)
    else:
        fps_idx = P.furthestsampling(center, offset, new_offset).long()
    return fps_idx, center[fps_idx, :], new_offset


class GeometryPlan:
    """levels: [(stride, nsample, num_sector)] of consecutive SurfaceAbstractionCD blocks starting at (center, offset);
    fp_k: neighbours of the feature-propagation interpolation between consecutive levels (None = not prefetched)."""

    def __init__(self, center, offset, levels, training, fp_k=3):
        dev = center.device
        main = torch.cuda.current_stream(dev)
        s_fps, s_knn = _side_streams(dev) if USE_SIDE_STREAMS else (main, main)
        start = torch.cuda.Event()
        start.record(main)
        self.sa, self.knn = {}, {}
        chain = []
        cur_c, cur_o = center, offset
        with torch.cuda.stream(s_fps):
            s_fps.wait_event(start)
            for stride, nsample, num_sector in levels:
                if stride > 1:
                    fps_idx, new_c, new_o = _sample(stride, cur_c, cur_o, num_sector, training)
                else:
                    fps_idx, new_c, new_o = None, cur_c, cur_o
                ev = torch.cuda.Event()
                ev.record(s_fps)
                chain.append(dict(center=cur_c, offset=cur_o, key=(stride, nsample, num_sector > 1 and training), fps_idx=fps_idx,
                                  new_center=new_c, new_offset=new_o, ev_fps=ev))
                cur_c, cur_o = new_c, new_o
        with torch.cuda.stream(s_knn):
            s_knn.wait_event(start)
            for lv in chain:
                s_knn.wait_event(lv["ev_fps"])
                lv["group_idx"], _ = P.knnquery(lv["key"][1], lv["center"], lv["new_center"], lv["offset"], lv["new_offset"])
                lv["ev"] = torch.cuda.Event()
                lv["ev"].record(s_knn)
            if fp_k is not None:
                for lv in reversed(chain):                      # coarse <- fine, in the order the decoder asks for them
                    if lv["fps_idx"] is None:
                        continue
                    idx, dist = P.knnquery(fp_k, lv["new_center"], lv["center"], lv["new_offset"], lv["offset"])
                    ev = torch.cuda.Event()
                    ev.record(s_knn)
                    self.knn[(fp_k, id(lv["new_center"]), id(lv["center"]))] = (idx, dist, ev, lv["new_center"], lv["center"])
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        for lv, nxt in zip(chain, chain[1:] + [None]):
            # the FPS that runs on the side stream while this level's shared MLP runs on the main one is the NEXT level's
            lv["sm_budget"] = 0
            if USE_SIDE_STREAMS and nxt is not None and nxt["fps_idx"] is not None and not nxt["key"][2]:
                ctas = _fps_ctas(P._sizes(P.host_offsets(nxt["offset"])))
                if ctas >= SM_BUDGET_MIN_CTAS:
                    lv["sm_budget"] = max(sms // 2, sms - ctas)
        for lv in chain:
            # side-stream allocations used by the main stream: tell the caching allocator
            for t in (lv["fps_idx"], lv["new_center"], lv["group_idx"]):
                if t is not None and t is not center:
                    t.record_stream(main)
            self.sa[id(lv["center"])] = lv
        for idx, dist, _ev, _a, _b in self.knn.values():
            idx.record_stream(main)
            dist.record_stream(main)

    def __enter__(self):
        global _ACTIVE_PLAN
        self._prev, _ACTIVE_PLAN = _ACTIVE_PLAN, self
        return self

    def __exit__(self, *exc):
        global _ACTIVE_PLAN
        _ACTIVE_PLAN = self._prev
        return False

    def level(self, center, stride, nsample, sectorized):
        lv = self.sa.get(id(center))
        if lv is None or lv["center"] is not center or lv["key"] != (stride, nsample, sectorized):
            return None
        torch.cuda.current_stream(center.device).wait_event(lv["ev"])
        return lv

    def sm_budget(self, center):
        """SMs left to the shared MLP of the level that starts at `center` (0 = all)."""
        lv = self.sa.get(id(center))
        return lv["sm_budget"] if (lv is not None and lv["center"] is center) else 0

    def neighbours(self, k, xyz, new_xyz):
        ent = self.knn.get((k, id(xyz), id(new_xyz)))
        if ent is None or ent[3] is not xyz or ent[4] is not new_xyz:
            return None
        torch.cuda.current_stream(xyz.device).wait_event(ent[2])
        return ent[0], ent[1]


def _sample_and_group(stride, nsample, center, normal, feature, offset, return_polar, num_sector, training):
    """ref: segmentation/modules/repsurface_utils.py:15-51.
    -> new_center [M,3], new_normal [M,Cn], rows [M*ns, C4] ([rel xyz, polar? | pad | normal, feature? | pad], csrc/group.cu
       group_rows_fwd), (first feature column, feature channels), new_offset."""
    lv = _ACTIVE_PLAN.level(center, stride, nsample, num_sector > 1 and training) if _ACTIVE_PLAN is not None else None
    if lv is not None:
        fps_idx, new_center, new_offset, group_idx = lv["fps_idx"], lv["new_center"], lv["new_offset"], lv["group_idx"]
    else:
        if stride > 1:
            fps_idx, new_center, new_offset = _sample(stride, center, offset, num_sector, training)
        else:
            fps_idx, new_center, new_offset = None, center, offset
        group_idx, _ = P.knnquery(nsample, center, new_center, offset, new_offset)
    # rows of `normal` at the sampled points: the packed grouping kernel with one sample per row (its backward scatters with
    # atomics; torch's advanced indexing would sort the indices in its backward)
    new_normal = P.grouping(normal.contiguous(), fps_idx.to(torch.int32).view(-1, 1)).view(fps_idx.shape[0], -1) \
        if fps_idx is not None else normal
    if not return_polar and FUSE_GATHER:
        # the row matrix is never built: the first GEMMs gather [xyz | normal | feature] rows of a per-point table with TMA
        rows = gather_rows(center, new_center, group_idx, normal, feature, nsample)
        return new_center, new_normal, rows, rows.layout, new_offset
    rows, layout = group_rows(center, new_center, group_idx, normal, feature, nsample, return_polar)
    return new_center, new_normal, rows, layout, new_offset


class SurfaceAbstractionCD(nn.Module):
    """ref: segmentation/modules/repsurface_utils.py:176-230.
    forward([center [N,3], normal [N,Cn], feature [N,C], offset [B]])
      -> [new_center [M,3], new_normal [M,Cn], new_feature [M,mlp[-1]], new_offset [B]]."""

    def __init__(self, stride, nsample, feat_channel, pos_channel, mlp, return_normal=True, return_polar=False,
                 num_sector=1):
        super().__init__()
        self.stride, self.nsample = stride, nsample
        self.return_normal, self.return_polar, self.num_sector = return_normal, return_polar, num_sector
        self.pos_channel = pos_channel
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        self.mlp_l0 = nn.Conv1d(self.pos_channel, mlp[0], 1)
        self.mlp_f0 = nn.Conv1d(feat_channel, mlp[0], 1)
        self.bn_l0 = nn.BatchNorm1d(mlp[0])
        self.bn_f0 = nn.BatchNorm1d(mlp[0])
        last = mlp[0]
        for out_channel in mlp[1:]:
            self.mlp_convs.append(nn.Conv1d(last, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm1d(out_channel))
            last = out_channel

    def forward(self, pos_nor_feat_off):
        center, normal, feature, offset = pos_nor_feat_off
        new_center, new_normal, rows, layout, new_offset = _sample_and_group(
            self.stride, self.nsample, center, normal, feature, offset, self.return_polar, self.num_sector,
            self.training)
        budget = _ACTIVE_PLAN.sm_budget(center) if (_ACTIVE_PLAN is not None and SM_BUDGET) else 0
        if budget:
            N.lib().rsb_tc_set_sm_budget(budget)
        try:
            new_feature = sa_mlp(rows, self.pos_channel, self, self.nsample, layout)
        finally:
            if budget:
                N.lib().rsb_tc_set_sm_budget(0)
        return [new_center, new_normal, new_feature, new_offset]


class SurfaceFeaturePropagationCD(nn.Module):
    """ref: segmentation/modules/repsurface_utils.py:233-284.
    forward([xyz1 [N,3], points1 [N,C1]|None, offset1], [xyz2 [M,3], points2 [M,C2], offset2]) -> [N, mlp[-1]]."""

    def __init__(self, prev_channel, skip_channel, mlp):
        super().__init__()
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        self.skip = skip_channel is not None
        self.mlp_f0 = nn.Linear(prev_channel, mlp[0])
        self.norm_f0 = nn.BatchNorm1d(mlp[0])
        if skip_channel is not None:
            self.mlp_s0 = nn.Linear(skip_channel, mlp[0])
            self.norm_s0 = nn.BatchNorm1d(mlp[0])
        last = mlp[0]
        for out_channel in mlp[1:]:
            self.mlp_convs.append(nn.Linear(last, out_channel))
            self.mlp_bns.append(nn.BatchNorm1d(out_channel))
            last = out_channel

    def forward(self, pos_feat_off1, pos_feat_off2):
        xyz1, points1, offset1 = pos_feat_off1
        xyz2, points2, offset2 = pos_feat_off2
        pre = _ACTIVE_PLAN.neighbours(3, xyz2, xyz1) if _ACTIVE_PLAN is not None else None
        idx, dist = pre if pre is not None else P.knnquery(3, xyz2, xyz1, offset2, offset1)   # coarse neighbours of every fine point
        weight = P._idw(dist).contiguous()
        coarse = tc.linear_bn(points2, self.mlp_f0, self.norm_f0, relu=False)   # projected BEFORE interpolation (:267)
        x = P._InterpApply.apply(coarse, idx, weight)
        if self.skip:
            x = x + tc.linear_bn(points1, self.mlp_s0, self.norm_s0, relu=False)
        x = F.relu(x)
        for lin, bn in zip(self.mlp_convs, self.mlp_bns):
            x = tc.linear_bn(x, lin, bn, relu=True)
        return x


class UmbrellaSurfaceConstructor(nn.Module):
    """ref: segmentation/modules/repsurface_utils.py:287-329.
    forward(center [N,3], offset [B]) -> [N,out_channel]."""

    def __init__(self, k, in_channel, out_channel, random_inv=True, sort='fix'):
        super().__init__()
        if sort not in (None, 'fix'):
            raise Exception('No such sorting method')
        self.k, self.random_inv, self.sort = k, random_inv, sort
        self.mlps = nn.Sequential(
            nn.Conv1d(in_channel, out_channel, 1, bias=True),
            nn.BatchNorm1d(out_channel),
            nn.ReLU(True),
            nn.Conv1d(out_channel, out_channel, 1, bias=True),
        )

    def forward(self, center, offset):
        with torch.no_grad():
            # all k neighbours are kept, the query itself included (the reference takes no [:, 1:] slice)
            idx, _ = P.knnquery(self.k, center, center, offset, offset)
            if self.random_inv:
                # same draw as the reference: numpy global RNG, one value per cloud (recons_utils.py:28-37)
                sizes = P._sizes(P.host_offsets(offset))
                if torch.cuda.is_current_stream_capturing():
                    # a host draw (and the copy out of its temporary pinned buffer) would be frozen into the graph: draw on the
                    # device, from torch's graph-aware generator, so that every replay flips afresh (same distribution)
                    sign = (torch.rand(offset.shape[0], device=center.device) < 0.5).float() * 2. - 1.
                else:
                    keep = np.random.rand(offset.shape[0]) < 0.5
                    sign = torch.from_numpy(np.where(keep, 1.0, -1.0).astype(np.float32))
                    sign = sign.pin_memory().to(center.device, non_blocking=True)
                flip = torch.repeat_interleave(sign, P.const_tensor(sizes, torch.int64, center.device),
                                               output_size=center.shape[0])
            else:
                flip = None
            # one kernel: azimuth sort, triangles, normals, centroids, polar form, plane constant, NaN repair
            fused = self.mlps[0].weight.shape[0] == 10 and self.mlps[0].weight.shape[1] == 10 and self.k <= 256
            ld = 10 if fused else 12                    # the generic path wants 16-byte aligned rows (TMA)
            feat = torch.empty(center.shape[0], self.k, ld, device=center.device)
            N.call("rsb_umbrella_features", center.shape[0], self.k, 0, 1 if self.sort == 'fix' else 0, 1,
                   center.contiguous(), idx, flip, feat, 10, ld)
        if fused:
            # both 10-channel layers, the BatchNorm, the ReLU and the sum over triangles in recomputing SIMT kernels
            return tc.umbrella_mlp_fused(feat, self.mlps[0], self.mlps[1], self.mlps[3])
        n, g, _ = feat.shape
        x = tc.linear_bn(feat.view(n * g, ld), self.mlps[0], self.mlps[1], relu=True)
        return tc.linear(x, self.mlps[3]).view(n, g, -1).sum(dim=1)
