"""Dense-layout RepSurf-U modules with the reference's constructor / forward signatures and
state_dict keys (classification/modules/repsurface_utils.py:186-307), running on the sm_100a
operator layer (`repsurf_b200.cls.pointops`).  CUDA only.

Drop-in: `from repsurf_b200.cls.modules import UmbrellaSurfaceConstructor, SurfaceAbstractionCD`
replaces `from modules.repsurface_utils import ...` in
classification/models/repsurf/repsurf_ssg_umb.py:8 (the `cuda=` argument is accepted and ignored:
there is only the CUDA path).

What differs from the reference implementation (results are the same):
  * FPS also emits the sampled coordinates (no separate gathering launch), no `.zero_()` pre-fills,
    no `torch.cuda.empty_cache()` after every op, no permute().contiguous() round trips of the
    grouped tensor: groups are built channel-first [B,C,m,ns] directly by the grouping kernel.
  * umbrella geometry is one kernel (csrc/umbrella.cu); every Conv/BatchNorm/ReLU runs on the tcgen05 row GEMMs
    (repsurf_b200.tc), in training and in eval mode.
"""
import torch
import torch.nn as nn

from . import pointops as P
from .. import tc
from ..mlp import group_rows, sa_mlp


def _grouped_inputs(npoint, radius, nsample, center_cf, normal, feature, return_normal, return_polar):
    """Sampling + ball-query grouping (reference sample_and_group, repsurface_utils.py:15-59).
    center_cf [B,3,N], normal [B,Cn,N], feature [B,Cf,N]|None ->
      new_center [B,3,m], new_normal [B,Cn,m], rows [B*m*ns, C] with C ordered
      [rel xyz(3), polar(3)?, normal(Cn)?, feature(Cf)?].
    Groups are gathered ROW-major (one contiguous C-vector per (centre, sample)) with the packed grouping
    kernel on a [B*N, C] view and globalised indices, so the shared MLP runs as plain GEMMs over rows."""
    B, _, N = center_cf.shape
    xyz = center_cf.transpose(1, 2).contiguous()                       # [B,N,3]
    fps_idx, new_xyz = P.furthestsampling_with_xyz(xyz, npoint)         # [B,m], [B,m,3]
    new_normal = P.gathering(normal.contiguous(), fps_idx)              # [B,Cn,m]
    idx = P.ballquery(radius, nsample, xyz, new_xyz)                    # [B,m,ns] local ids
    gidx = (idx + (torch.arange(B, device=idx.device, dtype=torch.int32) * N).view(B, 1, 1)).view(B * npoint, nsample)
    # one kernel builds the packed row matrix [rel xyz, polar | normal | feature]
    use_normal = feature is None or return_normal
    rows, layout = group_rows(xyz.view(B * N, 3), new_xyz.view(B * npoint, 3), gidx,
                              normal.transpose(1, 2).reshape(B * N, -1) if use_normal else None,
                              feature.transpose(1, 2).reshape(B * N, -1) if feature is not None else None,
                              nsample, return_polar)
    return new_xyz.transpose(1, 2).contiguous(), new_normal, rows, layout


def _all_inputs(center_cf, normal, feature, return_normal, return_polar):
    """group_all (reference sample_and_group_all, repsurface_utils.py:62-88): one group holding every point,
    RAW coordinates (not centre-relative), new_center = new_normal = zeros[B,3,1].  -> rows [B*N, C].
    Same row builder as the grouped levels: the group of cloud b is rows b*N .. (b+1)*N-1 around a zero centre
    (x - 0 is exact, so the rows are the raw coordinates and their polar form)."""
    B, _, N = center_cf.shape
    dev = center_cf.device
    new_center = torch.zeros(B, 3, 1, device=dev, dtype=center_cf.dtype)
    xyz = center_cf.transpose(1, 2).contiguous().view(B * N, 3)
    idx = torch.arange(B * N, device=dev, dtype=torch.int32)
    rows, layout = group_rows(xyz, new_center.view(B, 3), idx,
                              normal.transpose(1, 2).reshape(B * N, -1) if return_normal else None,
                              feature.transpose(1, 2).reshape(B * N, -1), N, return_polar)
    return new_center, new_center, rows, layout


class SurfaceAbstractionCD(nn.Module):
    """ref: classification/modules/repsurface_utils.py:186-249.
    forward(center [B,3,N], normal [B,Cn,N], feature [B,Cf,N] | None)
      -> (new_center [B,3,m], new_normal [B,Cn,m], new_feature [B,mlp[-1],m])."""

    def __init__(self, npoint, radius, nsample, feat_channel, pos_channel, mlp, group_all,
                 return_normal=True, return_polar=False, cuda=True):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.return_normal, self.return_polar = return_normal, return_polar
        self.pos_channel, self.group_all = pos_channel, group_all
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        self.mlp_l0 = nn.Conv2d(self.pos_channel, mlp[0], 1)
        self.mlp_f0 = nn.Conv2d(feat_channel, mlp[0], 1)
        self.bn_l0 = nn.BatchNorm2d(mlp[0])
        self.bn_f0 = nn.BatchNorm2d(mlp[0])
        last = mlp[0]
        for out_channel in mlp[1:]:
            self.mlp_convs.append(nn.Conv2d(last, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            last = out_channel

    def forward(self, center, normal, feature):
        B, N = center.shape[0], center.shape[2]
        if self.group_all:
            new_center, new_normal, rows, layout = _all_inputs(center, normal, feature, self.return_normal, self.return_polar)
            groups, ns = 1, N
        else:
            new_center, new_normal, rows, layout = _grouped_inputs(self.npoint, self.radius, self.nsample, center, normal,
                                                                   feature, self.return_normal, self.return_polar)
            groups, ns = self.npoint, self.nsample
        # channel de-differentiation: position and feature channels get their own first layer (:236-239)
        pooled = sa_mlp(rows, self.pos_channel, self, ns, layout)        # [B*groups, C']
        return new_center, new_normal, pooled.view(B, groups, -1).transpose(1, 2).contiguous()


class UmbrellaSurfaceConstructor(nn.Module):
    """ref: classification/modules/repsurface_utils.py:252-307.
    forward(center [B,3,N]) -> [B,in_channel,N]."""

    def __init__(self, k, in_channel, aggr_type='sum', return_dist=False, random_inv=True, cuda=True):
        super().__init__()
        self.k, self.return_dist, self.random_inv, self.aggr_type = k, return_dist, random_inv, aggr_type
        self.mlps = nn.Sequential(
            nn.Conv2d(in_channel, in_channel, 1, bias=False),
            nn.BatchNorm2d(in_channel),
            nn.ReLU(True),
            nn.Conv2d(in_channel, in_channel, 1, bias=True),
            nn.BatchNorm2d(in_channel),
            nn.ReLU(True),
            nn.Conv2d(in_channel, in_channel, 1, bias=True),
        )

    def forward(self, center):
        B, _, N = center.shape
        N_pts = N
        center_cf = center.contiguous()
        xyz = center_cf.transpose(1, 2).contiguous()
        with torch.no_grad():
            idx = P.knnquery(self.k, xyz, xyz)                                  # [B,N,k] local ids, self first
            if self.random_inv:
                # same draw as the reference: CPU generator, one per forward (recons_utils.py:49-51)
                if torch.cuda.is_current_stream_capturing():                    # a host draw would be frozen into the graph
                    sign = torch.randint(0, 2, (B, 1, 1), device=center.device).float() * 2. - 1.
                else:
                    sign = (torch.randint(0, 2, (B, 1, 1)).float() * 2. - 1.).to(center.device)
            else:
                sign = torch.ones(B, 1, 1, device=center.device)
            # one kernel: drop the query itself (:119), azimuth sort, triangles, normals, centroids, polar, NaN repair
            from .. import _native as _nat
            gidx = (idx + (torch.arange(B, device=idx.device, dtype=torch.int32) * N_pts).view(B, 1, 1)).contiguous()
            # triangle rows of pitch 12 floats: 9 or 10 descriptor channels + zero padding (16-byte aligned rows for TMA)
            feat = torch.empty(B, N_pts, self.k - 1, 12, device=center.device)
            _nat.call("rsb_umbrella_features", B * N_pts, self.k, 1, 0, 0, xyz.view(B * N_pts, 3), gidx.view(B * N_pts, self.k),
                      sign.expand(B, N_pts, 1).reshape(B * N_pts).contiguous(), feat, 10 if self.return_dist else 9, 12)
            G = self.k - 1
        x = tc.linear_bn(feat.view(B * N * G, 12), self.mlps[0], self.mlps[1], relu=True)
        x = tc.linear_bn(x, self.mlps[3], self.mlps[4], relu=True)
        x = tc.linear(x, self.mlps[6]).view(B, N, G, -1)
        if self.aggr_type == 'max':
            x = torch.max(x, 2)[0]
        elif self.aggr_type == 'avg':
            x = torch.mean(x, 2)
        else:
            x = torch.sum(x, 2)
        return x.transpose(1, 2).contiguous()
