"""CPU restatement of the RepSurf-U module layer + model stacks (torch fp32 on the host, oracle C for
the point operators).  TEST INFRASTRUCTURE ONLY (see oracle/pointops_oracle.c header): the checker for
the CUDA modules, and the reported CPU baseline of bench.py.  Never imported by `repsurf_b200`.

Pinned against the reference itself: oracle/make_golden.py runs the UNMODIFIED reference models from
/root/reference (through oracle/ref_loader.py) and tests/test_oracle_cpu.py checks this file against
those outputs with the same deterministic parameters.

Formulation: every shared MLP is evaluated on a row matrix [rows, C] (rows = all (cloud, centre, sample)
triples) with F.linear + F.batch_norm — the same arithmetic as the reference's 1x1 Conv2d/Conv1d +
BatchNorm2d/1d over [B,C,ns,m] / [M,C,ns] (train-mode batch statistics over all rows, biased variance for
normalisation, eps 1e-5, momentum 0.1).  Parameters live in torch containers with the reference's
state_dict keys and shapes.

Reference lines followed:
  cls  classification/modules/repsurface_utils.py:15-88 (sample_and_group[_all]), :112-132 (group_by_umbrella),
       :218-249 (SurfaceAbstractionCD.forward), :276-307 (UmbrellaSurfaceConstructor.forward)
       classification/modules/recons_utils.py:27-57,82-90,108-124,152-176; polar_utils.py:10-31
  seg  segmentation/modules/repsurface_utils.py:15-51, 71-98, 206-230, 257-284, 305-329
       segmentation/modules/recons_utils.py:10-45,48-56,74-90,117-138
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import oracle as O

# ------------------------------------------------------------------------------------------ helpers


def _rows_linear(x, layer):
    w = layer.weight
    return F.linear(x, w.reshape(w.shape[0], -1), layer.bias)


def _rows_bn(x, bn, training):
    if training and bn.track_running_stats:
        bn.num_batches_tracked += 1
    return F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, training, bn.momentum, bn.eps)


def _sphere(v):
    rho = v.pow(2).sum(-1, keepdim=True).sqrt()
    theta = torch.acos(v[..., 2:3] / rho)
    theta[rho == 0] = 0
    phi = torch.atan2(v[..., 1:2], v[..., 0:1])
    return torch.cat([rho, theta / np.pi, phi / (2 * np.pi) + .5], -1)


def _umbrella(offsets, flip, rotated_key, order):
    """offsets [..., G, 3] -> 10-channel triangle descriptors [..., G, 10] (see module docstring for lines)."""
    key = offsets
    if rotated_key:
        key = offsets @ torch.FloatTensor([[0.5, -0.5, 0.7071], [0.7071, 0.7071, 0.], [-0.5, 0.5, 0.7071]])
    phi = _sphere(key)[..., 2]
    order_idx = phi.argsort(dim=-1)
    a = torch.gather(offsets, -2, order_idx[..., None].expand_as(offsets))
    b = torch.roll(a, -1, dims=-2)
    tri = torch.stack([torch.zeros_like(a), a, b], dim=-2)              # [..., G, 3(points), 3]
    e1, e2 = tri[..., 1, :] - tri[..., 0, :], tri[..., 2, :] - tri[..., 0, :]
    n = torch.cross(e1, e2, dim=-1)
    n = n / torch.norm(n, dim=-1, keepdim=True)
    n = n * ((n[..., 0:1, 0] > 0).float() * 2. - 1.).unsqueeze(-1)
    n = n * flip
    c = tri.mean(dim=-2)
    polar = _sphere(c)
    pos = (n * c).sum(-1, keepdim=True) / torch.sqrt(torch.Tensor([3]))
    bad = torch.isnan(n).sum(-1) > 0
    first = torch.argmax((~bad).int(), dim=-1)
    for t in (n, c, pos):
        f = torch.gather(t, -2, first[..., None, None].expand(*first.shape, 1, t.shape[-1])).expand_as(t)
        t[bad] = f[bad]
    return torch.cat([c, polar, n, pos], -1) if order == "cls" else torch.cat([polar, n, pos, c], -1)


# ------------------------------------------------------------------------------------------ classification


class _ClsSA(nn.Module):
    def __init__(self, npoint, radius, nsample, feat_c, pos_c, mlp, group_all, polar):
        super().__init__()
        self.cfg = (npoint, radius, nsample, pos_c, group_all, polar)
        self.mlp_l0, self.mlp_f0 = nn.Conv2d(pos_c, mlp[0], 1), nn.Conv2d(feat_c, mlp[0], 1)
        self.bn_l0, self.bn_f0 = nn.BatchNorm2d(mlp[0]), nn.BatchNorm2d(mlp[0])
        self.mlp_convs = nn.ModuleList(nn.Conv2d(i, o, 1) for i, o in zip(mlp[:-1], mlp[1:]))
        self.mlp_bns = nn.ModuleList(nn.BatchNorm2d(o) for o in mlp[1:])

    def forward(self, center, normal, feature):
        npoint, radius, nsample, pos_c, group_all, polar = self.cfg
        B = center.shape[0]
        xyz = center.transpose(1, 2).contiguous()                      # [B,N,3]
        nrm = normal.transpose(1, 2)                                   # [B,N,Cn]
        fea = feature.transpose(1, 2) if feature is not None else None
        bidx = torch.arange(B)[:, None]
        if group_all:
            new_xyz = torch.zeros(B, 1, 3)
            new_nrm = new_xyz
            pos = xyz[:, None]
            if polar:
                pos = torch.cat([pos, _sphere(pos)], -1)
            g = torch.cat([pos, nrm[:, None], fea[:, None]], -1)        # [B,1,N,C]
        else:
            fidx = O.fps_dense(xyz.detach(), npoint).long()
            new_xyz, new_nrm = xyz[bidx, fidx], nrm[bidx, fidx]
            gidx = O.ballquery(radius, nsample, xyz.detach(), new_xyz.detach().contiguous()).long()
            bb = torch.arange(B)[:, None, None]
            pos = xyz[bb, gidx] - new_xyz[:, :, None]
            if polar:
                pos = torch.cat([pos, _sphere(pos)], -1)
            parts = [pos, nrm[bb, gidx]] + ([fea[bb, gidx]] if fea is not None else [])
            g = torch.cat(parts, -1)                                    # [B,m,ns,C]
        Bm, ns = g.shape[0] * g.shape[1], g.shape[2]
        rows = g.reshape(Bm * ns, -1)
        tr = self.training
        x = F.relu(_rows_bn(_rows_linear(rows[:, :pos_c], self.mlp_l0), self.bn_l0, tr)
                   + _rows_bn(_rows_linear(rows[:, pos_c:], self.mlp_f0), self.bn_f0, tr))
        for lin, bn in zip(self.mlp_convs, self.mlp_bns):
            x = F.relu(_rows_bn(_rows_linear(x, lin), bn, tr))
        x = x.view(B, -1, ns, x.shape[-1]).max(dim=2)[0]               # [B,m,C']
        return new_xyz.transpose(1, 2), new_nrm.transpose(1, 2), x.transpose(1, 2)


class _ClsUmbrella(nn.Module):
    def __init__(self, k, c, aggr, return_dist):
        super().__init__()
        self.k, self.aggr, self.return_dist = k, aggr, return_dist
        self.mlps = nn.Sequential(nn.Conv2d(c, c, 1, bias=False), nn.BatchNorm2d(c), nn.ReLU(True),
                                  nn.Conv2d(c, c, 1), nn.BatchNorm2d(c), nn.ReLU(True), nn.Conv2d(c, c, 1))

    def forward(self, center):
        B, _, N = center.shape
        xyz = center.transpose(1, 2).contiguous()
        idx = O.knn_dense(self.k, xyz)[:, :, 1:].long()
        off = xyz[torch.arange(B)[:, None, None], idx] - xyz[:, :, None]
        flip = (torch.randint(0, 2, (B, 1, 1)).float() * 2. - 1.).unsqueeze(-1)
        f = _umbrella(off, flip, False, "cls")
        if not self.return_dist:
            f = f[..., :9]
        G = f.shape[2]
        x = f.reshape(B * N * G, -1)
        tr = self.training
        x = F.relu(_rows_bn(_rows_linear(x, self.mlps[0]), self.mlps[1], tr))
        x = F.relu(_rows_bn(_rows_linear(x, self.mlps[3]), self.mlps[4], tr))
        x = _rows_linear(x, self.mlps[6]).view(B, N, G, -1)
        x = {"max": lambda t: t.max(2)[0], "avg": lambda t: t.mean(2)}.get(self.aggr, lambda t: t.sum(2))(x)
        return x.transpose(1, 2)


class ClsNet(nn.Module):
    """Same stack / state_dict keys as classification/models/repsurf/repsurf_ssg_umb.py:13-57."""

    def __init__(self, return_polar=True, return_dist=True, group_size=8, umb_pool="sum", num_class=15):
        super().__init__()
        pc = 6 if return_polar else 3
        self.surface_constructor = _ClsUmbrella(group_size + 1, 10, umb_pool, return_dist)
        self.sa1 = _ClsSA(512, 0.2, 32, 10, pc, [64, 64, 128], False, return_polar)
        self.sa2 = _ClsSA(128, 0.4, 64, 138, pc, [128, 128, 256], False, return_polar)
        self.sa3 = _ClsSA(None, None, None, 266, pc, [256, 512, 1024], True, return_polar)
        self.classfier = nn.Sequential(nn.Linear(1024, 512), nn.BatchNorm1d(512), nn.ReLU(True), nn.Dropout(0.4),
                                       nn.Linear(512, 256), nn.BatchNorm1d(256), nn.ReLU(True), nn.Dropout(0.4),
                                       nn.Linear(256, num_class))

    def forward(self, points, taps=None):
        c = points[:, :3, :]
        n = self.surface_constructor(c)
        if taps is not None:
            taps["umb"] = n
        c, n, f = self.sa1(c, n, None)
        if taps is not None:
            taps["sa1_center"], taps["sa1_feat"] = c, f
        c, n, f = self.sa2(c, n, f)
        c, n, f = self.sa3(c, n, f)
        if taps is not None:
            taps["sa3_feat"] = f
        return F.log_softmax(self.classfier(f.reshape(-1, 1024)), -1)


# ------------------------------------------------------------------------------------------ segmentation


def _strided(offset, stride):
    o = offset.tolist()
    acc, out, prev = 0, [], 0
    for v in o:
        acc += (v - prev) // stride
        out.append(acc)
        prev = v
    return torch.tensor(out, dtype=torch.int32)


class _SegSA(nn.Module):
    def __init__(self, stride, nsample, feat_c, pos_c, mlp, polar, num_sector=1):
        super().__init__()
        self.cfg = (stride, nsample, pos_c, polar, num_sector)
        self.mlp_l0, self.mlp_f0 = nn.Conv1d(pos_c, mlp[0], 1), nn.Conv1d(feat_c, mlp[0], 1)
        self.bn_l0, self.bn_f0 = nn.BatchNorm1d(mlp[0]), nn.BatchNorm1d(mlp[0])
        self.mlp_convs = nn.ModuleList(nn.Conv1d(i, o, 1) for i, o in zip(mlp[:-1], mlp[1:]))
        self.mlp_bns = nn.ModuleList(nn.BatchNorm1d(o) for o in mlp[1:])

    def forward(self, lvl):
        center, normal, feature, offset = lvl
        stride, nsample, pos_c, polar, num_sector = self.cfg
        if stride > 1:
            noff = _strided(offset, stride)
            if num_sector > 1 and self.training:
                fidx = O.sectorized_fps(center, offset, noff, num_sector)
            else:
                fidx = O.fps_packed(center, offset, noff).long()
            ncenter, nnormal = center[fidx], normal[fidx]
        else:
            ncenter, nnormal, noff = center, normal, offset
        gidx = O.knn_packed(nsample, center, ncenter.contiguous(), offset, noff)[0].long()
        pos = center[gidx] - ncenter[:, None]
        if polar:
            pos = torch.cat([pos, _sphere(pos)], -1)
        parts = [pos, normal[gidx]] + ([feature[gidx]] if feature is not None else [])
        g = torch.cat(parts, -1)                                        # [M,ns,C]
        M, ns = g.shape[:2]
        rows = g.reshape(M * ns, -1)
        tr = self.training
        x = F.relu(_rows_bn(_rows_linear(rows[:, :pos_c], self.mlp_l0), self.bn_l0, tr)
                   + _rows_bn(_rows_linear(rows[:, pos_c:], self.mlp_f0), self.bn_f0, tr))
        for lin, bn in zip(self.mlp_convs, self.mlp_bns):
            x = F.relu(_rows_bn(_rows_linear(x, lin), bn, tr))
        return [ncenter, nnormal, x.view(M, ns, -1).max(1)[0], noff]


class _SegFP(nn.Module):
    def __init__(self, prev_c, skip_c, mlp):
        super().__init__()
        self.skip = skip_c is not None
        self.mlp_f0, self.norm_f0 = nn.Linear(prev_c, mlp[0]), nn.BatchNorm1d(mlp[0])
        if self.skip:
            self.mlp_s0, self.norm_s0 = nn.Linear(skip_c, mlp[0]), nn.BatchNorm1d(mlp[0])
        self.mlp_convs = nn.ModuleList(nn.Linear(i, o) for i, o in zip(mlp[:-1], mlp[1:]))
        self.mlp_bns = nn.ModuleList(nn.BatchNorm1d(o) for o in mlp[1:])

    def forward(self, fine, coarse):
        xyz1, pts1, off1 = fine
        xyz2, pts2, off2 = coarse
        idx, dist = O.knn_packed(3, xyz2, xyz1, off2, off1)
        r = 1.0 / (dist + 1e-8)
        w = r / r.sum(1, keepdim=True)
        tr = self.training
        p2 = _rows_bn(self.mlp_f0(pts2), self.norm_f0, tr)
        x = torch.zeros(xyz1.shape[0], p2.shape[1])
        for i in range(3):
            x = x + p2[idx[:, i].long()] * w[:, i:i + 1]
        if self.skip:
            x = x + _rows_bn(self.mlp_s0(pts1), self.norm_s0, tr)
        x = F.relu(x)
        for lin, bn in zip(self.mlp_convs, self.mlp_bns):
            x = F.relu(_rows_bn(lin(x), bn, tr))
        return x


class _SegUmbrella(nn.Module):
    def __init__(self, k, cin, cout):
        super().__init__()
        self.k = k
        self.mlps = nn.Sequential(nn.Conv1d(cin, cout, 1), nn.BatchNorm1d(cout), nn.ReLU(True), nn.Conv1d(cout, cout, 1))

    def forward(self, center, offset):
        idx = O.knn_packed(self.k, center, center, offset, offset)[0].long()
        off = center[idx] - center[:, None]
        keep = np.random.rand(offset.shape[0]) < 0.5                    # numpy RNG, one draw per cloud
        o = [0] + offset.tolist()
        flip = torch.cat([torch.full((o[i + 1] - o[i], 1, 1), 1.0 if keep[i] else -1.0) for i in range(len(o) - 1)])
        f = _umbrella(off, flip, True, "seg")                           # [N,k,10]
        Np, G = f.shape[:2]
        x = f.reshape(Np * G, -1)
        x = F.relu(_rows_bn(_rows_linear(x, self.mlps[0]), self.mlps[1], self.training))
        return _rows_linear(x, self.mlps[3]).view(Np, G, -1).sum(1)


class SegNet(nn.Module):
    """Same stack / state_dict keys as segmentation/models/repsurf/repsurf_umb_ssg.py:13-63 (return_polar=False)."""

    def __init__(self, in_channel=6, num_class=13, group_size=8):
        super().__init__()
        self.sa1 = _SegSA(4, 32, in_channel + 10, 3, [32, 32, 64], False, num_sector=4)
        self.sa2 = _SegSA(4, 32, 74, 3, [64, 64, 128], False)
        self.sa3 = _SegSA(4, 32, 138, 3, [128, 128, 256], False)
        self.sa4 = _SegSA(4, 32, 266, 3, [256, 256, 512], False)
        self.fp4 = _SegFP(512, 256, [256, 256])
        self.fp3 = _SegFP(256, 128, [256, 256])
        self.fp2 = _SegFP(256, 64, [256, 128])
        self.fp1 = _SegFP(128, None, [128, 128, 128])
        self.classifier = nn.Sequential(nn.Linear(128, 128), nn.BatchNorm1d(128), nn.ReLU(True), nn.Dropout(0.5),
                                        nn.Linear(128, num_class))
        self.surface_constructor = _SegUmbrella(group_size + 1, 10, 10)

    def forward(self, inp, taps=None):
        coord, feat, offset = inp
        l0 = [coord, self.surface_constructor(coord, offset), torch.cat([coord, feat], 1), offset]
        if taps is not None:
            taps["umb"] = l0[1]
        l1 = self.sa1(l0)
        if taps is not None:
            taps["sa1_center"], taps["sa1_feat"] = l1[0], l1[2]
        l2 = self.sa2(l1)
        l3 = self.sa3(l2)
        l4 = self.sa4(l3)
        f3 = self.fp4([l3[0], l3[2], l3[3]], [l4[0], l4[2], l4[3]])
        f2 = self.fp3([l2[0], l2[2], l2[3]], [l3[0], f3, l3[3]])
        f1 = self.fp2([l1[0], l1[2], l1[3]], [l2[0], f2, l2[3]])
        f0 = self.fp1([l0[0], None, l0[3]], [l1[0], f1, l1[3]])
        return self.classifier(f0)


# ------------------------------------------------------------------------------------------ deterministic parameters


def det_fill_(model, seed=0):
    """Overwrite every parameter / buffer with values that depend only on (key, shape, seed), so that the
    reference model, this restatement and the CUDA model can be given identical weights without shipping a
    checkpoint.  BN weights ~ 1 +- 0.1, running_var in [0.5, 1.5], everything else ~ N(0, fan-in scaled)."""
    import zlib
    sd = model.state_dict()
    for key, t in sd.items():
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + seed) & 0x7FFFFFFF)
        if key.endswith("num_batches_tracked"):
            t.zero_()
        elif key.endswith("running_var"):
            t.copy_(torch.rand(t.shape, generator=g) + 0.5)
        elif key.endswith("running_mean"):
            t.copy_(torch.randn(t.shape, generator=g) * 0.1)
        elif t.dim() == 1 and key.endswith("weight"):      # every 1-D weight is a BatchNorm scale
            t.copy_(1.0 + 0.1 * torch.randn(t.shape, generator=g))
        elif t.dim() == 1:                                   # biases (conv / linear / BN)
            t.copy_(0.1 * torch.randn(t.shape, generator=g))
        else:
            fan_in = int(np.prod(t.shape[1:]))
            t.copy_(torch.randn(t.shape, generator=g) * math.sqrt(2.0 / fan_in))
    model.load_state_dict(sd)
    return model
