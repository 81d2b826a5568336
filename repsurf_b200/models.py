"""RepSurf-U model assemblies used as the HARNESS for tests and bench.py (the reference's model
files are out of scope as code — SURVEY.md §2 rows 7/12 — but /root/reference does not exist on the
GPU box, so the same layer stacks are declared here).  Attribute names and hyper-parameters follow
  classification/models/repsurf/repsurf_ssg_umb.py:13-57   (RepSurfCls; note the reference's `classfier` spelling)
  segmentation/models/repsurf/repsurf_umb_ssg.py:13-63     (RepSurfSeg)
so that a reference state_dict loads unchanged (strict=True)."""
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

from .cls import modules as cls_m
from .seg import modules as seg_m


def cls_args(**kw):
    """classification/scripts/scanobjectnn/repsurf_ssg_umb.sh"""
    a = dict(return_center=True, return_polar=True, return_dist=True, group_size=8, umb_pool="sum",
             cuda_ops=True, num_point=1024, num_class=15)
    a.update(kw)
    return types.SimpleNamespace(**a)


def seg_args(**kw):
    """segmentation/scripts/s3dis/train_repsurf_umb.sh + segmentation/tool/train.py:452-470"""
    a = dict(return_polar=False, in_channel=6, group_size=8, num_class=13)
    a.update(kw)
    return types.SimpleNamespace(**a)


class RepSurfCls(nn.Module):
    def __init__(self, args=None):
        super().__init__()
        args = args or cls_args()
        pos_c = 0 if not args.return_center else (6 if args.return_polar else 3)
        rs = 10
        self.init_nsample = args.num_point
        self.return_dist = args.return_dist
        self.surface_constructor = cls_m.UmbrellaSurfaceConstructor(args.group_size + 1, rs, return_dist=args.return_dist,
                                                                    aggr_type=args.umb_pool)
        self.sa1 = cls_m.SurfaceAbstractionCD(512, 0.2, 32, rs, pos_c, [64, 64, 128], False, return_polar=args.return_polar)
        self.sa2 = cls_m.SurfaceAbstractionCD(128, 0.4, 64, 128 + rs, pos_c, [128, 128, 256], False, return_polar=args.return_polar)
        self.sa3 = cls_m.SurfaceAbstractionCD(None, None, None, 256 + rs, pos_c, [256, 512, 1024], True, return_polar=args.return_polar)
        self.classfier = nn.Sequential(
            nn.Linear(1024, 512), nn.BatchNorm1d(512), nn.ReLU(True), nn.Dropout(0.4),
            nn.Linear(512, 256), nn.BatchNorm1d(256), nn.ReLU(True), nn.Dropout(0.4),
            nn.Linear(256, args.num_class))

    def forward(self, points):
        center = points[:, :3, :]
        normal = self.surface_constructor(center)
        center, normal, feature = self.sa1(center, normal, None)
        center, normal, feature = self.sa2(center, normal, feature)
        center, normal, feature = self.sa3(center, normal, feature)
        # same head (Linear-BN-ReLU-Dropout x2, Linear), GEMMs + BatchNorm on the tensor cores
        from . import tc
        c = self.classfier
        x = c[3](tc.linear_bn(feature.view(-1, 1024), c[0], c[1], relu=True))
        x = c[7](tc.linear_bn(x, c[4], c[5], relu=True))
        return F.log_softmax(tc.linear(x, c[8]), -1)


class RepSurfSeg(nn.Module):
    def __init__(self, args=None):
        super().__init__()
        args = args or seg_args()
        pos_c = 6 if args.return_polar else 3
        rs_in, rs = 10, 10
        self.sa1 = seg_m.SurfaceAbstractionCD(4, 32, args.in_channel + rs, pos_c, [32, 32, 64], True, args.return_polar, num_sector=4)
        self.sa2 = seg_m.SurfaceAbstractionCD(4, 32, 64 + rs, pos_c, [64, 64, 128], True, args.return_polar)
        self.sa3 = seg_m.SurfaceAbstractionCD(4, 32, 128 + rs, pos_c, [128, 128, 256], True, args.return_polar)
        self.sa4 = seg_m.SurfaceAbstractionCD(4, 32, 256 + rs, pos_c, [256, 256, 512], True, args.return_polar)
        self.fp4 = seg_m.SurfaceFeaturePropagationCD(512, 256, [256, 256])
        self.fp3 = seg_m.SurfaceFeaturePropagationCD(256, 128, [256, 256])
        self.fp2 = seg_m.SurfaceFeaturePropagationCD(256, 64, [256, 128])
        self.fp1 = seg_m.SurfaceFeaturePropagationCD(128, None, [128, 128, 128])
        self.classifier = nn.Sequential(
            nn.Linear(128, 128), nn.BatchNorm1d(128), nn.ReLU(True), nn.Dropout(0.5), nn.Linear(128, args.num_class))
        self.surface_constructor = seg_m.UmbrellaSurfaceConstructor(args.group_size + 1, rs_in, rs)

    def forward(self, pos_feat_off0):
        coord, feat, offset = pos_feat_off0
        # sampling / neighbour search of all four levels start now on side streams (they depend on coordinates only) and
        # overlap the umbrella constructor and the shared MLPs; the modules below pick the results up
        sas = (self.sa1, self.sa2, self.sa3, self.sa4)
        with seg_m.GeometryPlan(coord, offset, [(m.stride, m.nsample, m.num_sector) for m in sas], self.training):
            return self._forward(coord, feat, offset)

    def _forward(self, coord, feat, offset):
        l0 = [coord, self.surface_constructor(coord, offset), torch.cat([coord, feat], 1), offset]
        l1 = self.sa1(l0)
        l2 = self.sa2(l1)
        l3 = self.sa3(l2)
        l4 = self.sa4(l3)
        f3 = self.fp4([l3[0], l3[2], l3[3]], [l4[0], l4[2], l4[3]])
        f2 = self.fp3([l2[0], l2[2], l2[3]], [l3[0], f3, l3[3]])
        f1 = self.fp2([l1[0], l1[2], l1[3]], [l2[0], f2, l2[3]])
        f0 = self.fp1([l0[0], None, l0[3]], [l1[0], f1, l1[3]])
        # same head (Linear-BN-ReLU-Dropout-Linear), GEMMs + BatchNorm on the tensor cores (training and eval)
        from . import tc
        c = self.classifier
        return tc.linear(c[3](tc.linear_bn(f0, c[0], c[1], relu=True)), c[4])


class SmoothClsLoss(nn.Module):
    """Label-smoothed NLL over log-probabilities, the loss of classification/util/utils.py:55-69:
    target weight 1-eps on the true class and eps/(C-1) on every other class."""

    def __init__(self, smoothing_ratio=0.1):
        super().__init__()
        self.smoothing_ratio = smoothing_ratio

    def forward(self, log_prob, target):
        c = log_prob.size(1)
        off = self.smoothing_ratio / (c - 1)
        true_lp = log_prob.gather(1, target.view(-1, 1)).squeeze(1)
        per_sample = (1.0 - self.smoothing_ratio - off) * true_lp + off * log_prob.sum(dim=1)
        return -per_sample.mean()
