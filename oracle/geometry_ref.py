"""TEST INFRASTRUCTURE (never imported by repsurf_b200): umbrella-surface geometry as vectorised tensor math,
device-agnostic, forward-only.  Pinned to the unmodified reference's tensors by tests/test_oracle_cpu.py and used by the
GPU tests as a checker of csrc/umbrella.cu and csrc/group.cu on inputs the golden files do not cover.

Restates, for a whole umbrella at once, what the reference spreads over
  {classification,segmentation}/modules/polar_utils.py:10-31        (xyz2sphere)
  {classification,segmentation}/modules/recons_utils.py              (cal_normal, cal_center, cal_const, check_nan_umb)
  classification/modules/repsurface_utils.py:112-132                 (group_by_umbrella)
  segmentation/modules/repsurface_utils.py:71-98                     (_fixed_rotate, group_by_umbrella_v2)
"""
import math

import torch

_SQRT3 = math.sqrt(3.0)


def xyz2sphere(xyz, normalize=True):
    """(rho, theta, phi) of [..., 3] vectors; theta := 0 where rho == 0; normalised to [0,1] like
    polar_utils.py:27-29 (theta/pi, phi/(2pi)+0.5)."""
    x, y, z = xyz[..., 0:1], xyz[..., 1:2], xyz[..., 2:3]
    rho = torch.sqrt(torch.sum(xyz * xyz, dim=-1, keepdim=True))
    theta = torch.acos(z / rho)
    theta = torch.where(rho == 0, torch.zeros_like(theta), theta)
    phi = torch.atan2(y, x)
    if normalize:
        theta = theta / math.pi
        phi = phi / (2 * math.pi) + 0.5
    return torch.cat([rho, theta, phi], dim=-1)


# segmentation/modules/repsurface_utils.py:73 — fixed rotation applied only to the SORT KEY
_ROT = ((0.5, -0.5, 0.7071), (0.7071, 0.7071, 0.0), (-0.5, 0.5, 0.7071))


def umbrella_features(offsets, flip, rotate_key, order):
    """offsets: [..., G, 3] neighbour positions relative to the centre (unsorted).
    flip:    broadcastable to [..., 1, 1]: +-1 per cloud (the reference's random normal inversion).
    rotate_key: sort by the azimuth of the ROTATED offsets (segmentation 'fix' sort) or of the raw ones.
    order:   'cls' -> [centroid(3), polar(3), normal(3), pos(1)]   (classification/modules/repsurface_utils.py:290)
             'seg' -> [polar(3), normal(3), pos(1), centroid(3)]   (segmentation/modules/repsurface_utils.py:320)
    returns [..., G, 10].

    Triangle i of an umbrella = (centre, p_i, p_{i+1}) with p sorted by azimuth (cyclic).
    normal_i = unit(p_i x p_{i+1}), sign chosen so that triangle 0's x component is positive (NaN counts
    as not positive -> -1), times `flip`; centroid_i = (0 + p_i + p_{i+1}) / 3; pos_i = <normal_i, centroid_i>/sqrt(3).
    Triangles whose normal is NaN (degenerate) take normal / centroid / pos of the first non-NaN triangle of
    the same umbrella (check_nan_umb); the polar form is computed BEFORE that repair, as in the reference."""
    key_src = offsets
    if rotate_key:
        rot = torch.tensor(_ROT, dtype=offsets.dtype, device=offsets.device)
        key_src = offsets @ rot
    phi = torch.atan2(key_src[..., 1], key_src[..., 0]) / (2 * math.pi) + 0.5
    perm = phi.argsort(dim=-1)
    p = torch.gather(offsets, -2, perm.unsqueeze(-1).expand_as(offsets))
    p_next = torch.roll(p, -1, dims=-2)

    nor = torch.cross(p, p_next, dim=-1)
    unit = nor / torch.norm(nor, dim=-1, keepdim=True)
    sign = (unit[..., 0:1, 0:1] > 0).to(unit.dtype) * 2.0 - 1.0
    unit = unit * sign * flip

    centroid = (torch.zeros_like(p) + p + p_next) / 3
    polar = xyz2sphere(centroid)
    pos = torch.sum(unit * centroid, dim=-1, keepdim=True) / _SQRT3

    bad = torch.isnan(unit).any(dim=-1)                       # [..., G]
    first_ok = torch.argmax((~bad).int(), dim=-1, keepdim=True)  # [..., 1]

    def repair(t):
        first = torch.gather(t, -2, first_ok.unsqueeze(-1).expand(*first_ok.shape, t.shape[-1]))
        return torch.where(bad.unsqueeze(-1), first, t)

    unit, centroid, pos = repair(unit), repair(centroid), repair(pos)
    if order == "cls":
        return torch.cat([centroid, polar, unit, pos], dim=-1)
    return torch.cat([polar, unit, pos, centroid], dim=-1)
