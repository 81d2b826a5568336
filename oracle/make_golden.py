"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference through
oracle/ref_loader.py, point operators served by the C oracle) on seeded inputs.

Run in the build container only (needs /root/reference):  python -m oracle.make_golden
The fixtures pin (a) oracle/model_ref.py on CPU (tests/test_oracle_cpu.py) and (b) the CUDA modules
(tests/test_models_gpu.py) against the reference's own Python.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

from . import ref_loader as RL
from .model_ref import det_fill_

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
GRAD_KEYS_CLS = ["surface_constructor.mlps.0.weight", "surface_constructor.mlps.6.bias", "sa1.mlp_l0.weight",
                 "sa1.mlp_f0.weight", "sa1.bn_f0.weight", "sa2.mlp_convs.1.weight", "sa3.mlp_l0.bias",
                 "sa3.mlp_convs.1.weight"]
GRAD_KEYS_SEG = ["surface_constructor.mlps.0.weight", "surface_constructor.mlps.3.bias", "sa1.mlp_l0.weight",
                 "sa1.mlp_f0.weight", "sa2.bn_l0.weight", "sa4.mlp_convs.1.bias", "fp4.mlp_f0.weight",
                 "fp1.mlp_convs.1.weight", "classifier.4.weight"]


def _no_dropout(model):
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0


def _np(t):
    return t.detach().cpu().numpy()


def cls_inputs(B=6, N=1024, seed=11):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, N, generator=g) * 2 - 1
    y = torch.randint(0, 15, (B,), generator=g)
    return x, y


def cls_probe(shape, seed=13):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def seg_inputs(sizes=(10240, 6000), seed=12):
    g = torch.Generator().manual_seed(seed)
    n = sum(sizes)
    coord = torch.rand(n, 3, generator=g) * torch.tensor([8.0, 8.0, 3.0])
    o = np.cumsum(sizes)
    for a, b in zip([0] + list(o[:-1]), o):
        coord[a:b] -= coord[a:b].mean(0, keepdim=True)     # mean-centred per cloud (seg/util/data_util.py:62-63)
    feat = torch.randn(n, 3, generator=g)
    target = torch.randint(0, 13, (n,), generator=g)
    return coord.contiguous(), feat, torch.tensor(o, dtype=torch.int32), target


def golden_cls():
    x, y = cls_inputs()
    with RL.RefTree("cls") as t:
        Model = t.imp("models.repsurf.repsurf_ssg_umb").Model
        Loss = t.imp("util.utils").SmoothClsLoss
        model = det_fill_(Model(RL.cls_args()))
        _no_dropout(model)
        model.train()
        taps = {}
        model.surface_constructor.register_forward_hook(lambda m, i, o: taps.__setitem__("umb", o))
        model.sa1.register_forward_hook(lambda m, i, o: taps.update(sa1_center=o[0], sa1_feat=o[2]))
        model.sa3.register_forward_hook(lambda m, i, o: taps.update(sa3_feat=o[2]))
        torch.manual_seed(1234)
        out = model(x)
        loss = Loss()(out, y)
        # Gradients are probed on the 1024-d global feature (a fixed random projection of it): the classifier's
        # BatchNorm1d over only B samples is too ill-conditioned for a 1e-4-level gradient comparison.
        (taps["sa3_feat"] * cls_probe(taps["sa3_feat"].shape)).sum().backward()
        sd = model.state_dict()
        params = dict(model.named_parameters())
        np.savez_compressed(
            os.path.join(OUT, "cls_b6_n1024.npz"), x=_np(x), y=_np(y), umb=_np(taps["umb"][:, :, ::4]),
            sa1_center=_np(taps["sa1_center"]), sa1_feat=_np(taps["sa1_feat"][:, :, ::4]),
            sa3_feat=_np(taps["sa3_feat"]), out=_np(out), loss=_np(loss),
            bn_mean=_np(sd["sa1.bn_l0.running_mean"]), bn_var=_np(sd["sa2.mlp_bns.0.running_var"]),
            **{"grad:" + k: _np(params[k].grad) for k in GRAD_KEYS_CLS})
        print("cls golden: loss", float(loss))


def golden_seg():
    coord, feat, offset, target = seg_inputs()
    with RL.RefTree("seg") as t:
        Model = t.imp("models.repsurf.repsurf_umb_ssg").Model
        model = det_fill_(Model(RL.seg_args()))
        _no_dropout(model)
        model.train()
        taps = {}
        model.surface_constructor.register_forward_hook(lambda m, i, o: taps.__setitem__("umb", o))
        model.sa1.register_forward_hook(lambda m, i, o: taps.update(sa1_center=o[0], sa1_feat=o[2], sa1_offset=o[3]))
        np.random.seed(4321)
        out = model([coord, feat, offset])
        loss = nn.CrossEntropyLoss()(out, target)
        loss.backward()
        sd = model.state_dict()
        params = dict(model.named_parameters())
        np.savez_compressed(
            os.path.join(OUT, "seg_10240_6000.npz"), coord=_np(coord), feat=_np(feat), offset=_np(offset),
            target=_np(target), umb=_np(taps["umb"][::8]), sa1_center=_np(taps["sa1_center"]),
            sa1_feat=_np(taps["sa1_feat"][::4]), sa1_offset=_np(taps["sa1_offset"]), out=_np(out[::8]),
            loss=_np(loss), bn_mean=_np(sd["sa1.bn_l0.running_mean"]), bn_var=_np(sd["fp2.norm_s0.running_var"]),
            **{"grad:" + k: _np(params[k].grad) for k in GRAD_KEYS_SEG})
        print("seg golden: loss", float(loss))


def golden_kat():
    """Known-answer data from the reference's visualization/ assets: the clouds are stored in FPS order, so
    reference-semantics FPS on airplane_0001 returns 0,1,2,...."""
    pts = np.loadtxt(os.path.join(RL.REF_ROOT, "visualization", "airplane_0001.txt"), delimiter=",", dtype=np.float32)
    np.save(os.path.join(OUT, "airplane_xyz_4096.npy"), pts[:4096, :3].copy())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["cls", "seg", "kat"]
    if "cls" in which:
        golden_cls()
    if "seg" in which:
        golden_seg()
    if "kat" in which:
        golden_kat()
