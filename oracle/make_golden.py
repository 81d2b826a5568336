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
        # raw umbrella descriptors = the input of the umbrella MLP, [B, C, G, N] (repsurface_utils.py:291-293)
        model.surface_constructor.mlps[0].register_forward_pre_hook(lambda m, i: taps.__setitem__("umb_feat", i[0]))
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
            umb_feat=_np(taps["umb_feat"][:, :, :, ::4].permute(0, 3, 2, 1)),          # [B, N/4, G, C]
            sa1_center=_np(taps["sa1_center"]), sa1_feat=_np(taps["sa1_feat"][:, :, ::4]),
            sa3_feat=_np(taps["sa3_feat"]), out=_np(out), loss=_np(loss),
            bn_mean=_np(sd["sa1.bn_l0.running_mean"]), bn_var=_np(sd["sa2.mlp_bns.0.running_var"]),
            **{"grad:" + k: _np(params[k].grad) for k in GRAD_KEYS_CLS})
        print("cls golden: loss", float(loss))


def sector_edge_margin(coord, offset, num_sectors=4, min_points=10000):
    """Smallest |azimuth - inner sector edge| over the clouds that take the sectorized FPS path
    (seg/po/functions/pointops.py:70-78).  CPU and CUDA atan2 differ by an ulp (2.4e-7 near pi): inputs whose margin is
    far above that select the same sector on both, so FPS picks are comparable bit for bit."""
    m, start = float("inf"), 0
    for end in offset.tolist():
        pts = coord[start:end]
        if end - start >= min_points:
            ang = torch.atan2(pts[:, 0], pts[:, 1])
            edges = torch.linspace(float(ang.min()), float(ang.max()) + 1e-4, num_sectors + 1)[1:-1]
            m = min(m, float((ang[:, None] - edges[None, :]).abs().min()))
        start = end
    return m


SEG_SMALL = dict(name="seg_10240_6000.npz", sizes=(10240, 6000), seed=12, s_umb=8, s_feat=4, s_out=8)
# 8 clouds, every one above the sectorized-FPS threshold: sectorized FPS + grid kNN + feature propagation pinned together
SEG_BIG = dict(name="seg_8clouds.npz", sizes=(10240, 12000, 10000, 11111, 10240, 13000, 10500, 10240), seed=55,
               s_umb=64, s_feat=32, s_out=64)


def golden_seg(cfg=SEG_SMALL):
    coord, feat, offset, target = seg_inputs(cfg["sizes"], cfg["seed"])
    assert sector_edge_margin(coord, offset) > 1e-5, "pick another seed: a point sits on a sector edge"
    with RL.RefTree("seg") as t:
        Model = t.imp("models.repsurf.repsurf_umb_ssg").Model
        model = det_fill_(Model(RL.seg_args()))
        _no_dropout(model)
        model.train()
        taps = {}
        model.surface_constructor.register_forward_hook(lambda m, i, o: taps.__setitem__("umb", o))
        # raw umbrella descriptors = the input of the umbrella MLP, [N, C, G] (repsurface_utils.py:320-321)
        model.surface_constructor.mlps[0].register_forward_pre_hook(lambda m, i: taps.__setitem__("umb_feat", i[0]))
        model.sa1.register_forward_hook(lambda m, i, o: taps.update(sa1_center=o[0], sa1_feat=o[2], sa1_offset=o[3]))
        model.sa2.register_forward_hook(lambda m, i, o: taps.update(sa2_center=o[0]))
        np.random.seed(4321)
        out = model([coord, feat, offset])
        loss = nn.CrossEntropyLoss()(out, target)
        loss.backward()
        sd = model.state_dict()
        params = dict(model.named_parameters())
        su, sf, so = cfg["s_umb"], cfg["s_feat"], cfg["s_out"]
        big = len(cfg["sizes"]) > 2
        inputs = dict(coord_sum=np.float64(coord.double().sum()), feat_sum=np.float64(feat.double().sum())) if big else \
            dict(coord=_np(coord), feat=_np(feat), target=_np(target))
        np.savez_compressed(
            os.path.join(OUT, cfg["name"]), offset=_np(offset), umb=_np(taps["umb"][::su]),
            umb_feat=_np(taps["umb_feat"][::su].transpose(1, 2)),                      # [N/su, G, C]
            sa1_center=_np(taps["sa1_center"]), sa2_center=_np(taps["sa2_center"][::4]),
            sa1_feat=_np(taps["sa1_feat"][::sf]), sa1_offset=_np(taps["sa1_offset"]), out=_np(out[::so]),
            loss=_np(loss), bn_mean=_np(sd["sa1.bn_l0.running_mean"]), bn_var=_np(sd["fp2.norm_s0.running_var"]),
            **inputs, **{"grad:" + k: _np(params[k].grad) for k in GRAD_KEYS_SEG})
        print(cfg["name"], "loss", float(loss))


def golden_eval():
    """Eval-mode forward of the UNMODIFIED reference models (running statistics; plain FPS in sa1 of the segmentation
    network; the umbrella flip stays random, as in the reference) on the small inputs."""
    coord, feat, offset, target = seg_inputs(SEG_SMALL["sizes"], SEG_SMALL["seed"])
    with RL.RefTree("seg") as t:
        model = det_fill_(t.imp("models.repsurf.repsurf_umb_ssg").Model(RL.seg_args())).eval()
        taps = {}
        model.sa1.register_forward_hook(lambda m, i, o: taps.update(sa1_center=o[0], sa1_feat=o[2]))
        np.random.seed(4321)
        with torch.no_grad():
            out = model([coord, feat, offset])
        seg = dict(seg_out=_np(out[::8]), seg_sa1_center=_np(taps["sa1_center"]), seg_sa1_feat=_np(taps["sa1_feat"][::4]))
    x, y = cls_inputs()
    with RL.RefTree("cls") as t:
        model = det_fill_(t.imp("models.repsurf.repsurf_ssg_umb").Model(RL.cls_args())).eval()
        taps = {}
        model.sa3.register_forward_hook(lambda m, i, o: taps.update(sa3_feat=o[2]))
        torch.manual_seed(1234)
        with torch.no_grad():
            out = model(x)
        cls = dict(cls_out=_np(out), cls_sa3_feat=_np(taps["sa3_feat"]))
    np.savez_compressed(os.path.join(OUT, "eval_mode.npz"), **seg, **cls)
    print("eval golden written")


def golden_sample():
    """The reference's evaluation-time resampling `sample()` (classification/modules/pointnet2_utils.py:114-124, torch-native
    FPS with a random first pick) on a seeded batch, run unmodified on the CPU."""
    g = torch.Generator().manual_seed(21)
    x = torch.rand(4, 6, 2048, generator=g) * 2 - 1
    with RL.RefTree("cls") as t:
        sample = t.imp("modules.pointnet2_utils").sample
        torch.manual_seed(77)
        out = sample(1024, x.clone())
    np.savez_compressed(os.path.join(OUT, "cls_sample.npz"), x=_np(x), out=_np(out))
    print("sample golden written", tuple(out.shape))


def golden_keys():
    """state_dict keys + shapes of the UNMODIFIED reference models (checkpoint compatibility contract)."""
    import json
    out = {}
    with RL.RefTree("cls") as t:
        m = t.imp("models.repsurf.repsurf_ssg_umb").Model(RL.cls_args())
        out["cls"] = {k: list(v.shape) for k, v in m.state_dict().items()}
    with RL.RefTree("seg") as t:
        m = t.imp("models.repsurf.repsurf_umb_ssg").Model(RL.seg_args())
        out["seg"] = {k: list(v.shape) for k, v in m.state_dict().items()}
    json.dump(out, open(os.path.join(OUT, "reference_state_dict_keys.json"), "w"), indent=0, sort_keys=True)
    print("keys:", len(out["cls"]), "cls,", len(out["seg"]), "seg")


def golden_kat():
    """Known-answer data from the reference's visualization/ assets: the clouds are stored in FPS order, so
    reference-semantics FPS on airplane_0001 returns 0,1,2,...."""
    pts = np.loadtxt(os.path.join(RL.REF_ROOT, "visualization", "airplane_0001.txt"), delimiter=",", dtype=np.float32)
    np.save(os.path.join(OUT, "airplane_xyz_4096.npy"), pts[:4096, :3].copy())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["cls", "seg", "segbig", "eval", "keys", "kat", "sample"]
    if "sample" in which:
        golden_sample()
    if "eval" in which:
        golden_eval()
    if "cls" in which:
        golden_cls()
    if "seg" in which:
        golden_seg(SEG_SMALL)
    if "segbig" in which:
        golden_seg(SEG_BIG)
    if "keys" in which:
        golden_keys()
    if "kat" in which:
        golden_kat()
