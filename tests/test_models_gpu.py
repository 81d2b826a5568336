"""GPU suite: the CUDA module layer / model stacks versus the golden vectors generated from the UNMODIFIED
reference Python (oracle/make_golden.py), with the same deterministic parameters (oracle.model_ref.det_fill_)."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle.model_ref import det_fill_
from oracle.make_golden import cls_probe

pytestmark = pytest.mark.gpu
cuda = torch.device("cuda")

# Forward features: the reference's own torch fp32 arithmetic moves by ~4e-5 (max-norm, relative to the tensor's
# largest magnitude) when only its summation order changes (1 vs 8 CPU threads, measured in DESIGN.md), because
# train-mode BatchNorm chains amplify 1e-7 rounding noise.  Per-GEMM parity at 1e-5 is tested in test_mlp_gpu.py.
FWD_RTOL = 2e-4


def _no_dropout(model):
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0


def _close(a, b, rtol=FWD_RTOL, atol=1e-5):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-6)
    err = np.abs(a - b).max() / scale
    return err <= rtol + atol / scale, err


def _grad_close(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    nb = np.linalg.norm(b)
    if nb < 1e-2:
        return np.linalg.norm(a) < 5e-2, np.linalg.norm(a)
    cos = float(a @ b / (np.linalg.norm(a) * nb))
    rel = float(np.linalg.norm(a - b) / nb)
    return cos > 0.995 and rel < 0.10, (cos, rel)


def _rowwise_outliers(a, b, rtol=FWD_RTOL):
    """fraction of rows (last dim = channels) whose error exceeds the tolerance"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-6)
    bad = (np.abs(a - b).max(axis=-1) / scale) > rtol
    return float(bad.mean())


def test_cls_model_matches_reference_golden(golden_dir):
    from repsurf_b200.models import RepSurfCls, SmoothClsLoss
    g = np.load(os.path.join(golden_dir, "cls_b6_n1024.npz"))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = det_fill_(RepSurfCls())
    _no_dropout(model)
    model = model.to(cuda).train()
    taps = {}
    model.surface_constructor.register_forward_hook(lambda m, i, o: taps.__setitem__("umb", o))
    model.sa1.register_forward_hook(lambda m, i, o: taps.update(sa1_center=o[0], sa1_feat=o[2]))
    model.sa3.register_forward_hook(lambda m, i, o: taps.update(sa3_feat=o[2]))
    torch.manual_seed(1234)
    out = model(torch.from_numpy(g["x"]).to(cuda))
    loss = SmoothClsLoss()(out, torch.from_numpy(g["y"]).to(cuda))
    (taps["sa3_feat"] * cls_probe(taps["sa3_feat"].shape).to(cuda)).sum().backward()
    assert np.array_equal(taps["sa1_center"].detach().cpu().numpy(), g["sa1_center"])       # FPS picks: exact
    # umbrella: a point whose two neighbours have azimuths within one ulp may sort differently under CUDA's
    # atan2 -> allow isolated outlier points, everything else tight
    umb = taps["umb"].detach().cpu().numpy()[:, :, ::4]
    assert _rowwise_outliers(np.moveaxis(umb, 1, -1), np.moveaxis(g["umb"], 1, -1)) < 2e-3
    for name, got in (("sa1_feat", taps["sa1_feat"][:, :, ::4]), ("sa3_feat", taps["sa3_feat"])):
        ok, err = _close(got.detach().cpu().numpy(), g[name])
        assert ok, (name, err)
    for name, got in (("out", out), ("loss", loss)):
        ok, err = _close(got.detach().cpu().numpy(), g[name], rtol=2e-3)
        assert ok, (name, err)
    params = dict(model.named_parameters())
    for k in g.files:
        if k.startswith("grad:"):
            ok, err = _grad_close(params[k[5:]].grad.cpu().numpy(), g[k])
            assert ok, (k, err)
    sd = model.state_dict()
    assert _close(sd["sa1.bn_l0.running_mean"].cpu().numpy(), g["bn_mean"])[0]
    assert _close(sd["sa2.mlp_bns.0.running_var"].cpu().numpy(), g["bn_var"])[0]
    assert int(sd["sa1.bn_l0.num_batches_tracked"]) == 1


def test_seg_model_matches_reference_golden(golden_dir):
    from repsurf_b200.models import RepSurfSeg
    g = np.load(os.path.join(golden_dir, "seg_10240_6000.npz"))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = det_fill_(RepSurfSeg())
    _no_dropout(model)
    model = model.to(cuda).train()
    taps = {}
    model.surface_constructor.register_forward_hook(lambda m, i, o: taps.__setitem__("umb", o))
    model.sa1.register_forward_hook(lambda m, i, o: taps.update(sa1_center=o[0], sa1_feat=o[2], sa1_offset=o[3]))
    np.random.seed(4321)
    out = model([torch.from_numpy(g["coord"]).to(cuda), torch.from_numpy(g["feat"]).to(cuda),
                 torch.from_numpy(g["offset"]).to(cuda)])
    loss = nn.CrossEntropyLoss()(out, torch.from_numpy(g["target"]).to(cuda))
    loss.backward()
    assert np.array_equal(taps["sa1_offset"].cpu().numpy(), g["sa1_offset"])
    same_fps = np.array_equal(taps["sa1_center"].detach().cpu().numpy(), g["sa1_center"])
    assert _rowwise_outliers(taps["umb"].detach().cpu().numpy()[::8], g["umb"]) < 2e-3
    if not same_fps:
        pytest.skip("sector membership differs by a CPU/CUDA atan2 ulp; downstream tensors are not comparable")
    for name, got in (("sa1_feat", taps["sa1_feat"][::4]), ("out", out[::8]), ("loss", loss)):
        ok, err = _close(got.detach().cpu().numpy(), g[name], rtol=5e-4)
        assert ok, (name, err)
    params = dict(model.named_parameters())
    for k in g.files:
        if k.startswith("grad:"):
            ok, err = _grad_close(params[k[5:]].grad.cpu().numpy(), g[k])
            assert ok, (k, err)
    sd = model.state_dict()
    assert _close(sd["sa1.bn_l0.running_mean"].cpu().numpy(), g["bn_mean"])[0]
    assert _close(sd["fp2.norm_s0.running_var"].cpu().numpy(), g["bn_var"])[0]


def test_reference_state_dict_keys_load_strict():
    """Checkpoint compatibility: key names / shapes equal the oracle's (which mirror the reference's)."""
    from repsurf_b200.models import RepSurfCls, RepSurfSeg
    from oracle import model_ref as MR
    for mine, ref in ((RepSurfCls(), MR.ClsNet()), (RepSurfSeg(), MR.SegNet())):
        a, b = mine.state_dict(), ref.state_dict()
        assert list(a.keys()).sort() == list(b.keys()).sort()
        assert {k: tuple(v.shape) for k, v in a.items()} == {k: tuple(v.shape) for k, v in b.items()}
        mine.load_state_dict(b, strict=True)
