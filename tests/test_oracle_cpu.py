"""CPU suite (-m "not gpu"): pins the oracle against the golden vectors generated from the reference
(oracle/make_golden.py) and against the reference's in-repo known answers."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import oracle as O
from oracle import model_ref as MR


def _no_dropout(model):
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0


def _close(a, b, rtol=2e-4, atol=1e-5):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-6)
    return np.abs(a - b).max() <= atol + rtol * scale, np.abs(a - b).max() / scale


def _grad_close(a, b):
    """End-to-end gradients of this architecture are only piecewise continuous: every max-pool routes the
    gradient of a (cloud, channel) pair to ONE row, so a single arg-max flip caused by 1e-6-level forward
    noise moves O(1/B) of a weight row.  torch CPU run with 1 vs 8 threads already differs by 7-17 % (max-norm)
    on these tensors (measured, see DESIGN.md "tolerances"), so end-to-end gradients are compared by direction
    and Frobenius norm; tight per-op gradient checks live in tests/test_ops_gpu.py."""
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    nb = np.linalg.norm(b)
    if nb < 1e-2:   # parameters whose gradient is analytically zero (biases in front of a BatchNorm)
        return np.linalg.norm(a) < 5e-2, np.linalg.norm(a)
    cos = float(a @ b / (np.linalg.norm(a) * nb))
    rel = float(np.linalg.norm(a - b) / nb)
    return cos > 0.995 and rel < 0.10, (cos, rel)


def test_opt_n_threads_matches_reference_formula():
    # classification/modules/pointops/src/cuda_utils.h:15-18
    assert [O.opt_n_threads(n) for n in (1, 2, 3, 7, 8, 9, 511, 512, 1023, 1024, 1025, 40960)] == \
        [1, 2, 2, 4, 8, 8, 256, 512, 512, 1024, 1024, 1024]


def test_fps_known_answer_airplane(golden_dir):
    """visualization/airplane_0001.txt is stored in FPS order: reference-semantics FPS returns 0..m-1."""
    xyz = torch.from_numpy(np.load(os.path.join(golden_dir, "airplane_xyz_4096.npy")))[None].contiguous()
    idx = O.fps_dense(xyz[:, :2048].contiguous(), 1024)
    # the file was resampled with all 10000 points present, so only a prefix property holds on a subset:
    assert idx[0, 0] == 0
    full = O.fps_dense(xyz, 64)
    assert full[0, 0] == 0 and len(set(full[0].tolist())) == 64


def test_fps_tie_rule_bit_reversed_thread():
    """All points identical => every running distance is 0 after the first pick; the reference's tree keeps the
    lower slot, so with n = BS the winner is always index 0; with duplicated pairs the bit-reversed rule shows."""
    xyz = torch.zeros(1, 8, 3)
    assert O.fps_dense(xyz, 4)[0].tolist() == [0, 0, 0, 0]
    # two far points tie exactly: indices 5 and 6 at the same distance from point 0 (BS = 8).
    xyz = torch.zeros(1, 8, 3)
    xyz[0, 5, 0] = 1.0
    xyz[0, 6, 0] = -1.0
    # bitrev3(5)=5 (101->101), bitrev3(6)=3 (110->011) -> thread 6 wins the tie
    assert O.fps_dense(xyz, 2)[0].tolist() == [0, 6]


def test_ballquery_semantics():
    xyz = torch.tensor([[[0., 0, 0], [0.05, 0, 0], [1, 0, 0], [0.1, 0, 0], [0.19, 0, 0]]])
    q = torch.tensor([[[0., 0, 0], [5., 5, 5]]])
    idx = O.ballquery(0.2, 3, xyz, q)
    assert idx[0, 0].tolist() == [0, 1, 3]          # first 3 in index order, strict <
    assert idx[0, 1].tolist() == [0, 0, 0]          # empty ball -> zeros
    idx = O.ballquery(0.2, 8, xyz, q)
    assert idx[0, 0].tolist() == [0, 1, 3, 4, 0, 0, 0, 0]   # padded with the first hit


def test_knn_dense_stable_and_heap_sorted():
    torch.manual_seed(0)
    xyz = torch.rand(2, 300, 3)
    idx, d2 = O.knn_dense(9, xyz, return_dist2=True)
    assert (idx[:, :, 0] == torch.arange(300)[None]).all()           # self first (d2 = 0)
    assert (d2[:, :, 1:] >= d2[:, :, :-1]).all()
    hidx, hd2 = O.knn_heap_dense(9, xyz, return_dist2=True)
    assert torch.equal(hd2, d2)                                       # no ties in random data -> same order
    assert torch.equal(hidx, idx)
    # brute-force check
    D = ((xyz[:, :, None] - xyz[:, None]) ** 2).sum(-1)
    assert torch.equal(D.argsort(dim=-1, stable=True)[:, :, :9].int(), idx)


def test_knn_packed_respects_segments():
    torch.manual_seed(1)
    xyz = torch.rand(500, 3)
    off = torch.tensor([200, 500], dtype=torch.int32)
    idx, dist = O.knn_packed(4, xyz, xyz, off, off)
    assert (idx[:200] < 200).all() and (idx[200:] >= 200).all()
    assert torch.equal(idx[:, 0], torch.arange(500, dtype=torch.int32))


def test_sectorized_fps_shapes_and_membership():
    torch.manual_seed(2)
    xyz = torch.rand(12000 + 500, 3) - 0.5
    off = torch.tensor([12000, 12500], dtype=torch.int32)
    noff = torch.tensor([3000, 3125], dtype=torch.int32)
    idx = O.sectorized_fps(xyz, off, noff, 4)
    assert idx.dtype == torch.int64 and idx.shape[0] == 3125
    assert (idx[:3000] < 12000).all() and (idx[3000:] >= 12000).all()
    assert len(set(idx.tolist())) == 3125
    # sector-major order: azimuth sectors of the first cloud are contiguous runs of 750
    ang = torch.atan2(xyz[idx[:3000], 0], xyz[idx[:3000], 1])
    assert ang[:750].max() <= ang[750:1500].min() + 1e-6


@pytest.mark.timeout(600)
def test_model_ref_cls_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "cls_b6_n1024.npz"))
    model = MR.det_fill_(MR.ClsNet())
    _no_dropout(model)
    model.train()
    taps = {}
    torch.manual_seed(1234)
    out = model(torch.from_numpy(g["x"]), taps)
    from repsurf_b200.models import SmoothClsLoss
    loss = SmoothClsLoss()(out, torch.from_numpy(g["y"]))
    from oracle.make_golden import cls_probe
    (taps["sa3_feat"] * cls_probe(taps["sa3_feat"].shape)).sum().backward()
    assert np.array_equal(taps["sa1_center"].detach().numpy(), g["sa1_center"])   # FPS picks: exact
    for name, got in (("umb", taps["umb"][:, :, ::4]), ("sa1_feat", taps["sa1_feat"][:, :, ::4]),
                      ("sa3_feat", taps["sa3_feat"])):
        ok, err = _close(got.detach().numpy(), g[name])
        assert ok, (name, err)
    # The classifier's BatchNorm1d normalises over only B=6 samples, which amplifies fp32 summation-order
    # noise of the 1024-d feature: the head gets a looser bound and gradients are probed on sa3's output.
    for name, got in (("out", out), ("loss", loss)):
        ok, err = _close(got.detach().numpy(), g[name], rtol=2e-3)
        assert ok, (name, err)
    params = dict(model.named_parameters())
    for k in g.files:
        if k.startswith("grad:"):
            ok, err = _grad_close(params[k[5:]].grad.numpy(), g[k])
            assert ok, (k, err)
    sd = model.state_dict()
    assert _close(sd["sa1.bn_l0.running_mean"].numpy(), g["bn_mean"])[0]
    assert _close(sd["sa2.mlp_bns.0.running_var"].numpy(), g["bn_var"])[0]


@pytest.mark.timeout(600)
def test_model_ref_seg_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "seg_10240_6000.npz"))
    model = MR.det_fill_(MR.SegNet())
    _no_dropout(model)
    model.train()
    taps = {}
    np.random.seed(4321)
    out = model([torch.from_numpy(g["coord"]), torch.from_numpy(g["feat"]), torch.from_numpy(g["offset"])], taps)
    loss = nn.CrossEntropyLoss()(out, torch.from_numpy(g["target"]))
    loss.backward()
    assert np.array_equal(taps["sa1_center"].detach().numpy(), g["sa1_center"])   # sectorized FPS: exact
    for name, got in (("umb", taps["umb"][::8]), ("sa1_feat", taps["sa1_feat"][::4]), ("out", out[::8]), ("loss", loss)):
        ok, err = _close(got.detach().numpy(), g[name])
        assert ok, (name, err)
    params = dict(model.named_parameters())
    for k in g.files:
        if k.startswith("grad:"):
            ok, err = _grad_close(params[k[5:]].grad.numpy(), g[k])
            assert ok, (k, err)
    sd = model.state_dict()
    assert _close(sd["sa1.bn_l0.running_mean"].numpy(), g["bn_mean"])[0]
    assert _close(sd["fp2.norm_s0.running_var"].numpy(), g["bn_var"])[0]


def test_model_ref_state_dict_keys_match_reference(golden_dir):
    """The oracle's model restatement, like the CUDA models, must carry the reference's exact state_dict keys / shapes."""
    import json
    ref = json.load(open(os.path.join(golden_dir, "reference_state_dict_keys.json")))
    for mine, name in ((MR.ClsNet(), "cls"), (MR.SegNet(), "seg")):
        assert {k: list(v.shape) for k, v in mine.state_dict().items()} == ref[name]


def test_model_ref_eval_mode_matches_reference_golden(golden_dir):
    e = np.load(os.path.join(golden_dir, "eval_mode.npz"))
    g = np.load(os.path.join(golden_dir, "seg_10240_6000.npz"))
    model = MR.det_fill_(MR.SegNet()).eval()
    taps = {}
    np.random.seed(4321)
    with torch.no_grad():
        out = model([torch.from_numpy(g["coord"]), torch.from_numpy(g["feat"]), torch.from_numpy(g["offset"])], taps)
    assert np.array_equal(taps["sa1_center"].numpy(), e["seg_sa1_center"])      # plain FPS in eval mode: exact
    for name, got in (("seg_sa1_feat", taps["sa1_feat"][::4]), ("seg_out", out[::8])):
        ok, err = _close(got.numpy(), e[name])
        assert ok, (name, err)
    gc = np.load(os.path.join(golden_dir, "cls_b6_n1024.npz"))
    model = MR.det_fill_(MR.ClsNet()).eval()
    taps = {}
    torch.manual_seed(1234)
    with torch.no_grad():
        out = model(torch.from_numpy(gc["x"]), taps)
    for name, got in (("cls_sa3_feat", taps["sa3_feat"]), ("cls_out", out)):
        ok, err = _close(got.numpy(), e[name])
        assert ok, (name, err)


def test_geometry_ref_umbrella_matches_reference_tensors(golden_dir):
    """oracle/geometry_ref.py (vectorised umbrella restatement used as a checker by the GPU tests) against the umbrella
    descriptors of the UNMODIFIED reference (input of its umbrella MLP), same kNN lists from the C oracle."""
    from oracle import oracle as O
    from oracle.geometry_ref import umbrella_features
    g = np.load(os.path.join(golden_dir, "seg_10240_6000.npz"))
    coord, offset = torch.from_numpy(g["coord"]), torch.from_numpy(g["offset"])
    idx, _ = O.knn_packed(9, coord, coord, offset, offset)
    np.random.seed(4321)
    keep = np.random.rand(offset.shape[0]) < 0.5
    sizes = np.diff(np.concatenate([[0], g["offset"]]))
    flip = torch.from_numpy(np.repeat(np.where(keep, 1.0, -1.0).astype(np.float32), sizes))
    sel = torch.arange(0, coord.shape[0], 8)
    offs = coord[idx[sel].long()] - coord[sel][:, None]
    got = umbrella_features(offs, flip[sel].view(-1, 1, 1), rotate_key=True, order="seg").numpy()
    assert np.abs(got - g["umb_feat"]).max() < 1e-6
