"""Classification evaluation harness: sample() against the UNMODIFIED reference's output (golden) and against
a torch restatement of its FPS loop run on the GPU; augmentations against their formulas under the same seeds; the voting
loop end to end."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
cuda = torch.device("cuda")


def _torch_native_fps(xyz, npoint, farthest):
    """classification/modules/pointnet2_utils.py:62-75 with the first pick given (the reference draws it with torch.randint)"""
    B, N, _ = xyz.shape
    centroids = torch.zeros(B, npoint, dtype=torch.long, device=xyz.device)
    distance = torch.ones(B, N, device=xyz.device) * 1e10
    batch_indices = torch.arange(B, dtype=torch.long, device=xyz.device)
    for i in range(npoint):
        centroids[:, i] = farthest
        centroid = xyz[batch_indices, farthest, :].view(B, 1, 3)
        dist = torch.sum((xyz - centroid) ** 2, -1)
        mask = dist < distance
        distance[mask] = dist[mask]
        farthest = torch.max(distance, -1)[1]
    return centroids


def test_sample_matches_reference_golden(golden_dir):
    from repsurf_b200.cls import harness as H
    g = np.load(os.path.join(golden_dir, "cls_sample.npz"))
    torch.manual_seed(77)                                   # the reference draws the first picks from the global CPU generator
    out = H.sample(1024, torch.from_numpy(g["x"]).to(cuda))
    assert np.array_equal(out.cpu().numpy(), g["out"])


@pytest.mark.parametrize("B,C,n,m", [(8, 3, 1024, 256), (5, 6, 2048, 1024), (3, 3, 5000, 700), (2, 4, 12000, 300)])
def test_sample_matches_torch_native_fps_on_gpu(B, C, n, m):
    from repsurf_b200.cls import harness as H
    g = torch.Generator().manual_seed(B * n)
    x = (torch.rand(B, C, n, generator=g) * 2 - 1).to(cuda)
    torch.manual_seed(5)
    out, idx = H.sample_with_index(m, x)
    torch.manual_seed(5)
    farthest = torch.randint(0, n, (B,), dtype=torch.long).to(cuda)
    want = _torch_native_fps(x.permute(0, 2, 1)[:, :, :3].contiguous(), m, farthest)
    assert torch.equal(idx, want)
    assert torch.equal(out, torch.gather(x, 2, want[:, None, :].expand(-1, C, -1)))


def test_augmentations_follow_reference_formulas():
    from repsurf_b200.cls import harness as H
    import types
    x = torch.rand(6, 6, 512, device=cuda)
    torch.manual_seed(9)
    a = H.scale_point_cloud(x[:, :3].clone(), 0.5)
    torch.manual_seed(9)
    s = (torch.rand(6, 3, 1, device=cuda) * 2. - 1.) * 0.5 + 1.
    assert torch.equal(a, x[:, :3] * s)
    torch.manual_seed(10)
    b = H.shift_point_cloud(x[:, :3].clone(), 0.3)
    torch.manual_seed(10)
    assert torch.equal(b, x[:, :3] + (torch.rand(6, 3, 1, device=cuda) * 2. - 1.) * 0.3)
    args = types.SimpleNamespace(aug_scale=True, aug_shift=True, dataset="ScanObjectNN")
    torch.manual_seed(11)
    y = H.transform_point_cloud(x.clone(), args, H.get_aug_args(args))
    assert torch.equal(y[:, 3:], x[:, 3:]) and not torch.equal(y[:, :3], x[:, :3])


def test_voting_loop_runs_the_eval_model():
    from repsurf_b200.cls import harness as H
    from repsurf_b200.models import RepSurfCls
    from oracle.model_ref import det_fill_
    model = det_fill_(RepSurfCls()).to(cuda)
    g = torch.Generator().manual_seed(1)
    loader = [(torch.rand(4, 3, 2048, generator=g) * 2 - 1, torch.randint(0, 15, (4,), generator=g)) for _ in range(2)]
    torch.manual_seed(3)
    sing, vote = H.test(model, loader, num_class=15, num_point=1024, num_votes=3, total_num=8)
    assert 0.0 <= sing <= 1.0 and 0.0 <= vote <= 1.0
    torch.manual_seed(3)
    assert H.test(model, loader, num_class=15, num_point=1024, num_votes=3, total_num=8) == (sing, vote)
