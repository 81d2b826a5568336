"""INTEGRATION.md path B executed: the `pointops_cuda` stand-ins of integration/pointops_cuda_shim.py called with the
reference wrappers' calling convention (caller-allocated outputs, legacy argument order) against the C oracle."""
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
cuda = torch.device("cuda")


def test_cls_shim_functions_match_oracle():
    from integration.pointops_cuda_shim import cls_module
    pc = cls_module()
    g = torch.Generator().manual_seed(0)
    b, n, m, ns, c = 3, 700, 128, 16, 7
    xyz = torch.rand(b, n, 3, generator=g)
    dx = xyz.to(cuda)
    idx = torch.zeros(b, m, dtype=torch.int32, device=cuda)
    temp = torch.full((b, n), 1e10, device=cuda)
    pc.furthestsampling_cuda(b, n, m, dx, temp, idx)                                    # cls pointops.py:46
    assert torch.equal(idx.cpu(), O.fps_dense(xyz, m))
    new_xyz = torch.gather(xyz, 1, idx.cpu().long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    bq = torch.zeros(b, m, ns, dtype=torch.int32, device=cuda)
    pc.ballquery_cuda(b, n, m, 0.2, ns, new_xyz.to(cuda), dx, bq)                        # :221
    assert torch.equal(bq.cpu(), O.ballquery(0.2, ns, xyz, new_xyz))
    kn = torch.zeros(b, m, ns, dtype=torch.int32, device=cuda)
    d2 = torch.zeros(b, m, ns, device=cuda)
    pc.knnquery_cuda(b, n, m, ns, dx, new_xyz.to(cuda), kn, d2)                          # :315
    assert torch.equal(kn.cpu(), O.knn_dense(ns, xyz, new_xyz))
    feat = torch.randn(b, c, n, generator=g)
    out = torch.zeros(b, c, m, ns, device=cuda)
    pc.grouping_forward_cuda(b, c, n, m, ns, feat.to(cuda), kn, out)                     # :163
    assert torch.equal(out.cpu(), O.group_fwd(feat, kn.cpu()))
    gout = torch.zeros(b, c, m, device=cuda)
    pc.gathering_forward_cuda(b, c, n, m, feat.to(cuda), idx, gout)                      # :68
    assert torch.equal(gout.cpu(), O.gather_fwd(feat, idx.cpu()))
    with pytest.raises(RuntimeError):                                                    # errors come back as exceptions
        pc.knnquery_cuda(b, n, m, 1000, dx, new_xyz.to(cuda), kn, d2)


def test_seg_shim_functions_match_oracle():
    from integration.pointops_cuda_shim import seg_module
    ps = seg_module()
    g = torch.Generator().manual_seed(1)
    sizes, new_sizes = (900, 500), (225, 125)
    n, m, ns = sum(sizes), sum(new_sizes), 8
    xyz = torch.rand(n, 3, generator=g)
    off = torch.tensor([900, 1400], dtype=torch.int32)
    noff = torch.tensor([225, 350], dtype=torch.int32)
    idx = torch.zeros(m, dtype=torch.int32, device=cuda)
    tmp = torch.full((n,), 1e10, device=cuda)
    ps.furthestsampling_cuda(2, 900, xyz.to(cuda), off.to(cuda), noff.to(cuda), tmp, idx)            # seg pointops.py:44
    assert torch.equal(idx.cpu(), O.fps_packed(xyz, off, noff))
    new_xyz = xyz[idx.cpu().long()].contiguous()
    kn = torch.zeros(m, ns, dtype=torch.int32, device=cuda)
    d2 = torch.zeros(m, ns, device=cuda)
    ps.knnquery_cuda(m, ns, xyz.to(cuda), new_xyz.to(cuda), off.to(cuda), noff.to(cuda), kn, d2)    # :126 (squared distances)
    widx, wd = O.knn_packed(ns, xyz, new_xyz, off, noff, sqrt=False)
    assert torch.equal(kn.cpu(), widx) and torch.equal(d2.cpu(), wd)
    feat = torch.randn(n, 12, generator=g)
    out = torch.zeros(m, ns, 12, device=cuda)
    ps.grouping_forward_cuda(m, ns, 12, feat.to(cuda), kn, out)                                       # :147
    assert torch.equal(out.cpu(), O.group_packed_fwd(feat, kn.cpu()))
    sidx = torch.randint(0, n, (n, ns), generator=g, dtype=torch.int32)
    sub = torch.zeros(n, ns, 12, device=cuda)
    ps.subtraction_forward_cuda(n, ns, 12, feat.to(cuda), feat.to(cuda), sidx.to(cuda), sub)          # :201
    assert torch.equal(sub.cpu(), O.subtraction_fwd(feat, feat, sidx))
