"""GPU suite: the REFERENCE's own CUDA kernels (compiled unmodified into oracle/_ref by oracle/build_ref.sh)
versus (a) the CPU oracle — this is what pins the oracle's rounding / tie rules R1-R5 to the real reference —
and (b) the sm_100a kernels at sizes the CPU oracle would take too long for."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests import refcuda as R

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (R.available("cls") and R.available("seg")), reason="oracle/_ref not built")]
cuda = torch.device("cuda")


def _cloud(b, n, seed):
    return torch.rand(b, n, 3, generator=torch.Generator().manual_seed(seed)) * 2 - 1


def _lattice(b, n, seed, lo=-3, hi=4):
    return torch.randint(lo, hi, (b, n, 3), generator=torch.Generator().manual_seed(seed)).float() * 0.5


# ---- (a) oracle == reference CUDA -------------------------------------------------------------------------
@pytest.mark.parametrize("gen,b,n,m", [(_cloud, 3, 1024, 512), (_cloud, 2, 777, 300), (_lattice, 2, 640, 320),
                                       (_cloud, 1, 5000, 600), (_lattice, 1, 96, 50)])
def test_oracle_fps_dense_equals_reference_cuda(gen, b, n, m):
    xyz = gen(b, n, 1 + n)
    assert torch.equal(O.fps_dense(xyz, m), R.fps_dense(xyz.to(cuda), m).cpu())


def test_oracle_fps_packed_equals_reference_cuda():
    sizes = (900, 3000, 411)
    xyz = _cloud(1, sum(sizes), 3)[0].contiguous()
    off = torch.tensor(np.cumsum(sizes), dtype=torch.int32)
    noff = torch.tensor(np.cumsum([s // 4 for s in sizes]), dtype=torch.int32)
    assert torch.equal(O.fps_packed(xyz, off, noff), R.fps_packed(xyz.to(cuda), off.to(cuda), noff.to(cuda)).cpu())
    lat = _lattice(1, sum(sizes), 4)[0].contiguous()
    assert torch.equal(O.fps_packed(lat, off, noff), R.fps_packed(lat.to(cuda), off.to(cuda), noff.to(cuda)).cpu())


def test_oracle_ballquery_equals_reference_cuda():
    xyz = _cloud(2, 1024, 5)
    q = xyz[:, :300].contiguous()
    for r, ns in ((0.2, 32), (0.4, 64), (0.05, 8)):
        assert torch.equal(O.ballquery(r, ns, xyz, q), R.ballquery(r, ns, xyz.to(cuda), q.to(cuda)).cpu())


def test_oracle_knn_equals_reference_cuda():
    for gen in (_cloud, _lattice):
        xyz = gen(2, 800, 6)
        q = xyz[:, :200].contiguous()
        assert torch.equal(O.knn_dense(9, xyz, q), R.knn_dense(9, xyz.to(cuda), q.to(cuda)).cpu())
        widx, wd2 = O.knn_heap_dense(16, xyz, q, return_dist2=True)
        gidx, gd2 = R.knn_heap_dense(16, xyz.to(cuda), q.to(cuda))
        assert torch.equal(widx, gidx.cpu()) and torch.equal(wd2, gd2.cpu())
        wd, wi = O.nn3(q, xyz)
        gd, gi = R.nn3(q.to(cuda), xyz.to(cuda))
        assert torch.equal(wi, gi.cpu()) and torch.equal(wd, gd.cpu())


def test_oracle_knn_packed_equals_reference_cuda():
    sizes = (1500, 700)
    for gen in (_cloud, _lattice):
        xyz = gen(1, sum(sizes), 7)[0].contiguous()
        off = torch.tensor(np.cumsum(sizes), dtype=torch.int32)
        widx, wd2 = O.knn_packed(12, xyz, xyz, off, off, sqrt=False)
        gidx, gd2 = R.knn_packed(12, xyz.to(cuda), xyz.to(cuda), off.to(cuda), off.to(cuda))
        assert torch.equal(widx, gidx.cpu()) and torch.equal(wd2, gd2.cpu())


# ---- (b) sm_100a kernels == reference CUDA at full size -----------------------------------------------------
def test_fps_full_size_equals_reference_cuda():
    from repsurf_b200.seg import pointops as P
    B, N = 4, 40960
    xyz = (torch.rand(B * N, 3, generator=torch.Generator().manual_seed(8)) * torch.tensor([8.0, 8.0, 3.0])).to(cuda)
    off = P.make_offsets([N * (i + 1) for i in range(B)], cuda)
    noff = P.make_offsets([N // 4 * (i + 1) for i in range(B)], cuda)
    assert torch.equal(P.furthestsampling(xyz, off, noff), R.fps_packed(xyz, off, noff))


def test_knn_full_size_equals_reference_cuda():
    from repsurf_b200.seg import pointops as P
    B, N = 2, 40960
    xyz = (torch.rand(B * N, 3, generator=torch.Generator().manual_seed(9)) * torch.tensor([8.0, 8.0, 3.0])).to(cuda)
    off = P.make_offsets([N * (i + 1) for i in range(B)], cuda)
    for k in (9, 32):
        gidx, gdist = P.knnquery(k, xyz, xyz, off, off)
        ridx, rd2 = R.knn_packed(k, xyz, xyz, off, off)
        assert torch.equal(gidx, ridx)
        assert ((gdist.view(torch.int32).long() - torch.sqrt(rd2).view(torch.int32).long()).abs() <= 1).all()


def test_cls_ops_full_size_equal_reference_cuda():
    from repsurf_b200.cls import pointops as P
    xyz = _cloud(32, 1024, 10).to(cuda)
    fidx = P.furthestsampling(xyz, 512)
    assert torch.equal(fidx, R.fps_dense(xyz, 512))
    q = torch.gather(xyz, 1, fidx.long()[..., None].expand(-1, -1, 3)).contiguous()
    assert torch.equal(P.ballquery(0.2, 32, xyz, q), R.ballquery(0.2, 32, xyz, q))
    assert torch.equal(P.knnquery(9, xyz, xyz), R.knn_dense(9, xyz, xyz))
