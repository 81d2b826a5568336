"""CPU suite: the N>1 path (cloud sharding + single flat gradient all-reduce) with world_size 2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    from repsurf_b200.dist import FlatGrads, broadcast_module, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                      # different init per rank on purpose
    model = nn.Sequential(nn.Linear(6, 8), nn.ReLU(), nn.Linear(8, 3))
    broadcast_module(model)
    fg = FlatGrads(model.parameters())
    g = torch.Generator().manual_seed(7)
    clouds = torch.randn(8, 16, 6, generator=g)         # 8 "clouds" globally, identical on every rank
    lo, hi = shard_range(8, rank, world)
    fg.zero()
    model(clouds[lo:hi]).pow(2).mean().backward()
    flat = fg.allreduce_mean()
    assert all(torch.equal(p.grad.flatten(), v) for p, v in zip(fg.params, flat.split([p.numel() for p in fg.params])))
    q.put((rank, flat.clone(), torch.cat([p.data.flatten() for p in model.parameters()])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_flat_gradient_allreduce_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, g0, w0), (_, g1, w1) = res
    assert torch.equal(w0, w1)                          # broadcast made the replicas identical
    assert torch.allclose(g0, g1, rtol=0, atol=0)       # every rank holds the same averaged gradient
    # single-process reference: mean over the two shards' losses == gradient mean (equal shard sizes)
    torch.manual_seed(100)
    model = nn.Sequential(nn.Linear(6, 8), nn.ReLU(), nn.Linear(8, 3))
    g = torch.Generator().manual_seed(7)
    clouds = torch.randn(8, 16, 6, generator=g)
    (0.5 * (model(clouds[:4]).pow(2).mean() + model(clouds[4:]).pow(2).mean())).backward()
    ref = torch.cat([p.grad.flatten() for p in model.parameters()])
    assert torch.allclose(g0, ref, rtol=1e-5, atol=1e-7)


def test_shard_range_partitions_all_clouds():
    from repsurf_b200.dist import shard_range
    for world in (1, 2, 4, 8):
        got = [shard_range(64, r, world) for r in range(world)]
        assert got[0][0] == 0 and got[-1][1] == 64
        assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
