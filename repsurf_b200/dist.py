"""Data-parallel plumbing for the RepSurf-U path: shard by cloud, one process per GPU, ONE gradient all-reduce.

The reference trains segmentation with mp.spawn + DistributedDataParallel over NCCL
(segmentation/tool/train.py:115,137,145,482): batch split across ranks, per-GPU BatchNorm statistics
(--sync_bn off by default, :47-48), gradient mean over ranks.  The hot path has no other exchange, so the whole
multi-GPU story is: every rank owns B/G clouds and, after backward, the 3.9 MB (seg) / 5.9 MB (cls) of fp32
gradients are averaged.  Instead of DDP's bucketed hooks the gradients live in ONE flat buffer that autograd
accumulates into, and a single all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests) handles it.
"""
import torch
import torch.distributed as dist


class FlatGrads:
    """Views every parameter's .grad into one contiguous fp32 buffer."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), device=dev, dtype=torch.float32)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def zero(self):
        self.flat.zero_()

    def allreduce_mean(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat)
            self.flat.div_(dist.get_world_size())


def broadcast_module(module, src=0):
    """Same initial weights / buffers on every rank (what DDP does at construction)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for p in module.parameters():
            dist.broadcast(p.data, src)
        for b in module.buffers():
            dist.broadcast(b, src)


def shard_range(n_clouds_global, rank, world):
    """Contiguous block of clouds owned by `rank` (reference: batch_size // ngpus per process, train.py:137)."""
    per = n_clouds_global // world
    return rank * per, (rank + 1) * per
