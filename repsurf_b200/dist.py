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
    """One flat fp32 gradient buffer per step for ONE all-reduce (the reference's DistributedDataParallel buckets,
    segmentation/tool/train.py:163-170, collapsed into a single bucket: 3.9 MB seg / 5.9 MB cls).

    Gradients are produced by autograd as usual (`zero_grad(set_to_none=True)`: the first gradient of a parameter is
    adopted, not added), then packed with one batched `cat`, reduced, and copied back with one `_foreach_copy_` -
    a handful of launches, where accumulating into pre-assigned views costs one add kernel per parameter.  With one
    process there is nothing to reduce and nothing is packed."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.flat = None

    def zero(self):
        for p in self.params:
            p.grad = None

    def allreduce_mean(self, force_pack=False):
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if not multi and not force_pack:
            return None
        grads = []
        for p in self.params:
            if p.grad is None:                       # parameter unused this step: contributes zeros like DDP
                p.grad = torch.zeros_like(p)
            grads.append(p.grad)
        self.flat = torch.cat([g.reshape(-1) for g in grads])
        if multi:
            dist.all_reduce(self.flat)
            self.flat.div_(dist.get_world_size())
        o, views = 0, []
        for g in grads:
            views.append(self.flat[o:o + g.numel()].view_as(g))
            o += g.numel()
        torch._foreach_copy_(grads, views)
        return self.flat


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
