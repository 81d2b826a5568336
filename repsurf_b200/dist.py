"""Data-parallel plumbing for the RepSurf-U path: shard by cloud, one process per GPU, ONE gradient all-reduce.

The reference trains segmentation with mp.spawn + DistributedDataParallel over NCCL
(segmentation/tool/train.py:115,137,145,482): batch split across ranks, per-GPU BatchNorm statistics
(--sync_bn off by default, :47-48), gradient mean over ranks.  The hot path has no other exchange, so the whole
multi-GPU story is: every rank owns B/G clouds and, after backward, the 3.9 MB (seg) / 5.9 MB (cls) of fp32
gradients are averaged.  The gradients are packed into ONE flat buffer and reduced in two pieces that leave while backward is still
running (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


class FlatGrads:
    """One flat fp32 gradient buffer per step (the reference's DistributedDataParallel buckets,
    segmentation/tool/train.py:163-170; 3.9 MB seg / 5.9 MB cls), reduced in `buckets` contiguous pieces that leave while
    backward is still running.

    Gradients are produced by autograd as usual (`zero_grad(set_to_none=True)`: the first gradient of a parameter is
    adopted, not added).  Parameters are cut, in `parameters()` order, into runs of about equal bytes; backward produces the
    LAST run first (heads and propagation layers), and the moment every gradient of a run exists a post-accumulate hook packs
    the run with one `cat` into its slice of the flat buffer and starts an asynchronous all-reduce on it (NCCL's own stream
    on the GPU box, gloo's worker thread in the CPU tests), so only the first run's reduce - the encoder's first levels,
    whose gradients appear last - is exposed after backward.  `allreduce_mean` waits, divides and copies the averages back
    with one `_foreach_copy_`.  With one process there is nothing to reduce and nothing is packed or hooked.

    One backward per `zero()` is the fast path; a second backward before `allreduce_mean` (gradient accumulation) is noticed
    and answered with one plain reduce of everything at the end."""

    def __init__(self, params, buckets=2):
        self.params = [p for p in params if p.requires_grad]
        self.flat = None
        sizes = [p.numel() for p in self.params]
        self._off = [0]
        for n in sizes:
            self._off.append(self._off[-1] + n)
        total = max(self._off[-1], 1)
        buckets = max(1, min(buckets, len(self.params)))
        self._bucket = [min(buckets - 1, (self._off[i] + sizes[i] // 2) * buckets // total) for i in range(len(self.params))]
        self._members = [[i for i, b in enumerate(self._bucket) if b == k] for k in range(buckets)]
        self._members = [m for m in self._members if m]
        for k, m in enumerate(self._members):
            for i in m:
                self._bucket[i] = k
        self._hooked = False
        self._reset()

    @staticmethod
    def _multi():
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _reset(self):
        self._left = [len(m) for m in self._members]
        self._handles = [None] * len(self._members)
        self._dirty = False

    def _grads(self, members):
        out = []
        for i in members:
            p = self.params[i]
            if p.grad is None:                       # parameter unused this step: contributes zeros like DDP
                p.grad = torch.zeros_like(p)
            out.append(p.grad)
        return out

    def _slice(self, members):
        return self.flat[self._off[members[0]]:self._off[members[-1] + 1]]

    def _ensure_flat(self):
        p0 = self.params[0]
        if self.flat is None or self.flat.device != p0.device:
            self.flat = torch.empty(self._off[-1], dtype=p0.dtype, device=p0.device)

    def _start(self, k):
        members = self._members[k]
        self._ensure_flat()
        piece = self._slice(members)
        torch.cat([g.reshape(-1) for g in self._grads(members)], out=piece)
        self._handles[k] = dist.all_reduce(piece, async_op=True)

    def _hook(self, i):
        def fire(_param):
            k = self._bucket[i]
            if self._handles[k] is not None:         # a second backward after this run already left: redo everything at the end
                self._dirty = True
                return
            self._left[k] -= 1
            if self._left[k] == 0:
                self._start(k)
        return fire

    def zero(self):
        for p in self.params:
            p.grad = None
        self._reset()
        if not self._hooked and self._multi():
            for i, p in enumerate(self.params):
                p.register_post_accumulate_grad_hook(self._hook(i))
            self._hooked = True

    def allreduce_mean(self, force_pack=False):
        multi = self._multi()
        if not multi and not force_pack:
            return None
        self._ensure_flat()
        everything = list(range(len(self.params)))
        if multi:
            for k in range(len(self._members)):
                if self._handles[k] is None and not self._dirty:      # parameters without a gradient this step, or no hooks yet
                    self._start(k)
            for h in self._handles:
                if h is not None:
                    h.wait()
            if self._dirty:
                torch.cat([g.reshape(-1) for g in self._grads(everything)], out=self.flat)
                dist.all_reduce(self.flat)
            self.flat.div_(dist.get_world_size())
        else:
            torch.cat([g.reshape(-1) for g in self._grads(everything)], out=self.flat)
        grads = self._grads(everything)
        views = [self.flat[self._off[i]:self._off[i + 1]].view_as(g) for i, g in enumerate(grads)]
        torch._foreach_copy_(grads, views)
        self._reset()
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
