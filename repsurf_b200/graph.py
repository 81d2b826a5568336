"""A training step of a STATIC-shape batch captured once in a CUDA graph and replayed.

Both training steps of the path issue hundreds of kernels of a few microseconds each (dense classification ~300, packed
segmentation ~370 per step).  Issued eagerly the step is at the mercy of the host: classification needs longer to issue
(~7 ms) than to run (~6 ms), and the segmentation step (19.6 ms of GPU work) was measured at 25.3 ms on a GPU box whose host
issued launches half as fast as the others.  Every launch of this package goes to torch's current stream through the C-ABI
and no entry point of either path synchronises or reads device memory on the host, so forward + backward (+ optimizer step)
capture into one graph:
  * dense classification (`classification/tool/train_cls_scanobjectnn.py:212-234` is the loop it stands for): B x N fixed;
  * packed segmentation (`segmentation/tool/train.py:280-290`): the per-cloud sizes (offsets) must be the captured ones - the
    host turns them into launch plans (sector quotas, FPS cluster plans, grid sizes) while capturing.  The reference's loader
    crops every training cloud above `voxel_max` points to exactly `voxel_max` (`segmentation/util/data_util.py:46-48`), so
    batches of large scenes do have fixed offsets; a batch with other offsets needs its own capture (or the eager step).  The geometry plan's side streams fork from and join the capturing stream, so the
    overlap of sampling / neighbour search with the GEMMs is part of the graph.
With more than one process the gradient mean has to leave between backward and the optimizer: then forward + backward are
the graph and `after_backward` (the all-reduce) and the optimizer step are issued eagerly after every replay.
"""
import torch


class GraphedTrainStep:
    """step = GraphedTrainStep(model, criterion, optimizer, example_inputs, example_target)
    loss = step(inputs, target)          # copies into the static buffers, replays; `loss` is a static 0-dim tensor

    forward(model, static_inputs) -> network output; default `model(*static_inputs)`.
    fixed: indices of inputs that are never copied on a call because launch plans were derived from their VALUES while
        capturing (the offsets of the packed layout); a call checks `same_fixed(i, new)` (default: same tensor) instead.
    optimizer_in_graph=False: the graph ends after backward; each call then runs `after_backward()` (gradient all-reduce) and
        `optimizer.step()` eagerly on the gradients the replay has written.

    Same arithmetic as the eager step: the same kernels in the same order with the same launch plans.  Random draws inside
    the step (dropout; the umbrella's random flip, drawn on the device while capturing) come from torch's graph-aware CUDA
    generator, so each replay sees fresh values."""

    def __init__(self, model, criterion, optimizer, example_inputs, example_target, warmup=3, forward=None, fixed=(),
                 same_fixed=None, optimizer_in_graph=True, after_backward=None, prepare_static=None,
                 capture_error_mode="global"):
        dev = example_target.device
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.forward = forward if forward is not None else (lambda m, inp: m(*inp))
        self.fixed = set(fixed)
        self.same_fixed = same_fixed if same_fixed is not None else (lambda i, t: t is self.static_in[i])
        self.optimizer_in_graph, self.after_backward = optimizer_in_graph, after_backward
        self.static_in = [t.clone() for t in example_inputs]
        self.static_tgt = example_target.clone()
        if prepare_static is not None:
            prepare_static(self.static_in)            # e.g. register the host mirror of the static offset tensor
        self.params = [p for g in optimizer.param_groups for p in g["params"]]
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                     # lazily created state (momentum buffers, cached constants) first
            for _ in range(warmup):
                self._body()
                if not optimizer_in_graph:
                    self._tail()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import _native
        before = _native.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        # "thread_local" when a process group is alive: NCCL's watchdog thread polls the events of earlier collectives, which
        # the "global" mode forbids to EVERY thread for as long as the capture lasts
        with torch.cuda.graph(self.graph, capture_error_mode=capture_error_mode):
            self.loss = self._body()
        self.launches_per_step = _native.launch_count() - before      # C-ABI launches recorded in the graph
        # the gradient tensors every replay writes (an eager backward in between re-points p.grad elsewhere)
        self.static_grads = [p.grad for p in self.params]

    def _body(self):
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.criterion(self.forward(self.model, self.static_in), self.static_tgt)
        loss.backward()
        if self.optimizer_in_graph:
            self.optimizer.step()
        return loss

    def _tail(self):
        if self.after_backward is not None:
            self.after_backward()
        self.optimizer.step()

    def __call__(self, inputs, target):
        for i, (dst, src) in enumerate(zip(self.static_in, inputs)):
            if i in self.fixed:
                if not self.same_fixed(i, src):
                    raise RuntimeError(f"input {i} differs from the one this step was captured for (launch plans were derived "
                                       "from its values): capture a new GraphedTrainStep for it or run the eager step")
                continue
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        if self.static_tgt.data_ptr() != target.data_ptr():
            self.static_tgt.copy_(target, non_blocking=True)
        self.graph.replay()
        if not self.optimizer_in_graph:
            for p, g in zip(self.params, self.static_grads):
                p.grad = g
            self._tail()
        return self.loss


def graphed_seg_step(model, criterion, optimizer, coord, feat, offset, target, optimizer_in_graph=True, after_backward=None,
                     warmup=3, capture_error_mode="global"):
    """GraphedTrainStep for the packed segmentation model: inputs (coord [n,3], feat [n,C], offset int32 [B]), offsets fixed.
    A call `step([coord, feat, offset], target)` accepts any offset tensor with the captured VALUES (compared through the
    host mirror of `seg.pointops.host_offsets`: no device read for registered / cached offsets)."""
    from .seg import pointops as PS
    values = tuple(PS.host_offsets(offset))

    def prepare(static_in):
        PS.register_offsets(static_in[2], values)

    def same(i, t):
        return tuple(PS.host_offsets(t)) == values

    return GraphedTrainStep(model, criterion, optimizer, [coord, feat, offset], target, warmup=warmup,
                            forward=lambda m, inp: m([inp[0], inp[1], inp[2]]), fixed=(2,), same_fixed=same,
                            optimizer_in_graph=optimizer_in_graph, after_backward=after_backward, prepare_static=prepare,
                            capture_error_mode=capture_error_mode)
