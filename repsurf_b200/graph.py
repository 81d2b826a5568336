"""A training step of a STATIC-shape model captured once in a CUDA graph and replayed.

The dense classification path (B x N fixed, `classification/tool/train_cls_scanobjectnn.py:180-200` is the loop it stands
for) issues about 300 kernels of a few microseconds each per step; eager, the host needs longer to issue them (~7 ms) than
the GPU needs to run them (~6 ms).  Every launch of this package goes to torch's current stream through the C-ABI and no
entry point of the dense path synchronises or reads device memory on the host, so forward + backward + optimizer step
capture into one graph.  The packed segmentation path has data-dependent shapes (offsets, per-level sample counts) that the
host turns into launch plans, and it is GPU-bound anyway; it is not captured.
"""
import torch


class GraphedTrainStep:
    """step = GraphedTrainStep(model, criterion, optimizer, example_inputs, example_target)
    loss = step(inputs, target)          # copies into the static buffers, replays; `loss` is a static 0-dim tensor

    Same arithmetic as the eager step: the same kernels in the same order with the same launch plans.  Random draws inside
    the step (dropout; the umbrella's random flip, drawn on the device while capturing) come from torch's graph-aware CUDA
    generator, so each replay sees fresh values."""

    def __init__(self, model, criterion, optimizer, example_inputs, example_target, warmup=3):
        dev = example_target.device
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.static_in = [t.clone() for t in example_inputs]
        self.static_tgt = example_target.clone()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                     # lazily created state (momentum buffers, cached constants) first
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import _native
        before = _native.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._body()
        self.launches_per_step = _native.launch_count() - before      # C-ABI launches recorded in the graph

    def _body(self):
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.criterion(self.model(*self.static_in), self.static_tgt)
        loss.backward()
        self.optimizer.step()
        return loss

    def __call__(self, inputs, target):
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        if self.static_tgt.data_ptr() != target.data_ptr():
            self.static_tgt.copy_(target, non_blocking=True)
        self.graph.replay()
        return self.loss
