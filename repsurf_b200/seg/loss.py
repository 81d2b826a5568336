"""The criterion of the segmentation training step on the device: drop-in for `nn.CrossEntropyLoss(ignore_index=...)` as
segmentation/tool/train.py builds it (util/utils.py get_loss: weight=None, mean reduction), as ONE kernel over the logits
(csrc/scene.cu cross_entropy_kernel) - forward value and gradient in the same pass."""
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _native as N


class _CrossEntropy(Function):
    @staticmethod
    def forward(ctx, logits, target, ignore_index):
        assert logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 2 and logits.stride(1) == 1
        rows, nc = logits.shape
        target = target.contiguous()
        assert target.dtype == torch.int64 and target.shape[0] == rows
        ldg = (nc + 3) // 4 * 4                              # row-padded like the classifier output: the head's GEMMs take it as is
        gbuf = torch.empty(rows, ldg, dtype=torch.float32, device=logits.device)
        acc = torch.zeros(2, dtype=torch.float64, device=logits.device)
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        # the logits may be a [rows, nc] view of a row-padded buffer: addressed by (pointer, row pitch)
        N.call("rsb_cross_entropy_forward", rows, nc, logits.data_ptr(), logits.stride(0), target, int(ignore_index), gbuf, ldg, acc, loss)
        ctx.saved = (gbuf, acc, nc)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        gbuf, acc, nc = ctx.saved
        N.call("rsb_cross_entropy_backward", gbuf.numel(), gbuf, acc, g.contiguous().float())
        return gbuf[:, :nc], None, None


class CrossEntropyLoss(nn.Module):
    """nn.CrossEntropyLoss(weight=None, ignore_index=...) with mean reduction over [rows, classes] logits and int64 targets."""

    def __init__(self, weight=None, ignore_index=-100):
        super().__init__()
        if weight is not None:
            raise NotImplementedError("class weights are not used on the RepSurf path")
        self.ignore_index = ignore_index

    def forward(self, logits, target):
        return _CrossEntropy.apply(logits, target, self.ignore_index)
