"""Plain torch restatements used ONLY as checkers by the GPU tests (fp64 where the test asks for it)."""
import torch
import torch.nn.functional as F


def linear_rows(x, layer):
    w = layer.weight
    return F.linear(x, w.view(w.shape[0], -1), layer.bias)


def bn_rows(x, bn):
    """Train/eval BatchNorm over rows with the module's buffers (same side effects as calling the module)."""
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    use_batch = bn.training or bn.running_mean is None
    return F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, use_batch,
                        0.0 if bn.momentum is None else bn.momentum, bn.eps)


def _act(x, mask):
    """ReLU, or - when the kernel's own decision is given - multiplication by that 0/1 mask.  ReLU is only piecewise
    differentiable: an element whose pre-activation is within rounding of zero can be routed differently by an fp32 and
    an fp64 evaluation, which changes a whole gradient row; pinning the routing makes gradients comparable at 1e-4."""
    return F.relu(x) if mask is None else x * mask.to(x.dtype)


def sa_mlp_rows(rows, pos_channel, mod, nsample, arg=None, masks=None):
    """Channel-de-differentiated shared MLP + max-pool (classification/modules/repsurface_utils.py:233-247,
    segmentation/modules/repsurface_utils.py:217-229) over rows [G*nsample, C] -> [G, mlp[-1]].
    arg [G, C'] (optional): pool by these sample indices instead of the arg-max; masks (optional): one 0/1 tensor per
    layer replacing the ReLU decisions (both = the routing the kernel under test took)."""
    masks = masks if masks is not None else [None] * (1 + len(mod.mlp_convs))
    x = _act(bn_rows(linear_rows(rows[:, :pos_channel], mod.mlp_l0), mod.bn_l0)
             + bn_rows(linear_rows(rows[:, pos_channel:], mod.mlp_f0), mod.bn_f0), masks[0])
    for lin, bn, m in zip(mod.mlp_convs, mod.mlp_bns, masks[1:]):
        x = _act(bn_rows(linear_rows(x, lin), bn), m)
    x = x.view(-1, nsample, x.shape[-1])
    if arg is None:
        return x.max(dim=1)[0]
    return x.gather(1, arg.long().unsqueeze(1)).squeeze(1)


def kernel_relu_masks(grad_fn):
    """The ReLU decisions of repsurf_b200.tc._FusedSAMLP, recomputed exactly as its kernels do: z = fma(Y, sc, sh) per
    layer in fp32 from the stored pre-BatchNorm outputs (first layer: the sum of the two BatchNorm halves)."""
    Ys, coefs = grad_fn.saved[3], grad_fn.saved[4]
    out = []
    for l, (Y, co) in enumerate(zip(Ys, coefs)):
        sc, sh = co[0].double(), co[1].double()
        z = (Y.double() * sc + sh).float()            # one rounding of the exact product-sum, like fmaf
        if l == 0:
            C0 = Y.shape[1] // 2
            z = z[:, :C0] + z[:, C0:]
        out.append(z > 0)
    return out
