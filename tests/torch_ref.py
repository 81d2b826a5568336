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


def sa_mlp_rows(rows, pos_channel, mod, nsample, arg=None):
    """Channel-de-differentiated shared MLP + max-pool (classification/modules/repsurface_utils.py:233-247,
    segmentation/modules/repsurface_utils.py:217-229) over rows [G*nsample, C] -> [G, mlp[-1]].
    arg [G, C'] (optional): pool by these sample indices instead of the arg-max (a max-pool with the routing fixed)."""
    x = F.relu(bn_rows(linear_rows(rows[:, :pos_channel], mod.mlp_l0), mod.bn_l0)
               + bn_rows(linear_rows(rows[:, pos_channel:], mod.mlp_f0), mod.bn_f0))
    for lin, bn in zip(mod.mlp_convs, mod.mlp_bns):
        x = F.relu(bn_rows(linear_rows(x, lin), bn))
    x = x.view(-1, nsample, x.shape[-1])
    if arg is None:
        return x.max(dim=1)[0]
    return x.gather(1, arg.long().unsqueeze(1)).squeeze(1)
