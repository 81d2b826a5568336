"""Evaluation / augmentation harness of the classification tree on the device:
  classification/modules/pointnet2_utils.py:114-124        sample()  (torch-native FPS resampling of every evaluation batch)
  classification/modules/ptaug_utils.py:24-70               transform_point_cloud, shift_point_cloud, scale_point_cloud
  classification/tool/train_cls_scanobjectnn.py:70-108      test(): single + 10-vote (random anisotropic scaling) accuracy

`sample` runs ONE kernel per batch (csrc/harness.cu) with the torch-native FPS semantics (random start drawn like the
reference, un-fused squared distance, first-maximum picks); the augmentations are the reference's elementwise formulas with
the same random streams (torch.rand on the batch's device), so seeded runs reproduce the reference's draws."""
import torch

from .. import _native as N


def sample(nsample, feature, cuda=False):
    """feature [B, C, N] (xyz in channels 0..2) -> [B, C, nsample]: FPS resampling with a random first pick per cloud.
    `cuda` is accepted for signature compatibility (the reference switches between its two FPS implementations with it; the
    evaluation loop always calls it with the default)."""
    if not feature.is_cuda:
        raise RuntimeError("repsurf_b200 has no CPU path")
    B, C, n = feature.shape
    feature = feature.contiguous().float()
    farthest = torch.randint(0, n, (B,), dtype=torch.long).to(feature.device)       # same host draw as pointnet2_utils.py:66
    idx = torch.empty(B, nsample, dtype=torch.int64, device=feature.device)
    out = torch.empty(B, C, nsample, dtype=torch.float32, device=feature.device)
    N.call("rsb_fps_native_sample", B, C, n, nsample, feature, farthest, idx, out)
    return out


def sample_with_index(nsample, feature):
    """sample() that also returns the picked indices [B, nsample] int64 (for tests / inspection)."""
    B, C, n = feature.shape
    feature = feature.contiguous().float()
    farthest = torch.randint(0, n, (B,), dtype=torch.long).to(feature.device)
    idx = torch.empty(B, nsample, dtype=torch.int64, device=feature.device)
    out = torch.empty(B, C, nsample, dtype=torch.float32, device=feature.device)
    N.call("rsb_fps_native_sample", B, C, n, nsample, feature, farthest, idx, out)
    return out, idx


def get_aug_args(args):
    if args.dataset == 'ScanObjectNN':
        return {'scale_factor': 0.5, 'shift_factor': 0.3}
    raise Exception('No such dataset')


def shift_point_cloud(batch_data, shift_range=0.2):
    """B x C x N, shifted in place by one random offset per cloud and axis (ptaug_utils.py:41-51)."""
    shifts = (torch.rand(batch_data.shape[0], 3, 1, device=batch_data.device) * 2. - 1.) * shift_range
    batch_data += shifts
    return batch_data


def scale_point_cloud(batch_data, scale_range=0.2):
    """B x C x N, scaled in place by one random factor per cloud and axis (ptaug_utils.py:58-68)."""
    scales = (torch.rand(batch_data.shape[0], 3, 1, device=batch_data.device) * 2. - 1.) * scale_range + 1.
    batch_data *= scales
    return batch_data


def transform_point_cloud(batch, args, aug_args, train=True, label=None):
    """batch: B x 3/6 x N (ptaug_utils.py:24-34)."""
    if args.aug_scale:
        batch[:, 0:3] = scale_point_cloud(batch[:, 0:3], aug_args['scale_factor'])
    if args.aug_shift:
        batch[:, 0:3] = shift_point_cloud(batch[:, 0:3], shift_range=aug_args['shift_factor'])
    if label is not None:
        return batch, label
    return batch


def test(model, loader, num_class=15, num_point=1024, num_votes=10, total_num=1):
    """Single-pass and voted accuracy (train_cls_scanobjectnn.py:70-108): every batch is resampled to num_point by FPS,
    evaluated once as is and num_votes - 1 times under a random anisotropic scaling; the votes are averaged."""
    vote_correct = 0
    sing_correct = 0
    classifier = model.eval()
    dev = next(model.parameters()).device
    with torch.no_grad():
        for points, target in loader:
            points, target = points.to(dev), target.to(dev)
            points = sample(num_point, points)
            vote_pool = torch.zeros(target.shape[0], num_class, device=dev)
            sing_pred = None
            for i in range(num_votes):
                new_points = points.clone()
                if i > 0:
                    new_points[:, :3] = scale_point_cloud(new_points[:, :3])
                pred = classifier(new_points)
                if i == 0:
                    sing_pred = pred
                vote_pool += pred
            vote_pred = vote_pool / num_votes
            sing_correct += sing_pred.max(1)[1].eq(target.long()).sum()
            vote_correct += vote_pred.max(1)[1].eq(target.long()).sum()
    return int(sing_correct) / total_num, int(vote_correct) / total_num
