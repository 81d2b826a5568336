"""Whole-scene inference on the device: the evaluation loop of segmentation/tool/test_s3dis.py:186-238 and
the label median filter of segmentation/util/utils.py:235-245, on the sm_100a operators.

    votes = SceneVotes(n_points, num_class, device)
    for crop in crops:                                   # overlapping crops of up to voxel_max points (test_s3dis.py:131-159)
        votes.add(model([coord, feat, offset]), idx)     # softmax + scatter of the votes, one kernel
    pred = votes.decide()                                # argmax(pred / pred_count)
    pred = pc_median_filter_gpu(coord_all, pred, 32)     # kNN(32) over the whole scene as ONE ~10^6-point segment

The kNN of the filter is the same exact uniform-grid search the training path uses (csrc/knn_grid.cu); a scene is a
single segment far beyond the sizes where the reference's one-thread-per-query scan is practical."""
import torch

from .. import _native as N
from . import pointops as P


class SceneVotes:
    """pred [n, num_class] / pred_count [n] of test_s3dis.py:199-200, accumulated on the device."""

    def __init__(self, n_points, num_class, device):
        self.pred = torch.zeros(n_points, num_class, dtype=torch.float32, device=device)
        self.count = torch.zeros(n_points, dtype=torch.float32, device=device)
        self.num_class = num_class

    def add(self, logits, idx):
        """logits [rows, num_class] (model output of a batch of crops), idx [rows] = scene row of every crop point.
        pred[idx] += softmax(logits), pred_count[idx] += 1 (test_s3dis.py:208-213).  Rows that vote for the same point -
        crops of one batch overlap - ALL count (atomics); the reference's `pred[idx_part, :] += pred_part` keeps an
        arbitrary one of them in that case."""
        assert logits.is_cuda and logits.dtype == torch.float32 and logits.shape[1] == self.num_class and logits.stride(1) == 1
        idx = idx.to(device=logits.device, dtype=torch.int64).contiguous()
        # the classifier output is a [rows, num_class] view of a row-padded buffer: addressed by (pointer, row pitch)
        N.call("rsb_scene_vote", logits.shape[0], self.num_class, logits.data_ptr(), logits.stride(0), idx, self.pred, self.count)

    def decide(self):
        """argmax over classes of pred / pred_count -> int32 [n] (test_s3dis.py:217)."""
        out = torch.empty(self.pred.shape[0], dtype=torch.int32, device=self.pred.device)
        N.call("rsb_scene_decide", self.pred.shape[0], self.num_class, self.pred, self.count, out)
        return out


def label_median(coord, label, group_size=16, offset=None):
    """device version of pc_median_filter_gpu: int32 labels [N] -> int32 [N] (stays on the device)."""
    assert coord.is_contiguous()
    n = coord.shape[0]
    if offset is None:
        offset = P.make_offsets([n], coord.device)
    idx, _ = P.knnquery(group_size, coord, coord, offset, offset)
    out = torch.empty(n, dtype=torch.int32, device=coord.device)
    N.call("rsb_label_median", n, int(group_size), idx, label.to(torch.int32).contiguous(), out)
    return out


def pc_median_filter_gpu(coord, label, group_size=16):
    """Drop-in for segmentation/util/utils.py:235-245: coord [N,3], label [N] (device tensors) -> numpy [N] of the median
    label among each point's `group_size` nearest neighbours (itself included); the whole cloud is one segment."""
    return label_median(coord, label, group_size).cpu().numpy()


def scene_inference(model, coord_parts, feat_parts, idx_parts, n_points, num_class, batch_size=1, filter_k=None, coord_all=None):
    """The loop of test_s3dis.py:199-224 for prepared crops: *_parts are lists of per-crop tensors (device or host), idx_parts
    the scene rows of each crop.  Returns int32 labels [n_points] on the device (median-filtered when filter_k is set)."""
    dev = next(model.parameters()).device
    votes = SceneVotes(n_points, num_class, dev)
    model.eval()
    for s in range(0, len(idx_parts), batch_size):
        cs = [c.to(dev, non_blocking=True) for c in coord_parts[s:s + batch_size]]
        fs = [f.to(dev, non_blocking=True) for f in feat_parts[s:s + batch_size]]
        ids = torch.cat([i.to(dev, non_blocking=True).long() for i in idx_parts[s:s + batch_size]])
        sizes, run = [], 0
        for c in cs:
            run += c.shape[0]
            sizes.append(run)
        offset = P.make_offsets(sizes, dev)
        with torch.no_grad():
            votes.add(model([torch.cat(cs).contiguous(), torch.cat(fs).contiguous(), offset]), ids)
    label = votes.decide()
    if filter_k:
        label = label_median(coord_all.to(dev).contiguous(), label, filter_k)
    return label
