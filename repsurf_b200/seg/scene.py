"""Whole-scene inference on the device: the evaluation loop of segmentation/tool/test_s3dis.py:186-238 and
the label median filter of segmentation/util/utils.py:235-245, on the sm_100a operators.

    idx_data = scene_parts(coord, voxel_size)            # data_load: member i of every voxel -> part i (test_s3dis.py:123-129)
    idx, coords, feats, sizes = data_process(coord, feat, idx_data, voxel_max)   # covering nearest crops (test_s3dis.py:131-159)
    votes = SceneVotes(n_points, num_class, device)
    for crop in crops:                                   # batches of crops of up to voxel_max points
        votes.add(model([coord, feat, offset]), idx)     # softmax + scatter of the votes, one kernel
    pred = votes.decide()                                # argmax(pred / pred_count)
    pred = pc_median_filter_gpu(coord_all, pred, 32)     # kNN(32) over the whole scene as ONE ~10^6-point segment

The kNN of the filter is the same exact uniform-grid search the training path uses (csrc/knn_grid.cu); a scene is a
single segment far beyond the sizes where the reference's one-thread-per-query scan is practical."""
import numpy as np
import torch

from .. import _native as N
from . import pointops as P


def scene_parts(coord, voxel_size):
    """data_load of test_s3dis.py:123-129 on the device: coord [n,3] float32 -> list of int64 row tensors; part i holds member
    i % count of every occupied voxel (so every point is in at least one part, one point per voxel in each)."""
    from . import datapath as D
    coord = coord.contiguous()
    if not voxel_size:
        return [torch.arange(coord.shape[0], dtype=torch.int64, device=coord.device)]
    idx_sort, count, start, cmax = D.voxel_runs(coord - coord.min(0)[0], voxel_size)
    n_vox = count.shape[0]
    parts = []
    for i in range(cmax):
        out = torch.empty(n_vox, dtype=torch.int64, device=coord.device)
        N.call("rsb_voxel_pick", n_vox, start, count, torch.full((n_vox,), i, dtype=torch.int64, device=coord.device), idx_sort, out)
        parts.append(out)
    return parts


def crop_plan(coord_part, voxel_max):
    """The covering loop of test_s3dis.py:143-158 for one part (coord_part [n,3] float32 on the device, n > voxel_max): repeat
    {seed = argmin of the priorities, crop = the voxel_max rows nearest to it, raise the priorities of the crop by
    (1 - d / d_max)^2} until every row has been in a crop.  Returns the crops (int64 rows of coord_part, ascending distance).
    The initial priorities are the reference's draw, np.random.rand(n) * 1e-3.  Per crop: 6 launches + one stable sort and ONE
    4-byte read-back (the loop ends when the covered count reaches n); the seed index stays on the device."""
    coord_part = coord_part.contiguous()
    n, dev = coord_part.shape[0], coord_part.device
    assert coord_part.dtype == torch.float32 and 0 < voxel_max < n
    priority = torch.from_numpy(np.random.rand(n) * 1e-3).to(dev)
    covered = torch.zeros(n, dtype=torch.int32, device=dev)
    n_covered = torch.zeros(1, dtype=torch.int32, device=dev)
    work = torch.empty(2, dtype=torch.int64, device=dev)
    dist = torch.empty(n, dtype=torch.float32, device=dev)
    crops = []
    while True:
        N.call("rsb_argmin_f64", n, priority, work)
        N.call("rsb_seed_distance_dev", n, coord_part, work[1:2], dist)
        crop = torch.sort(dist, stable=True)[1][:voxel_max].contiguous()
        N.call("rsb_crop_update", voxel_max, crop, dist, priority, covered, n_covered)
        crops.append(crop)
        if int(n_covered) >= n:
            return crops
        if len(crops) > 2 * n:       # every seed's priority rises by 1 when it is used, so 2n seeds cover any finite cloud
            raise RuntimeError("crop_plan does not converge (non-finite coordinates?)")


def input_normalize(coord, feat, data_norm='mean', color_mean=None, color_std=None):
    """test_s3dis.py:162-175 on device tensors (out of place)."""
    if data_norm == 'mean':
        coord = coord - coord.double().mean(0).float()
    elif data_norm == 'min':
        coord = coord - coord.min(0)[0]
    else:
        raise Exception('No such data norm type')
    feat = feat / torch.tensor(255., device=feat.device)
    if color_mean is not None and color_std is not None:
        feat = (feat - torch.as_tensor(color_mean, device=feat.device, dtype=feat.dtype)) / torch.as_tensor(color_std, device=feat.device, dtype=feat.dtype)
    return coord.contiguous(), feat.contiguous()


def data_process(coord, feat, idx_data, voxel_max, data_norm='mean', color_mean=None, color_std=None):
    """test_s3dis.py:131-159: parts -> crops of at most voxel_max points, normalised.  Returns (idx_list, coord_list, feat_list,
    offset_list) like the reference: scene rows, centred coordinates, scaled colours and size of every crop, in its order."""
    idx_list, coord_list, feat_list, offset_list = [], [], [], []
    for idx_part in idx_data:
        coord_part, feat_part = coord[idx_part], feat[idx_part]
        if voxel_max and coord_part.shape[0] > voxel_max:
            for crop in crop_plan(coord_part, voxel_max):
                c, f = input_normalize(coord_part[crop], feat_part[crop], data_norm, color_mean, color_std)
                idx_list.append(idx_part[crop]), coord_list.append(c), feat_list.append(f), offset_list.append(int(crop.shape[0]))
        else:
            c, f = input_normalize(coord_part, feat_part, data_norm, color_mean, color_std)
            idx_list.append(idx_part), coord_list.append(c), feat_list.append(f), offset_list.append(int(idx_part.shape[0]))
    return idx_list, coord_list, feat_list, offset_list


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


def infer_scene(model, coord, feat, num_class, voxel_size=0.04, voxel_max=80000, batch_size=12, filter_k=None,
                data_norm='mean', color_mean=None, color_std=None):
    """One scene of test_s3dis.py:186-238 end to end on the device: coord [n,3] (scene coordinates), feat [n,3] (0..255 colours)
    -> int32 labels [n].  data_load (voxel parts) -> data_process (covering crops, normalisation) -> batches of `batch_size`
    crops through the eval-mode model -> softmax votes -> decision -> optional kNN(filter_k) label median filter."""
    dev = next(model.parameters()).device
    coord = coord.to(dev, dtype=torch.float32).contiguous()
    feat = feat.to(dev, dtype=torch.float32).contiguous()
    idx_data = scene_parts(coord, voxel_size)
    idx_list, coord_list, feat_list, _ = data_process(coord, feat, idx_data, voxel_max, data_norm, color_mean, color_std)
    return scene_inference(model, coord_list, feat_list, idx_list, coord.shape[0], num_class, batch_size=batch_size,
                           filter_k=filter_k, coord_all=coord)
