"""Input pipeline of the segmentation tree on the device: the per-cloud preparation that feeds the packed
layout, for clouds that are already resident in HBM.
  segmentation/modules/voxelize_utils.py:4-58       fnv_hash_vec, voxelize (grid subsampling, one random point per voxel)
  segmentation/util/data_util.py:28-73              data_prepare (grid sampling, nearest crop to voxel_max, shuffle, centring,
                                                    colour normalisation)
  segmentation/util/data_util.py:15-25              collate_fn (concatenate clouds, cumulative offsets)

Random draws come from numpy's global generator in the reference's order and with its arguments (np.random.randint /
np.random.shuffle), so a seeded run makes the same choices.  Two things are implementation-defined in the reference and pinned
here: np.argsort's order among EQUAL keys (numpy's default sort is not stable; here points of a voxel are in ascending index
order, i.e. a stable sort) and the dtype of `coord / np.array(voxel_size)` (float32 under the NumPy 1.x the reference was
written for; NumPy 2 promotes it to float64) - the fp32 reading is used.  Host synchronisation: one read-back per cloud (the
number of occupied voxels sizes everything after it and is an argument of the reference's own random draw)."""
import numpy as np
import torch

from .. import _native as N


def _keys(coord, voxel_size):
    n = coord.shape[0]
    cmin = torch.full((3,), float("inf"), dtype=torch.float32, device=coord.device)
    N.call("rsb_coord_min", n, coord, cmin)
    key = torch.empty(n, dtype=torch.int64, device=coord.device)
    N.call("rsb_voxel_keys", n, coord, cmin, float(voxel_size), key)
    return key


def fnv_hash_vec(coord, voxel_size):
    """FNV64-1A keys of floor((coord - coord.min(0)) / voxel_size) as the reference computes them, returned as the uint64
    values in an int64 tensor (bit pattern)."""
    return _keys(coord.contiguous(), voxel_size) ^ torch.tensor(-2 ** 63, dtype=torch.int64, device=coord.device)


def ravel_hash_vec(coord, voxel_size):
    """Keys of voxelize_utils.py:20-35 for floor(coord / voxel_size): the row-major rank of a point's voxel inside the
    bounding box of the occupied voxels, int64 [n] (device)."""
    coord = coord.contiguous()
    n = coord.shape[0]
    cmin = torch.full((3,), float("inf"), dtype=torch.float32, device=coord.device)
    cmax = torch.full((3,), float("-inf"), dtype=torch.float32, device=coord.device)
    N.call("rsb_coord_min", n, coord, cmin)
    N.call("rsb_coord_max", n, coord, cmax)
    key = torch.empty(n, dtype=torch.int64, device=coord.device)
    N.call("rsb_voxel_keys_ravel", n, coord, cmin, cmax, float(voxel_size), key)
    return key


def voxel_runs(coord, voxel_size, hash_type='fnv'):
    """Sort the points by voxel key: (idx_sort int64 [n], count int32 [n_voxels], start int32 [>= n_voxels] = first sorted
    position of every voxel, largest count).  One read-back (the number of occupied voxels sizes what follows)."""
    coord = coord.contiguous()
    n = coord.shape[0]
    dev = coord.device
    key = ravel_hash_vec(coord, voxel_size) if hash_type == 'ravel' else _keys(coord, voxel_size)
    key_sort, idx_sort = torch.sort(key, stable=True)            # cub radix sort; stable = ascending index inside a voxel
    scratch = torch.empty((n + 1023) // 1024, dtype=torch.int32, device=dev)
    start = torch.empty(n, dtype=torch.int32, device=dev)
    scalars = torch.zeros(2, dtype=torch.int32, device=dev)
    N.call("rsb_voxel_runs", n, key_sort, scratch, start, scalars[0:1])
    n_vox = int(scalars[0])                                       # read-back: sizes everything that follows
    count = torch.empty(n_vox, dtype=torch.int32, device=dev)
    N.call("rsb_voxel_counts", n_vox, n, start, count, scalars[1:2])
    return idx_sort, count, start, int(scalars[1])


def voxelize(coord, voxel_size=0.05, hash_type='fnv', mode=0):
    """coord [n,3] float32 (device).  mode 0 (train): int64 [n_voxels] = one randomly chosen point per occupied voxel, voxels
    in ascending key order (voxelize_utils.py:53-56).  mode 1 (val): (idx_sort int64 [n], count int32 [n_voxels]).
    hash_type 'fnv' (the pipeline's default) or 'ravel', as in the reference."""
    idx_sort, count, start, cmax = voxel_runs(coord, voxel_size, hash_type)
    if mode != 0:
        return idx_sort, count
    n_vox = count.shape[0]
    draw = torch.from_numpy(np.random.randint(0, cmax, n_vox)).to(coord.device)          # same draw as voxelize_utils.py:53
    out = torch.empty(n_vox, dtype=torch.int64, device=coord.device)
    N.call("rsb_voxel_pick", n_vox, start, count, draw, idx_sort, out)
    return out


def nearest_crop(coord, init_idx, voxel_max):
    """indices of the voxel_max points closest to coord[init_idx], ascending distance (data_util.py:46-48)."""
    coord = coord.contiguous()
    d = torch.empty(coord.shape[0], dtype=torch.float32, device=coord.device)
    N.call("rsb_seed_distance", coord.shape[0], coord, int(init_idx), d)
    return torch.sort(d, stable=True)[1][:voxel_max]


def data_prepare(coord, feat, label, voxel_size=0.04, voxel_max=80000, split='train', data_norm='mean', dataset='S3DIS',
                 rgb_mean=None, rgb_std=None, shuffle_index=True):
    """Device version of data_util.data_prepare for one cloud (coord [n,3], feat [n,3], label [n] | None on the device; the
    coordinate / colour augmentations of the reference's transform objects are applied by the caller beforehand)."""
    if voxel_size:
        sel = voxelize(coord - coord.min(0)[0], voxel_size)
        coord, feat = coord[sel], feat[sel]
        label = label[sel] if label is not None else None
    if split != 'val' and voxel_max and coord.shape[0] > voxel_max:
        init_idx = np.random.randint(coord.shape[0]) if 'train' in split else coord.shape[0] // 2
        crop = nearest_crop(coord, init_idx, voxel_max)
        coord, feat = coord[crop], feat[crop]
        label = label[crop] if label is not None else None
    if shuffle_index:
        shuf = np.arange(coord.shape[0])
        np.random.shuffle(shuf)
        shuf = torch.from_numpy(shuf).to(coord.device)
        coord, feat = coord[shuf], feat[shuf]
        label = label[shuf] if label is not None else None
    if data_norm == 'mean':
        coord = coord - coord.double().mean(0).float()
    elif data_norm == 'min':
        coord = coord - coord.min(0)[0]
    if dataset in ('S3DIS', 'ScanNet'):
        feat = feat / torch.tensor(255., device=feat.device)      # a tensor divisor: true division (a python scalar would
                                                                    # become a multiplication by 1/255 on CUDA, one ulp off numpy)
        if rgb_mean is not None and rgb_std is not None:
            feat = (feat - torch.as_tensor(rgb_mean, device=feat.device, dtype=feat.dtype)) / torch.as_tensor(rgb_std, device=feat.device, dtype=feat.dtype)
    return coord.contiguous().float(), feat.contiguous().float(), (label.long() if label is not None else None)


def collate_fn(batch):
    """[(coord, feat, label)] -> (coord [sum n,3], feat, label | None, offset int32 [B] cumulative ends), data_util.py:15-25;
    the offsets come with their host mirror registered, so the model's first layer does not read them back."""
    from . import pointops as P
    coord, feat, label = list(zip(*batch))
    sizes, run = [], 0
    for c in coord:
        run += c.shape[0]
        sizes.append(run)
    offset = P.make_offsets(sizes, coord[0].device)
    return torch.cat(coord), torch.cat(feat), (torch.cat(label) if label[0] is not None else None), offset
