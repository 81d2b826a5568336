"""TEST INFRASTRUCTURE: numpy restatement of the reference's segmentation input pipeline with its two implementation-defined
points pinned (stable argsort; float32 `coord / voxel_size`), used to check repsurf_b200/seg/datapath.py.
Follows segmentation/modules/voxelize_utils.py:4-58 and segmentation/util/data_util.py:28-73."""
import numpy as np


def fnv_hash_vec(arr):
    arr = arr.astype(np.uint64)
    h = np.uint64(14695981039346656037) * np.ones(arr.shape[0], dtype=np.uint64)
    for j in range(arr.shape[1]):
        h *= np.uint64(1099511628211)
        h = np.bitwise_xor(h, arr[:, j])
    return h


def voxelize(coord, voxel_size=0.05, mode=0):
    discrete = np.floor(coord.astype(np.float32) / np.float32(voxel_size))
    key = fnv_hash_vec(discrete)
    idx_sort = np.argsort(key, kind='stable')
    _, count = np.unique(key[idx_sort], return_counts=True)
    if mode == 0:
        idx_select = np.cumsum(np.insert(count, 0, 0)[0:-1]) + np.random.randint(0, count.max(), count.size) % count
        return idx_sort[idx_select]
    return idx_sort, count


def data_prepare(coord, feat, label, voxel_size, voxel_max, split='train', shuffle_index=True):
    if voxel_size:
        u = voxelize(coord - np.min(coord, 0), voxel_size)
        coord, feat, label = coord[u], feat[u], label[u]
    if split != 'val' and voxel_max and coord.shape[0] > voxel_max:
        init_idx = np.random.randint(coord.shape[0]) if 'train' in split else coord.shape[0] // 2
        crop = np.argsort(np.sum(np.square(coord - coord[init_idx]), 1), kind='stable')[:voxel_max]
        coord, feat, label = coord[crop], feat[crop], label[crop]
    if shuffle_index:
        s = np.arange(coord.shape[0])
        np.random.shuffle(s)
        coord, feat, label = coord[s], feat[s], label[s]
    coord = coord - np.mean(coord, 0)
    feat = feat / 255.
    return coord, feat, label
