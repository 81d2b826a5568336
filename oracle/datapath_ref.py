"""TEST INFRASTRUCTURE: numpy restatement of the reference's segmentation input pipeline with its two implementation-defined
points pinned (stable argsort; float32 `coord / voxel_size`), used to check repsurf_b200/seg/datapath.py.
Follows segmentation/modules/voxelize_utils.py:4-58, segmentation/util/data_util.py:28-73 and, for whole-scene inference,
segmentation/tool/test_s3dis.py:114-159 (data_load's voxel parts, data_process's covering crops)."""
import numpy as np


def fnv_hash_vec(arr):
    arr = arr.astype(np.uint64)
    h = np.uint64(14695981039346656037) * np.ones(arr.shape[0], dtype=np.uint64)
    for j in range(arr.shape[1]):
        h *= np.uint64(1099511628211)
        h = np.bitwise_xor(h, arr[:, j])
    return h


def ravel_hash_vec(arr):
    """voxelize_utils.py:20-35: row-major rank inside the occupied bounding box"""
    arr = arr - arr.min(0)
    arr = arr.astype(np.uint64)
    ext = arr.max(0).astype(np.uint64) + np.uint64(1)
    keys = np.zeros(arr.shape[0], dtype=np.uint64)
    for j in range(arr.shape[1] - 1):
        keys += arr[:, j]
        keys *= ext[j + 1]
    keys += arr[:, -1]
    return keys


def voxelize(coord, voxel_size=0.05, hash_type='fnv', mode=0):
    discrete = np.floor(coord.astype(np.float32) / np.float32(voxel_size))
    key = ravel_hash_vec(discrete) if hash_type == 'ravel' else fnv_hash_vec(discrete)
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


def scene_parts(coord, voxel_size):
    """test_s3dis.py:123-129 (data_load): part i takes member i % count of every voxel; every point is in at least one part"""
    idx_sort, count = voxelize(coord - np.min(coord, 0), voxel_size, mode=1)
    first = np.cumsum(np.insert(count, 0, 0)[0:-1])
    return [idx_sort[first + i % count] for i in range(count.max())]


def crop_plan(coord_part, voxel_max):
    """test_s3dis.py:143-158: nearest crops of voxel_max points around the lowest-priority point until all are covered.
    Returns the crops as rows of coord_part.  fp32 coordinates (the .npy dtype decides in the reference); stable argsort."""
    n = coord_part.shape[0]
    coord_p, covered, crops = np.random.rand(n) * 1e-3, np.zeros(n, dtype=bool), []
    while not covered.all():
        init_idx = np.argmin(coord_p)
        dist = np.sum(np.power(coord_part - coord_part[init_idx], 2), 1)
        idx_crop = np.argsort(dist, kind='stable')[:voxel_max]
        d = dist[idx_crop]
        coord_p[idx_crop] += np.square(1 - d / np.max(d))
        covered[idx_crop] = True
        crops.append(idx_crop)
    return crops


def data_process(coord, feat, idx_data, voxel_max, color_mean=None, color_std=None):
    """test_s3dis.py:131-159 with data_norm='mean': (idx_list, coord_list, feat_list, offset_list)"""
    def norm(c, f):
        c = c - np.mean(c, 0)
        f = f / 255.
        if color_mean is not None and color_std is not None:
            f = (f - color_mean) / color_std
        return c, f
    idx_list, coord_list, feat_list, offset_list = [], [], [], []
    for idx_part in idx_data:
        coord_part, feat_part = coord[idx_part], feat[idx_part]
        crops = crop_plan(coord_part, voxel_max) if (voxel_max and coord_part.shape[0] > voxel_max) else [np.arange(idx_part.shape[0])]
        for c in crops:
            cs, fs = norm(coord_part[c], feat_part[c])
            idx_list.append(idx_part[c]), coord_list.append(cs), feat_list.append(fs), offset_list.append(c.size)
    return idx_list, coord_list, feat_list, offset_list
