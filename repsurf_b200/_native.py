"""ctypes binding of librepsurf_b200.so (the C-ABI declared in include/repsurf_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError is
raised.  torch is used for device memory and streams only (data_ptr / current stream handle).
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librepsurf_b200.so")
_lib = None

_i = ctypes.c_int
_f = ctypes.c_float
_p = ctypes.c_void_p
_l = ctypes.c_long

# name -> argtypes (the trailing cudaStream_t is appended automatically)
_SIGS = {
    "rsb_furthestsampling_dense": [_i, _i, _i, _p, _p, _p, _p],
    "rsb_gathering_forward": [_i, _i, _i, _i, _p, _p, _p],
    "rsb_gathering_backward": [_i, _i, _i, _i, _p, _p, _p],
    "rsb_ballquery": [_i, _i, _i, _f, _i, _p, _p, _p],
    "rsb_knnquery_dense": [_i, _i, _i, _i, _p, _p, _p, _p],
    "rsb_knnquery_heap_dense": [_i, _i, _i, _i, _p, _p, _p, _p],
    "rsb_grouping_forward": [_i, _i, _i, _i, _i, _p, _p, _p],
    "rsb_grouping_backward": [_i, _i, _i, _i, _i, _p, _p, _p],
    "rsb_grouping_int_forward": [_i, _i, _i, _i, _i, _p, _p, _p],
    "rsb_nearestneighbor": [_i, _i, _i, _p, _p, _p, _p],
    "rsb_interpolation_forward": [_i, _i, _i, _i, _p, _p, _p, _p],
    "rsb_interpolation_backward": [_i, _i, _i, _i, _p, _p, _p, _p],
    "rsb_furthestsampling_packed": [_i, _i, _p, _p, _p, _p, _p, _p, _p],
    "rsb_furthestsampling_packed_bounded": [_i, _i, _i, _p, _p, _p, _p, _p, _p, _p],
    "rsb_subtraction_forward": [_i, _i, _i, _p, _p, _p, _p],
    "rsb_subtraction_backward": [_i, _i, _i, _p, _p, _p, _p],
    "rsb_aggregation_forward": [_i, _i, _i, _i, _p, _p, _p, _p, _p],
    "rsb_aggregation_backward": [_i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p],
    "rsb_scene_vote": [_l, _i, _p, _i, _p, _p, _p],
    "rsb_scene_decide": [_l, _i, _p, _p, _p],
    "rsb_label_median": [_l, _i, _p, _p, _p],
    "rsb_fps_native_sample": [_i, _i, _i, _i, _p, _p, _p, _p],
    "rsb_coord_min": [_l, _p, _p],
    "rsb_voxel_keys": [_l, _p, _p, _f, _p],
    "rsb_voxel_runs": [_l, _p, _p, _p, _p],
    "rsb_voxel_counts": [_i, _l, _p, _p, _p],
    "rsb_voxel_pick": [_i, _p, _p, _p, _p, _p],
    "rsb_seed_distance": [_l, _p, _l, _p],
    "rsb_coord_max": [_l, _p, _p],
    "rsb_voxel_keys_ravel": [_l, _p, _p, _p, _f, _p],
    "rsb_argmin_f64": [_l, _p, _p],
    "rsb_seed_distance_dev": [_l, _p, _p, _p],
    "rsb_crop_update": [_i, _p, _p, _p, _p, _p],
    "rsb_point_table": [_l, _i, _i, _i, _p, _p, _p, _p],
    "rsb_cross_entropy_forward": [_l, _i, _p, _i, _p, ctypes.c_longlong, _p, _i, _p, _p],
    "rsb_cross_entropy_backward": [_l, _p, _p, _p],
    "rsb_sector_split": [_i, _i, _i, _i, _p, _p, _p, _p, _p, _l, _p, _p, _p, _p],
    "rsb_sector_map_back": [_i, _p, _p, _p],
    "rsb_knnquery_packed": [_i, _i, _i, _p, _p, _p, _p, _p, _p, _i],
    "rsb_knnquery_grid": [_i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _l],
    "rsb_umbrella_features": [_l, _i, _i, _i, _i, _p, _p, _p, _p, _i, _i],
    "rsb_bn_eval_coef": [_i, _p, _p, _p, _p, _f, _p, _p, _p, _p],
    "rsb_bn_update_running": [_i, _l, _p, _p, _f, _p, _p, _p],
    "rsb_umbrella_mlp_stats": [_l, _i, _i, _p, _p, _p, _p],
    "rsb_umbrella_mlp_forward": [_l, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p],
    "rsb_umbrella_mlp_backward": [_l, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "rsb_group_rows_forward": [_l, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p],
    "rsb_group_rows_backward": [_l, _i, _i, _i, _i, _p, _p, _p, _p],
    "rsb_grouping_packed_forward": [_i, _i, _i, _p, _p, _p],
    "rsb_grouping_packed_backward": [_i, _i, _i, _p, _p, _p],
    "rsb_interpolation_packed_forward": [_i, _i, _i, _p, _p, _p, _p],
    "rsb_interpolation_packed_backward": [_i, _i, _i, _p, _p, _p, _p],
    "rsb_linear_tc_prep_weight": [_i, _i, _p, _i, _i, _p],
    "rsb_linear_tc_forward": [_l, _i, _i, _p, _i, _p, _p, _i, _p, _p, _p, _p],
    "rsb_gemm_rows": [_l, _i, _p, _p, _p],
    "rsb_gemm_wgrad": [_l, _p, _p, _p, _i],
    "rsb_bn_finalize": [_i, _l, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p],
    "rsb_pool_forward": [_l, _i, _i, _p, _i, _p, _p, _p, _p],
    "rsb_pool_backward_stats": [_l, _i, _i, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p],
    "rsb_bn_backward_coef": [_i, _l, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p],
    "rsb_bn_relu_backward": [_l, _i, _p, _i, _p, _i, _p, _i, _p, _p, _p, _p, _i, _p],
    "rsb_bn_apply": [_l, _i, _p, _i, _p, _p, _i, _p, _i],
    "rsb_pool_bn_backward_dense": [_l, _i, _i, _p, _p, _p, _i, _p, _p, _p],
}
EXPORTS = sorted(list(_SIGS) + ["rsb_abi_version", "rsb_last_error", "rsb_launch_count", "rsb_reset_launch_count",
                                "rsb_linear_tc_weight_floats", "rsb_knn_grid_workspace_bytes", "rsb_tc_set_generation", "rsb_tc_set_sm_budget", "rsb_fps_set_generation", "rsb_knn_grid_set_counters", "rsb_sector_split_workspace_bytes"])


def build(force=False):
    """Compile the library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    env = dict(os.environ)
    if force:
        env["FORCE"] = "1"
    subprocess.check_call(["bash", os.path.join(_HERE, "csrc", "build.sh")], env=env)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(repsurf_b200 has no CPU / PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args + [_p]
            fn.restype = _i
        L.rsb_last_error.restype = ctypes.c_char_p
        L.rsb_launch_count.restype = ctypes.c_ulonglong
        L.rsb_abi_version.restype = _i
        L.rsb_linear_tc_weight_floats.restype = _l
        L.rsb_linear_tc_weight_floats.argtypes = [_i, _i]
        L.rsb_knn_grid_workspace_bytes.restype = _l
        L.rsb_knn_grid_workspace_bytes.argtypes = [_i, _i]
        L.rsb_sector_split_workspace_bytes.restype = _l
        L.rsb_sector_split_workspace_bytes.argtypes = [_i, _i, _i]
        L.rsb_knn_grid_set_counters.restype = None
        L.rsb_knn_grid_set_counters.argtypes = [_p]
        L.rsb_fps_set_generation.restype = None
        L.rsb_fps_set_generation.argtypes = [_i]
        L.rsb_tc_set_generation.restype = None
        L.rsb_tc_set_generation.argtypes = [_i]
        L.rsb_tc_set_sm_budget.restype = None
        L.rsb_tc_set_sm_budget.argtypes = [_i]
        _lib = L
    return _lib


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("repsurf_b200 kernels take CUDA tensors (no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("repsurf_b200 kernels take contiguous tensors")
    return t.data_ptr()


_FN = {}


def call(name, *args):
    """Invoke a C-ABI entry on torch's current stream; tensors are passed as device pointers."""
    fn = _FN.get(name)
    if fn is None:
        fn = _FN[name] = getattr(lib(), name)
    conv = [(_ptr(a) if (a is None or isinstance(a, torch.Tensor)) else (ctypes.addressof(a) if isinstance(a, ctypes.Structure) else a)) for a in args]
    rc = fn(*conv, torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))
    if rc != 0:
        raise RuntimeError(f"{name} failed (cudaError {rc}): {lib().rsb_last_error().decode()}")


def launch_count():
    return int(lib().rsb_launch_count())


def reset_launch_count():
    lib().rsb_reset_launch_count()
