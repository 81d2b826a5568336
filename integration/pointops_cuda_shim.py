"""Path B of INTEGRATION.md as real code: stand-ins for the reference's pybind modules `pointops_cuda` (one per tree) that keep
the reference's own `modules/pointops/functions/pointops.py` untouched and route every native call to librepsurf_b200.so
through ctypes.  Usage on the reference side (before importing its pointops.py):

    import sys
    from integration.pointops_cuda_shim import cls_module          # or seg_module
    sys.modules["pointops_cuda"] = cls_module()

Function names, argument order and in-place output semantics are those of
  classification/modules/pointops/src/pointops_api.cpp:13-31      (12 functions)
  segmentation/modules/pointops/src/pointops_api.cpp:12-23         (10 functions)
Executed against the library by tests/test_integration_shim_gpu.py."""
import ctypes
import os
import types

import torch

_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "repsurf_b200", "librepsurf_b200.so")
_L = None


def _lib():
    global _L
    if _L is None:
        _L = ctypes.CDLL(_LIB)
        _L.rsb_last_error.restype = ctypes.c_char_p
        _L.rsb_knn_grid_workspace_bytes.restype = ctypes.c_long
    return _L


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ck(rc):
    if rc:
        raise RuntimeError(_lib().rsb_last_error().decode())


_i, _f, _l = ctypes.c_int, ctypes.c_float, ctypes.c_long


def cls_module():
    L = _lib()
    m = types.ModuleType("pointops_cuda")
    m.furthestsampling_cuda = lambda b, n, mm, xyz, temp, idx: _ck(L.rsb_furthestsampling_dense(_i(b), _i(n), _i(mm), _p(xyz), _p(temp), _p(idx), _p(None), _s()))
    m.gathering_forward_cuda = lambda b, c, n, mm, pts, idx, out: _ck(L.rsb_gathering_forward(_i(b), _i(c), _i(n), _i(mm), _p(pts), _p(idx), _p(out), _s()))
    m.gathering_backward_cuda = lambda b, c, n, mm, go, idx, gp: _ck(L.rsb_gathering_backward(_i(b), _i(c), _i(n), _i(mm), _p(go), _p(idx), _p(gp), _s()))
    m.ballquery_cuda = lambda b, n, mm, radius, ns, new_xyz, xyz, idx: _ck(L.rsb_ballquery(_i(b), _i(n), _i(mm), _f(radius), _i(ns), _p(new_xyz), _p(xyz), _p(idx), _s()))
    m.knnquery_cuda = lambda b, n, mm, ns, xyz, new_xyz, idx, d2: _ck(L.rsb_knnquery_dense(_i(b), _i(n), _i(mm), _i(ns), _p(xyz), _p(new_xyz), _p(idx), _p(d2), _s()))
    m.knnquery_heap_cuda = lambda b, n, mm, ns, xyz, new_xyz, idx, d2: _ck(L.rsb_knnquery_heap_dense(_i(b), _i(n), _i(mm), _i(ns), _p(xyz), _p(new_xyz), _p(idx), _p(d2), _s()))
    m.grouping_forward_cuda = lambda b, c, n, mm, ns, pts, idx, out: _ck(L.rsb_grouping_forward(_i(b), _i(c), _i(n), _i(mm), _i(ns), _p(pts), _p(idx), _p(out), _s()))
    m.grouping_backward_cuda = lambda b, c, n, mm, ns, go, idx, gp: _ck(L.rsb_grouping_backward(_i(b), _i(c), _i(n), _i(mm), _i(ns), _p(go), _p(idx), _p(gp), _s()))
    m.grouping_int_forward_cuda = lambda b, c, n, mm, ns, pts, idx, out: _ck(L.rsb_grouping_int_forward(_i(b), _i(c), _i(n), _i(mm), _i(ns), _p(pts), _p(idx), _p(out), _s()))
    m.nearestneighbor_cuda = lambda b, n, mm, unknown, known, d2, idx: _ck(L.rsb_nearestneighbor(_i(b), _i(n), _i(mm), _p(unknown), _p(known), _p(d2), _p(idx), _s()))
    m.interpolation_forward_cuda = lambda b, c, mm, n, pts, idx, w, out: _ck(L.rsb_interpolation_forward(_i(b), _i(c), _i(mm), _i(n), _p(pts), _p(idx), _p(w), _p(out), _s()))
    m.interpolation_backward_cuda = lambda b, c, n, mm, go, idx, w, gp: _ck(L.rsb_interpolation_backward(_i(b), _i(c), _i(n), _i(mm), _p(go), _p(idx), _p(w), _p(gp), _s()))
    return m


def seg_module():
    L = _lib()
    m = types.ModuleType("pointops_cuda")

    def furthestsampling_cuda(b, n_max, xyz, offset, new_offset, tmp, idx):
        _ck(L.rsb_furthestsampling_packed(_i(b), _i(n_max), _p(None), _p(xyz), _p(offset), _p(new_offset), _p(tmp), _p(idx), _p(None), _s()))

    def knnquery_cuda(mm, ns, xyz, new_xyz, offset, new_offset, idx, dist2):
        # the reference's launcher takes no cloud count: it is the length of the offset tensors
        _ck(L.rsb_knnquery_packed(_i(offset.shape[0]), _i(mm), _i(ns), _p(xyz), _p(new_xyz), _p(offset), _p(new_offset), _p(idx), _p(dist2), _i(0), _s()))
    m.furthestsampling_cuda = furthestsampling_cuda
    m.knnquery_cuda = knnquery_cuda
    m.grouping_forward_cuda = lambda mm, ns, c, inp, idx, out: _ck(L.rsb_grouping_packed_forward(_i(mm), _i(ns), _i(c), _p(inp), _p(idx), _p(out), _s()))
    m.grouping_backward_cuda = lambda mm, ns, c, go, idx, gi: _ck(L.rsb_grouping_packed_backward(_i(mm), _i(ns), _i(c), _p(go), _p(idx), _p(gi), _s()))
    m.interpolation_forward_cuda = lambda n, c, k, inp, idx, w, out: _ck(L.rsb_interpolation_packed_forward(_i(n), _i(c), _i(k), _p(inp), _p(idx), _p(w), _p(out), _s()))
    m.interpolation_backward_cuda = lambda n, c, k, go, idx, w, gi: _ck(L.rsb_interpolation_packed_backward(_i(n), _i(c), _i(k), _p(go), _p(idx), _p(w), _p(gi), _s()))
    m.subtraction_forward_cuda = lambda n, ns, c, a, b2, idx, out: _ck(L.rsb_subtraction_forward(_i(n), _i(ns), _i(c), _p(a), _p(b2), _p(idx), _p(out), _s()))

    def subtraction_backward_cuda(n, ns, c, idx, go, g1, g2):
        g2.zero_()                       # the reference's wrapper passes zeroed buffers (pointops.py:212-213); make it explicit
        _ck(L.rsb_subtraction_backward(_i(n), _i(ns), _i(c), _p(idx), _p(go), _p(g1), _p(g2), _s()))
    m.subtraction_backward_cuda = subtraction_backward_cuda
    m.aggregation_forward_cuda = lambda n, ns, c, wc, inp, pos, w, idx, out: _ck(L.rsb_aggregation_forward(_i(n), _i(ns), _i(c), _i(wc), _p(inp), _p(pos), _p(w), _p(idx), _p(out), _s()))

    def aggregation_backward_cuda(n, ns, c, wc, inp, pos, w, idx, go, g_in, g_pos, g_w):
        g_in.zero_()
        _ck(L.rsb_aggregation_backward(_i(n), _i(ns), _i(c), _i(wc), _p(inp), _p(pos), _p(w), _p(idx), _p(go), _p(g_in), _p(g_pos), _p(g_w), _s()))
    m.aggregation_backward_cuda = aggregation_backward_cuda
    return m
