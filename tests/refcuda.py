"""ctypes access to the REFERENCE's own CUDA kernels (oracle/_ref/libref_pointops_{cls,seg}.so, built by
oracle/build_ref.sh from /root/reference, unmodified).  Test infrastructure: the GPU-side comparator."""
import ctypes
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIBS = {}


def available(tree):
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", f"libref_pointops_{tree}.so"))


def lib(tree):
    if tree not in _LIBS:
        _LIBS[tree] = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", f"libref_pointops_{tree}.so"))
    return _LIBS[tree]


def _p(t):
    assert t.is_cuda and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


_i, _f = ctypes.c_int, ctypes.c_float


def _sync():
    torch.cuda.synchronize()


# ---- cls (dense) ------------------------------------------------------------------------------------
def fps_dense(xyz, m):
    b, n, _ = xyz.shape
    idx = torch.zeros(b, m, dtype=torch.int32, device=xyz.device)
    temp = torch.full((b, n), 1e10, dtype=torch.float32, device=xyz.device)
    _sync()
    lib("cls").furthestsampling_cuda_launcher(_i(b), _i(n), _i(m), _p(xyz), _p(temp), _p(idx))
    _sync()
    return idx


def ballquery(radius, nsample, xyz, new_xyz):
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.zeros(b, m, nsample, dtype=torch.int32, device=xyz.device)
    _sync()
    lib("cls").ballquery_cuda_launcher_fast(_i(b), _i(n), _i(m), _f(radius), _i(nsample), _p(new_xyz), _p(xyz), _p(idx),
                                            ctypes.c_void_p(0))
    _sync()
    return idx


def knn_dense(nsample, xyz, new_xyz):
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.zeros(b, m, nsample, dtype=torch.int32, device=xyz.device)
    d2 = torch.zeros(b, m, nsample, dtype=torch.float32, device=xyz.device)
    _sync()
    lib("cls").knnquery_cuda_launcher(_i(b), _i(n), _i(m), _i(nsample), _p(xyz), _p(new_xyz), _p(idx), _p(d2),
                                      ctypes.c_void_p(0))
    _sync()
    return idx


def knn_heap_dense(nsample, xyz, new_xyz):
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.zeros(b, m, nsample, dtype=torch.int32, device=xyz.device)
    d2 = torch.zeros(b, m, nsample, dtype=torch.float32, device=xyz.device)
    _sync()
    lib("cls").knnquery_heap_cuda_launcher(_i(b), _i(n), _i(m), _i(nsample), _p(xyz), _p(new_xyz), _p(idx), _p(d2),
                                           ctypes.c_void_p(0))
    _sync()
    return idx, d2


def nn3(unknown, known):
    b, n, _ = unknown.shape
    m = known.shape[1]
    d2 = torch.zeros(b, n, 3, dtype=torch.float32, device=unknown.device)
    idx = torch.zeros(b, n, 3, dtype=torch.int32, device=unknown.device)
    _sync()
    lib("cls").nearestneighbor_cuda_launcher_fast(_i(b), _i(n), _i(m), _p(unknown), _p(known), _p(d2), _p(idx))
    _sync()
    return d2, idx


# ---- seg (packed) -----------------------------------------------------------------------------------
def fps_packed(xyz, offset, new_offset):
    o = offset.tolist()
    n_max = max(b - a for a, b in zip([0] + o[:-1], o))
    idx = torch.zeros(int(new_offset[-1]), dtype=torch.int32, device=xyz.device)
    tmp = torch.full((xyz.shape[0],), 1e10, dtype=torch.float32, device=xyz.device)
    _sync()
    lib("seg").furthestsampling_cuda_launcher(_i(len(o)), _i(n_max), _p(xyz), _p(offset), _p(new_offset), _p(tmp), _p(idx))
    _sync()
    return idx


def knn_packed(nsample, xyz, new_xyz, offset, new_offset):
    m = new_xyz.shape[0]
    idx = torch.zeros(m, nsample, dtype=torch.int32, device=xyz.device)
    d2 = torch.zeros(m, nsample, dtype=torch.float32, device=xyz.device)
    _sync()
    lib("seg").knnquery_cuda_launcher(_i(m), _i(nsample), _p(xyz), _p(new_xyz), _p(offset), _p(new_offset), _p(idx), _p(d2))
    _sync()
    return idx, d2


def subtraction_fwd(in1, in2, idx):
    n, c = in1.shape
    ns = idx.shape[1]
    out = torch.zeros(n, ns, c, dtype=torch.float32, device=in1.device)
    _sync()
    lib("seg").subtraction_forward_cuda_launcher(_i(n), _i(ns), _i(c), _p(in1), _p(in2), _p(idx), _p(out))
    _sync()
    return out


def subtraction_bwd(grad_out, idx):
    n, ns, c = grad_out.shape
    g1 = torch.zeros(n, c, dtype=torch.float32, device=grad_out.device)
    g2 = torch.zeros(n, c, dtype=torch.float32, device=grad_out.device)
    _sync()
    lib("seg").subtraction_backward_cuda_launcher(_i(n), _i(ns), _i(c), _p(idx), _p(grad_out), _p(g1), _p(g2))
    _sync()
    return g1, g2


def aggregation_fwd(inp, pos, w, idx):
    n, ns, c = pos.shape
    out = torch.zeros(n, c, dtype=torch.float32, device=inp.device)
    _sync()
    lib("seg").aggregation_forward_cuda_launcher(_i(n), _i(ns), _i(c), _i(w.shape[-1]), _p(inp), _p(pos), _p(w), _p(idx), _p(out))
    _sync()
    return out


def aggregation_bwd(inp, pos, w, idx, grad_out):
    n, ns, c = pos.shape
    w_c = w.shape[-1]
    dev = inp.device
    g_in, g_pos, g_w = torch.zeros(n, c, device=dev), torch.zeros(n, ns, c, device=dev), torch.zeros(n, ns, w_c, device=dev)
    _sync()
    lib("seg").aggregation_backward_cuda_launcher(_i(n), _i(ns), _i(c), _i(w_c), _p(inp), _p(pos), _p(w), _p(idx), _p(grad_out),
                                                  _p(g_in), _p(g_pos), _p(g_w))
    _sync()
    return g_in, g_pos, g_w
