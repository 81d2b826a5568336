"""ctypes binding of the CPU oracle (oracle/pointops_oracle.c) on torch CPU tensors.

TEST INFRASTRUCTURE ONLY — see the header of pointops_oracle.c.  Imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs; never by the
product package `repsurf_b200`.

Two layers:
  * `lib()` + the thin `fps_dense(...)`-style helpers: allocate outputs exactly as the reference
    autograd wrappers do (cls/po/functions/pointops.py, seg/po/functions/pointops.py) and call C.
  * `fake_pointops_cuda(tree)`: a stand-in for the reference's pybind module `pointops_cuda`
    (cls/po/src/pointops_api.cpp:13-31, seg/po/src/pointops_api.cpp:12-23) with the same function
    names and argument orders, so the UNMODIFIED reference Python can be imported from
    /root/reference and run on CPU (used only by oracle/make_golden.py, in the build container).
"""
import ctypes
import os
import subprocess
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_c_int = ctypes.c_int
_c_float = ctypes.c_float
_vp = ctypes.c_void_p


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "pointops_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_opt_n_threads.restype = _c_int
        _LIB.orc_num_threads.restype = _c_int
    return _LIB


def _p(t):
    if t is None:
        return None
    assert t.device.type == "cpu" and t.is_contiguous(), "oracle takes contiguous CPU tensors"
    return _vp(t.data_ptr())


def num_threads():
    return int(lib().orc_num_threads())


def opt_n_threads(n):
    return int(lib().orc_opt_n_threads(_c_int(int(n))))


# ----------------------------------------------------------------------------------------------
# dense (classification) layout — cls/po/functions/pointops.py
# ----------------------------------------------------------------------------------------------
def fps_dense(xyz, m):
    b, n, _ = xyz.shape
    idx = torch.zeros(b, m, dtype=torch.int32)
    temp = torch.full((b, n), 1e10, dtype=torch.float32)
    lib().orc_fps_dense(_c_int(b), _c_int(n), _c_int(m), _p(xyz), _p(temp), _p(idx))
    return idx


def gather_fwd(features, idx):
    b, c, n = features.shape
    m = idx.shape[1]
    out = torch.empty(b, c, m, dtype=torch.float32)
    lib().orc_gather_fwd(_c_int(b), _c_int(c), _c_int(n), _c_int(m), _p(features), _p(idx), _p(out))
    return out


def gather_bwd(grad_out, idx, n):
    b, c, m = grad_out.shape
    g = torch.zeros(b, c, n, dtype=torch.float32)
    lib().orc_gather_bwd(_c_int(b), _c_int(c), _c_int(n), _c_int(m), _p(grad_out), _p(idx), _p(g))
    return g


def ballquery(radius, nsample, xyz, new_xyz):
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.zeros(b, m, nsample, dtype=torch.int32)
    lib().orc_ballquery(_c_int(b), _c_int(n), _c_int(m), _c_float(radius), _c_int(nsample), _p(new_xyz), _p(xyz),
                        _p(idx))
    return idx


def knn_dense(nsample, xyz, new_xyz=None, return_dist2=False):
    if new_xyz is None:
        new_xyz = xyz
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.zeros(b, m, nsample, dtype=torch.int32)
    d2 = torch.zeros(b, m, nsample, dtype=torch.float32)
    lib().orc_knn_dense(_c_int(b), _c_int(n), _c_int(m), _c_int(nsample), _p(xyz), _p(new_xyz), _p(idx), _p(d2))
    return (idx, d2) if return_dist2 else idx


def knn_heap_dense(nsample, xyz, new_xyz=None, return_dist2=False):
    if new_xyz is None:
        new_xyz = xyz
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.zeros(b, m, nsample, dtype=torch.int32)
    d2 = torch.zeros(b, m, nsample, dtype=torch.float32)
    lib().orc_knn_heap_dense(_c_int(b), _c_int(n), _c_int(m), _c_int(nsample), _p(xyz), _p(new_xyz), _p(idx),
                             _p(d2))
    return (idx, d2) if return_dist2 else idx


def group_fwd(features, idx):
    b, c, n = features.shape
    _, m, ns = idx.shape
    out = torch.empty(b, c, m, ns, dtype=torch.float32)
    lib().orc_group_fwd(_c_int(b), _c_int(c), _c_int(n), _c_int(m), _c_int(ns), _p(features), _p(idx), _p(out))
    return out


def group_bwd(grad_out, idx, n):
    b, c, m, ns = grad_out.shape
    g = torch.zeros(b, c, n, dtype=torch.float32)
    lib().orc_group_bwd(_c_int(b), _c_int(c), _c_int(n), _c_int(m), _c_int(ns), _p(grad_out), _p(idx), _p(g))
    return g


def nn3(unknown, known):
    b, n, _ = unknown.shape
    m = known.shape[1]
    d2 = torch.empty(b, n, 3, dtype=torch.float32)
    idx = torch.empty(b, n, 3, dtype=torch.int32)
    lib().orc_nn3(_c_int(b), _c_int(n), _c_int(m), _p(unknown), _p(known), _p(d2), _p(idx))
    return d2, idx


def interp_fwd(features, idx, weight):
    b, c, m = features.shape
    n = idx.shape[1]
    out = torch.empty(b, c, n, dtype=torch.float32)
    lib().orc_interp_fwd(_c_int(b), _c_int(c), _c_int(m), _c_int(n), _p(features), _p(idx), _p(weight), _p(out))
    return out


def interp_bwd(grad_out, idx, weight, m):
    b, c, n = grad_out.shape
    g = torch.zeros(b, c, m, dtype=torch.float32)
    lib().orc_interp_bwd(_c_int(b), _c_int(c), _c_int(n), _c_int(m), _p(grad_out), _p(idx), _p(weight), _p(g))
    return g


# ----------------------------------------------------------------------------------------------
# packed (segmentation) layout — seg/po/functions/pointops.py
# ----------------------------------------------------------------------------------------------
def _n_max(offset):
    o = offset.tolist()
    return max(b - a for a, b in zip([0] + o[:-1], o))


def fps_packed(xyz, offset, new_offset):
    """seg/po/functions/pointops.py:31-49 (FurthestSampling.forward)."""
    n, b = xyz.shape[0], offset.shape[0]
    idx = torch.zeros(int(new_offset[b - 1]), dtype=torch.int32)
    tmp = torch.full((n,), 1e10, dtype=torch.float32)
    lib().orc_fps_packed(_c_int(b), _c_int(_n_max(offset)), _p(xyz), _p(offset), _p(new_offset), _p(tmp), _p(idx))
    return idx


def sectorized_fps(xyz, offset, new_offset, num_sectors, min_points=10000):
    """Restates seg/po/functions/pointops.py:52-111 (SectorizedFurthestSampling.forward): per cloud
    with >= min_points points, azimuth = atan2(x, y) (sic: x first, :73), `num_sectors` equal-width
    bins over [min, max + 1e-4] built with torch.linspace (:74), half-open membership (:77),
    per-sector quota new_size // S with the remainder on the last sector (:84-85); one FPS per
    sector on the gathered coordinates; results mapped back through the index list (:105) => int64."""
    last = 0
    sizes, new_sizes, indices = [], [], []
    off = offset.tolist()
    noff = new_offset.tolist()
    for i in range(len(off)):
        size = off[i] - last
        s_cnt = 1 if size < min_points else num_sectors
        pts = xyz[last:last + size]
        angle = torch.atan2(pts[:, 0], pts[:, 1])
        edges = torch.linspace(angle.min(), angle.max() + 1e-4, s_cnt + 1)
        for s in range(s_cnt):
            sel = torch.where((angle >= edges[s]) & (angle < edges[s + 1]))[0] + last
            indices.append(sel)
            sizes.append(sel.shape[0])
        new_size = noff[i] - (noff[i - 1] if i > 0 else 0)
        quota = [new_size // s_cnt] * s_cnt
        quota[-1] += new_size % s_cnt
        new_sizes += quota
        last = off[i]
    sector_offset = torch.tensor(sizes, dtype=torch.long).cumsum(0).int()
    new_sector_offset = torch.tensor(new_sizes, dtype=torch.long).cumsum(0).int()
    indices = torch.cat(indices).long()
    sector_xyz = xyz[indices].contiguous()
    idx = fps_packed(sector_xyz, sector_offset, new_sector_offset)
    return indices[idx.long()]


def knn_packed(nsample, xyz, new_xyz, offset, new_offset, sqrt=True):
    """seg/po/functions/pointops.py:114-130: returns (idx int32 [m,ns], sqrt(dist2) [m,ns]); sqrt=False gives the
    kernel's raw squared distances."""
    if new_xyz is None:
        new_xyz = xyz
    m = new_xyz.shape[0]
    idx = torch.zeros(m, nsample, dtype=torch.int32)
    d2 = torch.zeros(m, nsample, dtype=torch.float32)
    lib().orc_knn_packed(_c_int(m), _c_int(nsample), _p(xyz), _p(new_xyz), _p(offset), _p(new_offset), _p(idx),
                         _p(d2))
    return idx, (torch.sqrt(d2) if sqrt else d2)


def group_packed_fwd(inp, idx):
    m, ns = idx.shape
    c = inp.shape[1]
    out = torch.empty(m, ns, c, dtype=torch.float32)
    lib().orc_group_packed_fwd(_c_int(m), _c_int(ns), _c_int(c), _p(inp), _p(idx), _p(out))
    return out


def group_packed_bwd(grad_out, idx, n):
    m, ns, c = grad_out.shape
    g = torch.zeros(n, c, dtype=torch.float32)
    lib().orc_group_packed_bwd(_c_int(m), _c_int(ns), _c_int(c), _p(grad_out), _p(idx), _p(g))
    return g


def interp_packed_fwd(inp, idx, weight):
    n, k = idx.shape
    c = inp.shape[1]
    out = torch.zeros(n, c, dtype=torch.float32)
    lib().orc_interp_packed_fwd(_c_int(n), _c_int(c), _c_int(k), _p(inp), _p(idx), _p(weight), _p(out))
    return out


def interp_packed_bwd(grad_out, idx, weight, m):
    n, c = grad_out.shape
    k = idx.shape[1]
    g = torch.zeros(m, c, dtype=torch.float32)
    lib().orc_interp_packed_bwd(_c_int(n), _c_int(c), _c_int(k), _p(grad_out), _p(idx), _p(weight), _p(g))
    return g


def subtraction_fwd(in1, in2, idx):
    n, c = in1.shape
    ns = idx.shape[1]
    out = torch.empty(n, ns, c, dtype=torch.float32)
    lib().orc_subtraction_fwd(_c_int(n), _c_int(ns), _c_int(c), _p(in1), _p(in2), _p(idx), _p(out))
    return out


def subtraction_bwd(grad_out, idx):
    n, ns, c = grad_out.shape
    g1, g2 = torch.zeros(n, c, dtype=torch.float32), torch.zeros(n, c, dtype=torch.float32)
    lib().orc_subtraction_bwd(_c_int(n), _c_int(ns), _c_int(c), _p(idx), _p(grad_out), _p(g1), _p(g2))
    return g1, g2


def aggregation_fwd(inp, pos, w, idx):
    n, ns, c = pos.shape
    out = torch.zeros(n, c, dtype=torch.float32)
    lib().orc_aggregation_fwd(_c_int(n), _c_int(ns), _c_int(c), _c_int(w.shape[-1]), _p(inp), _p(pos), _p(w), _p(idx), _p(out))
    return out


def aggregation_bwd(inp, pos, w, idx, grad_out):
    n, ns, c = pos.shape
    w_c = w.shape[-1]
    g_in, g_pos, g_w = torch.zeros(n, c), torch.zeros(n, ns, c), torch.zeros(n, ns, w_c)
    lib().orc_aggregation_bwd(_c_int(n), _c_int(ns), _c_int(c), _c_int(w_c), _p(inp), _p(pos), _p(w), _p(idx), _p(grad_out),
                              _p(g_in), _p(g_pos), _p(g_w))
    return g_in, g_pos, g_w


# ----------------------------------------------------------------------------------------------
# stand-in for the reference's pybind module (golden-vector generation only)
# ----------------------------------------------------------------------------------------------
def fake_pointops_cuda(tree):
    """Module object exposing the reference's native function names, backed by the C oracle.
    tree = 'cls' -> cls/po/src/pointops_api.cpp:13-31; 'seg' -> seg/po/src/pointops_api.cpp:12-23."""
    L = lib()
    mod = types.ModuleType("pointops_cuda")
    i = _c_int
    if tree == "cls":
        mod.furthestsampling_cuda = lambda b, n, m, xyz, temp, idx: L.orc_fps_dense(i(b), i(n), i(m), _p(xyz), _p(temp), _p(idx))
        mod.gathering_forward_cuda = lambda b, c, n, m, pts, idx, out: L.orc_gather_fwd(i(b), i(c), i(n), i(m), _p(pts), _p(idx), _p(out))
        mod.gathering_backward_cuda = lambda b, c, n, m, go, idx, gp: L.orc_gather_bwd(i(b), i(c), i(n), i(m), _p(go), _p(idx), _p(gp))
        mod.ballquery_cuda = lambda b, n, m, r, ns, new_xyz, xyz, idx: L.orc_ballquery(i(b), i(n), i(m), _c_float(r), i(ns), _p(new_xyz), _p(xyz), _p(idx))
        mod.knnquery_cuda = lambda b, n, m, ns, xyz, new_xyz, idx, d2: L.orc_knn_dense(i(b), i(n), i(m), i(ns), _p(xyz), _p(new_xyz), _p(idx), _p(d2))
        mod.knnquery_heap_cuda = lambda b, n, m, ns, xyz, new_xyz, idx, d2: L.orc_knn_heap_dense(i(b), i(n), i(m), i(ns), _p(xyz), _p(new_xyz), _p(idx), _p(d2))
        mod.grouping_forward_cuda = lambda b, c, n, m, ns, pts, idx, out: L.orc_group_fwd(i(b), i(c), i(n), i(m), i(ns), _p(pts), _p(idx), _p(out))
        mod.grouping_backward_cuda = lambda b, c, n, m, ns, go, idx, gp: L.orc_group_bwd(i(b), i(c), i(n), i(m), i(ns), _p(go), _p(idx), _p(gp))
        mod.grouping_int_forward_cuda = lambda b, c, n, m, ns, pts, idx, out: L.orc_group_int_fwd(i(b), i(c), i(n), i(m), i(ns), _p(pts), _p(idx), _p(out))
        mod.nearestneighbor_cuda = lambda b, n, m, unk, kn, d2, idx: L.orc_nn3(i(b), i(n), i(m), _p(unk), _p(kn), _p(d2), _p(idx))
        mod.interpolation_forward_cuda = lambda b, c, m, n, pts, idx, w, out: L.orc_interp_fwd(i(b), i(c), i(m), i(n), _p(pts), _p(idx), _p(w), _p(out))
        mod.interpolation_backward_cuda = lambda b, c, n, m, go, idx, w, gp: L.orc_interp_bwd(i(b), i(c), i(n), i(m), _p(go), _p(idx), _p(w), _p(gp))
    elif tree == "seg":
        mod.furthestsampling_cuda = lambda b, n_max, xyz, off, noff, tmp, idx: L.orc_fps_packed(i(b), i(int(n_max)), _p(xyz), _p(off), _p(noff), _p(tmp), _p(idx))
        mod.knnquery_cuda = lambda m, ns, xyz, new_xyz, off, noff, idx, d2: L.orc_knn_packed(i(m), i(ns), _p(xyz), _p(new_xyz), _p(off), _p(noff), _p(idx), _p(d2))
        mod.grouping_forward_cuda = lambda m, ns, c, inp, idx, out: L.orc_group_packed_fwd(i(m), i(ns), i(c), _p(inp), _p(idx), _p(out))
        mod.grouping_backward_cuda = lambda m, ns, c, go, idx, gi: L.orc_group_packed_bwd(i(m), i(ns), i(c), _p(go), _p(idx), _p(gi))
        mod.interpolation_forward_cuda = lambda n, c, k, inp, idx, w, out: L.orc_interp_packed_fwd(i(n), i(c), i(k), _p(inp), _p(idx), _p(w), _p(out))
        mod.interpolation_backward_cuda = lambda n, c, k, go, idx, w, gi: L.orc_interp_packed_bwd(i(n), i(c), i(k), _p(go), _p(idx), _p(w), _p(gi))
    else:
        raise ValueError(tree)
    return mod
