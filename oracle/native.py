"""ctypes front end of oracle/dusty_oracle.c (built into oracle/_build/liboracle.so by gcc)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "dusty_oracle.c")
LIB = os.path.join(HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        flags = ["-O3", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared"]
        try:        # fmaf() on the FMA unit when the host has one (same correctly rounded result as libm's)
            if " fma " in open("/proc/cpuinfo").read():
                flags.append("-mfma")
        except OSError:
            pass
        subprocess.check_call(["gcc", *flags, "-o", LIB, SRC, "-lm"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def nnsearch(xyz1, xyz2, rounding="cuda"):
    """Directed NN of xyz1 (b,n,3) in xyz2 (b,m,3) -> (dist (b,n) f32, idx (b,n) i32)."""
    a, c = _f(xyz1), _f(xyz2)
    b, n, _ = a.shape
    m = c.shape[1]
    dist = np.zeros((b, n), np.float32)
    idx = np.zeros((b, n), np.int32)
    fn = lib().oracle_nnsearch_cuda if rounding == "cuda" else lib().oracle_nnsearch_cpu
    fn(C.c_int(b), C.c_int(n), C.c_int(m), _p(a), _p(c), _p(dist), _p(idx))
    return dist, idx


def chamfer_forward(xyz1, xyz2, rounding="cuda"):
    d1, i1 = nnsearch(xyz1, xyz2, rounding)
    d2, i2 = nnsearch(xyz2, xyz1, rounding)
    return d1, d2, i1, i2


def chamfer_backward(xyz1, xyz2, g1, g2, idx1, idx2):
    a, c = _f(xyz1), _f(xyz2)
    b, n, _ = a.shape
    m = c.shape[1]
    gx1 = np.zeros_like(a)
    gx2 = np.zeros_like(c)
    i1 = np.ascontiguousarray(idx1, np.int32)
    i2 = np.ascontiguousarray(idx2, np.int32)
    g1, g2 = _f(g1), _f(g2)
    lib().oracle_chamfer_backward(C.c_int(b), C.c_int(n), C.c_int(m), _p(a), _p(c), _p(g1), _p(g2), _p(i1), _p(i2),
                                  _p(gx1), _p(gx2))
    return gx1, gx2


def pairwise_cd(A, B=None, rows=None, rounding="cuda"):
    """Matrix of compute_cd values; B=None is the symmetric case. rows=(begin,end)."""
    a = _f(A)
    sym = B is None
    c = a if sym else _f(B)
    na, pa, _ = a.shape
    nb, pb, _ = c.shape
    r0, r1 = rows if rows is not None else (0, na)
    M = np.zeros((na, nb), np.float32)
    lib().oracle_pairwise_cd(_p(a), C.c_int(na), C.c_int(pa), _p(c), C.c_int(nb), C.c_int(pb), C.c_int(r0),
                             C.c_int(r1), C.c_int(int(sym)), C.c_int(int(rounding == "cuda")), _p(M))
    return M


def fps(xyz, m):
    """Reference-exact farthest-point sampling of xyz (b,n,3) -> idx (b,m) i32."""
    a = _f(xyz)
    b, n, _ = a.shape
    idx = np.zeros((b, m), np.int32)
    lib().oracle_fps(C.c_int(b), C.c_int(n), C.c_int(m), _p(a), _p(idx))
    return idx


def gather_points(points, idx):
    pts = _f(points)
    b, c, n = pts.shape
    ii = np.ascontiguousarray(idx, np.int32)
    m = ii.shape[1]
    out = np.zeros((b, c, m), np.float32)
    lib().oracle_gather_points(C.c_int(b), C.c_int(c), C.c_int(n), C.c_int(m), _p(pts), _p(ii), _p(out))
    return out


def downsample_point_clouds(xyz, k):
    """(b,n,3) -> (b,k,3), reference fps/furthest_point_sampling.py:84-93."""
    a = _f(xyz)
    idx = fps(a, k)
    return np.take_along_axis(a, idx[:, :, None].astype(np.int64), axis=1), idx


def box_bound_violations(rows, cands):
    """(number of (row, candidate) pairs whose reference-rounded distance is below the kernels' box lower bound,
    the bound itself). The pruning in csrc/fps.cu and csrc/chamfer.cu is exact iff the count is always 0."""
    a, c = _f(rows).reshape(-1, 3), _f(cands).reshape(-1, 3)
    fn = lib().oracle_box_bound_violations
    fn.restype = C.c_longlong
    lb = C.c_float()
    bad = fn(C.c_int(len(a)), _p(a), C.c_int(len(c)), _p(c), C.byref(lb))
    return int(bad), float(lb.value)
