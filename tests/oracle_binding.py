"""ctypes binding of the CPU parity oracle (oracle/libame_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libame_oracle.so")

PRED_NAMES = ("FULL_2CP", "FULL_3CP", "HALF_2CP", "HALF_3CP")
CUS_PER_CTU = (201, 201, 284, 284)

CPMV_DTYPE = np.dtype([("nCPs", "<i4"), ("LTx", "<i4"), ("LTy", "<i4"), ("RTx", "<i4"),
                       ("RTy", "<i4"), ("LBx", "<i4"), ("LBy", "<i4")])
assert CPMV_DTYPE.itemsize == 28


class Opts(C.Structure):
    _fields_ = [("extra_grad_iter", C.c_int), ("fused_backsub", C.c_int),
                ("cvt_rule", C.c_int), ("threads", C.c_int)]


class RefList(C.Structure):
    _fields_ = [("refs", C.c_int * 4), ("is_lt", C.c_int * 4)]


class CpmvsC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("nCPs", "LTx", "LTy", "RTx", "RTy", "LBx", "LBy")]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "libame_oracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.oracle_num_ctus.restype = C.c_int
        L.oracle_lambda.restype = C.c_float
        L.oracle_lambda.argtypes = [C.c_int, C.c_int]
        L.oracle_rate_cost.argtypes = [C.c_int, C.c_float]
        L.oracle_scale_delta.argtypes = [C.c_double, C.c_int]
        L.oracle_ref_pass.argtypes = [C.POINTER(Opts), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        _lib = L
    return _lib


def default_opts(threads=0, extra_iter=0, fused_backsub=None, cvt_rule=None):
    o = Opts()
    lib().oracle_default_opts(C.byref(o))
    o.threads = threads
    o.extra_grad_iter = extra_iter
    if fused_backsub is not None:
        o.fused_backsub = fused_backsub
    if cvt_rule is not None:
        o.cvt_rule = cvt_rule
    return o


def num_ctus(W, H):
    return lib().oracle_num_ctus(W, H)


def ref_pass(ref, cur, lam, opts=None):
    """ref, cur: (H, W) uint16 planes.  Returns (costs[4], cpmvs[4]) numpy arrays."""
    ref = np.ascontiguousarray(ref, dtype=np.uint16)
    cur = np.ascontiguousarray(cur, dtype=np.uint16)
    H, W = ref.shape
    assert cur.shape == (H, W)
    n = num_ctus(W, H)
    costs = [np.zeros(n * k, dtype=np.int64) for k in CUS_PER_CTU]
    cpmvs = [np.zeros(n * k, dtype=CPMV_DTYPE) for k in CUS_PER_CTU]
    cp = (C.c_void_p * 4)(*[c.ctypes.data for c in costs])
    mp = (C.c_void_p * 4)(*[m.ctypes.data for m in cpmvs])
    o = opts if opts is not None else default_opts()
    lib().oracle_ref_pass(C.byref(o), ref.ctypes.data, cur.ctypes.data, W, H, C.c_float(lam), cp, mp)
    return costs, cpmvs


def lambda_for(qp, poc):
    return float(lib().oracle_lambda(qp, poc))


def delta_qp(qp, poc):
    return lib().oracle_compute_delta_qp(qp, poc)


def ref_lists(n_frames):
    """Reference POC lists for poc = 1..n_frames (newest first)."""
    st = RefList()
    lib().oracle_ref_list_init(C.byref(st))
    out = []
    lst = (C.c_int * 4)()
    for poc in range(1, n_frames + 1):
        n = lib().oracle_ref_list_step(C.byref(st), poc, lst)
        out.append([lst[i] for i in range(n)])
    return out


def predict_4x4(ref, px, py, mvx, mvy):
    ref = np.ascontiguousarray(ref, dtype=np.uint16)
    H, W = ref.shape
    out = (C.c_int * 16)()
    lib().oracle_predict_4x4(ref.ctypes.data_as(C.c_void_p), W, H, px, py, mvx, mvy, out)
    return np.array(out[:], dtype=np.int32).reshape(4, 4)


def satd_4x4(org, pred):
    a = (C.c_int * 16)(*[int(v) for v in np.asarray(org).reshape(-1)])
    b = (C.c_int * 16)(*[int(v) for v in np.asarray(pred).reshape(-1)])
    return lib().oracle_satd_4x4(a, b)


def affine_bits(nCP, c, p=(0, 0, 0, 0, 0, 0)):
    cc = CpmvsC(0, *c)
    pp = CpmvsC(0, *p)
    return lib().oracle_affine_bits(nCP, C.byref(cc), C.byref(pp))


def solve(M, n, fused=1):
    m = (C.c_double * 49)(*[float(v) for v in np.asarray(M, dtype=np.float64).reshape(-1)])
    out = (C.c_double * 6)()
    lib().oracle_solve(m, n, fused, out)
    return np.array(out[:n])


def cu_geometry(ha, k):
    out = (C.c_int * 4)()
    g = lib().oracle_cu_geometry(ha, k, out)
    return g, tuple(out[:])
