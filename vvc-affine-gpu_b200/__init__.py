"""vvc-affine-gpu_b200: B200-native affine motion-estimation search.

The product is the CUDA C-ABI library `libaffine_me.so` (include/affine_me.h) and the
drop-in CLI `bin/affine_b200`.  This module is a thin ctypes wrapper over the C ABI for
tests and benchmarks.  There is no fallback: if the library is missing or no CUDA device
is present, construction fails loudly.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AME_LIB") or os.path.join(HERE, "libaffine_me.so")  # AME_LIB: development A/B builds
CLI_PATH = os.path.join(HERE, "bin", "affine_b200")

PRED_NAMES = ("FULL_2CP", "FULL_3CP", "HALF_2CP", "HALF_3CP")
CPMV_DTYPE = np.dtype([("nCPs", "<i4"), ("LTx", "<i4"), ("LTy", "<i4"), ("RTx", "<i4"),
                       ("RTy", "<i4"), ("LBx", "<i4"), ("LBy", "<i4")])
OPT_CVT_RULE, OPT_FUSED_BACKSUB, OPT_EARLY_EXIT, OPT_REUSE_START, OPT_SHARE_FIRST, OPT_BIG_TMA, OPT_GROUP_BY_REF = 1, 2, 3, 4, 5, 6, 7
ROLE_CURRENT, ROLE_REFERENCE = 1, 2

# every symbol include/affine_me.h declares
EXPORTS = ("ame_num_ctus", "ame_create", "ame_destroy", "ame_result_len", "ame_set_option", "ame_upload_plane", "ame_upload_plane_ex",
           "ame_search", "ame_search_device", "ame_device_result", "ame_flush", "ame_sync", "ame_last_kernel_ms",
           "ame_timer_start", "ame_timer_stop", "ame_result_block_bytes", "ame_result_bind", "ame_exec_ns", "ame_prepare_plane", "ame_alloc_host", "ame_free_host", "ame_cu_geometry", "ame_last_error", "ame_version")


class AmeResult(C.Structure):
    _fields_ = [("cost", C.c_void_p * 4), ("cpmvs", C.c_void_p * 4)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run `make` (or __graft_entry__.build()) first; "
                               "there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.ame_last_error.restype = C.c_char_p
        L.ame_alloc_host.restype = C.c_void_p
        L.ame_alloc_host.argtypes = [C.c_uint64]
        L.ame_free_host.argtypes = [C.c_void_p]
        L.ame_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ame_destroy.argtypes = [C.c_void_p]
        L.ame_result_len.argtypes = [C.c_void_p, C.c_int]
        L.ame_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ame_upload_plane.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ame_upload_plane_ex.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ame_prepare_plane.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ame_search.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.POINTER(AmeResult)]
        L.ame_search_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int]
        L.ame_device_result.argtypes = [C.c_void_p, C.c_int, C.POINTER(AmeResult)]
        L.ame_flush.argtypes = [C.c_void_p]
        L.ame_sync.argtypes = [C.c_void_p]
        L.ame_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.ame_timer_start.argtypes = [C.c_void_p]
        L.ame_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.ame_cu_geometry.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.ame_result_block_bytes.restype = C.c_uint64
        L.ame_result_block_bytes.argtypes = [C.c_void_p]
        L.ame_result_bind.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(AmeResult)]
        L.ame_exec_ns.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        _lib = L
    return _lib


class AmeError(RuntimeError):
    pass


def _check(rc):
    if rc < 0:
        raise AmeError("affine_me error %d: %s" % (rc, lib().ame_last_error().decode()))
    return rc


def num_ctus(W, H):
    return lib().ame_num_ctus(W, H)


def cu_geometry(pred, k):
    out = (C.c_int * 4)()
    g = lib().ame_cu_geometry(pred, k, out)
    return g, tuple(out[:])


class PinnedArray:
    """numpy view of cudaHostAlloc'ed memory."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(shape)) * self.dtype.itemsize
        self.ptr = lib().ame_alloc_host(self.nbytes)
        if not self.ptr:
            raise AmeError("ame_alloc_host(%d) failed" % self.nbytes)
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().ame_free_host(self.ptr)
            self.ptr = None


class HostResult:
    """Pinned destination of one search: cost[p] int64, cpmvs[p] CPMV_DTYPE, in one block laid out like the library's
    device-side result block (ame_result_bind), so that a search comes back with one device-to-host copy.
    separate=True allocates eight independent arrays instead (the general ame_result case)."""

    def __init__(self, ctx, separate=False):
        self._pins = []
        self.cost, self.cpmvs = [], []
        self.c = AmeResult()
        if separate:
            for p in range(4):
                n = ctx.result_len(p)
                a = PinnedArray((n,), np.int64)
                b = PinnedArray((n,), CPMV_DTYPE)
                self._pins += [a, b]
                self.cost.append(a.array)
                self.cpmvs.append(b.array)
                self.c.cost[p] = a.ptr
                self.c.cpmvs[p] = b.ptr
            return
        nbytes = int(lib().ame_result_block_bytes(ctx.h))
        blk = PinnedArray((nbytes,), np.uint8)
        self._pins.append(blk)
        _check(lib().ame_result_bind(ctx.h, blk.ptr, C.byref(self.c)))
        for p in range(4):
            n = ctx.result_len(p)
            o = self.c.cost[p] - blk.ptr
            self.cost.append(blk.array[o:o + 8 * n].view(np.int64))
            o = self.c.cpmvs[p] - blk.ptr
            self.cpmvs.append(blk.array[o:o + 28 * n].view(CPMV_DTYPE))

    def free(self):
        self.cost, self.cpmvs = [], []
        for p in self._pins:
            p.free()
        self._pins = []


class AffineME:
    """One context = one GPU.  Mirrors the C ABI one to one."""

    def __init__(self, width, height, device=0, num_slots=8, max_in_flight=8):
        self.W, self.H = width, height
        h = C.c_void_p()
        _check(lib().ame_create(C.byref(h), device, width, height, num_slots, max_in_flight))
        self.h = h
        self._keep = []
        if os.environ.get("AME_BIG_TMA", "") != "":  # development A/B switch (tools/, bench.py)
            self.set_option(OPT_BIG_TMA, int(os.environ["AME_BIG_TMA"]))
        if os.environ.get("AME_GROUP_BY_REF", "") != "":
            self.set_option(OPT_GROUP_BY_REF, int(os.environ["AME_GROUP_BY_REF"]))

    def close(self):
        if self.h:
            lib().ame_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def result_len(self, pred):
        return _check(lib().ame_result_len(self.h, pred))

    def set_option(self, opt, value):
        _check(lib().ame_set_option(self.h, opt, value))

    def upload(self, slot, plane, roles=ROLE_CURRENT | ROLE_REFERENCE):
        """plane: (H, W) uint16 numpy array or PinnedArray; kept alive until sync()."""
        arr = plane.array if isinstance(plane, PinnedArray) else np.ascontiguousarray(plane, dtype=np.uint16)
        assert arr.shape == (self.H, self.W), arr.shape
        self._keep.append(arr)
        _check(lib().ame_upload_plane_ex(self.h, slot, arr.ctypes.data, roles))

    def prepare(self, slot, roles):
        """Re-runs the plane preparation (block order / edge replication + horizontal filter stage) on the resident plane."""
        _check(lib().ame_prepare_plane(self.h, slot, roles))

    def search(self, cur_slot, ref_slot, lam, result, extra_iters=0):
        _check(lib().ame_search(self.h, cur_slot, ref_slot, C.c_float(lam), extra_iters, C.byref(result.c)))

    def search_device(self, cur_slot, ref_slot, lam, result_index, extra_iters=0):
        _check(lib().ame_search_device(self.h, cur_slot, ref_slot, C.c_float(lam), extra_iters, result_index))

    def flush(self):
        _check(lib().ame_flush(self.h))

    def sync(self):
        _check(lib().ame_sync(self.h))
        self._keep = []

    def last_kernel_ms(self):
        ms, n = C.c_float(), C.c_int()
        _check(lib().ame_last_kernel_ms(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def timer_start(self):
        _check(lib().ame_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        _check(lib().ame_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def exec_ns(self, reset=False):
        """Device nanoseconds per prediction type (FULL_2CP, FULL_3CP, HALF_2CP, HALF_3CP) since the last reset."""
        out = (C.c_double * 4)()
        _check(lib().ame_exec_ns(self.h, out, 1 if reset else 0))
        return list(out)

    def ref_pass(self, ref, cur, lam, extra_iters=0):
        """Convenience: one search on two host planes -> (costs[4], cpmvs[4]) numpy copies."""
        res = HostResult(self)
        try:
            self.upload(0, cur, ROLE_CURRENT)
            self.upload(1, ref, ROLE_REFERENCE)
            self.search(0, 1, lam, res, extra_iters)
            self.sync()
            return [c.copy() for c in res.cost], [m.copy() for m in res.cpmvs]
        finally:
            res.free()
