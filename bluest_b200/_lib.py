"""ctypes binding of libbluest_b200.so (include/bluest_b200.h).

The library is the product: if it is missing, or no CUDA device is visible, every call
raises -- there is no CPU fallback anywhere in this package.
"""
import ctypes
import os
import weakref

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbluest_b200.so")

BLU_OK, BLU_ERR_ARG, BLU_ERR_CUDA, BLU_ERR_STATE, BLU_ERR_NOMEM, BLU_ERR_NODEVICE = range(6)
FLAG_TINY, FLAG_NO_MODEL0, FLAG_PARTIAL, FLAG_PEER_TIMEOUT = 1, 2, 4, 8
BUF_M, BUF_PHI, BUF_PINV, BUF_GRAD, BUF_U, BUF_V, BUF_HESS, BUF_CINV, BUF_SCAL = range(9)

c_int, c_i64, c_dbl, c_uint = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_uint
p_dbl, p_i64, p_void = ctypes.POINTER(c_dbl), ctypes.POINTER(c_i64), ctypes.c_void_p

# every symbol declared in include/bluest_b200.h: (restype, argtypes)
SIGNATURES = {
    "blu_last_error": (ctypes.c_char_p, []),
    "blu_device_count": (c_int, []),
    "blu_version": (ctypes.c_char_p, []),
    "blu_ctx_create": (c_int, [c_int, c_int, c_int, p_i64, p_i64, ctypes.POINTER(p_void)]),
    "blu_ctx_destroy": (c_int, [p_void]),
    "blu_ctx_clone": (c_int, [p_void, ctypes.POINTER(p_void)]),
    "blu_ctx_set_covariance": (c_int, [p_void, p_dbl, c_dbl, p_i64]),
    "blu_ctx_set_invcovs": (c_int, [p_void, c_int, p_dbl]),
    "blu_ctx_get_invcovs": (c_int, [p_void, c_int, p_dbl]),
    "blu_ctx_assemble_psi": (c_int, [p_void, p_dbl]),
    "blu_get_phi": (c_int, [p_void, p_dbl, c_dbl, p_dbl]),
    "blu_variance": (c_int, [p_void, p_dbl, c_dbl, p_dbl, ctypes.POINTER(c_uint)]),
    "blu_variance_GH": (c_int, [p_void, p_dbl, c_dbl, p_dbl, p_dbl, p_void, ctypes.POINTER(c_uint)]),
    "blu_variance_GH_begin": (c_int, [p_void, p_dbl, c_dbl, c_int, p_void]),
    "blu_variance_GH_end": (c_int, [p_void, p_dbl, p_dbl, ctypes.POINTER(c_uint)]),
    "blu_cleanup_matrix": (c_int, [p_void, p_dbl, c_dbl, c_int, p_dbl, ctypes.POINTER(c_uint)]),
    "blu_blue_estimator": (c_int, [p_void, p_dbl, p_dbl, p_dbl, p_dbl, p_dbl, ctypes.POINTER(c_uint)]),
    "blu_candidate_variances": (c_int, [p_void, p_dbl, c_int, p_i64, p_i64, c_i64, c_dbl, p_dbl]),
    "blu_batch_create": (c_int, [ctypes.POINTER(p_void), c_int, ctypes.POINTER(p_i64), c_i64, ctypes.POINTER(p_void)]),
    "blu_batch_eval": (c_int, [p_void, p_dbl, c_int, c_dbl, p_dbl, ctypes.POINTER(c_uint), p_dbl]),
    "blu_batch_destroy": (c_int, [p_void]),
    "blu_kkt_solve": (c_int, [p_void, c_int, c_dbl, c_int, p_dbl, p_dbl, p_dbl, p_dbl, p_dbl, p_dbl, p_dbl, ctypes.POINTER(ctypes.c_float)]),
    "blu_eval_device": (c_int, [p_void, p_void, c_dbl, c_int, c_int]),
    "blu_ctx_sync": (c_int, [p_void]),
    "blu_ctx_device_ptr": (c_int, [p_void, c_int, ctypes.POINTER(p_void), p_i64]),
    "blu_ctx_stream": (c_int, [p_void, ctypes.POINTER(p_void)]),
    "blu_ctx_last_result": (c_int, [p_void, p_dbl, ctypes.POINTER(c_uint)]),
    "blu_ctx_last_timing": (c_int, [p_void, ctypes.POINTER(ctypes.c_float)]),
    "blu_ctx_last_launches": (c_int, [p_void]),
    "blu_ctx_set_option": (c_int, [p_void, ctypes.c_char_p, c_int]),
    "blu_ctx_get_option": (c_int, [p_void, ctypes.c_char_p, ctypes.POINTER(c_int)]),
    "blu_ctx_timing_log": (c_int, [p_void, c_int]),
    "blu_ctx_timing_read": (c_int, [p_void, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(c_int)]),
    "blu_ctx_last_stamps": (c_int, [p_void, ctypes.POINTER(ctypes.c_uint64)]),
    "blu_ctx_graph_begin": (c_int, [p_void]),
    "blu_ctx_graph_end": (c_int, [p_void, ctypes.POINTER(c_int)]),
    "blu_ctx_graph_launch": (c_int, [p_void, c_int, c_int]),
    "blu_ctx_save_result": (c_int, [p_void, p_void, p_void]),
    "blu_ctx_set_grad_output": (c_int, [p_void, p_void]),
    "blu_ctx_set_slice": (c_int, [p_void, c_i64, c_i64]),
    "blu_ctx_peer_handle": (c_int, [p_void, p_void]),
    "blu_ctx_peer_connect": (c_int, [p_void, c_int, c_int, p_void]),
    "blu_shard_eval_fused": (c_int, [p_void, p_void, c_dbl, c_int, c_int]),
    "blu_shard_phi": (c_int, [p_void, p_void]),
    "blu_shard_finish": (c_int, [p_void, c_dbl, c_int, c_int]),
    "blu_shard_hess": (c_int, [p_void, c_i64, c_i64]),
    "blu_shard_hv_partial": (c_int, [p_void, p_void, p_void]),
    "blu_shard_hv_apply": (c_int, [p_void, p_void, p_void]),
    "blu_variance_GH_factored": (c_int, [p_void, p_dbl, c_dbl, p_dbl, p_dbl, ctypes.POINTER(c_uint)]),
    "blu_hess_matvec": (c_int, [p_void, p_dbl, c_int, p_dbl]),
    "blu_hess_matvec_device": (c_int, [p_void, p_void, p_void]),
    "blu_pilot_covariance": (c_int, [c_int, p_void, c_i64, c_int, c_int, p_dbl, p_dbl, p_dbl, ctypes.POINTER(ctypes.c_float)]),
    "blu_pilot_sums": (c_int, [c_int, p_void, c_i64, c_int, c_int, c_i64, c_int, c_int, p_void, p_void, c_int, ctypes.POINTER(ctypes.c_float)]),
    "blu_pilot_finalize": (c_int, [p_dbl, c_i64, c_int, c_int, p_dbl, p_dbl, p_dbl, p_dbl, p_dbl, p_dbl]),
    "blu_assemble_psi_c": (c_int, [p_dbl, c_int, c_int, c_int, p_i64, p_dbl]),
    "blu_objectiveK_c": (c_int, [p_dbl, c_int, c_int, c_int, p_dbl, p_i64, p_dbl]),
    "blu_objectiveK_c_i64": (c_int, [p_dbl, c_int, c_int, c_int, p_i64, p_i64, p_dbl]),
    "blu_cleanupK_c": (c_int, [p_dbl, c_int, c_int, c_int, p_i64, p_dbl, p_dbl]),
    "blu_gradK_c": (c_int, [p_dbl, c_int, c_int, c_int, p_i64, p_dbl, p_dbl]),
    "blu_hessKQ_c": (c_int, [p_dbl, c_int, c_int, c_int, c_int, c_int, p_i64, p_i64, p_dbl, p_dbl, p_dbl]),
    "blu_host_alloc": (c_int, [ctypes.c_size_t, ctypes.POINTER(p_void)]),
    "blu_host_free": (c_int, [p_void]),
}


class BluError(RuntimeError):
    """A libbluest_b200 call returned a non-zero status."""

    def __init__(self, code, msg):
        super().__init__("libbluest_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError("bluest_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)          # AttributeError here == header/library mismatch
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != BLU_OK:
        raise BluError(rc, lib().blu_last_error().decode())


def device_count():
    return lib().blu_device_count()


def require_device():
    if device_count() <= 0:
        raise BluError(BLU_ERR_NODEVICE, "no CUDA device visible: bluest_b200 has no CPU fallback")


def dptr(a):
    return a.ctypes.data_as(p_dbl)


def iptr(a):
    return a.ctypes.data_as(p_i64)


def f64(a, n=None, name="array"):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.size != n:
        raise ValueError("%s has %d entries, expected %d" % (name, a.size, n))
    return a


class _PinnedBlock:
    def __init__(self, nbytes):
        self.ptr = p_void()
        check(lib().blu_host_alloc(nbytes, ctypes.byref(self.ptr)))
        self.nbytes = nbytes

    def free(self):
        if self.ptr:
            lib().blu_host_free(self.ptr)
            self.ptr = p_void()


class PinnedPool:
    """Page-locked host buffers for the dense Hessian (8 L^2 bytes): pinning 8.6 GB costs about a
    second, so blocks are recycled once the numpy array handed to the caller is garbage."""

    def __init__(self):
        self.free_blocks = {}

    def empty(self, shape):
        n = int(np.prod(shape)) * 8
        blocks = self.free_blocks.setdefault(n, [])
        blk = blocks.pop() if blocks else _PinnedBlock(max(n, 8))
        buf = (ctypes.c_char * max(n, 8)).from_address(blk.ptr.value)
        arr = np.frombuffer(buf, dtype=np.float64, count=int(np.prod(shape))).reshape(shape)
        weakref.finalize(buf, self._give_back, n, blk)
        return arr

    def _give_back(self, n, blk):
        # keep at least two blocks per size, and as many as fit in 256 MB (a MOSAP evaluation hands out
        # one Hessian per output; freeing and re-pinning them every call would cost milliseconds)
        lst = self.free_blocks.setdefault(n, [])
        if len(lst) < 2 or (len(lst) + 1) * n <= (256 << 20):
            lst.append(blk)
        else:
            blk.free()

    def clear(self):
        for lst in self.free_blocks.values():
            for blk in lst:
                blk.free()
        self.free_blocks = {}


pinned_pool = PinnedPool()
