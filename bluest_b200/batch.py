"""Batched evaluation of small problems: P problems x B sample vectors in one kernel launch (``blu_batch_*``).

The device form of the reference's Python loops over outputs (mosap.py:86-100) and over the instances of a budget /
tolerance sweep: every (problem, sample vector) pair is one CTA that assembles Phi, inverts it, takes the variance
and the gradient.  Used by ``MOSAP.variances`` / ``MOSAP.variance_GH(nohess=True)`` and ``evaluate_many``."""
import ctypes

import numpy as np

from . import _lib
from ._lib import check, dptr, lib


class Batch:
    def __init__(self, saps, maps=None, Lm=None):
        """saps: the problems (``SAP`` objects on one device).  maps: per problem the int64 index vector that picks
        its sample vector out of a shared one of length ``Lm`` (``MOSAP.mappings``); None: the input of one evaluation
        is the concatenation of the problems' own vectors."""
        self.saps = list(saps)
        self.P = len(self.saps)
        self.Ls = [int(s.L) for s in self.saps]
        self.pre = np.concatenate([[0], np.cumsum(self.Ls)]).astype(np.int64)
        arr = (ctypes.c_void_p * self.P)(*[s._ctx for s in self.saps])
        self._maps = None
        mp = None
        if maps is not None:
            self._maps = [np.ascontiguousarray(mm, dtype=np.int64) for mm in maps]
            mp = (ctypes.POINTER(ctypes.c_int64) * self.P)(*[mm.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)) for mm in self._maps])
        self.Lm = int(Lm) if maps is not None else int(self.pre[-1])
        self._h = ctypes.c_void_p()
        check(lib().blu_batch_create(arr, self.P, mp, self.Lm, ctypes.byref(self._h)))

    def eval(self, M, delta=0.0, grad=True):
        """M: (B, Lm) or (Lm,) sample vectors.  Returns (var (P,B), flags (P,B), grads) with ``grads[p]`` a (B, L_p)
        array (None when ``grad`` is False).  ``var`` is inf and the gradient row inf where max|m| < 0.05 (misc.py:464,484)."""
        M = np.ascontiguousarray(np.atleast_2d(np.asarray(M, dtype=np.float64)))
        if M.shape[1] != self.Lm:
            raise ValueError("sample vectors have %d entries, expected %d" % (M.shape[1], self.Lm))
        B = M.shape[0]
        var = np.empty((self.P, B)); flags = np.empty((self.P, B), dtype=np.uint32)
        g = np.empty(int(self.pre[-1]) * B) if grad else None
        check(lib().blu_batch_eval(self._h, dptr(M), B, float(delta), dptr(var), flags.ctypes.data_as(ctypes.POINTER(ctypes.c_uint)),
                                   dptr(g) if grad else None))
        grads = None
        if grad:
            grads = [g[int(self.pre[p]) * B:int(self.pre[p + 1]) * B].reshape(B, self.Ls[p]) for p in range(self.P)]
        return var, flags, grads

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().blu_batch_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def evaluate_many(sap, M, delta=0.0, grad=True):
    """B sample vectors of ONE problem in one launch (the instances of a budget sweep, the trial points of a line
    search): returns (var (B,), flags (B,), grad (B, L) or None)."""
    b = getattr(sap, "_batch1", None)
    if b is None:
        b = sap._batch1 = Batch([sap])
    var, flags, grads = b.eval(M, delta=delta, grad=grad)
    return var[0], flags[0], (grads[0] if grad else None)
