"""Level-1 drop-ins for the reference's ``_cmisc_bluest`` module (bluest/cmisc.cpp:99-110) and
for the thin numpy wrappers around it in bluest/misc.py:600-629.

Same names, argument order and in-place "+=" convention as the reference; the work is done by
CUDA kernels behind libbluest_b200.so.  Unlike pybind11's silent copy on a dtype mismatch
(a float32 output array is silently NOT updated), wrong dtypes raise TypeError.
"""
import numpy as np

from ._lib import check, dptr, iptr, lib


def _out(a, name):
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.flags.writeable):
        raise TypeError("%s must be a writable C-contiguous float64 array (it is updated in place)" % name)
    return a


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _g(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def assemble_psi_c(psi, N, k, Lk, groupsk, invcovsk):
    groupsk, invcovsk = _g(groupsk), _f(invcovsk)
    check(lib().blu_assemble_psi_c(dptr(_out(psi, "psi")), int(N), int(k), int(Lk), iptr(groupsk), dptr(invcovsk)))


def objectiveK_c(PHI, N, k, Lk, mk, groupsk, invcovsk):
    groupsk, invcovsk = _g(groupsk), _f(invcovsk)
    mk = np.ascontiguousarray(mk)
    if mk.dtype == np.int64:                                  # the `long int` overload, cmisc.cpp:105
        check(lib().blu_objectiveK_c_i64(dptr(_out(PHI, "PHI")), int(N), int(k), int(Lk), iptr(mk), iptr(groupsk), dptr(invcovsk)))
    else:
        mk = _f(mk)
        check(lib().blu_objectiveK_c(dptr(_out(PHI, "PHI")), int(N), int(k), int(Lk), dptr(mk), iptr(groupsk), dptr(invcovsk)))


def cleanupK_c(X, k, Lk, groupsk, invcovsk, invPHI_0):
    groupsk, invcovsk, x = _g(groupsk), _f(invcovsk), _f(invPHI_0)
    check(lib().blu_cleanupK_c(dptr(_out(X, "X")), int(x.size), int(k), int(Lk), iptr(groupsk), dptr(invcovsk), dptr(x)))


def gradK_c(grad, k, Lk, groupsk, invcovsk, invPHI_0):
    groupsk, invcovsk, x = _g(groupsk), _f(invcovsk), _f(invPHI_0)
    check(lib().blu_gradK_c(dptr(_out(grad, "grad")), int(x.size), int(k), int(Lk), iptr(groupsk), dptr(invcovsk), dptr(x)))


def hessKQ_c(hess, N, k, q, Lk, Lq, groupsk, groupsq, invcovsk, invcovsq, invPHI):
    groupsk, groupsq = _g(groupsk), _g(groupsq)
    invcovsk, invcovsq, invPHI = _f(invcovsk), _f(invcovsq), _f(invPHI)
    check(lib().blu_hessKQ_c(dptr(_out(hess, "hess")), int(N), int(k), int(q), int(Lk), int(Lq), iptr(groupsk), iptr(groupsq),
                             dptr(invcovsk), dptr(invcovsq), dptr(invPHI)))


# ---- the wrappers of bluest/misc.py:600-629 ---------------------------------------------------
def assemble_psi(N, k, Lk, groupsk, invcovsk):
    psi = np.zeros((N * N, Lk), order='C')
    assemble_psi_c(psi.ravel(order='C'), N, k, Lk, np.asarray(groupsk).ravel(order='C'), invcovsk)
    return psi


def cleanupK(k, Lk, groupsk, invcovsk, invPHI):
    N = invPHI.shape[0]
    X = np.zeros((N, Lk), order='C')
    cleanupK_c(X.ravel(order='C'), k, Lk, np.asarray(groupsk).ravel(order='C'), invcovsk, invPHI[0])
    return X


def objectiveK(N, k, Lk, mk, groupsk, invcovsk):
    """Correct-arity version of misc.py:612-616 (the reference's wrapper drops ``N`` and raises)."""
    PHI = np.zeros((N * N,))
    objectiveK_c(PHI, N, k, Lk, mk, np.asarray(groupsk).ravel(order='C'), invcovsk)
    return PHI


def gradK(k, Lk, groupsk, invcovsk, invPHI):
    grad = np.zeros((Lk,))
    gradK_c(grad, k, Lk, np.asarray(groupsk).ravel(order='C'), invcovsk, invPHI[0])
    return grad


def hessKQ(k, q, Lk, Lq, groupsk, groupsq, invcovsk, invcovsq, invPHI):
    N = invPHI.shape[0]
    hess = np.zeros((Lk, Lq), order='C')
    hessKQ_c(hess.ravel(order='C'), N, k, q, Lk, Lq, np.asarray(groupsk).ravel(order='C'), np.asarray(groupsq).ravel(order='C'),
             invcovsk, invcovsq, np.ascontiguousarray(invPHI).ravel(order='C'))
    return hess
