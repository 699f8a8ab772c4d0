"""Pilot-sample covariance (kernel 4): the Gram step behind
``BLUEProblem.estimate_missing_covariances`` (blue_models.py:326-346, blue_fn.py:159-167)."""
import ctypes

import numpy as np

from ._lib import check, dptr, lib


def pilot_covariance(Y, device=0, return_ms=False):
    """Y: (n, N) samples (numpy array, or a CUDA float64 torch tensor for a device-resident
    matrix).  Returns (sumse (N,), sumsc (N,N), C_hat (N,N)) with
    C_hat = sumsc/n - outer(sumse, sumse)/n^2   (blue_models.py:333, biased one-pass formula)."""
    on_device = hasattr(Y, "data_ptr")
    if on_device:
        assert Y.is_cuda and Y.is_contiguous() and Y.dim() == 2 and str(Y.dtype) == "torch.float64"
        n, N = Y.shape
        ptr = ctypes.c_void_p(int(Y.data_ptr()))
        device = Y.device.index or 0
    else:
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        n, N = Y.shape
        ptr = ctypes.c_void_p(Y.ctypes.data)
    s1 = np.empty(N); S2 = np.empty((N, N)); Ch = np.empty((N, N))
    ms = ctypes.c_float(0.0)
    check(lib().blu_pilot_covariance(device, ptr, int(n), int(N), int(on_device), dptr(s1), dptr(S2), dptr(Ch), ctypes.byref(ms)))
    if return_ms:
        return s1, S2, Ch, ms.value
    return s1, S2, Ch


def fill_missing_covariances(adjacency, C_hat, rtol=1.0e-7):
    """The bookkeeping that follows the Gram step in ``estimate_missing_covariances`` (blue_models.py:340-346):
    every edge of the model graph whose weight is still unknown (NaN in the adjacency, diagonal included)
    receives the pilot estimate ``C_hat[i,j]``; pairs whose estimated correlation is below ``rtol`` in
    magnitude are marked uncorrelated (``inf``).  Known entries, and 0 entries (models that are never
    coupled), are left alone.  Returns a new adjacency matrix."""
    A = np.array(adjacency, dtype=np.float64, copy=True)
    C_hat = np.asarray(C_hat, dtype=np.float64)
    unknown = np.argwhere(np.isnan(A))
    for i, j in unknown:
        if i > j:
            continue
        value = C_hat[i, j]
        if abs(C_hat[i, j] / np.sqrt(C_hat[i, i] * C_hat[j, j])) < rtol:
            value = np.inf                                   # mark as uncorrelated
        A[i, j] = A[j, i] = value
    return A


def estimate_missing_covariances(Y, adjacency, device=0):
    """Pilot samples ``Y`` (n, N) -> the completed adjacency: Gram contraction on the device, then
    ``fill_missing_covariances``.  Returns (adjacency, C_hat)."""
    _, _, C_hat = pilot_covariance(Y, device=device)
    return fill_missing_covariances(adjacency, C_hat), C_hat
