"""Pilot-sample covariance (kernel 4): the Gram step behind
``BLUEProblem.estimate_missing_covariances`` (blue_models.py:326-346, blue_fn.py:159-167)."""
import ctypes

import numpy as np

from ._lib import check, dptr, lib


def _producer_stream(Y):
    """The torch stream a device-resident Y was (possibly still is being) written on: the Gram kernel runs on its
    own stream and is ordered behind it (stream contract of the *_device entry points: work queued on the CURRENT
    torch stream of Y's device before the call is waited for)."""
    import torch
    return ctypes.c_void_p(int(torch.cuda.current_stream(Y.device).cuda_stream))


def pilot_sums(Y, device=0, telescoped=False, dist=None, group=None, return_ms=False):
    """Y: (n, N) or (n_out, n, N) samples -- numpy, or a contiguous CUDA float64 torch tensor.  One launch for all
    outputs.  Returns ``sums`` (n_out, N*N + N): per output the Gram matrix then the column sums (of Y, or of the
    telescoped differences Z_0 = Y_0, Z_j = Y_j - Y_{j-1} when ``telescoped``), and the total sample count.

    ``dist`` (torch.distributed, optional): every rank passes ITS rows of the sample matrix; the per-rank sums are
    added with one all-reduce of (N*N + N) n_out doubles over NCCL (the reduction of blue_fn.py:177-187)."""
    on_device = hasattr(Y, "data_ptr")
    if on_device:
        assert Y.is_cuda and Y.is_contiguous() and Y.dim() in (2, 3) and str(Y.dtype) == "torch.float64"
        shape = tuple(Y.shape)
        ptr = ctypes.c_void_p(int(Y.data_ptr()))
        device = Y.device.index or 0
        producer = _producer_stream(Y)
    else:
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        shape = Y.shape
        ptr = ctypes.c_void_p(Y.ctypes.data)
        producer = None
    n_out = 1 if len(shape) == 2 else int(shape[0])
    n, N = int(shape[-2]), int(shape[-1])
    ms = ctypes.c_float(0.0)
    world = dist.get_world_size(group) if dist is not None else 1
    if world > 1:
        import torch
        dsums = torch.empty((n_out, N * N + N), dtype=torch.float64, device="cuda:%d" % device)
        check(lib().blu_pilot_sums(device, ptr, n, N, n_out, 0, int(on_device), int(bool(telescoped)), producer,
                                   ctypes.c_void_p(int(dsums.data_ptr())), 1, ctypes.byref(ms)))
        cnt = torch.tensor([float(n)], dtype=torch.float64, device=dsums.device)
        dist.all_reduce(dsums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
        sums = dsums.cpu().numpy()
        n_total = int(round(float(cnt.item())))
    else:
        sums = np.empty((n_out, N * N + N))
        check(lib().blu_pilot_sums(device, ptr, n, N, n_out, 0, int(on_device), int(bool(telescoped)), producer,
                                   ctypes.c_void_p(sums.ctypes.data), 0, ctypes.byref(ms)))
        n_total = n
    if return_ms:
        return sums, n_total, ms.value
    return sums, n_total


def finalize_sums(sums, n_total, N, telescoped=False):
    """The reference's quantities from the (reduced) sums, per output: dict of arrays with a leading output axis --
    ``sumse`` (N), ``sumsc`` (N,N), ``C_hat`` (N,N) (blue_models.py:333), ``sumsd1`` / ``sumsd2`` (N,N; entries
    i < j, blue_fn.py:147-157) and ``dV`` (N,N; i < j, NaN elsewhere; blue_models.py:339)."""
    sums = np.ascontiguousarray(sums, dtype=np.float64).reshape(-1, N * N + N)
    No = sums.shape[0]
    out = {k: np.empty((No, N) if k == "sumse" else (No, N, N)) for k in ("sumse", "sumsc", "C_hat", "sumsd1", "sumsd2", "dV")}
    for o in range(No):
        check(lib().blu_pilot_finalize(dptr(sums[o]), int(n_total), int(N), int(bool(telescoped)), dptr(out["sumse"][o]), dptr(out["sumsc"][o]),
                                       dptr(out["C_hat"][o]), dptr(out["sumsd1"][o]), dptr(out["sumsd2"][o]), dptr(out["dV"][o])))
    return out


def pilot_statistics(Y, device=0, dist=None, group=None, telescoped=True):
    """Everything ``estimate_missing_covariances`` takes from ``blue_fn(..., compute_mlmc_differences=True)``
    (blue_models.py:331-339) for all outputs in one pass over the samples: see ``finalize_sums``."""
    sums, n_total = pilot_sums(Y, device=device, telescoped=telescoped, dist=dist, group=group)
    N = int(Y.shape[-1])
    return finalize_sums(sums, n_total, N, telescoped=telescoped)


class PilotAccumulator:
    """The accumulation loop of ``blue_fn`` (blue_fn.py:115-167) for model outputs that are produced in batches of N1
    samples and stay on the GPU: every ``add`` runs the Gram kernel on one batch and adds its (N*N + N) n_out sums to a
    device-resident running total -- nothing but those sums ever leaves HBM.  ``finalize`` all-reduces the totals over
    the ranks (blue_fn.py:177-187) and returns what ``finalize_sums`` returns.

        acc = PilotAccumulator(N, n_outputs)
        for it in range(0, n, N1):
            Ps = model(samples[it:it + N1])          # (n_outputs, N1, N) CUDA float64 tensor
            acc.add(Ps)
        stats = acc.finalize(dist)                   # sumse, sumsc, C_hat, sumsd1, sumsd2, dV
    """

    def __init__(self, N, n_outputs=1, device=0, telescoped=True):
        import torch
        self.N, self.n_out, self.device, self.telescoped = int(N), int(n_outputs), int(device), bool(telescoped)
        self.sums = torch.zeros((self.n_out, self.N * self.N + self.N), dtype=torch.float64, device="cuda:%d" % self.device)
        self._tmp = torch.empty_like(self.sums)
        self.n = 0

    def add(self, Y):
        """Y: (n_out, n_batch, N) or (n_batch, N) contiguous CUDA float64 tensor (host arrays are uploaded)."""
        import torch
        if not hasattr(Y, "data_ptr"):
            Y = torch.from_numpy(np.ascontiguousarray(Y, dtype=np.float64)).to(self.sums.device)
        assert Y.is_cuda and Y.is_contiguous() and str(Y.dtype) == "torch.float64"
        shape = tuple(Y.shape)
        n_out = 1 if len(shape) == 2 else int(shape[0])
        if n_out != self.n_out or int(shape[-1]) != self.N:
            raise ValueError("batch of shape %s does not match %d outputs x %d models" % (shape, self.n_out, self.N))
        nb = int(shape[-2])
        check(lib().blu_pilot_sums(self.device, ctypes.c_void_p(int(Y.data_ptr())), nb, self.N, self.n_out, 0, 1, int(self.telescoped),
                                   _producer_stream(Y), ctypes.c_void_p(int(self._tmp.data_ptr())), 1, None))
        self.sums += self._tmp                     # the call above returned after its own stream finished
        self.n += nb

    def finalize(self, dist=None, group=None):
        import torch
        sums, n = self.sums, self.n
        if dist is not None and dist.get_world_size(group) > 1:
            sums = sums.clone()
            cnt = torch.tensor([float(n)], dtype=torch.float64, device=sums.device)
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
            n = int(round(float(cnt.item())))
        return finalize_sums(sums.cpu().numpy(), n, self.N, telescoped=self.telescoped)


def pilot_covariance(Y, device=0, return_ms=False):
    """Y: (n, N) samples (numpy array, or a CUDA float64 torch tensor for a device-resident
    matrix).  Returns (sumse (N,), sumsc (N,N), C_hat (N,N)) with
    C_hat = sumsc/n - outer(sumse, sumse)/n^2   (blue_models.py:333, biased one-pass formula)."""
    sums, n_total, ms = pilot_sums(Y, device=device, telescoped=False, return_ms=True)
    N = int(Y.shape[-1])
    r = finalize_sums(sums, n_total, N, telescoped=False)
    if return_ms:
        return r["sumse"][0], r["sumsc"][0], r["C_hat"][0], ms
    return r["sumse"][0], r["sumsc"][0], r["C_hat"][0]


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
