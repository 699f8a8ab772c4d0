"""Group-sharded evaluation across the GPUs of one box.

Each rank owns a contiguous, work-balanced slice of the flat group enumeration and the packed
inverses of that slice only.  One evaluation is three local phases with two small exchanges:

    shard_phi      partial Phi of the slice                     (kernel 2, local)
    ALL-REDUCE     N*N (+33 indicator) doubles                  <- the one real exchange of the path
    shard_finish   delta*I, pinv, variance (redundant on every rank); gradient, U, V of the slice
    ALL-GATHER     U, V rows (N*L doubles each) -- only when the dense Hessian is wanted
    shard_hess     an equal-rows panel of H against all columns (kernel 3b, local; H stays sharded)

The Hessian OPERATOR needs no gather at all (and is the only form of the Hessian at N = 20):

    evaluate_factors(m)        U, V rows of the own slice stay on their rank
    hess_matvec(p):  hv_partial   t_r = sum over own rows of p_i u_i            (local)
                     ALL-REDUCE   32 doubles
                     hv_apply     (H p)_i = v_i . t for the own rows            (local)

The choreography is engine-agnostic: ``GpuEngine`` drives a ``blu_ctx`` through the C ABI and
``torch.distributed`` (NCCL over NVLink); the CPU tests plug a numpy engine into the same class and
run it under gloo with world_size 2.
"""
import ctypes

import numpy as np

from . import _lib
from .groups import balanced_slices, stream_cost


class ShardedEvaluator:
    def __init__(self, engine, sizes, rank, world, dist=None, group=None, fused=False, replicate_front=False, weight=stream_cost,
                 set_slice=True):
        """engine: object with shard_phi / shard_finish / shard_hess / buffers (see GpuEngine).
        sizes = [L1..LK].  fused=True: the Phi all-reduce runs inside the finish kernel over NVLink
        peer memory (engine.connect_peers) instead of a separate NCCL call.
        set_slice=False: the engine is a second evaluation lane (``SAP.clone()``) that already carries its
        parent's slice; only the peer connection of the lane is made."""
        self.engine = engine
        self.fused = bool(fused)
        # replicate_front: every rank evaluates Phi, pinv, gradient and U,V for ALL groups (tens of
        # microseconds where a dense Hessian exists at all, N <= 17) and only the Hessian -- the
        # 8 L^2-byte part -- is split into row panels: no collective in the whole evaluation.
        self.replicate_front = bool(replicate_front)
        self.rank, self.world = rank, world
        self.dist, self.group = dist, group
        self.L = int(sum(sizes))
        self.slices = balanced_slices(list(sizes), world, weight=weight)
        self.lo, self.hi = self.slices[rank]
        self.max_rows = max(hi - lo for lo, hi in self.slices)
        # Hessian row panels: equal row counts (equal bytes written), independent of the group slices
        cuts = [(self.L * r) // world for r in range(world + 1)]
        self.row_slices = [(cuts[r], cuts[r + 1]) for r in range(world)]
        self.rlo, self.rhi = self.row_slices[rank]
        if self.replicate_front:
            self.slices = [(0, self.L)] * world
            self.lo, self.hi = 0, self.L
        else:
            if set_slice:
                engine.set_slice(self.lo, self.hi)
            if self.fused:
                engine.connect_peers(rank, world, dist, group)

    def _phi_exchange_finish(self, m, delta, want_grad, want_uv):
        e = self.engine
        if self.fused:
            e.shard_eval_fused(m, delta, want_grad, want_uv)       # one kernel: reduce + NVLink all-reduce + pinv
        else:
            buf = e.shard_phi(m)                                  # (N*N + 40) partial sums + indicators
            self._all_reduce(buf)
            e.shard_finish(delta, want_grad, want_uv)

    # -- collectives (no-ops at world == 1) --------------------------------------------------
    def _all_reduce(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)

    def _all_gather_rows(self, full, width):
        """Every rank has filled rows [lo,hi) of ``full`` (rows x width, row-major, flat tensor);
        fetch the other ranks' rows.  The slices are uneven (balanced by work, not by count), so
        instead of a padded all-gather each owner broadcasts its slice IN PLACE: the views are
        contiguous, nothing is staged or copied."""
        if self.world == 1:
            return
        full2 = full[: (full.numel() // width) * width].view(-1, width)
        for r, (lo, hi) in enumerate(self.slices):
            if hi > lo:
                src = self.dist.get_global_rank(self.group, r) if self.group is not None else r
                self.dist.broadcast(full2[lo:hi], src=src, group=self.group)

    # -- one evaluation ----------------------------------------------------------------------
    def evaluate(self, m, delta=0.0, grad=True, hess=False, gather_grad=True):
        """m: full-length sample vector (every rank passes the same values).  Returns a dict with
        var, flags, and -- as the engine's buffers -- grad (full length if gather_grad, else only the
        own slice is valid) and the own Hessian row panel when hess=True."""
        e = self.engine
        if self.replicate_front:
            with e.stream_context():
                e.eval_full(m, delta, grad, hess)
                if hess:
                    e.shard_hess(self.rlo, self.rhi)
            var, flags = e.result()
            return dict(var=var, flags=flags, lo=self.lo, hi=self.hi, rlo=self.rlo, rhi=self.rhi)
        with e.stream_context():
            self._phi_exchange_finish(m, delta, grad or hess, hess)
            out = {}
            if grad and gather_grad:
                self._all_gather_rows(e.grad_buffer(), 1)
            if hess:
                self._all_gather_rows(e.u_buffer(), e.NP)
                self._all_gather_rows(e.v_buffer(), e.NP)
                e.shard_hess(self.rlo, self.rhi)
        var, flags = e.result()
        out.update(var=var, flags=flags, lo=self.lo, hi=self.hi, rlo=self.rlo, rhi=self.rhi)
        return out

    def evaluate_factors(self, m, delta=0.0):
        """Variance, gradient (own slice) and the U, V rows of the own slice: everything
        ``hess_matvec`` needs.  No gather -- the factors stay sharded."""
        e = self.engine
        with e.stream_context():
            if self.replicate_front:
                e.eval_full(m, delta, True, 2)                     # 2: U rows only -- the operator never reads V
            else:
                self._phi_exchange_finish(m, delta, True, 2)
        var, flags = e.result()
        return dict(var=var, flags=flags, lo=self.lo, hi=self.hi)

    def hess_matvec(self, p, gather=True):
        """H @ p with the Hessian factored and sharded by rows.  p: full-length vector (every rank
        passes the same values; device tensor, or numpy for a CPU engine).  Returns the engine's
        output buffer: rows [lo,hi) valid, all rows when ``gather``."""
        e = self.engine
        with e.stream_context():
            t = e.hv_partial(p)                                    # 32 doubles
            if not self.replicate_front:
                self._all_reduce(t)
            out = e.hv_apply(t)
            if gather and not self.replicate_front:
                self._all_gather_rows(out, 1)
        return out

    def evaluate_async(self, m, delta=0.0, grad=True, hess=False, gather_grad=True):
        """Same as evaluate() but without the final read-back of (var, flags): nothing on the host
        waits for the device, so evaluations can be queued back to back."""
        e = self.engine
        if self.replicate_front:
            with e.stream_context():
                e.eval_full(m, delta, grad, hess)
                if hess:
                    e.shard_hess(self.rlo, self.rhi)
            return None
        with e.stream_context():
            self._phi_exchange_finish(m, delta, grad or hess, hess)
            if grad and gather_grad:
                self._all_gather_rows(e.grad_buffer(), 1)
            if hess:
                self._all_gather_rows(e.u_buffer(), e.NP)
                self._all_gather_rows(e.v_buffer(), e.NP)
                e.shard_hess(self.rlo, self.rhi)
        return None


class GpuEngine:
    """blu_ctx-backed engine: every rank builds the context for ALL groups' tables (a few MB of
    ids/masks) but evaluates only its slice; buffers are torch tensors aliasing HBM."""

    def __init__(self, sap):
        import torch
        self.torch = torch
        self.sap = sap
        self.N = sap.N
        self.NP = 4 * ((sap.N + 3) // 4)
        self._ext = torch.cuda.ExternalStream(sap.stream(), device=sap.device)

    def stream_context(self):
        return self.torch.cuda.stream(self._ext)

    def set_slice(self, lo, hi):
        _lib.check(_lib.lib().blu_ctx_set_slice(self.sap._ctx, int(lo), int(hi)))

    def connect_peers(self, rank, world, dist=None, group=None):
        """Exchange the CUDA-IPC handles of the per-rank exchange buffers and map the peers."""
        h = ctypes.create_string_buffer(64)
        _lib.check(_lib.lib().blu_ctx_peer_handle(self.sap._ctx, h))
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, bytes(h.raw), group=group)
            blob = b"".join(handles)
        else:
            blob = bytes(h.raw)
        _lib.check(_lib.lib().blu_ctx_peer_connect(self.sap._ctx, int(world), int(rank), blob))

    def _m_ptr(self, m):
        if m is None:
            return None
        if hasattr(m, "data_ptr"):
            return ctypes.c_void_p(int(m.data_ptr()))
        self.sap.device_buffer(_lib.BUF_M).copy_(self.torch.from_numpy(np.ascontiguousarray(m, dtype=np.float64)), non_blocking=True)
        return None

    def eval_full(self, m, delta, want_grad, want_uv):
        """Whole-problem evaluation on this rank (no slice): Phi, pinv, variance, gradient and, if
        want_uv, the U,V factors -- but not the Hessian."""
        _lib.check(_lib.lib().blu_eval_device(self.sap._ctx, self._m_ptr(m), float(delta), int(bool(want_grad)), {0: 0, 1: 2, 2: 3}[int(want_uv)]))

    def shard_eval_fused(self, m, delta, want_grad, want_uv):
        _lib.check(_lib.lib().blu_shard_eval_fused(self.sap._ctx, self._m_ptr(m), float(delta), int(bool(want_grad)), int(want_uv)))

    def shard_phi(self, m):
        if m is not None:
            if hasattr(m, "data_ptr"):
                ptr = ctypes.c_void_p(int(m.data_ptr()))
            else:
                self.sap.device_buffer(_lib.BUF_M).copy_(self.torch.from_numpy(np.ascontiguousarray(m, dtype=np.float64)), non_blocking=True)
                ptr = None
        else:
            ptr = None
        _lib.check(_lib.lib().blu_shard_phi(self.sap._ctx, ptr))
        return self.sap.device_buffer(_lib.BUF_PHI)

    def shard_finish(self, delta, want_grad, want_uv):
        _lib.check(_lib.lib().blu_shard_finish(self.sap._ctx, float(delta), int(bool(want_grad)), int(want_uv)))

    def shard_hess(self, rlo, rhi):
        _lib.check(_lib.lib().blu_shard_hess(self.sap._ctx, int(rlo), int(rhi)))

    def hv_partial(self, p):
        """t = sum over the owned rows of p_i u_i  ->  (32,) device tensor (zero beyond N)."""
        torch = self.torch
        if not hasattr(self, "_hv_t"):
            dev = "cuda:%d" % self.sap.device
            self._hv_t = torch.zeros(32, dtype=torch.float64, device=dev)
            self._hv_out = torch.zeros(int(self.sap.L), dtype=torch.float64, device=dev)
            self._hv_p = torch.zeros(int(self.sap.L), dtype=torch.float64, device=dev)
        if not hasattr(p, "data_ptr"):
            self._hv_p.copy_(torch.from_numpy(np.ascontiguousarray(p, dtype=np.float64)), non_blocking=True)
            p = self._hv_p
        _lib.check(_lib.lib().blu_shard_hv_partial(self.sap._ctx, ctypes.c_void_p(int(p.data_ptr())), ctypes.c_void_p(int(self._hv_t.data_ptr()))))
        return self._hv_t

    def hv_apply(self, t):
        _lib.check(_lib.lib().blu_shard_hv_apply(self.sap._ctx, ctypes.c_void_p(int(t.data_ptr())), ctypes.c_void_p(int(self._hv_out.data_ptr()))))
        return self._hv_out

    def grad_buffer(self):
        return self.sap.device_buffer(_lib.BUF_GRAD)

    def u_buffer(self):
        return self.sap.device_buffer(_lib.BUF_U)

    def v_buffer(self):
        return self.sap.device_buffer(_lib.BUF_V)

    def hess_panel(self):
        return self.sap.device_buffer(_lib.BUF_HESS)

    def result(self):
        return self.sap.last_result()
