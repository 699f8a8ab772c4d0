"""SAP -- the sample-allocation problem of one output, on the B200.

Host-side mirror of the reference's ``bluest.sap.SAP`` (sap.py:52-143) for the hot path:
same constructor, same attributes, same closures with the same return conventions
(``get_phi``, ``variance``, ``variance_GH``, ``get_cleanup_matrix``), but the per-group
inverses live packed in HBM inside a ``blu_ctx`` and every closure is one call into
libbluest_b200.so.  ``psi`` and ``invcovs`` are materialised lazily for the host-side SDP
builders that read them (sap.py:250,328).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import BluError, check, dptr, f64, iptr, lib
from .groups import indicator_ES


class SAP(object):
    def __init__(self, C, K, groups, costs, verbose=True, device=0, invcovs=None, pivot_rtol=1e-10):
        """C (N,N) covariance; ``groups[k-1]`` = list of sorted model-index lists of size k
        (converted in place to int64 arrays, like sap.py:77); ``costs`` (L,).

        ``invcovs`` (optional): the reference's per-class flat inverse arrays to ingest instead of
        inverting on the device (parity testing of the later stages)."""
        _lib.require_device()
        self.verbose = verbose
        self.C = C
        self.N = C.shape[0]
        self.K = K
        self.costs = costs
        self.samples = None
        self.budget = None
        self.eps = None
        self.tot_cost = None
        self.device = device

        sizes = [0] + [len(groupsk) for groupsk in groups]
        flat = []
        for k in range(1, K + 1):
            gk = groups[k - 1]
            if isinstance(gk, np.ndarray):                       # already an (Lk,k) index array: taken as it is
                arr = np.ascontiguousarray(gk, dtype=np.int64).reshape(len(gk), k) if len(gk) else np.array([], dtype=np.int64)
            else:
                arr = np.array(gk, dtype=np.int64).reshape(len(gk), k) if len(gk) else np.array(gk, dtype=np.int64)
            groups[k - 1] = arr                                  # sap.py:77 mutates the caller's list too
            if len(gk):
                flat.append(arr.ravel())
        self.sizes = sizes
        self.groups = groups
        self._flattened_groups = None                            # sap.py:66-71: built on first use (a Python list per group)
        self.cumsizes = np.cumsum(sizes)
        self.L = self.cumsizes[-1]

        sz = np.array(sizes[1:], dtype=np.int64)
        gflat = np.ascontiguousarray(np.concatenate(flat)) if flat else np.zeros(0, dtype=np.int64)
        self._ctx = ctypes.c_void_p()
        check(lib().blu_ctx_create(device, self.N, K, iptr(sz), iptr(gflat), ctypes.byref(self._ctx)))
        self.n_fallback = 0
        if invcovs is None:
            nf = ctypes.c_int64(0)
            Cd = np.ascontiguousarray(np.asarray(C, dtype=np.float64))
            check(lib().blu_ctx_set_covariance(self._ctx, dptr(Cd), float(pivot_rtol), ctypes.byref(nf)))
            self.n_fallback = int(nf.value)
        else:
            for k in range(1, K + 1):
                if sizes[k]:
                    a = f64(invcovs[k - 1], sizes[k] * k * k, "invcovs[%d]" % (k - 1))
                    check(lib().blu_ctx_set_invcovs(self._ctx, k, dptr(a)))
        self._invcovs = None
        self._psi = None

        self._ES = None                                          # sap.py:89-95: N vectors of length L, built on first use
        self.get_variance_functions()

    # ---- lazily materialised host views ------------------------------------------------------
    @property
    def flattened_groups(self):
        """sap.py:66-71: every group as a Python list, flat (size-major) order."""
        if self._flattened_groups is None:
            self._flattened_groups = [g.tolist() for gk in self.groups for g in np.asarray(gk).reshape(len(gk), -1)]
        return self._flattened_groups

    @property
    def ES(self):
        """sap.py:89-95: ES[i][g] = int(model i is in group g)."""
        if self._ES is None:
            self._ES = indicator_ES([g for g in self.groups], self.N)
        return self._ES

    @property
    def e(self):
        return self.ES[0]

    @property
    def invcovs(self):
        """``invcovs[k-1]``: flat (Lk*k*k) float64, row-major [i][j][l] (sap.py:78-79)."""
        if self._invcovs is None:
            out = []
            for k in range(1, self.K + 1):
                Lk = self.sizes[k]
                if Lk == 0:
                    out.append(np.array([]))
                    continue
                a = np.empty(Lk * k * k)
                check(lib().blu_ctx_get_invcovs(self._ctx, k, dptr(a)))
                out.append(a)
            self._invcovs = out
        return self._invcovs

    @property
    def psi(self):
        """Dense (N^2, L) matrix of sap.py:129, assembled on the device on first use."""
        if self._psi is None:
            psi = np.empty((self.N * self.N, int(self.L)))
            check(lib().blu_ctx_assemble_psi(self._ctx, dptr(psi)))
            self._psi = psi
        return self._psi

    def _m(self, m):
        return f64(m, int(self.L), "m")

    def get_variance_functions(self):
        ctx, L, N = self._ctx, int(self.L), self.N

        def get_phi(m, delta=0):
            phi = np.empty((N, N))
            check(lib().blu_get_phi(ctx, dptr(self._m(m)), float(delta), dptr(phi)))
            return phi

        def variance(m, delta=0):
            var = ctypes.c_double(0.0); fl = ctypes.c_uint(0)
            check(lib().blu_variance(ctx, dptr(self._m(m)), float(delta), ctypes.byref(var), ctypes.byref(fl)))
            if fl.value & _lib.FLAG_TINY:
                return np.inf                                   # misc.py:464
            # misc.py:470 -- the model 0 must always be sampled
            assert not (fl.value & _lib.FLAG_NO_MODEL0)
            return var.value

        def variance_GH(m, delta=0, nohess=False):
            var = ctypes.c_double(0.0); fl = ctypes.c_uint(0)
            grad = np.empty(L)
            # page-locked destination only where it pays (pinning costs ~0.1 ms/MB once; a pageable D2H of a
            # big Hessian runs several times slower than the 57 GB/s of a pinned one)
            hess = None if nohess else (_lib.pinned_pool.empty((L, L)) if L * L * 8 >= (1 << 20) else np.empty((L, L)))
            hp = None if nohess else ctypes.c_void_p(hess.ctypes.data)
            check(lib().blu_variance_GH(ctx, dptr(self._m(m)), float(delta), ctypes.byref(var), dptr(grad), hp, ctypes.byref(fl)))
            if fl.value & _lib.FLAG_TINY:
                return np.inf, grad                             # 2-tuple, misc.py:484 (grad is inf*ones)
            return var.value, grad, hess

        def variance_GH_operator(m, delta=0):
            """``variance_GH`` with the Hessian returned as a ``scipy.sparse.linalg.LinearOperator``:
            same (var, grad, hess) triple and the same 2-tuple early-out as misc.py:479-505, but
            ``hess`` stays factored in HBM (H = V U^T) and ``hess @ p`` costs two passes over
            L x N doubles on the device instead of 8 L^2 bytes over PCIe.  The operator is valid until the next
            evaluation of this SAP that asks for a Hessian."""
            var = ctypes.c_double(0.0); fl = ctypes.c_uint(0)
            grad = np.empty(L)
            check(lib().blu_variance_GH_factored(ctx, dptr(self._m(m)), float(delta), ctypes.byref(var), dptr(grad), ctypes.byref(fl)))
            if fl.value & _lib.FLAG_TINY:
                return np.inf, grad
            return var.value, grad, self.hess_operator()

        def get_cleanup_matrix(m, delta=0, corrected=False):
            X = np.empty((N, L)); fl = ctypes.c_uint(0)
            check(lib().blu_cleanup_matrix(ctx, dptr(self._m(m)), float(delta), 1 if corrected else 0, dptr(X), ctypes.byref(fl)))
            if fl.value & _lib.FLAG_TINY:
                raise ValueError("No entry greater or equal than 1 found in m.")     # misc.py:510
            return X

        self.get_phi = get_phi
        self.variance = variance
        self.variance_GH = variance_GH
        self.variance_GH_operator = variance_GH_operator
        self.get_cleanup_matrix = get_cleanup_matrix

    # ---- split evaluation (several contexts in flight at once, used by MOSAP) -----------------
    def variance_GH_begin(self, m, delta=0, nohess=False, grad=True):
        L = int(self.L)
        hess = None if nohess else (_lib.pinned_pool.empty((L, L)) if L * L * 8 >= (1 << 20) else np.empty((L, L)))
        hp = None if hess is None else ctypes.c_void_p(hess.ctypes.data)
        check(lib().blu_variance_GH_begin(self._ctx, dptr(self._m(m)), float(delta), int(bool(grad)), hp))
        self._pending = (hess, grad)

    def variance_GH_end(self):
        hess, want_grad = self._pending
        self._pending = None
        var = ctypes.c_double(0.0); fl = ctypes.c_uint(0)
        grad = np.empty(int(self.L)) if want_grad else None
        check(lib().blu_variance_GH_end(self._ctx, ctypes.byref(var), None if grad is None else dptr(grad), ctypes.byref(fl)))
        return var.value, grad, hess, fl.value

    # ---- Hessian as an operator ---------------------------------------------------------------
    def hess_matvec(self, p):
        """H @ p for the Hessian of the last factored evaluation; p (L,) or (L,nvec)."""
        L = int(self.L)
        p = np.asarray(p, dtype=np.float64)
        if p.ndim == 1:
            pv = f64(p, L, "p"); nvec = 1
        else:
            if p.shape[0] != L:
                raise ValueError("p has %d rows, expected %d" % (p.shape[0], L))
            pv = np.ascontiguousarray(p.T); nvec = p.shape[1]
        out = np.empty((nvec, L))
        check(lib().blu_hess_matvec(self._ctx, dptr(pv), nvec, dptr(out)))
        return out[0] if p.ndim == 1 else np.ascontiguousarray(out.T)

    def hess_operator(self):
        """LinearOperator over the factors resident in HBM (symmetric: rmatvec = matvec)."""
        from scipy.sparse.linalg import LinearOperator
        L = int(self.L)
        return LinearOperator((L, L), matvec=self.hess_matvec, rmatvec=self.hess_matvec, matmat=self.hess_matvec,
                              rmatmat=self.hess_matvec, dtype=np.float64)

    def hess_matvec_device(self, d_p, d_out):
        """Asynchronous ``d_out = H @ d_p`` with both vectors in HBM (torch tensors or raw pointers)."""
        ptr = lambda t: ctypes.c_void_p(int(t.data_ptr()) if hasattr(t, "data_ptr") else int(t))
        check(lib().blu_hess_matvec_device(self._ctx, ptr(d_p), ptr(d_out)))

    # ---- device-resident interface (no host round trip of the big arrays) ---------------------
    def eval_device(self, d_m=None, delta=0.0, grad=True, hess=False):
        """Asynchronous evaluation with m already in HBM (``d_m``: device pointer / torch tensor /
        None to reuse the last uploaded m).  Results stay on the device (``device_buffer``)."""
        ptr = None
        if d_m is not None:
            ptr = ctypes.c_void_p(int(d_m.data_ptr()) if hasattr(d_m, "data_ptr") else int(d_m))
        check(lib().blu_eval_device(self._ctx, ptr, float(delta), int(bool(grad)), int(bool(hess))))

    # ---- structure-exploiting KKT solve for the SDP solvers (row f1) ----------------------------
    def sdp_linear_rows(self, budget_mode=True, max_model_samples=None):
        """The dense rows of the SDP's linear cone as sap.py:259-287 builds them (without the -I block) and the scale
        factor of sap.py:258: ``(Gx (nlin, n), scales, has_t)``, n = L + has_t."""
        L = int(self.L)
        w = np.asarray(self.costs, dtype=np.float64)
        e = np.asarray(self.e, dtype=np.float64)
        es, _ = self.get_max_sample_constraints(max_model_samples)
        scales = 1.0 / np.abs(self.psi).sum(axis=0).mean()
        if budget_mode:
            rows = [np.concatenate([[0.0], w]), -np.concatenate([[0.0], e])] + [-np.concatenate([[0.0], -np.asarray(ee, dtype=np.float64)]) for ee in es]
        else:
            rows = [-e] + [np.asarray(ee, dtype=np.float64) for ee in es]
        n = L + (1 if budget_mode else 0)
        return np.ascontiguousarray(np.array(rows, dtype=np.float64).reshape(len(rows), n)), float(scales), (1 if budget_mode else 0)

    def kkt_solve(self, has_t, scales, Gx, d, r, bx, bz, return_ms=False):
        """One KKT solve of the SDP's interior-point iteration on the device (``blu_kkt_solve``): the system
        [0 G^T; G -W^T W][ux; uz] = [bx; bz] with G = [-I; Gx; G1], W from the Nesterov-Todd scalings ``d`` (linear
        cone, length n + nlin) and ``r`` ((N+1, N+1), semidefinite block).  Returns (ux, uz)."""
        L, M = int(self.L), self.N + 1
        n = L + int(has_t)
        Gx = np.ascontiguousarray(Gx, dtype=np.float64).reshape(-1, n) if np.size(Gx) else np.zeros((0, n))
        nlin = Gx.shape[0]
        d = f64(d, n + nlin, "d"); r = f64(r, M * M, "r"); bx = f64(bx, n, "bx"); bz = f64(bz, n + nlin + M * M, "bz")
        ux = np.empty(n); uz = np.empty(n + nlin + M * M)
        ms = ctypes.c_float(0.0)
        check(lib().blu_kkt_solve(self._ctx, int(has_t), float(scales), int(nlin), dptr(Gx) if nlin else None, dptr(d), dptr(r), dptr(bx), dptr(bz),
                                  dptr(ux), dptr(uz), ctypes.byref(ms)))
        return (ux, uz, ms.value) if return_ms else (ux, uz)

    def clone(self):
        """A second evaluation lane on the same problem (``blu_ctx_clone``): shares this SAP's inverses and tables in
        HBM, owns its stream and per-evaluation buffers, so evaluations on the two overlap on the device.  Only the
        device-resident interface and the closures are meaningful on the clone; close it before this SAP."""
        import copy
        other = copy.copy(self)
        other._ctx = ctypes.c_void_p()
        check(lib().blu_ctx_clone(self._ctx, ctypes.byref(other._ctx)))
        other._parent = self                                     # keeps the owner of the shared buffers alive
        other._pending = None
        other.get_variance_functions()
        return other

    # ---- CUDA graphs over the device-resident calls --------------------------------------------
    def graph_begin(self):
        """Start recording the device-resident calls of this SAP (``eval_device``, ``save_result``,
        ``hess_matvec_device``, the sharded engine's calls) into a CUDA graph instead of running them."""
        check(lib().blu_ctx_graph_begin(self._ctx))

    def graph_end(self):
        """Stop recording; returns the graph id for ``graph_launch``."""
        gid = ctypes.c_int(-1)
        check(lib().blu_ctx_graph_end(self._ctx, ctypes.byref(gid)))
        return gid.value

    def graph_launch(self, graph_id, times=1):
        check(lib().blu_ctx_graph_launch(self._ctx, int(graph_id), int(times)))

    def save_result(self, d_var=None, d_flags=None):
        """Stream-ordered copy of the last enqueued evaluation's variance / flags into device memory
        (torch tensors or raw pointers)."""
        ptr = lambda t: None if t is None else ctypes.c_void_p(int(t.data_ptr()) if hasattr(t, "data_ptr") else int(t))
        check(lib().blu_ctx_save_result(self._ctx, ptr(d_var), ptr(d_flags)))

    def set_grad_output(self, d_grad=None):
        """Gradient destination of the following device evaluations (None: the context's own buffer)."""
        ptr = None if d_grad is None else ctypes.c_void_p(int(d_grad.data_ptr()) if hasattr(d_grad, "data_ptr") else int(d_grad))
        check(lib().blu_ctx_set_grad_output(self._ctx, ptr))

    def upload_m(self, m):
        """Copy a host m into the context's device buffer (BLU_BUF_M)."""
        import torch
        t = self.device_buffer(_lib.BUF_M)
        t.copy_(torch.from_numpy(self._m(m)))
        return t

    def sync(self):
        check(lib().blu_ctx_sync(self._ctx))

    def last_result(self):
        var = ctypes.c_double(0.0); fl = ctypes.c_uint(0)
        check(lib().blu_ctx_last_result(self._ctx, ctypes.byref(var), ctypes.byref(fl)))
        return var.value, fl.value

    def last_timing(self):
        ms = (ctypes.c_float * 4)()
        check(lib().blu_ctx_last_timing(self._ctx, ms))
        return {"phi_pinv_ms": ms[0], "grad_ms": ms[1], "hess_ms": ms[2], "total_ms": ms[3]}

    def timing_log(self, capacity):
        """Log the phase events of the next ``capacity`` device evaluations (no sync in between)."""
        check(lib().blu_ctx_timing_log(self._ctx, int(capacity)))
        self._timing_cap = int(capacity)

    def timing_read(self):
        """(n,4) array of ms: [phi+pinv, grad/U, Hessian, total] per logged evaluation."""
        cap = max(1, getattr(self, "_timing_cap", 1))
        buf = (ctypes.c_float * (4 * cap))(); n = ctypes.c_int(0)
        check(lib().blu_ctx_timing_read(self._ctx, buf, ctypes.byref(n)))
        return np.array(buf[:4 * n.value], dtype=np.float64).reshape(n.value, 4)

    def set_option(self, name, value):
        check(lib().blu_ctx_set_option(self._ctx, name.encode(), int(value)))

    def get_option(self, name):
        v = ctypes.c_int(0)
        check(lib().blu_ctx_get_option(self._ctx, name.encode(), ctypes.byref(v)))
        return v.value

    def last_launches(self):
        return int(lib().blu_ctx_last_launches(self._ctx))

    def device_ptr(self, which):
        p = ctypes.c_void_p(); n = ctypes.c_int64(0)
        check(lib().blu_ctx_device_ptr(self._ctx, which, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def device_buffer(self, which):
        """A torch float64 tensor aliasing one of the context's HBM buffers (no copy)."""
        import torch
        ptr, nbytes = self.device_ptr(which)
        if not ptr:
            raise BluError(_lib.BLU_ERR_STATE, "buffer %d not allocated yet" % which)

        class _Holder:       # __cuda_array_interface__ exporter keeping the SAP (hence the ctx) alive
            pass
        h = _Holder()
        h.owner = self
        h.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}
        return torch.as_tensor(h, device="cuda:%d" % self.device)

    def stream(self):
        p = ctypes.c_void_p()
        check(lib().blu_ctx_stream(self._ctx, ctypes.byref(p)))
        return p.value

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            lib().blu_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compute_BLUE_estimator(self, sums, samples=None):
        """sap.py:99-119 + misc.py:518-544.  ``sums[i]`` = the k sample sums of group i (flat group
        order, scalar outputs).  Returns (mu, var); inf when max|samples| < 0.05."""
        if samples is None:
            samples = self.samples
        flat = np.concatenate([np.asarray(s_, dtype=np.float64).ravel() for s_ in sums]) if len(sums) else np.zeros(0)
        expect = int(sum(self.sizes[k] * k for k in range(1, self.K + 1)))
        if flat.size != expect:
            raise ValueError("sums must hold k scalars per group (%d in total), got %d" % (expect, flat.size))
        flat = np.ascontiguousarray(flat)
        mu = ctypes.c_double(0.0); var = ctypes.c_double(0.0); fl = ctypes.c_uint(0)
        y = np.empty(self.N)
        check(lib().blu_blue_estimator(self._ctx, dptr(self._m(samples)), dptr(flat), ctypes.byref(mu), ctypes.byref(var), dptr(y), ctypes.byref(fl)))
        if fl.value & _lib.FLAG_TINY:
            return np.inf                                     # misc.py:519
        assert not (fl.value & _lib.FLAG_NO_MODEL0)           # misc.py:527
        self.last_y = y
        return mu.value, var.value

    # ---- host orchestration kept from the reference -------------------------------------------
    def integer_projection(self, samples, budget=None, eps=None, max_model_samples=None):
        """sap.py:145-187; the candidate variances are evaluated in one batched device call."""
        from .intproj import integer_projection
        return integer_projection(self, samples, budget=budget, eps=eps, max_model_samples=max_model_samples)

    def get_max_sample_constraints(self, max_model_samples):
        """sap.py:222-240."""
        from .constraints import max_sample_constraints
        return max_sample_constraints(self.ES, self.N, max_model_samples)

    def solve(self, budget=None, eps=None, solver="scipy", x0=None, continuous_relaxation=False, max_model_samples=None, solver_params=None, hess="dense", sparse_constraints=False):
        """Host-side driver kept from sap.py:189-220.  Only ``solver="scipy"`` (trust-constr
        iterating on the GPU closures) is provided here; the SDP solvers (cvxopt / cvxpy) and
        ipopt are third-party host code that consume ``self.psi`` and are not part of this package."""
        if budget is None and eps is None:
            raise ValueError("Need to specify either budget or RMSE tolerance")
        if solver != "scipy":
            raise ValueError("bluest_b200.SAP.solve provides solver='scipy'; for 'cvxopt'/'cvxpy'/'ipopt' hand "
                             "`sap.psi`, `sap.variance`, `sap.variance_GH` to the reference's own drivers (INTEGRATION.md)")
        samples = self.scipy_solve(budget=budget, eps=eps, x0=x0, max_model_samples=max_model_samples, hess=hess, sparse_constraints=sparse_constraints)
        if samples is None:
            self.samples = None
            return None
        if not continuous_relaxation:
            try:
                samples = self.integer_projection(samples, budget=budget, eps=eps, max_model_samples=max_model_samples)
            except AssertionError as e:                    # sap.py:209-213
                print(str(e))
                self.samples = None
                return None
        self.samples = samples
        self.budget = budget
        self.eps = eps
        self.tot_cost = samples @ self.costs
        return samples

    def scipy_solve(self, budget=None, eps=None, x0=None, max_model_samples=None, maxiter=1000, hess="dense", sparse_constraints=False):
        """sap.py:378-418 -- same constraints, tolerances and callbacks; the callbacks are the GPU closures.
        ``hess="operator"`` hands trust-constr the factored Hessian as a LinearOperator (it only ever
        multiplies by it) instead of the dense (L,L) array."""
        from .solvers import scipy_solve
        self.scipy_counters = {}
        res = scipy_solve(self, budget=budget, eps=eps, x0=x0, max_model_samples=max_model_samples, maxiter=maxiter,
                          verbose=self.verbose, counters=self.scipy_counters, hess=hess, sparse_constraints=sparse_constraints)
        self.scipy_result = res
        return res.x
