"""MOSAP -- multi-output wrapper (mirror of bluest/mosap.py:18-123 for the hot path).

One ``SAP`` (device context) per output, evaluated on ``m[mappings[n]]``; same constructor
signature, attributes (``SAPS``, ``mappings``, ``ES``, ``e``, ``sizes``, ``cumsizes``, ``L``) and
methods ``variances`` / ``variance_GH`` / ``get_cleanup_matrices`` as the reference.
"""
import numpy as np

from .groups import indicator_ES, mappings as build_mappings
from .sap import SAP


class BLUESTError(RuntimeError):
    pass


class MOSAP(object):
    ''' MOSAP, MultiObjectiveSampleAllocationProblem '''

    def __init__(self, C, K, Ks, groups, multi_groups, costs, multi_costs, verbose=True, device=0):
        """Same arguments as mosap.py:20: per-output covariances ``C``, the union ``groups`` (size-major) and the
        per-output ``multi_groups`` / ``multi_costs``.  ``groups[k]`` is converted to an int64 array in place
        (mosap.py:34), one device context per output is created, and ``mappings`` ties the two numberings."""
        self.verbose, self.n_outputs = verbose, len(C)
        self.C, self.N = C, C[0].shape[0]
        self.K, self.Ks = K, Ks
        self.costs, self.multi_groups, self.multi_costs = costs, multi_groups, multi_costs
        self.flattened_groups = [list(g) for k in range(K) for g in groups[k]]
        for k in range(K):
            groups[k] = np.array(groups[k], dtype=np.int64)
        self.groups = groups

        self.SAPS = [SAP(C[n], Ks[n], multi_groups[n], multi_costs[n], verbose=self.verbose, device=device)
                     for n in range(self.n_outputs)]

        self.sizes = [0] + [len(gk) for gk in groups]
        self.cumsizes = np.cumsum(self.sizes)
        self.L = self.cumsizes[-1]
        self.ES = indicator_ES(groups, self.N)
        self.e = self.ES[0]
        # m <-> groups, and m[mappings[n]] = m_n <-> multi_groups[n]   (mosap.py:54-67)
        self.mappings = build_mappings(groups, multi_groups)
        self.samples = self.budget = self.eps = self.tot_cost = None

    def check_input(self, budget, eps):
        """mosap.py:74-84: one tolerance per output; a scalar is broadcast, a sequence of the wrong length is an error."""
        if budget is None and eps is None:
            raise ValueError("Need to specify either budget or RMSE tolerance")
        if eps is None:
            return budget, eps
        if np.ndim(eps) == 0:
            return budget, np.full(self.n_outputs, eps, dtype=np.asarray(eps).dtype)
        if len(eps) != self.n_outputs:
            raise ValueError("eps must be a scalar or an array of tolerances")
        return budget, np.array(eps)

    def _small_batch(self):
        """One-launch evaluation of all outputs (bluest_b200/batch.py) when every output's problem is small enough
        that an evaluation is launch latency, not streaming (a CTA per output handles at most ~8k groups well)."""
        if not hasattr(self, "_batch"):
            self._batch = None
            if all(getattr(s, "_ctx", None) for s in self.SAPS) and max(int(s.L) for s in self.SAPS) <= 8192:
                from .batch import Batch
                self._batch = Batch(self.SAPS, maps=self.mappings, Lm=int(self.L))
        return self._batch

    def variances(self, m, delta=0):
        """mosap.py:86-89; all outputs in one kernel launch when they are small, else concurrently (one stream per
        context)."""
        from . import _lib
        bt = self._small_batch()
        if bt is not None:
            var, flags, _ = bt.eval(m, delta=delta, grad=False)
            out = []
            for n in range(self.n_outputs):
                if flags[n, 0] & _lib.FLAG_TINY:
                    out.append(np.inf)
                    continue
                assert not (flags[n, 0] & _lib.FLAG_NO_MODEL0)             # misc.py:470
                out.append(float(var[n, 0]))
            return out
        for n in range(self.n_outputs):
            self.SAPS[n].variance_GH_begin(m[self.mappings[n]], delta=delta, nohess=True, grad=False)
        out = []
        for n in range(self.n_outputs):
            var, _, _, fl = self.SAPS[n].variance_GH_end()
            if fl & _lib.FLAG_TINY:
                out.append(np.inf)
                continue
            assert not (fl & _lib.FLAG_NO_MODEL0)             # misc.py:470
            out.append(var)
        return out

    def variance_GH(self, m, nohess=False, delta=0):
        """mosap.py:91-100; without Hessians all outputs are one kernel launch (small problems), else all outputs in
        flight at once, results collected in output order."""
        from . import _lib
        bt = self._small_batch() if nohess else None
        if bt is not None:
            var, flags, grads = bt.eval(m, delta=delta, grad=True)
            if np.any(flags[:, 0] & _lib.FLAG_TINY):
                raise IndexError("tuple index out of range (max|m| < 0.05: misc.py:484 returns a 2-tuple)")
            return [float(v) for v in var[:, 0]], [np.array(g[0]) for g in grads], [None] * self.n_outputs
        for n in range(self.n_outputs):
            self.SAPS[n].variance_GH_begin(m[self.mappings[n]], delta=delta, nohess=nohess)
        variances, gradients, hessians = [], [], []
        for n in range(self.n_outputs):
            var, grad, hess, fl = self.SAPS[n].variance_GH_end()
            if fl & _lib.FLAG_TINY:
                # the reference's SAP returns a 2-tuple here and its MOSAP then fails on item[2]
                raise IndexError("tuple index out of range (max|m| < 0.05: misc.py:484 returns a 2-tuple)")
            variances.append(var); gradients.append(grad); hessians.append(hess)
        return variances, gradients, hessians

    def get_cleanup_matrices(self, m, delta=0):
        Xs = []
        for n in range(self.n_outputs):
            X = np.zeros((self.N, self.L))
            X[:, self.mappings[n]] = self.SAPS[n].get_cleanup_matrix(m[self.mappings[n]], delta=delta)
            Xs.append(X)
        return np.vstack(Xs)

    def compute_BLUE_estimators(self, sums, samples):
        """mosap.py:113-123."""
        out = []
        for n in range(self.n_outputs):
            sums_n = [sums[n][item] for item in self.mappings[n]]
            out.append(self.SAPS[n].compute_BLUE_estimator(sums_n, samples=samples[self.mappings[n]]))
        mus = [item[0] for item in out]
        Vars = np.array([item[1] for item in out])
        return mus, Vars

    # ---- host orchestration kept from the reference -------------------------------------------
    def output_indicators(self):
        """Per output n the (L,) vector that is ``e`` on the groups of output n and 0 elsewhere
        (mosap.py:136-141, 577-581)."""
        E = np.zeros((self.n_outputs, int(self.L)))
        for n in range(self.n_outputs):
            E[n, self.mappings[n]] = self.e[self.mappings[n]]
        return E

    def get_max_sample_constraints(self, max_model_samples):
        """mosap.py:326-344."""
        from .constraints import max_sample_constraints
        return max_sample_constraints(self.ES, self.N, max_model_samples)

    def cleanup_solution(self, m, delta=0, tol=0):
        """mosap.py:125-211 (see bluest_b200/cleanup.py): a sparser allocation of the same quality and cost."""
        from .cleanup import cleanup_solution
        return cleanup_solution(self, m, delta=delta, tol=tol)

    def integer_projection(self, samples, budget=None, eps=None, max_model_samples=None):
        """mosap.py:213-292; the candidate variances of every output come from one batched device call each."""
        from .intproj import integer_projection_multi
        return integer_projection_multi(self, samples, budget=budget, eps=eps, max_model_samples=max_model_samples)

    def solve(self, budget=None, eps=None, solver="scipy", x0=None, continuous_relaxation=False, max_model_samples=None,
              solver_params=None, hess="dense", sparse_constraints=False):
        """mosap.py:294-331.  ``solver="scipy"`` only (trust-constr on the GPU closures); the SDP
        solvers and ipopt are third-party host code consuming ``SAPS[n].psi``."""
        if budget is None and eps is None:
            raise ValueError("Need to specify either budget or RMSE tolerance")
        if solver != "scipy":
            raise ValueError("bluest_b200.MOSAP.solve provides solver='scipy'; for 'cvxopt'/'cvxpy'/'ipopt' hand the SAPS' "
                             "`psi` / closures to the reference's own drivers (INTEGRATION.md)")
        allocation = self.scipy_solve(budget=budget, eps=eps, x0=x0, max_model_samples=max_model_samples, hess=hess,
                                      sparse_constraints=sparse_constraints)
        if allocation is not None and not continuous_relaxation:
            try:
                allocation = self.integer_projection(allocation, budget=budget, eps=eps, max_model_samples=max_model_samples)
            except AssertionError as ex:                   # the reference reports the failed assertion and gives up (mosap.py:313-317)
                print(str(ex))
                allocation = None
        self.samples = allocation
        if allocation is None:
            return None
        self.budget, self.eps = budget, eps
        self.tot_cost = allocation @ self.costs
        for sap, mp in zip(self.SAPS, self.mappings):
            sap.samples = allocation[mp]
        return allocation

    def scipy_solve(self, budget=None, eps=None, x0=None, max_model_samples=None, maxiter=5000, hess="dense", sparse_constraints=False,
                    reference_eps_bound=True):
        """mosap.py:555-610 (see solvers.scipy_solve_multi for ``reference_eps_bound``)."""
        from .solvers import scipy_solve_multi
        self.scipy_counters = {}
        res = scipy_solve_multi(self, budget=budget, eps=eps, x0=x0, max_model_samples=max_model_samples, maxiter=maxiter,
                                verbose=self.verbose, counters=self.scipy_counters, hess=hess, sparse_constraints=sparse_constraints,
                                reference_eps_bound=reference_eps_bound)
        self.scipy_result = res
        return res.x
