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
        self.verbose = verbose
        self.n_outputs = len(C)
        self.C = C
        self.N = C[0].shape[0]
        self.K = K
        self.Ks = Ks
        self.costs = costs
        self.multi_groups = multi_groups
        self.multi_costs = multi_costs
        flattened_groups = []
        for k in range(K):
            flattened_groups += [list(g) for g in groups[k]]
            groups[k] = np.array(groups[k], dtype=np.int64)       # mosap.py:34
        self.flattened_groups = flattened_groups
        self.groups = groups

        self.SAPS = [SAP(C[n], Ks[n], multi_groups[n], multi_costs[n], verbose=self.verbose, device=device)
                     for n in range(self.n_outputs)]

        self.sizes = [0] + [len(groupsk) for groupsk in groups]
        self.cumsizes = np.cumsum(self.sizes)
        self.L = self.cumsizes[-1]
        self.ES = indicator_ES(groups, self.N)
        self.e = self.ES[0]
        # m <-> groups, and m[mappings[n]] = m_n <-> multi_groups[n]   (mosap.py:54-67)
        self.mappings = build_mappings(groups, multi_groups)

        self.samples = None
        self.budget = None
        self.eps = None
        self.tot_cost = None

    def check_input(self, budget, eps):
        if budget is None and eps is None:
            raise ValueError("Need to specify either budget or RMSE tolerance")
        if eps is not None:
            try:
                if len(eps) != self.n_outputs:
                    raise ValueError("eps must be a scalar or an array of tolerances")
                eps = np.array(eps)
            except TypeError:
                eps = np.array([eps for n in range(self.n_outputs)])
        return budget, eps

    def variances(self, m, delta=0):
        """mosap.py:86-89; the outputs are evaluated concurrently (one stream per context)."""
        from . import _lib
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
        """mosap.py:91-100; all outputs in flight at once, results collected in output order."""
        from . import _lib
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
        """mosap.py:333-351."""
        if max_model_samples is None:
            return [], []
        if not isinstance(max_model_samples, np.ndarray) or len(max_model_samples) != self.N:
            raise ValueError("The maximum number of model samples must be prescribed as a numpy array of the same length as the number of models.")
        if max_model_samples[0] < 1:
            raise ValueError("The high-fidelity model must be sampled at least once.")
        es, rhs = [], []
        for i in range(self.N):
            if np.isfinite(max_model_samples[i]):
                es.append(self.ES[i])
                rhs.append(int(np.round(max_model_samples[i])))
        return es, rhs

    def cleanup_solution(self, m, delta=0, tol=0):
        """mosap.py:125-211: walk along null-space directions of the stacked cleanup matrices that do
        not increase the cost, as far as positivity and the coverage constraints allow, while the
        largest output variance does not get worse -- a sparser allocation of the same quality.
        Like the reference it zeroes the entries of the caller's ``m`` that lie below ``tol``."""
        from scipy.linalg import null_space
        w = self.costs
        E = self.output_indicators()
        worst = lambda x: max(self.variances(x, delta=delta))
        idx = np.argwhere(m > tol).flatten()
        V0 = worst(m)
        smax = 0
        while len(idx) > self.N:
            idx = np.argwhere(m > tol).flatten()
            m[m < tol] = 0
            wr, Er = w[idx], E[:, idx]
            X = self.get_cleanup_matrices(m, delta=delta)[:, idx]
            NN = null_space(X)
            vals = wr @ NN
            signs = np.sign(vals)
            NN[:, signs > 0] *= -1                     # orient every direction so that it does not raise the cost
            vals[signs > 0] *= -1
            NN, vals = NN[:, abs(signs) > 0], vals[abs(signs) > 0]
            order = np.argsort(abs(vals))[::-1]        # steepest cost decrease first
            if len(vals) == 0:
                break
            em = Er @ m[idx]
            for i in order:
                t = NN[:, i]
                evals = Er @ t
                neg = np.argwhere(evals < 0).flatten()
                s1 = np.inf if len(neg) == 0 else min(abs(em[neg] - 1) / abs(evals[neg]))
                neg = np.argwhere(t < 0).flatten()
                s2 = np.inf if len(neg) == 0 else min(m[idx][neg] / abs(t[neg]))
                smax = max(min(s1, s2), 0)
                if smax > 5 * tol:
                    step = np.zeros_like(m); step[idx] = t
                    mnew = m + smax * step
                    V = worst(mnew)
                    if V < V0 or abs(V - V0) / abs(V0) < 1.0e-4:
                        m = mnew.copy()
                        break
                    smax = 0
            if smax <= 5 * tol:
                break
        m[m < tol] = 0
        return m

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
        samples = self.scipy_solve(budget=budget, eps=eps, x0=x0, max_model_samples=max_model_samples, hess=hess,
                                   sparse_constraints=sparse_constraints)
        if samples is None:
            self.samples = None
            return None
        if not continuous_relaxation:
            try:
                samples = self.integer_projection(samples, budget=budget, eps=eps, max_model_samples=max_model_samples)
            except AssertionError as ex:
                print(str(ex))
                self.samples = None
                return None
        self.samples = samples
        self.budget = budget
        self.eps = eps
        self.tot_cost = samples @ self.costs
        for n in range(self.n_outputs):
            self.SAPS[n].samples = samples[self.mappings[n]]
        return samples

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
