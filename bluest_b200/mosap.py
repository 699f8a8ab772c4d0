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
