"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by bluest_b200/).

CPU restatement (numpy + oracle/liboracle.so) of the reference's sample-allocation
hot path, written from the maths in SURVEY.md section 0.3 and following the
reference's semantics line by line where they matter for parity:

  enumerate_groups / group_costs   blue_models.py:462-501, :137-140 (complete graph
                                    and user group lists, size-major + lexicographic)
  mosap_mappings / indicator_ES    mosap.py:41-67
  SapOracle.__init__               sap.py:53-97  (per-group pinv, flat invcovs, int64 groups)
  SapOracle.psi                    sap.py:129 -> misc.py:600-604 -> cmisc.cpp:10-23
  get_phi / variance / variance_GH misc.py:453-505 (support selection, early-outs,
                                    pinv vs solve, ``hess += hess.T``)
  cleanup_matrix                   misc.py:507-516 -> cmisc.cpp:42-56 (assignment bug kept)
  pilot_covariance                 blue_fn.py:159-167 + blue_models.py:333

Parity status: PINNED.  tests/test_oracle.py checks every function here against
(a) the committed golden vectors in tests/golden/ that were produced by running the
real reference Python in the build container (tests/golden/make_golden.py), and
(b) live, against the reference itself through oracle/ref_shim.py whenever
/root/reference is present.

Who may import this: tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline /
``--impl reference`` legs -- as the checker or the timed CPU baseline, never as the
product path.
"""
import ctypes
import os
from itertools import combinations

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_i64 = ctypes.c_int64
_pd = ctypes.POINTER(ctypes.c_double)
_pi = ctypes.POINTER(ctypes.c_int64)


def lib():
    """Load oracle/liboracle.so (built by oracle/Makefile or __graft_entry__.build())."""
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.isfile(path):
            raise RuntimeError("oracle/liboracle.so missing: run `make -C oracle` or __graft_entry__.build()")
        L = ctypes.CDLL(path)
        L.orc_psi_class.argtypes = [_pd, _i64, _i64, _i64, _pi, _pd]
        L.orc_phi_class.argtypes = [_pd, _i64, _i64, _i64, _pd, _pi, _pd]
        L.orc_phi_class_i64.argtypes = [_pd, _i64, _i64, _i64, _pi, _pi, _pd]
        L.orc_grad_class.argtypes = [_pd, _i64, _i64, _pi, _pd, _pd]
        L.orc_cleanup_class.argtypes = [_pd, _i64, _i64, _pi, _pd, _pd]
        L.orc_ufactor_class.argtypes = [_pd, _i64, _i64, _pi, _pd, _pd]
        L.orc_hess_block.argtypes = [_pd, _i64, _i64, _i64, _i64, _i64, _pi, _pi, _pd, _pd, _pd]
        L.orc_pilot_cov.argtypes = [_pd, _i64, _i64, _pd, _pd, _pd]
        for f in ("orc_psi_class", "orc_phi_class", "orc_phi_class_i64", "orc_grad_class",
                  "orc_cleanup_class", "orc_ufactor_class", "orc_hess_block", "orc_pilot_cov"):
            getattr(L, f).restype = None
        _LIB = L
    return _LIB


def _d(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_pd)


def _l(a):
    assert a.dtype == np.int64 and a.flags.c_contiguous
    return a.ctypes.data_as(_pi)


# --------------------------------------------------------------------------------------
# group enumeration and integer index maps (bit-exact rows of SURVEY.md section 8a: a1, a2)
# --------------------------------------------------------------------------------------

def enumerate_groups(N, K=None):
    """Complete model graph: all subsets of size <= K, size-major, lexicographic inside a
    size class -- what blue_models.py:462-474 + ``groups[k].sort()`` (:500-501) yield."""
    K = N if K is None else min(K, N)
    return [[list(c) for c in combinations(range(N), k)] for k in range(1, K + 1)]


def enumerate_group_arrays(N, K=None):
    """``enumerate_groups`` as one (Lk,k) int64 array per size class -- same order, built without
    a Python list per group, so 20 models (1 048 575 groups) take seconds."""
    K = N if K is None else min(K, N)
    out = []
    for k in range(1, K + 1):
        flat = np.fromiter((v for c in combinations(range(N), k) for v in c), dtype=np.int64)
        out.append(flat.reshape(-1, k))
    return out


def batched_invcovs(C, group_arrays):
    """Per-class flat inverse arrays of sap.py:72-79 for WELL-CONDITIONED SPD blocks: one batched
    LAPACK inverse per size class instead of L calls of ``np.linalg.pinv`` (for such blocks pinv == inv
    to rounding; tests/test_oracle.py pins this against the per-group pinv of ``SapOracle`` and the
    golden reference inverses).  Lets the oracle cover 20 models (sap.py's own loop needs ~100 s there)."""
    C = np.asarray(C, dtype=np.float64)
    out = []
    for gk in group_arrays:
        if len(gk) == 0:
            out.append(np.array([]))
            continue
        sub = C[gk[:, :, None], gk[:, None, :]]                  # (Lk,k,k) blocks C[g,g]
        out.append(np.ascontiguousarray(np.linalg.inv(sub)).ravel())
    return out


def bin_user_groups(groups_list, N):
    """User-supplied flat list of groups -> sorted, binned by size (possibly empty
    classes), the way blue_models.py:476-491 does for a complete graph."""
    K = min(max(len(g) for g in groups_list), N)
    out = [[] for _ in range(K)]
    for g in groups_list:
        g = sorted(g)
        out[len(g) - 1].append(g)
    return out


def union_groups(multi_groups):
    """Union over outputs, then per-size lexicographic sort (blue_models.py:493-501)."""
    K = max(len(g) for g in multi_groups)
    groups = [[] for _ in range(K)]
    for mg in multi_groups:
        for k, gk in enumerate(mg):
            for g in gk:
                if g not in groups[k]:
                    groups[k].append(list(g))
    for k in range(K):
        groups[k].sort()
    return groups


def group_costs(groups, model_costs):
    """Cost of a group = sum of its models' costs (blue_models.py:137-140)."""
    mc = np.asarray(model_costs)
    return np.array([sum(mc[list(g)]) for gk in groups for g in gk])


def indicator_ES(groups, N):
    """ES[i][g] = 1 iff model i belongs to group g (sap.py:89-95, mosap.py:46-52)."""
    flat = [g for gk in groups for g in gk]
    ES = np.zeros((N, len(flat)), dtype=np.int64)
    for col, g in enumerate(flat):
        ES[list(g), col] = 1
    return [ES[i].copy() for i in range(N)]


def mosap_mappings(groups, multi_groups):
    """mappings[n][j] = flat position in the union ``groups`` of the j-th group of output n
    (mosap.py:54-67); here with a dictionary instead of the reference's O(L^2) scan."""
    sizes = [0] + [len(gk) for gk in groups]
    cum = np.cumsum(sizes)
    where = {}
    for k, gk in enumerate(groups):
        for j, g in enumerate(gk):
            where[tuple(int(v) for v in g)] = int(cum[k] + j)
    out = []
    for mg in multi_groups:
        out.append(np.array([where[tuple(int(v) for v in g)] for gk in mg for g in gk], dtype=np.int64))
    return out


# --------------------------------------------------------------------------------------
# SAP setup + closures (rows a3..a10)
# --------------------------------------------------------------------------------------

class SapOracle:
    """CPU statement of ``SAP.__init__`` + ``get_variance_functions`` (sap.py:53-143)."""

    def __init__(self, C, K, groups, costs=None, invcovs=None, with_ES=True):
        """``groups[k-1]``: lists of sorted model-index lists, or (Lk,k) int64 arrays.  ``invcovs``: use these
        flat per-class inverses instead of the per-group pinv loop.  ``with_ES=False`` skips the O(L N)
        indicator vectors (full-size checks that never read them)."""
        C = np.asarray(C, dtype=np.float64)
        self.C = C
        self.N = C.shape[0]
        self.K = K
        self.sizes = [0] + [len(gk) for gk in groups]
        self.cumsizes = np.cumsum(self.sizes)
        self.L = int(self.cumsizes[-1])
        self.costs = costs
        self.groups = []
        self.invcovs = []
        for k in range(1, K + 1):
            gk = np.array(groups[k - 1], dtype=np.int64).reshape(len(groups[k - 1]), k)
            self.groups.append(gk)
            if invcovs is not None:
                self.invcovs.append(np.ascontiguousarray(invcovs[k - 1], dtype=np.float64).ravel())
                continue
            blocks = [np.linalg.pinv(C[np.ix_(g, g)]) for g in gk]      # sap.py:72-74
            self.invcovs.append(np.concatenate([b.ravel() for b in blocks]) if blocks else np.array([]))
        self.ES = indicator_ES(groups, self.N) if with_ES else None
        self.e = self.ES[0] if with_ES else None
        self._psi = None

    # psi: (N^2, L), hstack over non-empty classes (sap.py:129)
    @property
    def psi(self):
        if self._psi is None:
            N = self.N
            cols = []
            for k in range(1, self.K + 1):
                Lk = self.sizes[k]
                if Lk == 0:
                    continue
                blk = np.zeros((N * N, Lk))
                lib().orc_psi_class(_d(blk), N, k, Lk, _l(self.groups[k - 1]), _d(self.invcovs[k - 1]))
                cols.append(blk)
            self._psi = np.hstack(cols)
        return self._psi

    def get_phi(self, m, delta=0.0):
        """misc.py:459-461, but through the sparse native loop (objectiveK_c) rather than the
        dense GEMV: identical maths, N^2 accumulators in group order."""
        N = self.N
        m = np.asarray(m)
        phi = np.zeros(N * N)
        for k in range(1, self.K + 1):
            Lk = self.sizes[k]
            if Lk == 0:
                continue
            mk = np.ascontiguousarray(m[self.cumsizes[k - 1]:self.cumsizes[k]])
            if mk.dtype == np.int64:
                lib().orc_phi_class_i64(_d(phi), N, k, Lk, _l(mk), _l(self.groups[k - 1]), _d(self.invcovs[k - 1]))
            else:
                mk = mk.astype(np.float64)
                lib().orc_phi_class(_d(phi), N, k, Lk, _d(mk), _l(self.groups[k - 1]), _d(self.invcovs[k - 1]))
        return delta * np.eye(N) + phi.reshape(N, N)

    def get_phi_dense(self, m, delta=0.0):
        """Literal misc.py:459-461: delta*I + reshape(psi @ m)."""
        N = self.N
        return delta * np.eye(N) + (self.psi @ m).reshape(N, N)

    def support(self, m):
        """misc.py:453-457: sorted unique models of the groups with |m_i| > 1e-6."""
        m = np.asarray(m)
        parts = []
        for k in range(self.K):
            mk = m[self.cumsizes[k]:self.cumsizes[k + 1]]
            parts.append(self.groups[k][np.abs(mk) > 1.0e-6].ravel())
        return np.unique(np.concatenate(parts))

    def variance(self, m, delta=0.0):
        """misc.py:463-477."""
        m = np.asarray(m)
        if np.abs(m).max() < 0.05:
            return np.inf
        phi = self.get_phi(m, delta)
        idx = self.support(m)
        assert idx.min() == 0
        sub = phi[np.ix_(idx, idx)]
        e0 = np.zeros(len(idx)); e0[0] = 1.0
        return np.linalg.solve(sub, e0)[0]

    def variance_GH(self, m, delta=0.0, nohess=False, hess_mode="reference"):
        """misc.py:479-505.  ``hess_mode``: "reference" runs the K^2 six-deep native loops
        (cmisc.cpp:74-97); "factored" uses H = 2 U^T Phi^+ U, same result
        to rounding, for sizes where the loops take hours."""
        m = np.asarray(m)
        L = len(m)
        if np.abs(m).max() < 0.05:
            return np.inf, np.inf * np.ones((L,))
        phi = self.get_phi(m, delta)
        P = np.linalg.pinv(phi)
        idx = self.support(m)
        var = np.linalg.pinv(phi[np.ix_(idx, idx)])[0, 0]
        x = np.ascontiguousarray(P[0])
        grad = np.zeros(L)
        for k in range(1, self.K + 1):
            Lk = self.sizes[k]
            if Lk == 0:
                continue
            gk = grad[self.cumsizes[k - 1]:self.cumsizes[k]]
            lib().orc_grad_class(_d(gk), k, Lk, _l(self.groups[k - 1]), _d(self.invcovs[k - 1]), _d(x))
        grad = -grad
        if nohess:
            return var, grad, None
        if hess_mode == "factored":
            U = self.ufactor(x)
            hess = U.T @ P @ U
        else:
            hess = np.zeros((L, L))
            Pf = np.ascontiguousarray(P).ravel()
            for k in range(1, self.K + 1):
                for q in range(1, self.K + 1):
                    Lk, Lq = self.sizes[k], self.sizes[q]
                    if Lk == 0 or Lq == 0:
                        continue
                    blk = np.zeros((Lk, Lq))
                    lib().orc_hess_block(_d(blk), self.N, k, q, Lk, Lq, _l(self.groups[k - 1]), _l(self.groups[q - 1]),
                                         _d(self.invcovs[k - 1]), _d(self.invcovs[q - 1]), _d(Pf))
                    hess[self.cumsizes[k - 1]:self.cumsizes[k], self.cumsizes[q - 1]:self.cumsizes[q]] = blk
        hess = hess + hess.T
        return var, grad, hess

    def ufactor(self, x):
        """U (N, L): column i is R_i^T Cinv_i R_i x  (SURVEY.md 0.3)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        cols = []
        for k in range(1, self.K + 1):
            Lk = self.sizes[k]
            if Lk == 0:
                continue
            blk = np.zeros((self.N, Lk))
            lib().orc_ufactor_class(_d(blk), k, Lk, _l(self.groups[k - 1]), _d(self.invcovs[k - 1]), _d(x))
            cols.append(blk)
        return np.hstack(cols)

    def cleanup_matrix(self, m, delta=0.0):
        """misc.py:507-516 with the native loop's assignment semantics (cmisc.cpp:51)."""
        m = np.asarray(m)
        if np.abs(m).max() < 0.05:
            raise ValueError("No entry greater or equal than 1 found in m.")
        phi = self.get_phi(m, delta)
        x = np.ascontiguousarray(np.linalg.pinv(phi)[0])
        cols = []
        for k in range(1, self.K + 1):
            Lk = self.sizes[k]
            blk = np.zeros((self.N, Lk))
            if Lk:
                lib().orc_cleanup_class(_d(blk), k, Lk, _l(self.groups[k - 1]), _d(self.invcovs[k - 1]), _d(x))
            cols.append(blk)
        return np.hstack(cols)


def blue_estimator(o, sums, samples):
    """compute_BLUE_estimator (sap.py:99-119) + PHIinvY0 (misc.py:518-544) on a SapOracle ``o``.
    sums[i] = the k sample sums of group i (flat group order).  Returns (mu, var, y)."""
    samples = np.asarray(samples)
    y = np.zeros(o.N)
    pos = 0
    for k in range(1, o.K + 1):
        for i in range(o.sizes[k]):
            Ci = o.invcovs[k - 1][k * k * i:k * k * (i + 1)].reshape(k, k)
            g = o.groups[k - 1][i]
            si = np.asarray(sums[pos], dtype=float)
            for j in range(k):
                for s in range(k):
                    y[g[j]] += Ci[j, s] * si[s]
            pos += 1
    if np.abs(samples).max() < 0.05:
        return np.inf, np.inf, y
    phi = o.get_phi(samples)
    idx = o.support(samples)
    assert idx.min() == 0
    P = np.linalg.pinv(phi[np.ix_(idx, idx)])
    mu = 0.0
    for j in range(len(idx)):
        mu += P[0, j] * y[idx[j]]
    return mu, P[0, 0], y


def pilot_covariance(Y):
    """Biased one-pass covariance of an (n, N) sample matrix, blue_models.py:333 with the
    default inner product of blue_fn.py:82-83.  Returns (s1, S2, C_hat)."""
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    n, N = Y.shape
    s1 = np.zeros(N); S2 = np.zeros((N, N)); C = np.zeros((N, N))
    lib().orc_pilot_cov(_d(Y), n, N, _d(s1), _d(S2), _d(C))
    return s1, S2, C


# --------------------------------------------------------------------------------------
# synthetic inputs shared by tests and bench
# --------------------------------------------------------------------------------------

def wishart_cov(N, seed=0):
    rng = np.random.RandomState(seed)
    A = rng.randn(N, 2 * N)
    return A @ A.T / (2 * N)


def dense_m(L, seed=0):
    rng = np.random.RandomState(1000 + seed)
    return 1.0 + 10.0 * rng.rand(L)


def sparse_m(L, N, seed=0, keep_first=True):
    """<= 2N non-zeros, typical optimiser output; group 0 ({model 0}) kept so model 0 is sampled."""
    rng = np.random.RandomState(2000 + seed)
    m = np.zeros(L)
    nz = rng.choice(L, size=min(2 * N, L), replace=False)
    m[nz] = np.ceil(50 * rng.rand(len(nz)))
    if keep_first:
        m[0] = 7.0
    return m
