"""Scope-table row f1: the structure-exploiting KKT solve of the SDP that ``SAP.cvxopt_solve`` builds (sap.py:242-307),
on the device, against the dense KKT solve of oracle/kkt.py -- the matrices G0, G1 are assembled exactly as the
reference assembles them for cvxopt.  cvxopt itself is absent from the image, so this pins the linear algebra of one
interior-point iteration, not a full cvxopt solve ("unverified against cvxopt")."""
import numpy as np
import pytest

import oracle as orc
import kkt as okkt
from conftest import maxrel

pytestmark = pytest.mark.gpu


def _problem(N, K, seed):
    import bluest_b200 as blu
    C = orc.wishart_cov(N, seed)
    groups = blu.enumerate_groups(N, K)
    costs = blu.group_costs(groups, 2.0 ** (N - np.arange(N)))
    costs = costs / costs.max()
    sap = blu.SAP(C, K, [[list(g) for g in gk] for gk in groups], costs, verbose=False)
    o = orc.SapOracle(C, K, orc.enumerate_groups(N, K))
    return sap, o, costs


# N >= 16 takes the 32-lanes-per-group form of the rows kernel, several CTAs per row range in the Gram kernel (more than 16
# tile blocks) and a capacitance matrix beyond what one SM's shared memory holds (the blocked Cholesky works in L2)
@pytest.mark.parametrize("N,K,budget_mode,seed", [(5, 5, True, 0), (6, 6, True, 1), (6, 6, False, 2), (8, 3, True, 3), (10, 10, True, 4), (10, 10, False, 5),
                                                  (17, 3, True, 6), (20, 2, False, 7), (16, 4, True, 8)])
def test_kkt_solve_matches_dense_kkt(N, K, budget_mode, seed):
    sap, o, costs = _problem(N, K, seed)
    L, M = o.L, N + 1
    G0, G1, Gx, scales, has_t = okkt.sdp_data(o.psi, costs, o.e.astype(float), N, budget_mode=budget_mode)
    Gx2, scales2, has_t2 = sap.sdp_linear_rows(budget_mode=budget_mode)
    assert has_t2 == has_t and np.array_equal(Gx2, Gx) and abs(scales2 - scales) <= 1e-13 * scales
    n, nlin = L + has_t, Gx.shape[0]
    rng = np.random.RandomState(100 + seed)
    d = 0.5 + 1.5 * rng.rand(n + nlin)                     # Nesterov-Todd scaling of the linear cone
    r = np.eye(M) + 0.1 * rng.randn(M, M)                  # scaling matrix of the semidefinite block (nonsingular)
    bx = rng.randn(n)
    Z = rng.randn(M, M); Z = Z + Z.T                       # the 's' part of a right-hand side is a symmetric matrix
    bz = np.concatenate([rng.randn(n + nlin), Z.ravel()])
    ux, uz = okkt.dense_kkt_solve(G0, G1, d, r, bx, bz)
    vx, vz, ms = sap.kkt_solve(has_t, scales, Gx, d, r, bx, bz, return_ms=True)
    # both solves are limited by the conditioning of the KKT matrix (~cond x eps); the residual check below is the sharp one
    assert maxrel(vx, ux) < 1e-9 and maxrel(vz, uz) < 1e-9
    # residual of the full KKT system with the device's solution (independent of the oracle's own solve)
    G = np.vstack([G0, G1])
    rrT = r @ r.T
    WtWuz = np.concatenate([d ** 2 * vz[:n + nlin], (rrT @ vz[n + nlin:].reshape(M, M) @ rrT).ravel()])
    assert maxrel(G.T @ vz, bx) < 1e-9
    assert np.max(np.abs(G @ vx - WtWuz - bz)) <= 1e-9 * max(np.max(np.abs(bz)), np.max(np.abs(WtWuz)))
    sap.close()


def test_kkt_solve_full_size_residual():
    """15 models, all 32767 groups (BASELINE config 3): the dense KKT matrix has 65 796 rows (34.6 GB) and is never
    formed; the device solution is checked through the residual of the reduced system, evaluated operator-wise."""
    import bluest_b200 as blu
    N = 15
    C = orc.wishart_cov(N, 0)
    ga = blu.enumerate_group_arrays(N)
    L = sum(len(g) for g in ga)
    costs = blu.group_costs(ga, 2.0 ** (N - np.arange(N))); costs = costs / costs.max()
    sap = blu.SAP(C, N, ga, costs, verbose=False)
    M = N + 1
    Gx, scales, has_t = sap.sdp_linear_rows(budget_mode=True)
    n, nlin = L + 1, Gx.shape[0]
    rng = np.random.RandomState(9)
    d = 0.1 + 10.0 * rng.rand(n + nlin)
    r = np.eye(M) + 0.2 * rng.randn(M, M)
    bx = rng.randn(n)
    Z = rng.randn(M, M); Z = Z + Z.T
    bz = np.concatenate([rng.randn(n + nlin), Z.ravel()])
    ux, uz, ms = sap.kkt_solve(has_t, scales, Gx, d, r, bx, bz, return_ms=True)
    # G^T uz = bx and G ux - W^T W uz = bz, with G1 applied through psi (59 MB dense on the host)
    psi = sap.psi
    def G1x(x):
        X = np.zeros((M, M)); X[:N, :N] = -scales * (psi @ x[1:]).reshape(N, N); X[N, N] -= x[0]
        return X
    def G1T(U):
        out = np.empty(n)
        out[0] = -U[N, N]
        out[1:] = -scales * (psi.T @ U[:N, :N].ravel())
        return out
    U = uz[n + nlin:].reshape(M, M)
    res1 = -uz[:n] + Gx.T @ uz[n:n + nlin] + G1T(U) - bx
    assert np.max(np.abs(res1)) <= 1e-8 * np.max(np.abs(bx))
    rrT = r @ r.T
    res2a = np.concatenate([-ux, Gx @ ux]) - d ** 2 * uz[:n + nlin] - bz[:n + nlin]
    res2b = G1x(ux) - rrT @ U @ rrT - Z
    scale = max(np.max(np.abs(bz)), np.max(np.abs(d ** 2 * uz[:n + nlin])))
    assert np.max(np.abs(res2a)) <= 1e-8 * scale and np.max(np.abs(res2b)) <= 1e-8 * scale
    print("kkt_solve at 15 models: device part %.2f ms" % ms)
    sap.close()
