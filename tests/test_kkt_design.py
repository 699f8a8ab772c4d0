"""Design validation for scope-table row f1 (CPU only, no product code involved): the diagonal + low-rank
(Woodbury) form of the SDP's reduced KKT system reproduces the dense solve on real SAP data."""
import os
import sys

import numpy as np

import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "lab"))


def test_woodbury_kkt_equals_dense_kkt():
    import kkt_woodbury as kk
    N, K = 6, 6
    groups = orc.enumerate_groups(N, K)
    o = orc.SapOracle(orc.wishart_cov(N, 0), K, groups)
    L = o.L
    w = np.array([sum(2.0 ** (N - np.array(g))) for gk in groups for g in gk]); w /= w.max()
    G0, G1 = kk.sdp_data_budget(o.psi, w, o.e.astype(float), N)
    rng = np.random.RandomState(0)
    d = 0.1 + rng.rand(L + 3) * 10                          # Nesterov-Todd scaling of the linear cone
    r = np.eye(N + 1) + 0.3 * rng.randn(N + 1, N + 1)       # scaling matrix of the semidefinite block (nonsingular)
    bx = rng.randn(L + 1)
    Z = rng.randn(N + 1, N + 1); Z = Z + Z.T                # the 's' part of a right-hand side is a symmetric matrix
    bz = np.concatenate([rng.randn(L + 3), Z.ravel()])
    ux, uz = kk.dense_kkt_solve(G0, G1, d, r, bx, bz)
    vx, vz = kk.woodbury_kkt_solve(G0, G1, d, r, bx, bz)
    assert np.max(np.abs(vx - ux)) <= 1e-9 * np.max(np.abs(ux))
    assert np.max(np.abs(vz - uz)) <= 1e-9 * np.max(np.abs(uz))
    # the low-rank part really is low rank: (N+1)(N+2)/2 + 2 directions at most
    Lam = np.linalg.inv(r @ r.T)
    Mlow = np.array([[np.trace(G1[:, i].reshape(N + 1, N + 1) @ Lam @ G1[:, j].reshape(N + 1, N + 1) @ Lam) for j in range(L + 1)] for i in range(L + 1)])
    assert np.linalg.matrix_rank(Mlow, tol=1e-9 * np.abs(Mlow).max()) <= (N + 1) * (N + 2) // 2
