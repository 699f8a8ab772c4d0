"""oracle/kkt.py -- TEST INFRASTRUCTURE ONLY (never imported by bluest_b200/).

Dense CPU statement of the KKT systems cvxopt's interior-point method solves for the semidefinite programme that
``SAP.cvxopt_solve`` builds (reference sap.py:242-307).  cvxopt is a third-party dependency of the reference that is
absent from this image (unpinned in the reference: README.md:34, docker/Dockerfile:282-295), so the algorithm is
restated from its published description -- ``cvxopt.solvers.conelp`` with no equality constraints solves, for the
Nesterov-Todd scaling W = blkdiag(diag(d), W_s), W_s vec(U) = vec(r^T U r),

    [ 0    G^T  ] [ux]   [bx]
    [ G  -W^T W ] [uz] = [bz]

once per iteration and right-hand side -- and parity of the device solver (``blu_kkt_solve``) is anchored on the
reference's own call site: the matrices G0, G1 are built exactly as sap.py:259-287 builds them (``sdp_data``).
Parity status: UNPINNED against cvxopt itself (package absent); pinned against this dense solve only.
"""
import numpy as np


def sdp_data(psi, w, e, N, budget_mode=True, extra_rows=None):
    """G0 (dense rows of the linear cone), G1 (the semidefinite block) and the scale factor of sap.py:258-287.

    budget_mode: variables [t, m/budget]; G0 = [-I; wt; -et (; sample-cap rows)], G1 = [-E_NN | -scales vec(pad Psi_i)].
    else       : variables m;             G0 = [-I; -et (; ...)],                 G1 = -scales vec(pad Psi_i)."""
    L = psi.shape[1]
    scales = 1.0 / np.abs(psi).sum(axis=0).mean()                       # sap.py:258
    M = N + 1
    has_t = 1 if budget_mode else 0
    n = L + has_t
    rows = []
    if budget_mode:
        rows.append(np.concatenate([[0.0], w]))                         # wt, sap.py:260
        rows.append(-np.concatenate([[0.0], e]))                        # -et, sap.py:261,269
    else:
        rows.append(-np.asarray(e, dtype=float))                        # sap.py:281
    if extra_rows is not None:
        rows.extend(np.asarray(r_, dtype=float) for r_ in extra_rows)
    Gx = np.array(rows, dtype=float).reshape(len(rows), n)
    G0 = np.vstack([-np.eye(n), Gx])
    G1 = np.zeros((M * M, n))
    for i in range(L):
        X = np.zeros((M, M)); X[:N, :N] = psi[:, i].reshape(N, N)       # a zero row/column after every N entries, sap.py:272,284
        G1[:, i + has_t] = -scales * X.ravel()
    if budget_mode:
        G1[M * M - 1, 0] = -1.0                                         # sap.py:273: the t column
    return G0, G1, Gx, scales, has_t


def dense_kkt_solve(G0, G1, d, r, bx, bz):
    """Assemble and solve the full KKT matrix (the work cvxopt's built-in solvers do by dense factorisations)."""
    n = G0.shape[1]
    G = np.vstack([G0, G1])
    rrT = r @ r.T
    WtW = np.zeros((G.shape[0], G.shape[0]))
    WtW[:len(d), :len(d)] = np.diag(np.asarray(d) ** 2)
    WtW[len(d):, len(d):] = np.kron(rrT, rrT)              # vec(r r^T U r r^T) = (rr^T (x) rr^T) vec(U), U symmetric
    K = np.block([[np.zeros((n, n)), G.T], [G, -WtW]])
    sol = np.linalg.solve(K, np.concatenate([bx, bz]))
    return sol[:n], sol[n:]
