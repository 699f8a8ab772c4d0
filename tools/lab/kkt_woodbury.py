"""Design study for scope-table row f1 (NOT part of the product, CPU only): the reduced KKT system of the
SDP that ``SAP.cvxopt_solve`` hands to ``cvxopt.solvers.sdp`` (sap.py:242-307) is diagonal + low rank, so one
interior-point iteration needs an (N+1)^2+2 capacitance matrix instead of a dense (L+1) x (L+1) factorisation.

Budget mode (sap.py:259-275), variables x = [t, m/budget] in R^(L+1):
    G0 = [-I; wt^T; -et^T]                       linear cone ('l'), L+3 rows
    G1 = [ -E_NN | -scale * vec(pad(Psi_i)) ]    one (N+1) x (N+1) semidefinite block ('s')
cvxopt's conelp with no equality constraints solves, per iteration and right-hand side,
    [ 0    G^T  ] [ux]   [bx]
    [ G  -W^T W ] [uz] = [bz]
with the Nesterov-Todd scaling W = blkdiag(diag(d), W_s), W_s vec(U) = vec(r^T U r).  Eliminating uz:
    M ux = bx + G^T (W^T W)^-1 bz,      M = G^T (W^T W)^-1 G
    M = diag(d_0..L^-2) + wt wt^T / d_w^2 + et et^T / d_e^2 + [ tr(X_i Lam X_j Lam) ]_ij,   Lam = (r r^T)^-1
The last term is A^T A with A = (Lam^1/2 (x) Lam^1/2) G1, of rank <= (N+1)(N+2)/2, hence (Woodbury)
    M^-1 b = D^-1 b - D^-1 B (I + B^T D^-1 B)^-1 B^T D^-1 b,     B = [A^T | wt/d_w | et/d_e].
Cost per iteration: the capacitance matrix I + B^T D^-1 B is a weighted Gram contraction over the groups,
((N+1)^2+2)^2 L flops (2.2 GFlop at 15 models -- a DMMA kernel of the blu_gram family over the packed inverses)
against L^3/3 = 1.2e13 for the dense factorisation the reference's call implies.

What is missing to build it: cvxopt itself (absent from the image), to check the ``kktsolver(W)`` callback
protocol (in-place x, z; z returned scaled by W) and the result against the reference's solve.
"""
import numpy as np


def sdp_data_budget(psi, w, e, N):
    """G0, G1 of sap.py:259-275 as dense arrays (scale = 1/mean column 1-norm of psi, sap.py:258)."""
    L = psi.shape[1]
    scales = 1.0 / np.abs(psi).sum(axis=0).mean()
    wt = np.concatenate([[0.0], w]); et = np.concatenate([[0.0], e])
    G0 = np.vstack([-np.eye(L + 1), wt, -et])
    G1 = np.zeros(((N + 1) ** 2, L + 1))
    for i in range(L):
        X = np.zeros((N + 1, N + 1)); X[:N, :N] = psi[:, i].reshape(N, N)
        G1[:, i + 1] = -scales * X.ravel()
    G1[(N + 1) ** 2 - 1, 0] = -1.0
    return G0, G1


def dense_kkt_solve(G0, G1, d, r, bx, bz):
    """Reference answer: assemble and solve the full KKT matrix."""
    n = G0.shape[1]
    G = np.vstack([G0, G1])
    rrT = r @ r.T
    WtW = np.zeros((G.shape[0], G.shape[0]))
    WtW[:len(d), :len(d)] = np.diag(d ** 2)
    WtW[len(d):, len(d):] = np.kron(rrT, rrT)              # vec(r r^T U r r^T) = (rr^T (x) rr^T) vec(U), U symmetric
    K = np.block([[np.zeros((n, n)), G.T], [G, -WtW]])
    sol = np.linalg.solve(K, np.concatenate([bx, bz]))
    return sol[:n], sol[n:]


def woodbury_kkt_solve(G0, G1, d, r, bx, bz):
    """Same solution through diagonal + low-rank structure; never forms an (L+1) x (L+1) matrix."""
    n = G0.shape[1]
    Np1 = r.shape[0]
    Lam = np.linalg.inv(r @ r.T)
    ev, Q = np.linalg.eigh(Lam)
    Lh = (Q * np.sqrt(ev)) @ Q.T                            # Lam^1/2
    di2 = d ** -2.0
    D = di2[:n]                                             # from the -I rows
    wt, et = G0[n], -G0[n + 1]
    # A = (Lam^1/2 (x) Lam^1/2) G1, column by column: vec(Lam^1/2 X_i Lam^1/2)
    A = np.stack([(Lh @ G1[:, i].reshape(Np1, Np1) @ Lh).ravel() for i in range(n)], axis=1)
    B = np.concatenate([A.T, (wt * np.sqrt(di2[n]))[:, None], (et * np.sqrt(di2[n + 1]))[:, None]], axis=1)
    # right-hand side of the reduced system
    bz0, bz1 = bz[:len(d)], bz[len(d):]
    rhs = bx + G0.T @ (di2 * bz0) + G1.T @ (Lam @ bz1.reshape(Np1, Np1) @ Lam).ravel()
    DinvB = B / D[:, None]
    cap = np.eye(B.shape[1]) + B.T @ DinvB                  # ((N+1)^2 + 2)^2: the only dense factorisation
    ux = rhs / D - DinvB @ np.linalg.solve(cap, DinvB.T @ rhs)
    uz0 = di2 * (G0 @ ux - bz0)
    uz1 = (Lam @ ((G1 @ ux) - bz1).reshape(Np1, Np1) @ Lam).ravel()
    return ux, np.concatenate([uz0, uz1])
