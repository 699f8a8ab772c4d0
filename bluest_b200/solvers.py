"""Host-side solver driver kept from the reference (sap.py:378-418): scipy ``trust-constr`` iterating
on the objective closures.  Works on any object exposing ``L, costs, e, variance, variance_GH,
get_max_sample_constraints`` -- the B200-backed SAP in production, an oracle-backed stand-in when
bench.py times the CPU reference with the identical driver."""
import numpy as np


def _scaled(op, p):
    """p * H for a LinearOperator H (the eps-mode constraint Hessian of sap.py:415)."""
    from scipy.sparse.linalg import LinearOperator
    mv = lambda v: p * op.matvec(v)
    return LinearOperator(op.shape, matvec=mv, rmatvec=mv, dtype=op.dtype)


def scipy_solve(problem, budget=None, eps=None, x0=None, max_model_samples=None, maxiter=1000, verbose=False, counters=None,
                hess="dense", sparse_constraints=False):
    """``hess``: "dense" -- the reference's callbacks (sap.py:410,416: a dense (L,L) array per Hessian
    evaluation); "operator" -- ``problem.variance_GH_operator`` (Hessian factored in HBM, handed to
    trust-constr as a LinearOperator; its projected-CG only multiplies by it).
    ``sparse_constraints``: pass the linear constraint rows (sap.py:403-406) as scipy.sparse matrices.
    With the reference's dense rows trust-constr turns the bounds into a dense (L,L) identity and
    QR-factorises (L+2, 2L+2) Jacobians every iteration -- 8.6 GB and O(L^3) at 15 models; with sparse
    rows it walks the same iterates through a sparse LU.  Both options together make the 32767-group
    problem solvable at all (seconds) and leave the closures as the only O(L) work per iteration."""
    from scipy.optimize import Bounds, LinearConstraint, NonlinearConstraint, minimize
    if budget is None and eps is None:
        raise ValueError("Need to specify either budget or RMSE tolerance")
    delta = 0
    L = int(problem.L)
    w = problem.costs
    e = problem.e
    es, rhs = problem.get_max_sample_constraints(max_model_samples)
    cnt = counters if counters is not None else {}
    cnt.setdefault("f", 0); cnt.setdefault("g", 0); cnt.setdefault("H", 0)

    def fg(x):
        cnt["g"] += 1
        return problem.variance_GH(x, nohess=True, delta=delta)[:-1]

    if hess not in ("dense", "operator"):
        raise ValueError("hess must be 'dense' or 'operator'")
    hess_mode = hess

    def hess(x):
        cnt["H"] += 1
        if hess_mode == "operator":
            return problem.variance_GH_operator(x, delta=delta)[-1]
        return problem.variance_GH(x, delta=delta)[-1]

    def var(x):
        cnt["f"] += 1
        return problem.variance(x, delta=delta)

    def jac(x):
        cnt["g"] += 1
        return problem.variance_GH(x, nohess=True, delta=delta)[1]

    if sparse_constraints:
        import scipy.sparse as sps
        row = lambda v: sps.csr_matrix(np.asarray(v, dtype=np.float64).reshape(1, -1))
    else:
        row = lambda v: v
    constraint1 = Bounds(0.0 * np.ones((L,)), np.inf * np.ones((L,)), keep_feasible=True)
    constraint3 = LinearConstraint(row(e), 1, np.inf, keep_feasible=True)
    constraint4 = [LinearConstraint(row(ee), -np.inf, rr) for ee, rr in zip(es, rhs)]
    opts = {"factorization_method": None, "disp": False, "maxiter": maxiter, "verbose": 3 * int(verbose)}
    if budget is not None:
        constraint2 = LinearConstraint(row(w), -np.inf, budget)
        if x0 is None:
            x0 = np.ceil(10 * abs(np.random.randn(L)))
        res = minimize(fg, x0, jac=True, hess=hess, bounds=constraint1, constraints=[constraint2, constraint3] + constraint4,
                       method="trust-constr", options=opts, tol=1.0e-8)
    else:
        epsq = eps ** 2
        constraint2 = NonlinearConstraint(var, epsq, epsq, jac=(lambda x: row(jac(x))) if sparse_constraints else jac, hess=lambda x, p: hess(x) * p if hess_mode == "dense" else _scaled(hess(x), p))
        if x0 is None:
            x0 = np.ceil(eps ** -2 * np.random.rand(L))
        wn = w / np.linalg.norm(w)
        res = minimize(lambda x: [wn @ x, wn], x0, jac=True, hessp=lambda x, p: np.zeros((len(x),)), bounds=constraint1,
                       constraints=[constraint2, constraint3] + constraint4, method="trust-constr", options=opts, tol=1.0e-10)
    return res
