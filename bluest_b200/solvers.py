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


def scipy_solve_multi(problem, budget=None, eps=None, x0=None, max_model_samples=None, maxiter=5000, verbose=False, counters=None,
                      hess="dense", sparse_constraints=False, reference_eps_bound=True):
    """Multi-output driver kept from the reference (mosap.py:555-610): trust-constr on
    budget mode:  min t  s.t.  t >= V_n(m[mappings[n]]) for every output, w.m <= budget   (x = [t, m])
    eps mode:     min w.m/|w|  s.t.  V_n(m[mappings[n]]) <= eps_n^2
    with every output covered (e_n.m >= 1), m >= 0 and the optional per-model sample caps.  ``problem``
    is a MOSAP-like object (``SAPS``, ``mappings``, ``costs``, ``e``, ``L``, ``variances``,
    ``check_input``, ``get_max_sample_constraints``).

    The reference hands scipy each output's gradient / Hessian in that output's OWN group numbering
    (length L_n), which only fits when every output uses every group; here they are scattered to the
    union numbering through ``mappings[n]`` (identical whenever the reference's call is well-formed).
    ``hess="operator"`` / ``sparse_constraints``: as in ``scipy_solve``.

    ``reference_eps_bound`` (default True): in eps mode the reference bounds EVERY output's variance by
    the LAST output's tolerance -- the upper bound at mosap.py:605 reads ``epsq[n]`` with the ``n`` left
    over from the loop at mosap.py:577, not the comprehension's ``nn``.  Kept by default so that the
    drop-in returns the reference's allocations; False bounds output n by its own eps_n**2."""
    from scipy.optimize import Bounds, LinearConstraint, NonlinearConstraint, minimize
    from scipy.sparse.linalg import LinearOperator
    if hess not in ("dense", "operator"):
        raise ValueError("hess must be 'dense' or 'operator'")
    budget, eps = problem.check_input(budget, eps)
    delta = 1.0e-15
    L, No = int(problem.L), problem.n_outputs
    w, e, mappings, SAPS = problem.costs, problem.e, problem.mappings, problem.SAPS
    ES, rhs = problem.get_max_sample_constraints(max_model_samples)
    cnt = counters if counters is not None else {}
    cnt.setdefault("f", 0); cnt.setdefault("g", 0); cnt.setdefault("H", 0)
    if sparse_constraints:
        import scipy.sparse as sps
        row = lambda v: sps.csr_matrix(np.asarray(v, dtype=np.float64).reshape(1, -1))
    else:
        row = lambda v: v
    es = []
    for n in range(No):
        ee = np.zeros((L,)); ee[mappings[n]] = e[mappings[n]]
        es.append(ee)
    off = 1 if budget is not None else 0                   # budget mode carries the epigraph variable t in x[0]
    sign = -1.0 if budget is not None else 1.0             # constraint is t - V_n >= 0 there, V_n <= eps^2 here

    def var_n(x, n):
        cnt["f"] += 1
        return SAPS[n].variance(x[off:][mappings[n]], delta=delta)

    def grad_n(x, n):
        cnt["g"] += 1
        g = np.zeros(L + off)
        g[off:][mappings[n]] = sign * SAPS[n].variance_GH(x[off:][mappings[n]], nohess=True, delta=delta)[1]
        if off:
            g[0] = 1.0
        return row(g)

    def hess_n(x, p, n):
        cnt["H"] += 1
        mp = mappings[n]
        if hess == "dense":
            H = np.zeros((L + off, L + off))
            H[np.ix_(off + mp, off + mp)] = sign * SAPS[n].variance_GH(x[off:][mp], delta=delta)[2]
            return H * p
        op = SAPS[n].variance_GH_operator(x[off:][mp], delta=delta)[2]
        scale = sign * float(np.asarray(p).ravel()[0])

        def mv(v):
            v = np.asarray(v, dtype=np.float64).ravel()
            out = np.zeros(L + off)
            out[off + mp] = scale * op.matvec(np.ascontiguousarray(v[off + mp]))
            return out
        return LinearOperator((L + off, L + off), matvec=mv, rmatvec=mv, dtype=np.float64)

    opts = {"factorization_method": None, "disp": False, "maxiter": maxiter, "verbose": 3 * int(verbose)}
    pad = (lambda v: np.concatenate([[0], v])) if off else (lambda v: v)
    bounds = Bounds(0.0 * np.ones((L + off,)), np.inf * np.ones((L + off,)), keep_feasible=True)
    cover = [LinearConstraint(row(pad(ee)), 1, np.inf, keep_feasible=True) for ee in es]
    caps = [LinearConstraint(row(pad(ees)), -np.inf, rr) for ees, rr in zip(ES, rhs)]
    if budget is not None:
        cost = LinearConstraint(row(pad(w)), -np.inf, budget)
        epi = [NonlinearConstraint(lambda x, n=n: x[0] - var_n(x, n), 0, np.inf, jac=lambda x, n=n: grad_n(x, n),
                                   hess=lambda x, p, n=n: hess_n(x, p, n)) for n in range(No)]
        if x0 is None:
            x0 = np.ceil(budget * abs(np.random.randn(L)))
        if len(x0) == L:
            x0 = np.concatenate([[max(problem.variances(x0, delta=delta))], x0])
        eee = np.zeros((L + 1,)); eee[0] = 1
        res = minimize(lambda x: (x[0], eee), x0, jac=True, hessp=lambda x, p: np.zeros((len(x),)), bounds=bounds,
                       constraints=[cost] + cover + epi + caps, method="trust-constr", options=opts, tol=1.0e-7)
        res.x = res.x[1:]
    else:
        epsq = eps ** 2
        ub = (lambda n: epsq[No - 1]) if reference_eps_bound else (lambda n: epsq[n])
        acc = [NonlinearConstraint(lambda x, n=n: var_n(x, n), -np.inf, ub(n), jac=lambda x, n=n: grad_n(x, n),
                                   hess=lambda x, p, n=n: hess_n(x, p, n)) for n in range(No)]
        if x0 is None:
            x0 = np.ceil(np.linalg.norm(eps) ** -2 * np.random.rand(L))
        wn = w / np.linalg.norm(w)
        res = minimize(lambda x: [wn @ x, wn], x0, jac=True, hessp=lambda x, p: np.zeros((len(x),)), bounds=bounds,
                       constraints=acc + cover + caps, method="trust-constr", options=opts, tol=1.0e-7)
    return res
