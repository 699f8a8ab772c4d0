"""Integer projection of a continuous sample allocation ("next" row f2 of the scope table).

Host-side candidate enumeration with the reference's exact rules (which groups are rounded, the
order of the floor/ceil combinations, the feasibility filters and the tie-breaking of
``best_closest_integer_solution_BLUE``, misc.py:141-165, 313-382) around ONE batched device call:
for all surviving candidates at once ``Phi_c = Phi(base) + sum_t ms[t,c] Psi_idx[t]`` and
``V_c = pinv(Phi_c, hermitian, rcond=1e-10)[0,0]`` (misc.py:368-369) --
``blu_candidate_variances``.  The dense ``psi`` matrix the reference multiplies with is never formed.
"""
import numpy as np

from ._lib import check, dptr, iptr, lib


def feasible_integer_bounds(sol, N, e=None):
    """Which entries of ``sol`` get a floor/ceil choice, and in which order (misc.py:141-165)."""
    L = len(sol)
    order = np.argsort(sol)[-int(1.2 * N):]
    idx = np.array([i for i in order if sol[i] > 1.0e-8])
    if e is not None:
        if sum(e > 0.99) == 0:
            val = 1 / sum(e) / 2
            while sum(e > val) == 0:
                val /= 2
        else:
            val = 0.99
        idx2 = np.argwhere(e > val).flatten()
        by_size = np.argsort(sol[e > val])[::-1]
        idx2 = idx2[by_size[:N]]
        idx = np.unique(np.concatenate([idx, idx2]))
    lb = np.zeros((L,), dtype=int); ub = np.zeros((L,), dtype=int)
    lb[idx] = np.floor(sol).astype(int)[idx]
    ub[idx] = np.ceil(sol).astype(int)[idx]
    idx = idx[np.argsort(lb[idx])[::-1]]
    return lb[idx], ub[idx], idx


def _combinations(lb, ub):
    """(LL, 2^LL) matrix whose column c takes ub[t] where bit t of c is set, else lb[t]
    (misc.py:322-325: unpackbits + bnds[combs, ee].T)."""
    LL = len(lb)
    c = np.arange(2 ** LL, dtype=int)
    bits = (c[None, :] >> np.arange(LL, dtype=int)[:, None]) & 1
    return np.where(bits.astype(bool), ub[:, None], lb[:, None])


def candidate_variances(sap, base, idx, ms, rcond=1.0e-10):
    """V_c for every column of ``ms`` on the device; ``base`` = integer allocation with the idx entries zeroed."""
    basephi = np.ascontiguousarray(sap.get_phi(base))
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    ms = np.ascontiguousarray(ms, dtype=np.int64)
    Vs = np.empty(ms.shape[1])
    if ms.shape[1]:
        check(lib().blu_candidate_variances(sap._ctx, dptr(basephi), int(len(idx)), iptr(idx), iptr(ms), int(ms.shape[1]), float(rcond), dptr(Vs)))
    return Vs


def best_closest_integer_solution_BLUE(sap, sol, budget=None, eps=None, max_samples_info=([], [])):
    """misc.py:313-382 with ``sap`` in place of (psi, w, e)."""
    w, e = sap.costs, sap.e
    ES, rhs = max_samples_info
    lb, ub, idx = feasible_integer_bounds(sol, sap.N, e=e)
    LL = len(idx)
    if LL > 24:
        raise ValueError('Too many dimensions to brute-force it')
    ms = _combinations(lb, ub)

    val = np.round(sol).astype(int)
    baseval = val.copy(); baseval[idx] = 0
    basecost = w @ baseval
    basee = e @ baseval
    base_checks = [ees @ baseval for ees in ES]

    if basee < 1:                                            # model 0 must be sampled at least once
        keep = np.argwhere(basee + e[idx] @ ms >= 1).flatten()
        if len(keep) == 0:
            return None, np.inf
        ms = ms[:, keep]
    if len(ES) > 0:
        if any(b > rr for b, rr in zip(base_checks, rhs)):
            return None, np.inf
        checks = [b + ees[idx] @ ms for b, ees in zip(base_checks, ES)]
        keep = np.argwhere(np.all([c <= rr for c, rr in zip(checks, rhs)], axis=0)).flatten()
        if len(keep) == 0:
            return None, np.inf
        ms = ms[:, keep]
    if budget is not None and basecost > budget:
        return None, np.inf

    costs = basecost + w[idx] @ ms
    if budget is not None:
        ms = ms[:, np.argwhere(costs <= 1.0001 * budget).flatten()][:, ::-1]
    else:
        ms = ms[:, np.argsort(costs)[::-1]]
    if np.prod(ms.shape) == 0:
        return None, np.inf

    Vs = candidate_variances(sap, baseval, idx, ms, rcond=1.0e-10)

    if budget is not None:
        i = np.argmin(Vs)
    else:
        ok = np.argwhere(Vs <= 1.0001 * eps ** 2).flatten()
        if len(ok) == 0:
            return None, np.inf
        i = ok[-1]
    val[idx] = ms[:, i]
    return val, Vs[i]


def integer_projection(sap, samples, budget=None, eps=None, max_model_samples=None):
    """SAP.integer_projection (sap.py:145-187), including its fallback ladder."""
    if budget is None and eps is None:
        raise ValueError("Need to specify either budget or RMSE tolerance")
    ss = samples.copy()
    es, rhs = sap.get_max_sample_constraints(max_model_samples)
    out, fval = best_closest_integer_solution_BLUE(sap, ss, budget=budget, eps=eps, max_samples_info=(es, rhs))
    if np.isinf(fval):
        # sap.py:163-170 retries with the SAME budget/eps (the enlarged ones are computed but not
        # passed on), so the retries cannot succeed; one repeat keeps the behaviour without the noise.
        out, fval = best_closest_integer_solution_BLUE(sap, ss, budget=budget, eps=eps, max_samples_info=(es, rhs))
    if np.isinf(fval):
        if max_model_samples is not None and not all(np.ceil(ss) @ ee <= rr for ee, rr in zip(es, rhs)):
            out = np.floor(ss)
            if not out @ sap.e >= 1.0:
                out = np.ceil(ss)
        else:
            out = np.ceil(ss)
    return out.astype(int)
