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


# ---------------------------------------------------------------------------------------------
# multi-output projection (misc.py:177-311, mosap.py:213-292)
# ---------------------------------------------------------------------------------------------
LL_MAX_MULTI = 15            # misc.py:188: at most 2^15 brute-force candidates, the rest randomised


def _multi_helper(mosap, sol, budget, eps, lb, ub, idx, max_samples_info=([], [])):
    """best_closest_integer_solution_BLUE_multi_helper (misc.py:228-311) with the outputs' device
    contexts in place of ``psis``: one batched ``blu_candidate_variances`` call per output, default
    pinv cutoff (misc.py:294)."""
    from functools import reduce
    ES, rhs = max_samples_info
    w, e, mappings = mosap.costs, mosap.e, mosap.mappings
    No = len(mappings)
    ms = _combinations(lb, ub)

    val = np.round(sol).astype(int)
    baseval = val.copy(); baseval[idx] = 0
    basecost = w @ baseval
    basees = [e[mappings[n]] @ baseval[mappings[n]] for n in range(No)]
    base_checks = [ees @ baseval for ees in ES]

    # position of every brute-forced group inside each output's own group list (built once per MOSAP: the
    # randomised search calls this helper up to 250 times)
    pos = getattr(mosap, "_intproj_positions", None)
    if pos is None or len(pos) != No:
        pos = [{int(g): j for j, g in enumerate(mappings[n])} for n in range(No)]
        try:
            mosap._intproj_positions = pos
        except AttributeError:
            pass
    redmaps = [np.array([i for i in range(len(idx)) if int(idx[i]) in pos[n]], dtype=int) for n in range(No)]
    idxs = [np.array([pos[n][int(g)] for g in idx if int(g) in pos[n]], dtype=np.int64) for n in range(No)]

    es = []
    for n in range(No):
        if basees[n] < 1:
            es.append(np.argwhere(basees[n] + e[idx][redmaps[n]] @ ms[redmaps[n], :] >= 1).flatten())
    if len(es) == 0:
        return None, np.inf                              # misc.py:264 (sic: also when every output is already covered)
    es = np.unique(np.concatenate(es))
    ms = ms[:, es]

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
        ind = np.argwhere(costs <= 1.0001 * budget).flatten()
        if len(ind) > 0:
            ms = ms[:, ind][:, ::-1]
        else:
            return None, np.inf
    else:
        ms = ms[:, np.argsort(costs)[::-1]]
    if np.prod(ms.shape) == 0:
        return None, np.inf

    Vs = []
    for n in range(No):
        sap = mosap.SAPS[n]
        if len(idxs[n]):
            Vs.append(candidate_variances(sap, baseval[mappings[n]], idxs[n], ms[redmaps[n], :], rcond=1.0e-15))
        else:                                            # no brute-forced group belongs to this output
            v = candidate_variances(sap, baseval[mappings[n]], np.zeros(1, dtype=np.int64), np.zeros((1, 1), dtype=np.int64), rcond=1.0e-15)
            Vs.append(np.full(ms.shape[1], v[0]))
    V_max = reduce(np.maximum, Vs)

    if budget is not None:
        i = np.argmin(V_max)
    else:
        i = np.argwhere(reduce(np.logical_and, [Vs[n] <= 1.0001 * eps[n] ** 2 for n in range(No)])).flatten()
        if len(i) > 0:
            i = i[-1]
        else:
            return None, np.inf
    val[idx] = ms[:, i]
    return val, V_max[i]


def best_closest_integer_solution_BLUE_multi(mosap, sol, budget=None, eps=None, max_samples_info=([], []), verbose=False):
    """misc.py:177-226.  Up to 15 groups are brute-forced; with more, the reference fixes a random
    floor/ceil choice for the surplus groups and retries up to 250 times.  The calls into
    ``np.random`` (one ``permutation`` and one ``randint`` per trial) are made in the reference's
    order, so a caller that seeds ``np.random`` gets the reference's allocation."""
    lb_full, ub_full, idx_full = feasible_integer_bounds(sol, mosap.N, e=mosap.e)
    LL = len(idx_full)
    if LL <= LL_MAX_MULTI:
        return _multi_helper(mosap, sol, budget, eps, lb_full, ub_full, idx_full, max_samples_info=max_samples_info)
    if verbose:
        print('WARNING! Too many dimensions to brute-force it. Randomising search. Note: result might not be optimal.')
    best_val, best_fval = None, np.inf
    trial = 0
    while best_val is None and trial < 250:
        trial += 1
        sample = np.random.permutation(LL)
        brute_force, random_choice = sample[:LL_MAX_MULTI], sample[LL_MAX_MULTI:]
        idx, lb, ub = idx_full[brute_force], lb_full[brute_force], ub_full[brute_force]
        LR = LL - LL_MAX_MULTI
        r_idx = idx_full[random_choice]
        r_bnds = np.vstack([lb_full[random_choice], ub_full[random_choice]])
        comb = np.random.randint(2, size=LR)
        r_sol = sol.copy()
        r_sol[r_idx] = r_bnds[comb, np.arange(LR)]
        best_val, best_fval = _multi_helper(mosap, r_sol, budget, eps, lb, ub, idx, max_samples_info=max_samples_info)
    if trial >= 100 and best_val is None:                # misc.py:221 (sic: 100, not 250)
        return None, np.inf
    return best_val, best_fval


def integer_projection_multi(mosap, samples, budget=None, eps=None, max_model_samples=None):
    """MOSAP.integer_projection (mosap.py:213-292) with its whole fallback ladder."""
    if budget is None and eps is None:
        raise ValueError("Need to specify either budget or RMSE tolerance")

    def increase_tolerance(budget, eps, fac):
        b = None if budget is None else budget * (1 + fac)
        e_ = None if eps is None else np.sqrt(np.array(eps) ** 2 * (1 + fac))
        return b, e_

    ss = samples.copy()
    ES, rhs = mosap.get_max_sample_constraints(max_model_samples)
    proj = lambda s, b, e_: best_closest_integer_solution_BLUE_multi(mosap, s, budget=b, eps=e_, max_samples_info=(ES, rhs), verbose=mosap.verbose)
    # STEP 0: standard
    out, fval = proj(ss, budget, eps)
    # STEP 1: cleanup + standard
    if np.isinf(fval):
        css = mosap.cleanup_solution(ss)
        out, fval = proj(css, budget, eps)
    # STEP 2: increase tolerances
    if np.isinf(fval):
        for i in reversed(range(4)):
            nb, ne = increase_tolerance(budget, eps, 10. ** -i)
            out, fval = proj(ss, nb, ne)
            if np.isinf(fval):
                out, fval = proj(css, nb, ne)
            if not np.isinf(fval):
                break
    # STEP 3: round up / down
    if np.isinf(fval):
        maxvar = lambda s: max(mosap.variances(s))
        ssf = np.floor(ss); ss = np.ceil(ss)
        cssf = np.floor(css); css = np.ceil(css)
        tot_ss, tot_css = ss @ mosap.costs, css @ mosap.costs
        var_ss, var_css = maxvar(ss), maxvar(css)
        covered = lambda s: all(s[mosap.mappings[n]] @ mosap.e[mosap.mappings[n]] >= 1 for n in range(mosap.n_outputs))
        pick_up = lambda: (ss if tot_ss < tot_css else css) if eps is None else (ss if var_ss < var_css else css)
        if max_model_samples is not None:
            if all(ss @ ees <= rr for ees, rr in zip(ES, rhs)):
                out = ss
            elif all(css @ ees <= rr for ees, rr in zip(ES, rhs)):
                out = css
            elif covered(ssf):
                out = ssf
            elif covered(cssf):
                out = cssf
            else:
                out = pick_up()
        else:
            out = pick_up()
    return out.astype(int)
