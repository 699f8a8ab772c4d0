"""Sparsification of a multi-output allocation (mosap.py:125-211).

Given an allocation m with more than N non-zero groups, the stacked cleanup matrices X (N*n_outputs x L,
column i = what group i contributes to every output's estimator) have a null space on the support of m:
moving along a null direction leaves every output's variance unchanged to first order.  Directions are
oriented so that the cost does not rise, tried from the steepest cost decrease down, and followed as far
as positivity of m and the coverage constraints e_n . m >= 1 allow; a step is kept when the largest output
variance does not get worse (relative 1e-4).  Repeats until no direction makes progress or at most N groups
remain.  Behaviour (including zeroing the caller's entries below ``tol`` in place) follows the reference."""
import numpy as np


def _largest_feasible_step(direction, coverage_rate, coverage_now, m_support):
    """How far m + s * direction stays non-negative and keeps every output covered."""
    losing = np.argwhere(coverage_rate < 0).flatten()
    s_cover = np.inf if len(losing) == 0 else min(abs(coverage_now[losing] - 1) / abs(coverage_rate[losing]))
    shrinking = np.argwhere(direction < 0).flatten()
    s_positive = np.inf if len(shrinking) == 0 else min(m_support[shrinking] / abs(direction[shrinking]))
    return max(min(s_cover, s_positive), 0)


def cleanup_solution(mosap, m, delta=0, tol=0):
    from scipy.linalg import null_space
    costs = mosap.costs
    coverage = mosap.output_indicators()                        # (n_outputs, L): e restricted to each output's groups
    worst = lambda x: max(mosap.variances(x, delta=delta))
    support = np.argwhere(m > tol).flatten()
    v_start = worst(m)
    step = 0
    while len(support) > mosap.N:
        support = np.argwhere(m > tol).flatten()
        m[m < tol] = 0
        cov_s = coverage[:, support]
        basis = null_space(mosap.get_cleanup_matrices(m, delta=delta)[:, support])
        slope = costs[support] @ basis
        uphill = np.sign(slope) > 0
        basis[:, uphill] *= -1                                  # every direction now lowers (or keeps) the cost
        slope[uphill] *= -1
        moving = abs(np.sign(slope)) > 0
        basis, slope = basis[:, moving], slope[moving]
        if len(slope) == 0:
            break
        coverage_now = cov_s @ m[support]
        for i in np.argsort(abs(slope))[::-1]:                  # steepest cost decrease first
            direction = basis[:, i]
            step = _largest_feasible_step(direction, cov_s @ direction, coverage_now, m[support])
            if step > 5 * tol:
                full = np.zeros_like(m); full[support] = direction
                trial = m + step * full
                v_trial = worst(trial)
                if v_trial < v_start or abs(v_trial - v_start) / abs(v_start) < 1.0e-4:
                    m = trial.copy()
                    break
                step = 0
        if step <= 5 * tol:
            break
    m[m < tol] = 0
    return m
