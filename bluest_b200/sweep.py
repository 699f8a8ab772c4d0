"""Budget / tolerance sweeps of independent SAP instances across the GPUs of one box.

BASELINE config 3 ("budget sweep of 64 SAP instances across GPUs"): the
instances share the covariance, the groups and the costs and differ only in the budget (or the
tolerance), so they are embarrassingly parallel -- every rank builds its own context on its own GPU
(15.7 MB of packed inverses at 15 models) and solves a contiguous share of the instances; there is
no collective on the data path, only one gather of the small results at the end.

The per-instance work is exactly ``SAP.solve`` (``sap.py:189-220``: continuous solve, then the
integer projection unless ``continuous_relaxation``).  ``make_problem`` is a factory so that the CPU
tests can drive the same code with an oracle-backed stand-in under gloo.
"""
import numpy as np


def split_instances(n, world, rank):
    """Contiguous, balanced share [lo, hi) of n instances for ``rank`` of ``world``."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank %r/%r" % (world, rank))
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


def solve_sweep(make_problem, budgets=None, epss=None, x0=None, dist=None, group=None, rank=0, world=1,
                continuous_relaxation=False, max_model_samples=None, solve_kwargs=None):
    """Solve one SAP instance per entry of ``budgets`` (or ``epss``); returns, on every rank, the list of
    results in instance order: dicts with ``budget`` / ``eps``, ``samples``, ``variance``, ``cost``, ``rank``.

    make_problem() -> an object with ``solve(budget=, eps=, x0=, continuous_relaxation=, max_model_samples=, **kw)``,
    ``variance(m)`` and ``costs`` -- ``bluest_b200.SAP`` bound to this rank's GPU in production.
    x0: None, one start vector for all instances, or a callable ``x0(i, value) -> vector``.
    dist: ``torch.distributed`` (initialised) or None for a single process."""
    if (budgets is None) == (epss is None):
        raise ValueError("Need to specify either budgets or RMSE tolerances (not both)")
    values = np.atleast_1d(np.asarray(budgets if budgets is not None else epss, dtype=np.float64))
    key = "budget" if budgets is not None else "eps"
    if dist is not None:
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = split_instances(len(values), world, rank)
    problem = make_problem() if hi > lo else None
    kw = dict(solve_kwargs or {})
    mine = []
    for i in range(lo, hi):
        start = x0(i, values[i]) if callable(x0) else (None if x0 is None else np.array(x0, dtype=np.float64, copy=True))
        samples = problem.solve(**{key: float(values[i])}, x0=start, continuous_relaxation=continuous_relaxation,
                                max_model_samples=max_model_samples, **kw)
        if samples is None:
            mine.append({"index": i, key: float(values[i]), "samples": None, "variance": np.inf, "cost": np.inf, "rank": rank})
            continue
        samples = np.asarray(samples)
        mine.append({"index": i, key: float(values[i]), "samples": samples, "variance": float(problem.variance(samples)),
                     "cost": float(samples @ problem.costs), "rank": rank})
    if dist is None or world == 1:
        return mine
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    out = [r for part in gathered for r in part]
    out.sort(key=lambda r: r["index"])
    return out
