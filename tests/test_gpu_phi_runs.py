"""Phi(m) of the run-length streaming kernel (blu_phi.cuh) against the oracle on the inputs that exercise its
bookkeeping: dense, sparse and clustered-zero sample vectors (runs that skip unsampled groups), SHUFFLED group
orders (target changes are detected from the masks, not assumed from the enumeration order), both ring depths,
and bit-reproducibility from call to call.  Reference: psi @ m of misc.py:459-461 / sap.py:73 (group order)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def blu():
    import bluest_b200
    if bluest_b200.device_count() <= 0:
        pytest.fail("no CUDA device")
    return bluest_b200


CASES = [(10, 10, False, 2), (12, 5, False, 2), (9, 9, True, 2), (14, 14, False, 4), (24, 2, False, 2), (8, 3, True, 4), (25, 25, False, 2)]


@pytest.mark.parametrize("N,K,shuffle,stages", CASES)
def test_phi_runs_random_vectors(blu, N, K, shuffle, stages):
    C = orc.wishart_cov(N, N)
    rng = np.random.RandomState(N * 7 + K)
    if N == 25:                                          # groups of 24 and 25 members take the generic walk (more than 9 steps)
        import itertools
        allg = orc.enumerate_groups(N, 2)
        groups = [allg[0], allg[1]] + [[]] * 21 + [[list(c) for c in itertools.combinations(range(N), 24)], [list(range(N))]]
    else:
        groups = orc.enumerate_groups(N, K)
    if shuffle:
        groups = [[gk[i] for i in rng.permutation(len(gk))] for gk in groups]
    L = sum(len(g) for g in groups)
    o = orc.SapOracle(C, K, groups)
    sap = blu.SAP(C, K, [[list(g) for g in gk] for gk in groups], np.ones(L), verbose=False)
    sap.set_option("phi_stages", stages)
    for trial in range(12):
        m = rng.rand(L) * 10 ** rng.uniform(-3, 3, L)
        m[rng.rand(L) < (0.0, 0.5, 0.9, 0.99)[trial % 4]] = 0.0
        if trial >= 8:                                   # whole runs of consecutive groups unsampled
            for _ in range(20):
                a = rng.randint(L)
                m[a:a + rng.randint(1, 70)] = 0.0
        if trial == 11:
            m[:] = 0.0
            m[rng.randint(L)] = 3.0
        p, q = sap.get_phi(m), o.get_phi(m)
        assert np.max(np.abs(p - q)) <= 1e-12 * max(np.max(np.abs(q)), 1e-300), (trial,)
        assert np.array_equal(p, sap.get_phi(m))         # fixed summation order: bit-identical from call to call
    sap.close()
