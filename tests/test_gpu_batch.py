"""Batched evaluation of small problems (blu_batch_*): P problems x B sample vectors in one launch must reproduce the
per-problem closures and the oracle, including the reference's edge cases (mosap.py:86-100, misc.py:453-505)."""
import numpy as np
import pytest

import oracle as orc
from conftest import maxrel

pytestmark = pytest.mark.gpu
TOL = 1e-12


def test_mosap_4x10_one_launch():
    """BASELINE config 4: 4 outputs x 10 models, all 1023 groups each, shared groups."""
    import bluest_b200 as blu
    N, No = 10, 4
    groups = blu.enumerate_groups(N)
    L = sum(len(g) for g in groups)
    Cs = [orc.wishart_cov(N, 10 + n) for n in range(No)]
    copy = lambda gs: [[list(g) for g in gk] for gk in gs]
    mos = blu.MOSAP(Cs, N, [N] * No, copy(groups), [copy(groups) for _ in range(No)], np.ones(L), [np.ones(L)] * No, verbose=False)
    assert mos._small_batch() is not None
    oracles = [orc.SapOracle(Cs[n], N, orc.enumerate_groups(N)) for n in range(No)]
    for m in (orc.dense_m(L, 0), orc.sparse_m(L, N, 2), np.round(orc.dense_m(L, 3)).astype(np.int64)):
        Vs = mos.variances(m)
        vs, gs, hs = mos.variance_GH(m, nohess=True)
        tol_g = TOL if np.count_nonzero(m) > 2 * N else 1e-9       # singular Phi: SURVEY 8a'
        for n in range(No):
            vo, go, _ = oracles[n].variance_GH(m[mos.mappings[n]], nohess=True)
            assert abs(Vs[n] - vo) <= TOL * vo and abs(vs[n] - vo) <= TOL * vo
            assert maxrel(gs[n], go) < tol_g and hs[n] is None
            v1, g1, _ = mos.SAPS[n].variance_GH(m[mos.mappings[n]], nohess=True)      # the per-context path
            assert abs(v1 - vs[n]) <= 1e-13 * vo and maxrel(gs[n], g1) < 1e-13
    assert mos.variances(0.01 * np.ones(L)) == [np.inf] * No                              # misc.py:464
    with pytest.raises(IndexError):
        mos.variance_GH(0.01 * np.ones(L), nohess=True)                                  # the reference fails the same way
    for s in mos.SAPS:
        s.close()


def test_evaluate_many_sweep_and_ragged_problems():
    """64 sample vectors of one problem in one launch; then a batch of problems of different sizes without maps."""
    import bluest_b200 as blu
    N = 10
    C = orc.wishart_cov(N, 1)
    groups = blu.enumerate_groups(N)
    L = sum(len(g) for g in groups)
    sap = blu.SAP(C, N, [[list(g) for g in gk] for gk in groups], np.ones(L), verbose=False)
    o = orc.SapOracle(C, N, orc.enumerate_groups(N))
    rng = np.random.RandomState(0)
    M = np.array([orc.dense_m(L, 100 + j) * 10.0 ** (3.0 * j / 63.0) for j in range(64)])
    M[5] = 0.0; M[5, :3] = 0.01                           # tiny: inf
    M[6] = orc.sparse_m(L, N, 4)
    var, flags, grad = blu.evaluate_many(sap, M)
    assert var.shape == (64,) and grad.shape == (64, L)
    for j in (0, 1, 6, 31, 63):
        vo, go, _ = o.variance_GH(M[j], nohess=True)
        assert abs(var[j] - vo) <= TOL * vo and maxrel(grad[j], go) < (1e-9 if j == 6 else TOL)
    assert np.isinf(var[5]) and np.all(np.isinf(grad[5])) and flags[5] & 1
    again = blu.evaluate_many(sap, M)
    assert np.array_equal(again[0], var) and np.array_equal(again[2], grad)              # deterministic
    # ragged batch: problems with different N, K and group counts, concatenated inputs
    N2, K2 = 7, 3
    C2 = orc.wishart_cov(N2, 5)
    g2 = blu.enumerate_groups(N2, K2)
    L2 = sum(len(g) for g in g2)
    sap2 = blu.SAP(C2, K2, [[list(g) for g in gk] for gk in g2], np.ones(L2), verbose=False)
    o2 = orc.SapOracle(C2, K2, orc.enumerate_groups(N2, K2))
    bt = blu.Batch([sap, sap2])
    Min = np.concatenate([M[:3], np.array([orc.dense_m(L2, j) for j in range(3)])], axis=1)
    var, flags, grads = bt.eval(Min)
    for j in range(3):
        vo, go, _ = o.variance_GH(Min[j, :L], nohess=True)
        assert abs(var[0, j] - vo) <= TOL * vo and maxrel(grads[0][j], go) < TOL
        vo, go, _ = o2.variance_GH(Min[j, L:], nohess=True)
        assert abs(var[1, j] - vo) <= TOL * vo and maxrel(grads[1][j], go) < TOL
    bt.close(); sap2.close(); sap.close()
