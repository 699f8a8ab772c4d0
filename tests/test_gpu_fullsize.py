"""Parity at BASELINE.json's full size (15 models, all 32767 groups): the oracle still finishes the
setup, Phi, variance and gradient in seconds, so those are compared in full; the 8.59 GB Hessian is
checked through size-independent properties (exact symmetry, sampled rows against the factored form
built from the oracle's U and pinv, directional finite differences of the gradient)."""
import numpy as np
import pytest

import oracle as orc
from conftest import maxrel

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def problem():
    import bluest_b200 as blu
    if blu.device_count() <= 0:
        pytest.fail("no CUDA device")
    N = 15
    C = orc.wishart_cov(N, 0)
    groups = orc.enumerate_groups(N)
    L = sum(len(g) for g in groups)
    o = orc.SapOracle(C, N, groups)
    sap = blu.SAP(C, N, [[list(g) for g in gk] for gk in groups], np.ones(L), verbose=False)
    yield N, L, o, sap
    sap.close()


def test_group_enumeration_and_inverses_full_size(problem):
    N, L, o, sap = problem
    assert L == 2 ** N - 1 == 32767 and sap.n_fallback == 0
    flat = np.concatenate([g.ravel() for g in sap.groups])
    assert np.array_equal(flat, np.concatenate([g.ravel() for g in o.groups]))          # bit-exact indexing
    for k in range(N):
        assert maxrel(sap.invcovs[k], o.invcovs[k]) < TOL


def test_phi_variance_gradient_full_size(problem):
    N, L, o, sap = problem
    m1, m2 = orc.dense_m(L, 0), orc.dense_m(L, 1)
    assert maxrel(sap.get_phi(m1), o.get_phi(m1)) < TOL
    # linearity of Phi in m (delta = 0)
    assert maxrel(sap.get_phi(2.5 * m1 - 0.5 * m2), 2.5 * sap.get_phi(m1) - 0.5 * sap.get_phi(m2)) < 1e-13
    vo, go, _ = o.variance_GH(m1, nohess=True)
    v, g, _ = sap.variance_GH(m1, nohess=True)
    assert abs(v - vo) <= TOL * vo and maxrel(g, go) < TOL
    assert abs(sap.variance(m1) - o.variance(m1)) <= TOL * vo
    ms = orc.sparse_m(L, N, 3)
    vo, go, _ = o.variance_GH(ms, nohess=True)
    v, g, _ = sap.variance_GH(ms, nohess=True)
    assert abs(v - vo) <= TOL * vo and maxrel(g, go) < 1e-9


def test_dense_hessian_full_size_properties(problem):
    N, L, o, sap = problem
    m = orc.dense_m(L, 0)
    v, g, H = sap.variance_GH(m)
    assert H.shape == (L, L)
    # exact symmetry of the full 8.59 GB matrix (blockwise to bound the temporaries)
    for r0 in range(0, L, 4096):
        blk = H[r0:r0 + 4096]
        assert np.array_equal(blk, H[:, r0:r0 + 4096].T)
    # sampled rows against 2 U^T pinv(Phi) U from the oracle
    P = np.linalg.pinv(o.get_phi(m))
    U = o.ufactor(np.ascontiguousarray(P[0]))              # (N, L)
    rows = np.array([0, 1, 14, 15, 119, 120, 4095, 4096, 16383, 20000, 32703, 32704, 32766])
    ref = 2.0 * (U[:, rows].T @ P @ U)
    assert maxrel(H[rows], ref) < TOL
    assert maxrel(np.diag(H), 2.0 * np.einsum("al,ab,bl->l", U, P, U)) < TOL
    # the gradient is the derivative of the variance, the Hessian of the gradient (central differences)
    rng = np.random.RandomState(5)
    d = rng.randn(L); d /= np.linalg.norm(d)
    h = 1e-3
    vp, gp, _ = sap.variance_GH(m + h * d, nohess=True)
    vm, gm, _ = sap.variance_GH(m - h * d, nohess=True)
    assert abs((vp - vm) / (2 * h) - g @ d) <= 1e-6 * abs(g @ d)
    Hd = H @ d
    assert maxrel((gp - gm) / (2 * h), Hd) < 1e-5
    del H


def test_hessian_operator_full_size(problem):
    """15 models: the operator must reproduce the dense 8.59 GB Hessian times p without forming it."""
    N, L, o, sap = problem
    m = orc.dense_m(L, 0)
    P = np.linalg.pinv(o.get_phi(m))
    U = o.ufactor(np.ascontiguousarray(P[0]))
    vo, go, _ = o.variance_GH(m, nohess=True)
    v, g, op = sap.variance_GH_operator(m)
    assert abs(v - vo) <= TOL * vo and maxrel(g, go) < TOL
    rng = np.random.RandomState(11)
    for _ in range(3):
        p = rng.randn(L)
        assert maxrel(op @ p, 2.0 * (U.T @ (P @ (U @ p)))) < TOL
    # H is positive semi-definite (2 U^T pinv(Phi) U): p^T H p >= 0, and = 0 never for random p
    p = rng.randn(L)
    assert p @ (op @ p) > 0
    # directional second derivative of the variance
    d = rng.randn(L); d /= np.linalg.norm(d)
    h = 1e-3
    _, gp, _ = sap.variance_GH(m + h * d, nohess=True)
    _, gm, _ = sap.variance_GH(m - h * d, nohess=True)
    v2, g2, op2 = sap.variance_GH_operator(m)
    assert maxrel((gp - gm) / (2 * h), op2 @ d) < 1e-5


def test_phi_variance_gradient_20_models():
    """20 models, all 1 048 575 groups (BASELINE config 5) against an INDEPENDENT value: the oracle with its
    per-class batched LAPACK inverses (pinned against the reference's pinv blocks in tests/test_oracle.py) and the
    restated native loops for Phi and the gradient.  Inverses, Phi, variance and gradient in full."""
    import bluest_b200 as blu
    N = 20
    C = orc.wishart_cov(N, 0)
    ga = orc.enumerate_group_arrays(N)
    o = orc.SapOracle(C, N, ga, invcovs=orc.batched_invcovs(C, ga), with_ES=False)
    L = o.L
    assert L == 2 ** N - 1
    groups = blu.enumerate_groups(N)
    sap = blu.SAP(C, N, groups, np.ones(L), verbose=False)
    assert sap.n_fallback == 0
    for k in (1, 2, 7, 10, 13, 19, 20):                      # sampled size classes of the 881 MB of inverses
        assert maxrel(sap.invcovs[k - 1], o.invcovs[k - 1]) < TOL
    assert np.array_equal(np.asarray(sap.groups[9]), ga[9])  # bit-exact enumeration of the largest class
    sap._invcovs = None
    for seed in (0, 1):
        m = orc.dense_m(L, seed)
        assert maxrel(sap.get_phi(m), o.get_phi(m)) < TOL
        vo, go, _ = o.variance_GH(m, nohess=True)
        v, g, _ = sap.variance_GH(m, nohess=True)
        assert abs(v - vo) <= TOL * vo
        assert maxrel(g, go) < TOL
        assert abs(sap.variance(m) - o.variance(m)) <= TOL * vo
    ms = orc.sparse_m(L, N, 3)                               # an optimiser-like iterate: 2N non-zeros, singular Phi
    vo, go, _ = o.variance_GH(ms, nohess=True)
    v, g, _ = sap.variance_GH(ms, nohess=True)
    assert abs(v - vo) <= TOL * vo and maxrel(g, go) < 1e-9
    sap.close()


def test_hessian_operator_20_models():
    """20 models, 1 048 575 groups: the dense Hessian (8.8 TB) cannot exist; the operator is checked
    against central differences of the gradient and against the oracle's factors on sampled rows."""
    import bluest_b200 as blu
    N = 20
    C = orc.wishart_cov(N, 0)
    groups = blu.enumerate_groups(N)
    L = sum(len(g) for g in groups)
    assert L == 2 ** N - 1
    sap = blu.SAP(C, N, groups, np.ones(L), verbose=False)
    m = orc.dense_m(L, 0)
    v, g, op = sap.variance_GH_operator(m)
    rng = np.random.RandomState(3)
    d = rng.randn(L); d /= np.linalg.norm(d)
    Hd = op @ d
    h = 1e-2
    _, gp, _ = sap.variance_GH(m + h * d, nohess=True)
    _, gm, _ = sap.variance_GH(m - h * d, nohess=True)
    assert maxrel((gp - gm) / (2 * h), Hd) < 1e-5
    assert d @ Hd > 0
    # factors against numpy on the device's own Phi: U rows of sampled groups, t = U^T d, V = 2 U pinv(Phi)
    phi = sap.get_phi(m)
    P = np.linalg.pinv(phi)
    x = P[0]
    Ud = sap.device_buffer(blu._lib.BUF_U).cpu().numpy()
    NP = 4 * ((N + 3) // 4)
    Ud = Ud[: (Ud.size // NP) * NP].reshape(-1, NP)[:L, :N]
    flat_rows = [0, 19, 20, 209, 210, 524287, 524288, 1048574]
    inv = sap.invcovs
    for i in flat_rows:
        k = int(np.searchsorted(sap.cumsizes, i, side="right"))
        j = i - int(sap.cumsizes[k - 1])
        grp = np.asarray(sap.groups[k - 1][j])
        Ci = inv[k - 1][j * k * k:(j + 1) * k * k].reshape(k, k)
        u = np.zeros(N); u[grp] = Ci @ x[grp]
        assert maxrel(Ud[i], u) < 1e-11
    t = Ud.T @ d
    assert maxrel(Hd[flat_rows], 2.0 * (Ud[flat_rows] @ (P @ t))) < 1e-11
    sap.close()
