"""Pins oracle/ (the CPU checker) against the real reference: the committed golden vectors
(tests/golden/, produced by tests/golden/make_golden.py from the reference Python) and, when
/root/reference is present, the live reference through oracle/ref_shim.py.  CPU only."""
import os

import numpy as np
import pytest

import oracle as orc
import ref_shim
from conftest import GOLDEN, maxrel

TOL = 1e-12     # north_star: FP64 outputs agree to relative error <= 1e-12


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


def _case(d, tag):
    K = int(d[f"{tag}/K"])
    groups = [d[f"{tag}/groups{k+1}"].tolist() for k in range(K)]
    invcovs = [d[f"{tag}/invcovs{k+1}"] for k in range(K)]
    return d[f"{tag}/C"], K, groups, invcovs


SYN_TAGS = ["N4K4", "N6K6", "N8K4", "N8K8"]


@pytest.mark.parametrize("tag", SYN_TAGS + ["ragged"])
def test_setup_matches_reference(tag):
    """a3/a4: per-group pinv, flat layout, psi."""
    d = _load("synthetic.npz")
    C, K, groups, invcovs = _case(d, tag)
    o = orc.SapOracle(C, K, groups)
    assert o.sizes == d[f"{tag}/sizes"].tolist()
    for k in range(K):
        assert o.invcovs[k].shape == invcovs[k].shape
        if invcovs[k].size:
            assert maxrel(o.invcovs[k], invcovs[k]) < TOL
    assert o.psi.shape == d[f"{tag}/psi"].shape
    assert maxrel(o.psi, d[f"{tag}/psi"]) < TOL
    assert np.array_equal(o.e, d[f"{tag}/e"])
    # with the reference's own inverses psi must be bit-identical (pure scatter)
    o2 = orc.SapOracle(C, K, groups, invcovs=invcovs)
    assert np.array_equal(o2.psi, d[f"{tag}/psi"])


@pytest.mark.parametrize("tag", ["N6K6", "N8K8"])
def test_batched_inverses_match_reference(tag):
    """The vectorised per-class inverse (one LAPACK batch per size class) that lets the oracle reach 20 models:
    pinned against the reference's own per-group ``np.linalg.pinv`` blocks (golden) and the array enumeration
    against the list one."""
    d = _load("synthetic.npz")
    C, K, groups, invcovs = _case(d, tag)
    ga = [np.array(gk, dtype=np.int64).reshape(len(gk), k + 1) for k, gk in enumerate(groups)]
    inv = orc.batched_invcovs(C, ga)
    for k in range(K):
        assert inv[k].shape == invcovs[k].shape
        assert maxrel(inv[k], invcovs[k]) < TOL
    N = C.shape[0]
    for a, l in zip(orc.enumerate_group_arrays(N, K), orc.enumerate_groups(N, K)):
        assert a.dtype == np.int64 and np.array_equal(a, np.array(l, dtype=np.int64))
    o = orc.SapOracle(C, K, ga, invcovs=inv, with_ES=False)
    for j in range(int(d[f"{tag}/n_m"])):
        m = d[f"{tag}/m{j}"]; delta = float(d[f"{tag}/delta{j}"])
        assert maxrel(o.get_phi(m, delta), d[f"{tag}/phi{j}"]) < TOL


@pytest.mark.parametrize("tag", SYN_TAGS)
def test_closures_match_reference(tag):
    """a5-a10 on dense / sparse / few-model / tiny / integer / threshold / delta / no-model-0 m."""
    d = _load("synthetic.npz")
    C, K, groups, invcovs = _case(d, tag)
    o = orc.SapOracle(C, K, groups, invcovs=invcovs)
    for j in range(int(d[f"{tag}/n_m"])):
        m = d[f"{tag}/m{j}"]; delta = float(d[f"{tag}/delta{j}"])
        assert maxrel(o.get_phi(m, delta), d[f"{tag}/phi{j}"]) < TOL
        assert maxrel(o.get_phi_dense(m, delta), d[f"{tag}/phi{j}"]) < TOL
        vref = float(d[f"{tag}/variance{j}"])
        if np.isnan(vref):
            with pytest.raises(AssertionError):
                o.variance(m, delta)
        elif np.isinf(vref):
            assert np.isinf(o.variance(m, delta))
        else:
            assert abs(o.variance(m, delta) - vref) <= TOL * abs(vref)
        res = o.variance_GH(m, delta)
        assert len(res) == int(d[f"{tag}/gh_len{j}"])          # 2-tuple early-out (misc.py:484)
        if len(res) == 2:
            assert np.isinf(res[0]) and np.all(np.isinf(res[1]))
            with pytest.raises(ValueError):
                o.cleanup_matrix(m, delta)
            continue
        mf = m.astype(float)
        full_rank = len(o.support(m)) == o.N and np.min(np.abs(mf[np.abs(mf) > 0])) > 1e-5
        tol = TOL if full_rank else 1e-9       # singular Phi: pinv junk at 1e-19 level differs (8a')
        assert abs(res[0] - float(d[f"{tag}/gh_var{j}"])) <= TOL * abs(float(d[f"{tag}/gh_var{j}"]))
        assert maxrel(res[1], d[f"{tag}/gh_grad{j}"]) < tol
        assert maxrel(res[2], d[f"{tag}/gh_hess{j}"]) < tol
        fac = o.variance_GH(m, delta, hess_mode="factored")
        assert maxrel(fac[2], d[f"{tag}/gh_hess{j}"]) < tol
        assert maxrel(o.cleanup_matrix(m, delta), d[f"{tag}/cleanup{j}"]) < tol


def test_ragged_phi_and_variance():
    d = _load("synthetic.npz")
    C, K, groups, invcovs = _case(d, "ragged")
    o = orc.SapOracle(C, K, groups, invcovs=invcovs)
    m = d["ragged/m0"]
    assert maxrel(o.get_phi(m), d["ragged/phi0"]) < TOL
    assert abs(o.variance(m) - float(d["ragged/variance0"])) <= TOL * abs(float(d["ragged/variance0"]))
    # the oracle (unlike the reference, which raises IndexError here) handles the empty class
    v, g, h = o.variance_GH(m)
    assert g.shape == (len(m),) and h.shape == (len(m), len(m))


def test_tutorial_known_answers():
    """tutorials/01_tutorial.ipynb:331-334 and :459-461: printed errors and exact costs."""
    d = _load("tutorial.npz")
    N = 5
    groups = orc.enumerate_groups(N)
    costs = orc.group_costs(groups, d["model_costs"])
    o = orc.SapOracle(d["C"], N, groups, costs)
    for a in range(2):
        m = d[f"m{a}"]
        v = o.variance(m)
        assert abs(v - float(d[f"ref_variance{a}"])) <= TOL * v
        assert abs(np.sqrt(v) - float(d[f"printed_error{a}"])) < 5e-9          # 8 printed digits
        assert m @ costs == float(d[f"printed_cost{a}"]) == float(d[f"ref_cost{a}"])
        vg, g, _ = o.variance_GH(m.astype(float), nohess=True)
        assert abs(vg - float(d[f"ref_gh_var{a}"])) <= 1e-10 * vg
        # singular Phi (sparse integer allocation): gradient entries of supported groups agree
        sup = np.abs(d[f"ref_grad{a}"]) > 1e-12 * np.abs(d[f"ref_grad{a}"]).max()
        assert maxrel(g[sup], d[f"ref_grad{a}"][sup]) < 1e-9


def test_hodgkin_huxley_allocation():
    """Stored K=7 allocation of the Hodgkin-Huxley example: per-output variances."""
    d = _load("hodgkin.npz")
    M, K, No = int(d["M"]), int(d["K"]), int(d["n_outputs"])
    groups = orc.enumerate_groups(M, K)
    samples = d["samples"]
    expect = [8.9059e-4, 1.00004e-3, 9.4788e-4, 5.4767e-4, 5.5831e-4]
    for n in range(No):
        C = d[f"C{n}"]
        o = orc.SapOracle(C, K, groups)
        v = o.variance(samples)
        # ill-conditioned (cond 1e9..5e10): the pinv of each group is only reproducible to cond*eps
        assert abs(v - float(d[f"variance{n}"])) <= 1e-6 * v
        assert abs(np.sqrt(v / C[0, 0]) - expect[n]) < 5e-8
        assert np.sqrt(v / C[0, 0]) <= 1e-3 * 1.0001
    # with the reference's inverses ingested the result is exact to rounding
    inv3 = [d[f"invcovs3_{k+1}"] for k in range(K)]
    o = orc.SapOracle(d["C3"], K, groups, invcovs=inv3)
    assert maxrel(o.get_phi(samples), d["phi3"]) < TOL
    idx = o.support(samples)
    cond = np.linalg.cond(d["phi3"][np.ix_(idx, idx)])     # the solve amplifies re-association by cond(Phi)
    assert abs(o.variance(samples) - float(d["variance3"])) <= max(TOL, 10 * cond * np.finfo(float).eps) * float(d["variance3"])
    flat_costs = orc.group_costs(groups, d["costs"])
    assert abs(samples @ flat_costs - float(d["total_cost"])) < 1e-6


def test_matern_ill_conditioned():
    d = _load("matern.npz")
    C, K, groups, invcovs = _case(d, "matern")
    o = orc.SapOracle(C, K, groups, invcovs=invcovs)
    m = d["matern/m0"]
    assert maxrel(o.get_phi(m), d["matern/phi0"]) < TOL
    v, g, h = o.variance_GH(m)
    cond_phi = np.linalg.cond(d["matern/phi0"])
    tol = max(TOL, 50 * cond_phi * np.finfo(float).eps)
    assert abs(v - float(d["matern/gh_var0"])) <= tol * abs(v)
    assert maxrel(g, d["matern/gh_grad0"]) < tol
    # The Hessian is NOT reproducible at this conditioning (cond(Phi)=3e8, cond(C_k) up to 1.5e9):
    # the reference's six-deep loop, the same loop without -ffast-math and the factored form
    # 2 U^T Phi^+ U disagree with each other by O(1) -- entries (4e-9) are pure cancellation noise.
    assert h.shape == d["matern/gh_hess0"].shape


@pytest.mark.parametrize("tag,N,K", [("complete_N4_K4", 4, 4), ("complete_N6_K3", 6, 3)])
def test_enumeration_complete_graph_bit_exact(tag, N, K):
    """a1: networkx clique enumeration on a complete graph == size-major lexicographic subsets."""
    d = _load("enumeration.npz")
    groups = orc.enumerate_groups(N, K)
    assert int(d[f"{tag}/K"]) == K
    for k in range(K):
        assert np.array_equal(np.array(groups[k], dtype=np.int64), d[f"{tag}/groups{k+1}"])
    ES = orc.indicator_ES(groups, N)
    assert np.array_equal(np.array(ES), d[f"{tag}/ES"])
    maps = orc.mosap_mappings(groups, [groups])
    assert np.array_equal(maps[0], d[f"{tag}/mapping0"])


def test_enumeration_two_outputs_union_and_mappings():
    """a1/a2: union over outputs with different coupling graphs, sort, mappings, ES."""
    d = _load("enumeration.npz")
    tag = "two_outputs"
    K = int(d[f"{tag}/K"]); Ks = d[f"{tag}/Ks"].tolist()
    multi = [[d[f"{tag}/multi{n}_groups{k+1}"].tolist() for k in range(Ks[n])] for n in range(2)]
    groups = orc.union_groups(multi)
    assert len(groups) == K
    for k in range(K):
        assert np.array_equal(np.array(groups[k], dtype=np.int64).reshape(-1, k + 1), d[f"{tag}/groups{k+1}"])
    maps = orc.mosap_mappings(groups, multi)
    for n in range(2):
        assert np.array_equal(maps[n], d[f"{tag}/mapping{n}"])
    assert np.array_equal(np.array(orc.indicator_ES(groups, 6)), d[f"{tag}/ES"])
    # MOSAP.variances / variance_GH on m[mappings[n]] (a11)
    m = d[f"{tag}/m"]
    for n in range(2):
        o = orc.SapOracle(d[f"{tag}/Cwish{n}"], Ks[n], multi[n])
        mn = m[maps[n]]
        assert abs(o.variance(mn) - d[f"{tag}/variances"][n]) <= TOL * d[f"{tag}/variances"][n]
        assert maxrel(o.variance_GH(mn, nohess=True)[1], d[f"{tag}/grad{n}"]) < TOL


def test_pilot_covariance():
    """a12: one-pass biased covariance from the reference's accumulation loop."""
    d = _load("pilot.npz")
    s1, S2, C = orc.pilot_covariance(d["Y"])
    assert maxrel(s1, d["sumse"]) < TOL
    assert maxrel(S2, d["sumsc"]) < TOL
    assert maxrel(C, d["C_hat"]) < 1e-11        # cancellation in S2/n - s1 s1^T/n^2


def test_c_loops_match_reference_binary():
    """The compiled reference cmisc.cpp (oracle/_ref, -ffast-math) vs the plain-C restatement."""
    if ref_shim._ref_so() is None:
        pytest.skip("oracle/_ref not built")
    import importlib, sys
    sys.path.insert(0, os.path.dirname(ref_shim._ref_so()))
    cm = importlib.import_module("_cmisc_bluest")
    N, k, q = 7, 3, 4
    rng = np.random.RandomState(0)
    C = orc.wishart_cov(N, 5)
    groups = orc.enumerate_groups(N, 4)
    o = orc.SapOracle(C, 4, groups)
    gk, gq = o.groups[k - 1], o.groups[q - 1]
    ck, cq = o.invcovs[k - 1], o.invcovs[q - 1]
    Lk, Lq = len(gk), len(gq)
    P = np.linalg.pinv(o.get_phi(orc.dense_m(o.L)))
    L = orc.lib()
    psi_a = np.zeros(N * N * Lk); psi_b = np.zeros((N * N, Lk))
    cm.assemble_psi_c(psi_a, N, k, Lk, gk.ravel(), ck)
    L.orc_psi_class(orc._d(psi_b), N, k, Lk, orc._l(gk), orc._d(ck))
    assert np.array_equal(psi_a.reshape(N * N, Lk), psi_b)
    mk = 1 + rng.rand(Lk)
    pa = np.zeros(N * N); pb = np.zeros(N * N)
    cm.objectiveK_c(pa, N, k, Lk, mk, gk.ravel(), ck)
    L.orc_phi_class(orc._d(pb), N, k, Lk, orc._d(mk), orc._l(gk), orc._d(ck))
    assert maxrel(pb, pa) < 1e-14
    ga = np.zeros(Lk); gb = np.zeros(Lk); x = np.ascontiguousarray(P[0])
    cm.gradK_c(ga, k, Lk, gk.ravel(), ck, x)
    L.orc_grad_class(orc._d(gb), k, Lk, orc._l(gk), orc._d(ck), orc._d(x))
    assert maxrel(gb, ga) < 1e-14
    ha = np.zeros(Lk * Lq); hb = np.zeros((Lk, Lq))
    cm.hessKQ_c(ha, N, k, q, Lk, Lq, gk.ravel(), gq.ravel(), ck, cq, np.ascontiguousarray(P).ravel())
    L.orc_hess_block(orc._d(hb), N, k, q, Lk, Lq, orc._l(gk), orc._l(gq), orc._d(ck), orc._d(cq), orc._d(np.ascontiguousarray(P).ravel()))
    assert maxrel(hb.ravel(), ha) < 1e-13
    xa = np.zeros(N * Lk); xb = np.zeros((N, Lk))
    cm.cleanupK_c(xa, k, Lk, gk.ravel(), ck, x)
    L.orc_cleanup_class(orc._d(xb), k, Lk, orc._l(gk), orc._d(ck), orc._d(x))
    assert np.array_equal(xa.reshape(N, Lk), xb)


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")
def test_live_reference_random_cases():
    """Oracle vs the reference Python itself on fresh random inputs (build container only)."""
    ns = ref_shim.load()
    for N, K, seed in [(5, 5, 11), (7, 3, 12), (9, 9, 13)]:
        C = orc.wishart_cov(N, seed)
        groups = orc.enumerate_groups(N, K)
        L = sum(len(g) for g in groups)
        sap = ns.sap.SAP(C.copy(), K, [[list(g) for g in gk] for gk in groups], np.ones(L), verbose=False)
        o = orc.SapOracle(C, K, groups)
        for m in (orc.dense_m(L, seed), orc.sparse_m(L, N, seed)):
            assert maxrel(o.get_phi(m), sap.get_phi(m)) < TOL
            assert abs(o.variance(m) - sap.variance(m)) <= TOL * abs(sap.variance(m))
            a = o.variance_GH(m); b = sap.variance_GH(m)
            full = len(o.support(m)) == N
            tol = TOL if full else 1e-9
            assert abs(a[0] - b[0]) <= TOL * abs(b[0])
            assert maxrel(a[1], b[1]) < tol and maxrel(a[2], b[2]) < tol


@pytest.mark.parametrize("tag,N", [("N6K6", 6), ("N9K4", 9)])
def test_blue_estimator_matches_reference(tag, N):
    """"next" row f3: compute_BLUE_estimator (sap.py:99-119) + PHIinvY0 (misc.py:518-544)."""
    d = _load("estimator.npz")
    K = int(d[f"{tag}/K"])
    groups = orc.enumerate_groups(N, K)
    inv = [d[f"{tag}/invcovs{k+1}"] for k in range(K)]
    o = orc.SapOracle(d[f"{tag}/C"], K, groups, invcovs=inv)
    flat = [g for gk in groups for g in gk]
    sums, pos = [], 0
    for g in flat:
        sums.append(d[f"{tag}/sums"][pos:pos + len(g)]); pos += len(g)
    mu, var, _ = orc.blue_estimator(o, sums, d[f"{tag}/samples"])
    assert abs(mu - float(d[f"{tag}/mu"])) <= 1e-12 * abs(mu)
    assert abs(var - float(d[f"{tag}/var"])) <= 1e-12 * var
