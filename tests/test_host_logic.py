"""Host-side integer logic of the package (group enumeration, unions, mappings, indicator vectors,
slice balancing) against the golden vectors of the reference -- bit-exact, CPU only."""
import os

import numpy as np
import pytest

import bluest_b200 as blu
import oracle as orc
from conftest import GOLDEN, maxrel


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.mark.parametrize("tag,N,K", [("complete_N4_K4", 4, 4), ("complete_N6_K3", 6, 3)])
def test_complete_graph_enumeration(tag, N, K):
    d = _load("enumeration.npz")
    groups = blu.enumerate_groups(N, K)
    cl = blu.enumerate_cliques(np.ones((N, N)), K)
    for k in range(K):
        assert np.array_equal(np.array(groups[k], dtype=np.int64), d[f"{tag}/groups{k+1}"])
        assert np.array_equal(np.array(cl[k], dtype=np.int64), d[f"{tag}/groups{k+1}"])
    assert np.array_equal(np.array(blu.indicator_ES([np.array(g) for g in groups], N)), d[f"{tag}/ES"])
    assert np.array_equal(blu.mappings(groups, [groups])[0], d[f"{tag}/mapping0"])


def test_two_output_cliques_union_mappings():
    d = _load("enumeration.npz")
    tag = "two_outputs"
    K = int(d[f"{tag}/K"]); Ks = d[f"{tag}/Ks"].tolist()
    multi = []
    for n in range(2):
        mine = blu.enumerate_cliques(d[f"{tag}/adj{n}"], 4)
        assert len(mine) == Ks[n]
        for k in range(Ks[n]):
            assert np.array_equal(np.array(mine[k], dtype=np.int64), d[f"{tag}/multi{n}_groups{k+1}"])
        multi.append(mine)
    groups = blu.union_groups(multi)
    assert len(groups) == K
    for k in range(K):
        assert np.array_equal(np.array(groups[k], dtype=np.int64).reshape(-1, k + 1), d[f"{tag}/groups{k+1}"])
    maps = blu.mappings(groups, multi)
    for n in range(2):
        assert np.array_equal(maps[n], d[f"{tag}/mapping{n}"])
    assert np.array_equal(np.array(blu.indicator_ES([np.array(g) for g in groups], 6)), d[f"{tag}/ES"])
    with pytest.raises(AssertionError):
        blu.mappings(groups, [[[[7]]]])             # a group that is not in the union (mosap.py:60)


def test_disconnected_component_is_filtered():
    """blue_models.py:468: cliques outside model 0's connected component are dropped."""
    A = np.zeros((5, 5)); A[:3, :3] = 1; A[3:, 3:] = 1
    cl = blu.enumerate_cliques(A, 5)
    flat = [tuple(g) for gk in cl for g in gk]
    assert flat == [(0,), (1,), (2,), (0, 1), (0, 2), (1, 2), (0, 1, 2)]


def test_group_costs_and_matches_oracle():
    mc = 2.0 ** (6 - np.arange(6))
    groups = blu.enumerate_groups(6, 3)
    assert np.array_equal(blu.group_costs(groups, mc), orc.group_costs(groups, mc))
    assert blu.enumerate_groups(7) == orc.enumerate_groups(7)


@pytest.mark.parametrize("N,world", [(10, 2), (15, 8), (12, 3)])
def test_balanced_slices(N, world):
    from math import comb
    sizes = [comb(N, k) for k in range(1, N + 1)]
    sl = blu.balanced_slices(sizes, world)
    assert sl[0][0] == 0 and sl[-1][1] == sum(sizes)
    assert all(sl[i][1] == sl[i + 1][0] for i in range(world - 1))
    work = np.concatenate([np.full(s, (k + 1) ** 2) for k, s in enumerate(sizes)])
    per = [work[lo:hi].sum() for lo, hi in sl]
    assert max(per) / (sum(per) / world) < 1.02          # within 2% of perfect balance


def test_pinned_pool_is_lazy_on_cpu():
    from bluest_b200 import _lib
    assert _lib.pinned_pool.free_blocks == {} or isinstance(_lib.pinned_pool.free_blocks, dict)


@pytest.mark.parametrize("tag,N", [("N6K6", 6), ("N8K3", 8)])
def test_integer_projection_host_logic(tag, N, monkeypatch):
    """Row f2: candidate enumeration / filtering / tie-breaking against the reference's result, with the
    device call replaced by a numpy stand-in built on the oracle (CPU-only check of the host logic)."""
    from bluest_b200 import intproj
    d = _load("intproj.npz")
    K = int(d[f"{tag}/K"])
    groups = orc.enumerate_groups(N, K)
    o = orc.SapOracle(d[f"{tag}/C"], K, groups, invcovs=[d[f"{tag}/invcovs{k+1}"] for k in range(K)])
    sol = d[f"{tag}/sol"]
    lb, ub, idx = intproj.feasible_integer_bounds(sol, N, e=o.e)
    assert np.array_equal(lb, d[f"{tag}/lb"]) and np.array_equal(ub, d[f"{tag}/ub"]) and np.array_equal(idx, d[f"{tag}/idx"])

    class FakeSap:
        pass
    sap = FakeSap()
    sap.N, sap.costs, sap.e, sap.L = N, d[f"{tag}/w"], o.e, o.L
    sap.get_max_sample_constraints = lambda mm: ([], [])

    def fake_candidates(sap_, base, idx_, ms, rcond=1e-10):
        phis = np.stack([o.get_phi(base) + sum(ms[t, c] * o.psi[:, idx_[t]].reshape(N, N) for t in range(len(idx_))) for c in range(ms.shape[1])])
        return np.linalg.pinv(phis, hermitian=True, rcond=rcond)[:, 0, 0]
    monkeypatch.setattr(intproj, "candidate_variances", fake_candidates)
    val, fval = intproj.best_closest_integer_solution_BLUE(sap, sol, budget=float(d[f"{tag}/budget"]))
    assert np.array_equal(val, d[f"{tag}/budget_val"]) and abs(fval - float(d[f"{tag}/budget_fval"])) <= 1e-10 * fval
    val, fval = intproj.best_closest_integer_solution_BLUE(sap, sol, eps=float(d[f"{tag}/eps"]))
    assert np.array_equal(val, d[f"{tag}/eps_val"]) and abs(fval - float(d[f"{tag}/eps_fval"])) <= 1e-10 * fval
    assert np.array_equal(intproj.integer_projection(sap, sol, budget=float(d[f"{tag}/budget"])), d[f"{tag}/projection_budget"])


def test_graph_data_file_roundtrip(tmp_path):
    """Row f4: the reference's .npz graph format (blue_models.py:265-299) and the covariance view of
    get_covariance (blue_models.py:166-179)."""
    from bluest_b200 import io
    g = io.load_graph_data(os.path.join(GOLDEN, "hh_graph_data.npz"))
    d = _load("hodgkin.npz")
    assert g["M"] == 12 and g["n_outputs"] == 5 and len(g["C"]) == 5
    for n in range(5):
        assert np.array_equal(g["C"][n], d[f"C{n}"])           # complete graph: no NaN / inf to translate
    assert np.array_equal(g["costs"], d["costs"])
    A = g["adjacency"][0].copy(); A[0, 5] = A[5, 0] = 0.0; A[1, 2] = A[2, 1] = np.inf
    io.save_graph_data(str(tmp_path / "g.npz"), g["costs"], [A], g["SG"][:1])
    h = io.load_graph_data(str(tmp_path / "g.npz"))
    assert np.isnan(h["C"][0][0, 5]) and h["C"][0][1, 2] == 0.0 and h["M"] == 12
    with pytest.raises(ValueError):
        io.load_graph_data(str(tmp_path / "g.npz"), n_outputs=3)


class _OracleProblem:
    """Oracle-backed stand-in with the attribute surface of bluest_b200.SAP that the shared scipy driver
    needs (the GPU tests run the same driver on the real SAP)."""

    def __init__(self, N, K, seed=0):
        self.o = orc.SapOracle(orc.wishart_cov(N, seed), K, orc.enumerate_groups(N, K))
        self.L, self.N, self.e = self.o.L, N, self.o.e
        from bluest_b200.groups import enumerate_groups, group_costs
        self.costs = group_costs(enumerate_groups(N, K), 2.0 ** (N - np.arange(N)))

    def variance(self, m, delta=0):
        return self.o.variance(m, delta)

    def variance_GH(self, m, delta=0, nohess=False):
        return self.o.variance_GH(m, delta, nohess=nohess, hess_mode="factored")

    def variance_GH_operator(self, m, delta=0):
        from scipy.sparse.linalg import LinearOperator
        v, g, _ = self.o.variance_GH(m, delta, nohess=True)
        P = np.linalg.pinv(self.o.get_phi(m, delta))
        U = self.o.ufactor(np.ascontiguousarray(P[0]))
        mv = lambda p: 2.0 * (U.T @ (P @ (U @ p)))
        return v, g, LinearOperator((self.L, self.L), matvec=mv, rmatvec=mv, dtype=np.float64)

    def get_max_sample_constraints(self, mm):
        return [], []


@pytest.mark.parametrize("mode", ["budget", "eps"])
def test_scipy_driver_operator_and_sparse_constraints_walk_the_same_iterates(mode):
    """solvers.scipy_solve: the reference's dense driver (sap.py:387-418), the Hessian-operator driver
    and the sparse-constraint driver must reach the same allocation from the same feasible x0."""
    from bluest_b200.solvers import scipy_solve
    p = _OracleProblem(6, 4)
    x0 = np.ceil(10 * abs(np.random.RandomState(0).randn(p.L)))
    if mode == "budget":
        kw = dict(budget=float(x0 @ p.costs) / 0.9)
    else:
        kw = dict(eps=float(np.sqrt(p.variance(x0))))
    ref = scipy_solve(p, x0=x0.copy(), **kw)
    for opts in (dict(hess="operator"), dict(sparse_constraints=True), dict(hess="operator", sparse_constraints=True)):
        cnt = {}
        r = scipy_solve(p, x0=x0.copy(), counters=cnt, **opts, **kw)
        assert r.status == ref.status
        if mode == "budget":
            assert abs(r.fun - ref.fun) <= 1e-4 * abs(ref.fun)          # gtol stop on a flat objective
            assert abs(r.x @ p.costs - ref.x @ p.costs) <= 1e-4 * kw["budget"]
        else:
            assert abs(p.variance(r.x) - kw["eps"] ** 2) <= 1e-6 * kw["eps"] ** 2
            assert abs(r.x @ p.costs - ref.x @ p.costs) <= 1e-3 * (ref.x @ p.costs)
        assert cnt["H"] > 0
    with pytest.raises(ValueError):
        scipy_solve(p, x0=x0, hess="banded", **kw)


# ---------------------------------------------------------------------------------------------
# multi-output host orchestration (mosap.py:125-331, misc.py:177-311) with oracle-backed outputs
# ---------------------------------------------------------------------------------------------
class _OracleOutput:
    """Stand-in for one output's device context: the oracle behind the few SAP members MOSAP touches."""

    def __init__(self, C, K, groups):
        self.o = orc.SapOracle(C, K, groups)
        self.N, self.L, self.e = self.o.N, self.o.L, self.o.e
        self.samples = None

    def get_phi(self, m, delta=0):
        return self.o.get_phi(m, delta)

    def variance(self, m, delta=0):
        return self.o.variance(m, delta)

    def get_cleanup_matrix(self, m, delta=0):
        return self.o.cleanup_matrix(m, delta)


def _oracle_mosap(d, tag):
    """A real bluest_b200.MOSAP object (host logic under test) whose outputs are oracle stand-ins."""
    from bluest_b200.groups import indicator_ES, mappings as build_mappings
    K = int(d[f"{tag}/K"]); Ks = d[f"{tag}/Ks"].tolist(); No = int(d[f"{tag}/n_outputs"])
    groups = [d[f"{tag}/groups{k+1}"].tolist() for k in range(K)]
    multi = [[d[f"{tag}/multi{n}_groups{k+1}"].tolist() for k in range(Ks[n])] for n in range(No)]
    mos = object.__new__(blu.MOSAP)
    mos.verbose = False
    mos.n_outputs = No
    mos.C = [d[f"{tag}/C{n}"] for n in range(No)]
    mos.N = mos.C[0].shape[0]
    mos.K, mos.Ks, mos.costs = K, Ks, d[f"{tag}/w"]
    mos.groups = [np.array(g, dtype=np.int64) for g in groups]
    mos.SAPS = [_OracleOutput(mos.C[n], Ks[n], multi[n]) for n in range(No)]
    mos.sizes = [0] + [len(g) for g in groups]
    mos.cumsizes = np.cumsum(mos.sizes)
    mos.L = mos.cumsizes[-1]
    mos.ES = indicator_ES(groups, mos.N)
    mos.e = mos.ES[0]
    mos.mappings = build_mappings(groups, multi)
    mos.variances = lambda m, delta=0: [mos.SAPS[n].variance(np.asarray(m)[mos.mappings[n]], delta) for n in range(No)]
    return mos


def _fake_candidates(sap_, base, idx_, ms, rcond=1e-10):
    o = sap_.o
    N = o.N
    phis = (o.get_phi(base).reshape(-1, 1) + o.psi[:, idx_] @ ms).T.reshape(-1, N, N)
    return np.linalg.pinv(phis, hermitian=True, rcond=rcond)[:, 0, 0]


@pytest.mark.parametrize("tag,name", [("two_outputs", "small"), ("shared_N8K3", "big")])
def test_multi_output_integer_projection_host_logic(tag, name, monkeypatch):
    """misc.py:177-311 / mosap.py:213-292: brute-force (<= 15 groups) and randomised (> 15, seeded)
    projection, the fallback ladder and the max-sample caps -- allocations equal to the reference's."""
    from bluest_b200 import intproj
    d = _load("mosap.npz")
    mos = _oracle_mosap(d, tag)
    monkeypatch.setattr(intproj, "candidate_variances", _fake_candidates)
    sol = d[f"{tag}/{name}_sol"]
    lb, ub, idx = intproj.feasible_integer_bounds(sol, mos.N, e=mos.e)
    assert np.array_equal(idx, d[f"{tag}/{name}_idx"])
    assert (len(idx) > intproj.LL_MAX_MULTI) == (name == "big")
    budget = float(d[f"{tag}/{name}_budget"]); eps = d[f"{tag}/{name}_eps"]
    np.random.seed(1234)
    val, fval = intproj.best_closest_integer_solution_BLUE_multi(mos, sol.copy(), budget=budget)
    assert np.array_equal(val, d[f"{tag}/{name}_budget_val"]) and abs(fval - float(d[f"{tag}/{name}_budget_fval"])) <= 1e-10 * fval
    np.random.seed(1234)
    val, fval = intproj.best_closest_integer_solution_BLUE_multi(mos, sol.copy(), eps=eps)
    assert np.array_equal(val, d[f"{tag}/{name}_eps_val"]) and abs(fval - float(d[f"{tag}/{name}_eps_fval"])) <= 1e-10 * fval
    np.random.seed(1234)
    assert np.array_equal(mos.integer_projection(sol.copy(), budget=budget), d[f"{tag}/{name}_projection_budget"])
    np.random.seed(1234)
    assert np.array_equal(mos.integer_projection(sol.copy(), eps=eps), d[f"{tag}/{name}_projection_eps"])
    np.random.seed(1234)
    assert np.array_equal(mos.integer_projection(sol.copy(), budget=budget, max_model_samples=d[f"{tag}/{name}_caps"]), d[f"{tag}/{name}_projection_caps"])
    with pytest.raises(ValueError):
        mos.integer_projection(sol)
    with pytest.raises(ValueError):
        mos.get_max_sample_constraints(np.ones(mos.N + 1))
    with pytest.raises(ValueError):
        mos.get_max_sample_constraints(np.zeros(mos.N))


def test_multi_output_cleanup_solution_host_logic():
    """mosap.py:125-211: the sparsified allocation equals the reference's, its cost is not higher and
    the largest output variance not worse."""
    d = _load("mosap.npz")
    tag = "two_outputs"
    mos = _oracle_mosap(d, tag)
    m = d[f"{tag}/cleanup_m"]
    X = mos.get_cleanup_matrices(m.copy())
    assert np.max(np.abs(X - d[f"{tag}/cleanup_X"])) <= 1e-12 * np.max(np.abs(d[f"{tag}/cleanup_X"]))
    out = mos.cleanup_solution(m.copy())
    ref = d[f"{tag}/cleanup_result"]
    assert np.array_equal(out > 0, ref > 0)
    assert np.max(np.abs(out - ref)) <= 1e-8 * np.max(np.abs(ref))
    assert (out > 0).sum() < (m > 0).sum()
    assert out @ mos.costs <= m @ mos.costs * (1 + 1e-12)
    assert max(mos.variances(out)) <= max(d[f"{tag}/cleanup_variances_before"]) * (1 + 1e-4)


def test_group_costs_and_mappings_vectorised_equal_the_reference_forms():
    """groups.group_costs must be bit-identical to ``sum(model_costs[group])`` (blue_models.py:137-140)
    and groups.mappings to the reference's linear scan (mosap.py:54-67), also for ragged per-output
    group lists, integer costs and missing groups."""
    from bluest_b200.groups import enumerate_groups, group_costs, mappings
    rng = np.random.RandomState(0)
    N = 11
    groups = enumerate_groups(N, 9)
    mc = rng.rand(N) * 10
    ref = np.array([sum(mc[list(g)]) for gk in groups for g in gk])
    assert np.array_equal(group_costs(groups, mc), ref)
    mci = rng.randint(1, 100, size=N)
    refi = np.array([sum(mci[list(g)]) for gk in groups for g in gk])
    got = group_costs(groups, mci)
    assert np.array_equal(got, refi) and got.dtype.kind == "i"
    assert group_costs([[], [[0, 1]]], mc)[0] == mc[0] + mc[1]
    # ragged outputs: random subsets of every size class, shuffled order inside a class
    flat = [g for gk in groups for g in gk]
    multi = []
    for n in range(3):
        mg = []
        for gk in groups:
            take = [gk[i] for i in rng.permutation(len(gk))[: max(1, len(gk) // (n + 2))]]
            mg.append(take)
        multi.append(mg)
    maps = mappings(groups, multi)
    for n in range(3):
        want = np.array([flat.index(g) for gk in multi[n] for g in gk])
        assert np.array_equal(maps[n], want) and maps[n].dtype == np.int64
    with pytest.raises(AssertionError):
        mappings(enumerate_groups(N, 2), [[[[0]], [[0, 1]], [[0, 1, 2]]]])
    with pytest.raises(AssertionError):
        mappings([[[0], [1]], [[0, 1]]], [[[[2]]]])


@pytest.mark.parametrize("mode", ["budget", "eps"])
def test_multi_output_scipy_driver_variants(mode):
    """solvers.scipy_solve_multi on oracle-backed outputs: the dense driver of mosap.py:555-610, the
    Hessian-operator driver and sparse constraint rows reach the same optimum; the reference's
    eps-bound quirk (mosap.py:605) is reproduced by default and can be switched off."""
    from bluest_b200.solvers import scipy_solve_multi
    d = _load("mosap.npz")
    tag = "solve_N5K3"
    mos = _oracle_mosap(d, tag)
    for s_ in mos.SAPS:
        s_.variance_GH = (lambda o: (lambda m, delta=0, nohess=False: o.variance_GH(m, delta, nohess=nohess, hess_mode="factored")))(s_.o)

        def make_op(o):
            def variance_GH_operator(m, delta=0):
                from scipy.sparse.linalg import LinearOperator
                v, g, _ = o.variance_GH(m, delta, nohess=True)
                P = np.linalg.pinv(o.get_phi(m, delta)); U = o.ufactor(np.ascontiguousarray(P[0]))
                mv = lambda p: 2.0 * (U.T @ (P @ (U @ p)))
                return v, g, LinearOperator((o.L, o.L), matvec=mv, rmatvec=mv, dtype=np.float64)
            return variance_GH_operator
        s_.variance_GH_operator = make_op(s_.o)
    w = d[f"{tag}/w"]; x0 = d[f"{tag}/x0"]
    if mode == "budget":
        kw = dict(budget=float(d[f"{tag}/budget"]))
        ref_x = d[f"{tag}/continuous_budget"]
    else:
        kw = dict(eps=d[f"{tag}/eps"])
        ref_x = d[f"{tag}/continuous_eps"]
    res = scipy_solve_multi(mos, x0=x0.copy(), **kw)
    # the driver on the oracle's closures against the reference's own run (golden): same iterates up to solver tolerance
    assert np.max(np.abs(res.x - ref_x)) <= 2e-2 * np.max(np.abs(ref_x))
    worst = lambda x: max(mos.variances(x))
    for opts in (dict(hess="operator"), dict(hess="operator", sparse_constraints=True)):
        r2 = scipy_solve_multi(mos, x0=x0.copy(), **opts, **kw)
        if mode == "budget":
            assert abs(worst(r2.x) - worst(res.x)) <= 1e-4 * worst(res.x)
        else:
            assert abs(r2.x @ w - res.x @ w) <= 1e-3 * (res.x @ w)
    if mode == "eps":
        fixed = scipy_solve_multi(mos, x0=x0.copy(), reference_eps_bound=False, **kw)
        assert np.all(np.array(mos.variances(fixed.x)) <= kw["eps"] ** 2 * (1 + 1e-6))
        assert np.all(np.array(mos.variances(res.x)) <= kw["eps"][-1] ** 2 * (1 + 1e-6))
    with pytest.raises(ValueError):
        scipy_solve_multi(mos, x0=x0.copy(), hess="sparse", **kw)


def test_mosap_constructor_and_solve_end_to_end_on_oracle_contexts(monkeypatch):
    """The real ``bluest_b200.MOSAP`` class -- constructor, ``check_input``, ``solve(solver="scipy")`` with its
    integer projection -- with the per-output device contexts replaced by oracle-backed stand-ins: the integer
    allocation equals the one the REFERENCE's MOSAP produced from the same x0 (golden, mosap.npz)."""
    import bluest_b200.mosap as mosap_mod
    from bluest_b200 import intproj

    class StubSAP(_OracleOutput):
        def __init__(self, C, K, groups, costs, verbose=True, device=0):
            super().__init__(C, K, [[list(g) for g in gk] for gk in groups])

        def variance_GH(self, m, delta=0, nohess=False):
            return self.o.variance_GH(m, delta, nohess=nohess, hess_mode="factored")

    monkeypatch.setattr(mosap_mod, "SAP", StubSAP)
    monkeypatch.setattr(intproj, "candidate_variances", _fake_candidates)
    d = _load("mosap.npz")
    tag = "solve_N5K3"
    K = int(d[f"{tag}/K"]); Ks = d[f"{tag}/Ks"].tolist(); No = int(d[f"{tag}/n_outputs"])
    groups = [d[f"{tag}/groups{k+1}"].tolist() for k in range(K)]
    multi = [[d[f"{tag}/multi{n}_groups{k+1}"].tolist() for k in range(Ks[n])] for n in range(No)]
    w = d[f"{tag}/w"]
    mos = blu.MOSAP([d[f"{tag}/C{n}"] for n in range(No)], K, Ks, groups, multi, w, [w] * No, verbose=False)
    monkeypatch.setattr(mos, "variances", lambda m, delta=0: [mos.SAPS[n].variance(np.asarray(m)[mos.mappings[n]], delta) for n in range(No)], raising=False)
    assert isinstance(groups[0], np.ndarray) and groups[0].dtype == np.int64           # converted in place, mosap.py:34
    assert mos.L == 25 and mos.flattened_groups[:3] == [[0], [1], [2]] and mos.samples is None
    b, e = mos.check_input(None, 0.1)
    assert b is None and np.array_equal(e, [0.1] * No)
    with pytest.raises(ValueError):
        mos.check_input(None, [0.1, 0.2])
    with pytest.raises(ValueError):
        mos.check_input(None, None)
    np.random.seed(1234)
    ints = mos.solve(budget=float(d[f"{tag}/budget"]), x0=d[f"{tag}/x0"].copy(), continuous_relaxation=False)
    assert np.array_equal(ints, d[f"{tag}/integer_budget"])
    assert mos.tot_cost == ints @ w and mos.budget == float(d[f"{tag}/budget"])
    for n in range(No):
        assert np.array_equal(mos.SAPS[n].samples, ints[mos.mappings[n]])


@pytest.mark.parametrize("tag", ["default_K3", "default_K7", "user_groups", "user_multi"])
def test_setup_solver_group_bookkeeping_equals_the_reference(tag):
    """Row a1: io.prepare_groups against what the REAL ``BLUEProblem.setup_solver`` hands to the MOSAP
    constructor (blue_models.py:453-509; golden captured there): default clique enumeration on non-complete
    graphs with a cut-off model, user ``groups`` (unsorted members, non-cliques, groups outside model 0's
    component -- dropped; size classes left empty) and per-output ``multi_groups``.  Integer work: bit-exact."""
    from bluest_b200 import io
    d = _load("setup.npz")
    graph = io.load_graph_data(os.path.join(GOLDEN, "setup_graph_data.npz"))
    assert graph["M"] == 7 and graph["n_outputs"] == 2
    assert np.isnan(graph["C"][0][1, 4]) and graph["C"][0][2, 5] == 0.0                 # get_covariance view: 0 -> NaN, inf -> 0
    kw = {"default_K3": dict(K=3), "default_K7": dict(K=7),
          "user_groups": dict(K=3, groups=[[3, 0], [0], [2, 1, 0], [1, 4], [4, 1, 0], [6], [0, 6], [5, 2, 0], [3], [0, 3, 2, 1], [2]]),
          "user_multi": dict(K=3, multi_groups=[[[0], [1, 0], [2, 0, 1], [4, 3]], [[0], [0, 3], [2, 0], [6, 0], [1, 2], [5]]])}[tag]
    K, Ks, groups, multi = io.prepare_groups(graph, **kw)
    assert K == int(d[f"{tag}/K"]) and list(Ks) == d[f"{tag}/Ks"].tolist()
    assert len(groups) == K
    for k in range(K):
        assert np.array_equal(np.array(groups[k], dtype=np.int64).reshape(-1, k + 1), d[f"{tag}/groups{k+1}"])
    for n in range(2):
        assert len(multi[n]) == int(d[f"{tag}/n_classes{n}"])
        for k in range(len(multi[n])):
            assert np.array_equal(np.array(multi[n][k], dtype=np.int64).reshape(-1, k + 1), d[f"{tag}/multi{n}_groups{k+1}"])
        assert np.array_equal(blu.group_costs(multi[n], graph["costs"]), d[f"{tag}/multi_costs{n}"])
    assert np.array_equal(blu.group_costs(groups, graph["costs"]), d[f"{tag}/costs"])
    if tag == "user_groups":
        assert kw["groups"][0] == [0, 3]                                               # sorted in place, like the reference
    assert io.is_subclique(graph["adjacency"][0], [0, 2, 5]) and not io.is_subclique(graph["adjacency"][0], [0, 1, 4])
    with pytest.raises(ValueError):
        io.prepare_groups(graph, multi_groups=[[[0]]])


def test_setup_solver_front_end_on_oracle_contexts(monkeypatch):
    """io.setup_solver (blue_models.py:448-538) end to end with user-supplied groups on a non-complete two-output
    graph; the device contexts are oracle-backed stand-ins that speak the split begin/end evaluation protocol,
    so the REAL MOSAP.variances / solve / integer projection run.  Checks the reference's ``blue_data`` contract."""
    import bluest_b200.mosap as mosap_mod
    from bluest_b200 import _lib, intproj, io

    class StubSAP(_OracleOutput):
        def __init__(self, C, K, groups, costs, verbose=True, device=0):
            super().__init__(C, K, [[list(g) for g in gk] for gk in groups[:K]])

        def variance_GH(self, m, delta=0, nohess=False):
            return self.o.variance_GH(m, delta, nohess=nohess, hess_mode="factored")

        def variance_GH_begin(self, m, delta=0, nohess=False, grad=True):
            self._pending = (np.asarray(m, dtype=float), delta, nohess, grad)

        def variance_GH_end(self):
            m, delta, nohess, grad = self._pending
            if np.abs(m).max() < 0.05:
                return np.inf, None, None, _lib.FLAG_TINY
            if not grad:
                return self.o.variance(m, delta), None, None, 0
            v, g, h = self.o.variance_GH(m, delta, nohess=nohess, hess_mode="factored")
            return v, g, h, 0

    monkeypatch.setattr(mosap_mod, "SAP", StubSAP)
    monkeypatch.setattr(intproj, "candidate_variances", _fake_candidates)
    graph = io.load_graph_data(os.path.join(GOLDEN, "setup_graph_data.npz"))
    user = [[3, 0], [0], [2, 1, 0], [1, 4], [4, 1, 0], [6], [0, 6], [5, 2, 0], [3], [0, 3, 2, 1], [2], [1], [0, 1]]
    eps = [0.25 * np.sqrt(graph["C"][n][0, 0]) for n in range(2)]
    np.random.seed(5)
    mosap, blue = io.setup_solver(graph, K=3, eps=eps, groups=[list(g) for g in user])
    assert set(blue) == {"models", "samples", "errors", "total_cost"}
    assert len(blue["models"]) == len(blue["samples"]) == int((mosap.samples > 0).sum())
    assert all(s > 0 for s in blue["samples"]) and mosap.samples.dtype.kind == "i"
    picked = [mosap.flattened_groups[i] for i in np.flatnonzero(mosap.samples)]
    assert blue["models"] == picked
    assert blue["total_cost"] == mosap.samples @ mosap.costs
    assert np.all(blue["errors"] <= np.array(eps) * np.sqrt(1.0001) * (1 + 1e-9))          # projection accepts V <= 1.0001 eps^2
    assert np.allclose(blue["errors"] ** 2, mosap.variances(mosap.samples))
    # budget wins when both are given (blue_models.py:450); neither is an error
    with pytest.raises(ValueError):
        io.setup_solver(graph, K=3)


def test_fill_missing_covariances_matches_the_reference_rule():
    """blue_models.py:340-346 restated with networkx exactly as the reference walks the graph edges."""
    import networkx as nx
    from bluest_b200.pilot import fill_missing_covariances
    rng = np.random.RandomState(0)
    N = 6
    B = rng.randn(N, 3 * N); C_hat = B @ B.T / (3 * N)
    C_hat[1, 4] = C_hat[4, 1] = 1e-9 * np.sqrt(C_hat[1, 1] * C_hat[4, 4])     # practically uncorrelated pair
    A = np.full((N, N), np.nan)
    A[0, 5] = A[5, 0] = 0.0                      # never coupled: no edge
    A[2, 3] = A[3, 2] = 0.77                     # known already
    A[2, 2] = 1.5
    # the reference's loop
    G = nx.from_numpy_array(A.copy())
    for i, j, c in G.edges(data=True):
        if np.isnan(c["weight"]):
            if abs(C_hat[i, j] / np.sqrt(C_hat[i, i] * C_hat[j, j])) < 1.0e-7:
                G[i][j]["weight"] = np.inf
            else:
                G[i][j]["weight"] = C_hat[i, j]
    want = nx.adjacency_matrix(G).toarray()
    got = fill_missing_covariances(A, C_hat)
    assert np.array_equal(got, want)
    assert np.isinf(got[1, 4]) and got[0, 5] == 0.0 and got[2, 3] == 0.77 and got[2, 2] == 1.5 and got[0, 0] == C_hat[0, 0]
    assert np.isnan(A[0, 1])                       # the input is not modified


def _np_sums(Y, telescoped):
    """(N*N + N) sums of one (n, N) sample matrix, plain or telescoped, in numpy."""
    X = Y.copy()
    if telescoped:
        X[:, 1:] = Y[:, 1:] - Y[:, :-1]
    return np.concatenate([(X.T @ X).ravel(), X.sum(axis=0)])


@pytest.mark.parametrize("telescoped", [False, True])
def test_pilot_finalize_reproduces_reference_blue_fn(telescoped):
    """``blu_pilot_finalize`` (host arithmetic of the C ABI, no GPU needed) on sums formed in numpy reproduces what the
    REAL reference ``blue_fn(compute_mlmc_differences=True)`` + blue_models.py:333,339 produced (tests/golden/pilot_mlmc.npz):
    sumse, sumsc, C_hat, the MLMC difference sums and dV, for both outputs, from the plain and from the telescoped form."""
    from bluest_b200.pilot import finalize_sums
    d = np.load(os.path.join(GOLDEN, "pilot_mlmc.npz"))
    Y = d["Y"]
    No, n, M = Y.shape
    sums = np.array([_np_sums(Y[o], telescoped) for o in range(No)])
    r = finalize_sums(sums, n, M, telescoped=telescoped)
    iu = np.triu_indices(M, 1)
    for tag in ("one", "batch"):
        assert maxrel(r["sumse"], d[tag + "/sumse"]) < 1e-13
        assert maxrel(r["sumsc"], d[tag + "/sumsc"]) < 1e-13
        assert maxrel(r["C_hat"], d[tag + "/C_hat"]) < 1e-12
        for o in range(No):
            assert maxrel(r["sumsd1"][o][iu], d[tag + "/sumsd1"][o][iu]) < 1e-12
            assert maxrel(r["sumsd2"][o][iu], d[tag + "/sumsd2"][o][iu]) < 1e-12
            rel = np.abs(r["dV"][o][iu] - d[tag + "/dV"][o][iu]) / np.abs(d[tag + "/dV"][o][iu])
            assert rel.max() < (1e-13 if telescoped else 1e-11)          # element-wise: the telescoped form has no cancellation
            assert np.all(np.isnan(r["dV"][o][np.tril_indices(M)]))


def test_sample_file_matches_reference_blue_fn(tmp_path):
    """Row f4: the sample snapshot written for batched model outputs has the keys, values and append semantics of the
    file the real reference ``blue_fn(filename=...)`` writes (blue_fn.py:97-104, 132-145, 189-222), including its
    one-input-copy-per-saved-output quirk."""
    import importlib
    import ref_shim
    if not ref_shim.available():
        pytest.skip("reference checkout not present")
    ref_shim.load(with_models=True)
    bf = importlib.import_module("bluest.blue_fn")
    from bluest_b200 import io
    rng = np.random.RandomState(0)
    No, n, ls = 2, 12, [0, 2, 5]
    Y = rng.randn(No, 2 * n, len(ls)); X = rng.randn(2 * n, len(ls))

    class P:
        def __init__(self):
            self.at = 0

        def evaluate(self, ls_, samples):
            k = self.at; self.at += 1
            return [[Y[o][k, i] for i in range(len(ls_))] for o in range(No)]
    cnt = {"k": 0}

    def sampler(ls_):
        k = cnt["k"]; cnt["k"] += 1
        return [X[k, i] for i in range(len(ls_))]
    f_ref, f_our = str(tmp_path / "ref.npz"), str(tmp_path / "our.npz")
    prob = P()
    bf.blue_fn(ls, n, prob, sampler=sampler, N1=1, No=No, verbose=False, filename=f_ref)
    io.save_sample_file(f_our, ls, Y[:, :n], X[:n])
    a = np.load(io.sample_file_name(f_ref, ls), allow_pickle=True); b = np.load(io.sample_file_name(f_our, ls), allow_pickle=True)
    assert sorted(a.keys()) == sorted(b.keys())
    for k in a.keys():
        assert np.array_equal(np.asarray(a[k], dtype=float).ravel(), np.asarray(b[k], dtype=float).ravel()), k
    # appending (the reference's own append path asserts on a list-vs-array comparison, blue_fn.py:209, and cannot be
    # exercised; the intended semantics -- lists extended, n_samples added up -- are checked on the round trip)
    io.save_sample_file(f_our, ls, Y[:, n:], X[n:])
    vals, ins, ns = io.load_sample_file(f_our, ls)
    assert ns == 2 * n and np.array_equal(vals, Y) and np.array_equal(ins, X)
