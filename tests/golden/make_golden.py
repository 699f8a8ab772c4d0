"""tests/golden/make_golden.py -- regenerates tests/golden/*.npz by running the REAL reference.

Runs only in the build container (needs /root/reference and `make -C oracle ref`); the
vectors it writes are committed and travel to the GPU box.  It imports the reference's
own ``bluest.sap.SAP`` / ``bluest.mosap.MOSAP`` / ``bluest.misc`` through
oracle/ref_shim.py (stubbed mpi4py/cvxpy/cvxopt) and records inputs and
outputs of the sample-allocation hot path.

    python tests/golden/make_golden.py

Files written
  synthetic.npz   Wishart covariances N=4,6,8: invcovs, psi, Phi, variance,
                  gradient, Hessian, cleanup matrix, for dense / sparse / tiny / integer m and delta>0
  tutorial.npz    the 5x5 covariance printed in tutorials/01_tutorial.ipynb:204-208 with the two
                  stored MLBLUE allocations (:331-334, :459-461) and the printed errors / costs
  hodgkin.npz     examples/paper_examples/hodgkin-huxley/model_graph_data.npz (M=12, 5 outputs,
                  cond 1e9-5e10) with the stored K=7 allocation (samples.npz): per-output variances
  matern.npz      restrictions_matern covariance (M=7, cond 1.5e9), K=7: reference inverses + outputs
  enumeration.npz clique enumeration / union / mappings / ES from BLUEProblem.setup_solver code
                  path (networkx) for complete and non-complete model graphs
  estimator.npz   SAP.compute_BLUE_estimator (sap.py:99-119) on random per-group sample sums
  intproj.npz     get_feasible_integer_bounds / best_closest_integer_solution_BLUE (misc.py:141-165,313-382)
  solve.npz       SAP.solve(solver="scipy") end to end: continuous solution and integer allocation
  mosap.npz       MOSAP cleanup matrices / cleanup_solution, multi-output integer projection (brute force and
                  randomised, misc.py:177-311), MOSAP.integer_projection, MOSAP.scipy_solve from a fixed x0
  setup.npz       group bookkeeping of BLUEProblem.setup_solver (default enumeration, user groups / multi_groups)
                  captured at the reference's MOSAP constructor call; setup_graph_data.npz is its input graph file
  pilot.npz       pilot-sample sums and C_hat computed with the reference's accumulation loop
  pilot_mlmc.npz  the reference's blue_fn with MLMC differences, two outputs (sums, difference sums, C_hat, dV)
"""
import os
import sys
from itertools import combinations

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF = ref_shim.REF_DEFAULT


def all_groups(N, K):
    return [[list(c) for c in combinations(range(N), k)] for k in range(1, K + 1)]


def wishart(N, seed):
    rng = np.random.RandomState(seed)
    A = rng.randn(N, 2 * N)
    return A @ A.T / (2 * N)


def record_case(ns, out, tag, C, K, groups, m_list, deltas, with_gh=True):
    """Run the reference SAP on (C, groups) and store everything the hot path produces."""
    groups_ref = [[list(g) for g in gk] for gk in groups]
    L = sum(len(gk) for gk in groups)
    sap = ns.sap.SAP(C.copy(), K, groups_ref, np.ones(L), verbose=False)
    out[f"{tag}/C"] = C
    out[f"{tag}/K"] = np.int64(K)
    out[f"{tag}/sizes"] = np.array(sap.sizes, dtype=np.int64)
    for k in range(K):
        out[f"{tag}/groups{k+1}"] = np.asarray(sap.groups[k], dtype=np.int64).reshape(-1, k + 1)
        out[f"{tag}/invcovs{k+1}"] = np.asarray(sap.invcovs[k], dtype=np.float64)
    out[f"{tag}/psi"] = sap.psi
    out[f"{tag}/e"] = np.asarray(sap.e, dtype=np.int64)
    out[f"{tag}/n_m"] = np.int64(len(m_list))
    for j, (m, delta) in enumerate(zip(m_list, deltas)):
        out[f"{tag}/m{j}"] = m
        out[f"{tag}/delta{j}"] = np.float64(delta)
        out[f"{tag}/phi{j}"] = sap.get_phi(m, delta=delta)
        try:
            out[f"{tag}/variance{j}"] = np.float64(sap.variance(m, delta=delta))
        except AssertionError:
            out[f"{tag}/variance{j}"] = np.float64(np.nan)      # model 0 not in the support
        if not with_gh:
            continue
        res = sap.variance_GH(m, delta=delta)
        out[f"{tag}/gh_len{j}"] = np.int64(len(res))
        out[f"{tag}/gh_var{j}"] = np.float64(res[0])
        out[f"{tag}/gh_grad{j}"] = np.asarray(res[1], dtype=np.float64)
        if len(res) == 3:
            out[f"{tag}/gh_hess{j}"] = res[2]
            out[f"{tag}/cleanup{j}"] = sap.get_cleanup_matrix(m, delta=delta)
    return sap


def make_synthetic(ns):
    out = {}
    for N, K, seed in [(4, 4, 0), (6, 6, 1), (8, 4, 2), (8, 8, 3)]:
        C = wishart(N, seed)
        groups = all_groups(N, K)
        L = sum(len(g) for g in groups)
        rng = np.random.RandomState(100 + seed)
        m_dense = 1.0 + 10.0 * rng.rand(L)
        m_sparse = np.zeros(L)
        nz = rng.choice(L, size=min(2 * N, L), replace=False)
        m_sparse[nz] = np.ceil(50 * rng.rand(len(nz)))
        m_sparse[0] = 7.0
        # only a few models supported: groups {0}, {1}, {0,1}
        m_few = np.zeros(L); m_few[0] = 3.0; m_few[1] = 2.0; m_few[N] = 5.0
        m_tiny = 0.01 * np.ones(L)                                     # early-out: max|m| < 0.05
        m_int = np.round(m_dense).astype(np.int64)                     # integer m after rounding
        m_thr = m_dense.copy(); m_thr[1] = 5e-7; m_thr[2] = 2e-6       # support threshold 1e-6
        m_no0 = np.zeros(L); m_no0[1] = 4.0; m_no0[2] = 3.0            # model 0 not sampled
        ms = [m_dense, m_sparse, m_few, m_tiny, m_int, m_thr, m_dense, m_no0]
        ds = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1e-6, 0.0]
        record_case(ns, out, f"N{N}K{K}", C, K, groups, ms, ds)
    # a non-complete, user-style group list with an EMPTY size class (k=2 missing)
    N = 6
    C = wishart(N, 7)
    groups = [[[0], [2], [5]], [], [[0, 1, 2], [0, 3, 5], [1, 2, 4]], [[0, 1, 2, 3], [0, 2, 4, 5]]]
    L = sum(len(g) for g in groups)
    rng = np.random.RandomState(5)
    # NOTE: with an empty class the reference's variance_GH itself raises IndexError (pybind
    # ``.data(0)`` on the empty invcovs array, misc.py:620), so only psi/Phi/variance are pinned.
    record_case(ns, out, "ragged", C, 4, groups, [1.0 + 3.0 * rng.rand(L)], [0.0], with_gh=False)
    np.savez_compressed(os.path.join(OUT, "synthetic.npz"), **out)
    print("synthetic.npz:", len(out), "arrays")


def make_tutorial(ns):
    C = np.array([[2.53500581, 2.3583669, 2.36312599, 1.66331444, 0.72897923],
                  [2.3583669, 2.27345322, 2.07702441, 1.67029559, 0.80520217],
                  [2.36312599, 2.07702441, 2.40127866, 1.37242096, 0.47072499],
                  [1.66331444, 1.67029559, 1.37242096, 1.29027193, 0.67275196],
                  [0.72897923, 0.80520217, 0.47072499, 0.67275196, 1.56637342]])
    N = 5
    model_costs = np.array([2.0 ** (N - i) for i in range(N)])
    groups = all_groups(N, N)
    flat = [g for gk in groups for g in gk]
    L = len(flat)
    costs = np.array([model_costs[g].sum() for g in flat])
    sap = ns.sap.SAP(C.copy(), N, [[list(g) for g in gk] for gk in groups], costs, verbose=False)
    out = {"C": C, "model_costs": model_costs}
    allocs = [([[2], [3], [0, 2, 3], [1, 2, 3], [0, 1, 2, 3]], [5938, 6797, 1, 493, 43], 0.01592249, 91120),
              ([[1], [0, 3], [0, 1, 2, 3, 4]], [9882, 1794, 305], 0.01592225, 241606)]
    for a, (gs, ns_, err, cost) in enumerate(allocs):
        m = np.zeros(L, dtype=np.int64)
        for g, n in zip(gs, ns_):
            m[flat.index(g)] = n
        out[f"m{a}"] = m
        out[f"printed_error{a}"] = np.float64(err)
        out[f"printed_cost{a}"] = np.float64(cost)
        out[f"ref_variance{a}"] = np.float64(sap.variance(m))
        out[f"ref_cost{a}"] = np.float64(m @ costs)
        v, g_, _ = sap.variance_GH(m.astype(float), nohess=True)
        out[f"ref_gh_var{a}"] = np.float64(v)
        out[f"ref_grad{a}"] = g_
        print("tutorial", a, np.sqrt(out[f"ref_variance{a}"]), out[f"ref_cost{a}"])
    np.savez_compressed(os.path.join(OUT, "tutorial.npz"), **out)


def make_hodgkin(ns):
    d = np.load(os.path.join(REF, "examples/paper_examples/hodgkin-huxley/model_graph_data.npz"))
    s = np.load(os.path.join(REF, "examples/paper_examples/hodgkin-huxley/samples.npz"))
    M = int(d["M"]); No = int(d["n_outputs"]); K = 7
    groups = all_groups(M, K)
    L = sum(len(g) for g in groups)
    samples = s["samples"]
    assert len(samples) == L
    out = {"M": np.int64(M), "K": np.int64(K), "n_outputs": np.int64(No), "costs": d["costs"], "samples": samples}
    for n in range(No):
        C = d[f"C{n}"]
        sap = ns.sap.SAP(C.copy(), K, [[list(g) for g in gk] for gk in groups], np.ones(L), verbose=False)
        out[f"C{n}"] = C
        out[f"variance{n}"] = np.float64(sap.variance(samples))
        v, g, _ = sap.variance_GH(samples.astype(float), nohess=True)
        out[f"gh_var{n}"] = np.float64(v)
        out[f"grad{n}"] = g
        out[f"phi{n}"] = sap.get_phi(samples)
        if n == 3:                                   # worst-conditioned output only (size)
            for k in range(K):
                out[f"invcovs{n}_{k+1}"] = np.asarray(sap.invcovs[k])
        print("hodgkin", n, np.sqrt(out[f"variance{n}"] / C[0, 0]), np.linalg.cond(C))
    flat_costs = np.array([d["costs"][g].sum() for gk in groups for g in gk])
    out["total_cost"] = np.float64(samples @ flat_costs)
    np.savez_compressed(os.path.join(OUT, "hodgkin.npz"), **out)


def make_matern(ns):
    d = np.load(os.path.join(REF, "examples/paper_examples/restrictions_matern/restrictions_matern_model_data.npz"))
    C = d["C0"]
    N = C.shape[0]
    out = {}
    groups = all_groups(N, N)
    L = sum(len(g) for g in groups)
    rng = np.random.RandomState(11)
    record_case(ns, out, "matern", C, N, groups, [1.0 + 10.0 * rng.rand(L)], [0.0])
    out["cond"] = np.float64(np.linalg.cond(C))
    np.savez_compressed(os.path.join(OUT, "matern.npz"), **out)
    print("matern cond", out["cond"])


def make_enumeration(ns):
    """Drive the clique enumeration exactly the way BLUEProblem.setup_solver does
    (blue_models.py:458-501) on stand-alone networkx graphs."""
    import networkx as nx
    bm = ns.blue_models
    out = {}

    def enumerate_ref(graphs, K):
        Ks, multi_groups = [], []
        M = graphs[0].number_of_nodes()
        K = min(K, M)
        for G in graphs:
            groups = [[] for _ in range(K)]
            SG = G.subgraph(nx.node_connected_component(G, 0))      # model-0 component (check_graphs)
            for clique in nx.enumerate_all_cliques(G):
                kn = len(clique)
                if kn > K:
                    break
                if all(node in SG for node in clique):
                    groups[kn - 1].append(clique)
            groups = [item for item in groups if len(item) > 0]
            multi_groups.append(groups)
            Ks.append(min(K, len(groups)))
        K = max(Ks)
        groups = [[] for _ in range(K)]
        for n in range(len(graphs)):
            for k in range(Ks[n]):
                for group in multi_groups[n][k]:
                    if group not in groups[k]:
                        groups[k].append(group)
        for k in range(K):
            groups[k].sort()
        return K, Ks, groups, multi_groups

    def store(tag, K, Ks, groups, multi_groups, C_list):
        out[f"{tag}/K"] = np.int64(K)
        out[f"{tag}/Ks"] = np.array(Ks, dtype=np.int64)
        out[f"{tag}/n_outputs"] = np.int64(len(multi_groups))
        for k in range(K):
            out[f"{tag}/groups{k+1}"] = np.array(groups[k], dtype=np.int64).reshape(-1, k + 1)
        for n, mg in enumerate(multi_groups):
            for k in range(Ks[n]):
                out[f"{tag}/multi{n}_groups{k+1}"] = np.array(mg[k], dtype=np.int64).reshape(-1, k + 1)
            out[f"{tag}/adj{n}"] = C_list[n]
        # MOSAP mappings / ES through the reference constructor
        L = sum(len(g) for g in groups)
        N = C_list[0].shape[0]
        Cs = [wishart(N, 40 + n) for n in range(len(multi_groups))]
        mosap = ns.mosap.MOSAP([c.copy() for c in Cs], K, list(Ks), [[list(g) for g in gk] for gk in groups],
                               [[[list(g) for g in gk] for gk in mg] for mg in multi_groups],
                               np.ones(L), [np.ones(sum(len(g) for g in mg)) for mg in multi_groups], verbose=False)
        for n in range(len(multi_groups)):
            out[f"{tag}/mapping{n}"] = np.asarray(mosap.mappings[n], dtype=np.int64)
            out[f"{tag}/Cwish{n}"] = Cs[n]
        out[f"{tag}/ES"] = np.array(mosap.ES, dtype=np.int64)
        rng = np.random.RandomState(3)
        m = 1.0 + 5.0 * rng.rand(L)
        out[f"{tag}/m"] = m
        out[f"{tag}/variances"] = np.array(mosap.variances(m), dtype=np.float64)
        vs, gs, _ = mosap.variance_GH(m, nohess=True)
        for n in range(len(multi_groups)):
            out[f"{tag}/grad{n}"] = gs[n]

    # complete graphs: N=4 (K=4), N=6 (K=3)
    for N, K in [(4, 4), (6, 3)]:
        A = np.ones((N, N))
        G = nx.from_numpy_array(A)
        res = enumerate_ref([G], K)
        store(f"complete_N{N}_K{K}", *res, [A])
    # two outputs with different, non-complete coupling graphs on 6 models
    A0 = np.ones((6, 6)); A0[0, 5] = A0[5, 0] = 0; A0[2, 4] = A0[4, 2] = 0
    A1 = np.ones((6, 6)); A1[1, 3] = A1[3, 1] = 0
    G0 = nx.from_numpy_array(A0); G1 = nx.from_numpy_array(A1)
    res = enumerate_ref([G0, G1], 4)
    store("two_outputs", *res, [A0, A1])
    np.savez_compressed(os.path.join(OUT, "enumeration.npz"), **out)
    print("enumeration.npz:", len(out), "arrays")


def make_estimator(ns):
    """SAP.compute_BLUE_estimator (sap.py:99-119 -> misc.PHIinvY0) on random per-group sample sums."""
    out = {}
    for tag, N, K, seed in [("N6K6", 6, 6, 31), ("N9K4", 9, 4, 32)]:
        C = wishart(N, seed)
        groups = all_groups(N, K)
        flat = [g for gk in groups for g in gk]
        L = len(flat)
        rng = np.random.RandomState(seed)
        sap = ns.sap.SAP(C.copy(), K, [[list(g) for g in gk] for gk in groups], np.ones(L), verbose=False)
        samples = np.zeros(L, dtype=np.int64)
        nz = rng.choice(L, size=2 * N, replace=False)
        samples[nz] = rng.randint(1, 200, size=len(nz)); samples[0] = 50
        sums = [samples[i] * (1.0 + 0.1 * rng.randn(len(flat[i]))) for i in range(L)]
        mu, var = sap.compute_BLUE_estimator(sums, samples=samples)
        out[f"{tag}/C"] = C; out[f"{tag}/K"] = np.int64(K); out[f"{tag}/samples"] = samples
        out[f"{tag}/sums"] = np.concatenate(sums); out[f"{tag}/mu"] = np.float64(mu); out[f"{tag}/var"] = np.float64(var)
        for k in range(K):
            out[f"{tag}/invcovs{k+1}"] = np.asarray(sap.invcovs[k])
        print("estimator", tag, mu, var)
    np.savez_compressed(os.path.join(OUT, "estimator.npz"), **out)


def make_intproj(ns):
    """best_closest_integer_solution_BLUE (misc.py:313-382) in budget and eps mode."""
    out = {}
    for tag, N, K, seed in [("N6K6", 6, 6, 41), ("N8K3", 8, 3, 42)]:
        C = wishart(N, seed)
        groups = all_groups(N, K)
        flat = [g for gk in groups for g in gk]
        L = len(flat)
        model_costs = 2.0 ** (N - np.arange(N))
        w = np.array([model_costs[g].sum() for g in flat])
        sap = ns.sap.SAP(C.copy(), K, [[list(g) for g in gk] for gk in groups], w, verbose=False)
        rng = np.random.RandomState(seed)
        sol = np.zeros(L)
        nz = rng.choice(L, size=N + 3, replace=False)
        sol[nz] = 0.3 + 40 * rng.rand(len(nz)); sol[0] = 2.6
        out[f"{tag}/C"] = C; out[f"{tag}/K"] = np.int64(K); out[f"{tag}/w"] = w; out[f"{tag}/sol"] = sol
        for k in range(K):
            out[f"{tag}/invcovs{k+1}"] = np.asarray(sap.invcovs[k])
        lb, ub, idx = ns.misc.get_feasible_integer_bounds(sol, N, e=sap.e)
        out[f"{tag}/lb"] = lb; out[f"{tag}/ub"] = ub; out[f"{tag}/idx"] = idx
        budget = float(w @ np.round(sol)) * 1.01
        val, fval = ns.misc.best_closest_integer_solution_BLUE(sol, sap.psi, w, sap.e, budget=budget)
        out[f"{tag}/budget"] = np.float64(budget); out[f"{tag}/budget_val"] = val; out[f"{tag}/budget_fval"] = np.float64(fval)
        eps = float(np.sqrt(sap.variance(np.round(sol).astype(int)) * 1.02))
        val2, fval2 = ns.misc.best_closest_integer_solution_BLUE(sol, sap.psi, w, sap.e, eps=eps)
        out[f"{tag}/eps"] = np.float64(eps); out[f"{tag}/eps_val"] = val2; out[f"{tag}/eps_fval"] = np.float64(fval2)
        ip = sap.integer_projection(sol, budget=budget)
        out[f"{tag}/projection_budget"] = np.asarray(ip)
        print("intproj", tag, len(idx), fval, fval2, int(np.abs(val - val2).sum()))
    np.savez_compressed(os.path.join(OUT, "intproj.npz"), **out)


def make_solve(ns):
    """End-to-end SAP.solve(solver="scipy") of the reference (the only solver runnable without
    cvxopt/cvxpy/ipopt): continuous solution and integer allocation from a fixed x0."""
    out = {}
    C5 = np.array([[2.53500581, 2.3583669, 2.36312599, 1.66331444, 0.72897923],
                   [2.3583669, 2.27345322, 2.07702441, 1.67029559, 0.80520217],
                   [2.36312599, 2.07702441, 2.40127866, 1.37242096, 0.47072499],
                   [1.66331444, 1.67029559, 1.37242096, 1.29027193, 0.67275196],
                   [0.72897923, 0.80520217, 0.47072499, 0.67275196, 1.56637342]])
    for tag, C, K, seed in [("tutorial", C5, 5, 51), ("N6K3", wishart(6, 52), 3, 52)]:
        N = C.shape[0]
        groups = all_groups(N, K)
        flat = [g for gk in groups for g in gk]
        L = len(flat)
        model_costs = 2.0 ** (N - np.arange(N))
        w = np.array([model_costs[g].sum() for g in flat])
        sap = ns.sap.SAP(C.copy(), K, [[list(g) for g in gk] for gk in groups], w, verbose=False)
        x0 = np.ceil(10 * abs(np.random.RandomState(seed).randn(L)))
        budget = 100.0 * w.max()
        cont = sap.solve(budget=budget, solver="scipy", x0=x0.copy(), continuous_relaxation=True)
        ints = sap.solve(budget=budget, solver="scipy", x0=x0.copy(), continuous_relaxation=False)
        out[f"{tag}/C"] = C; out[f"{tag}/K"] = np.int64(K); out[f"{tag}/w"] = w; out[f"{tag}/x0"] = x0
        out[f"{tag}/budget"] = np.float64(budget); out[f"{tag}/continuous"] = cont; out[f"{tag}/integer"] = np.asarray(ints)
        out[f"{tag}/variance"] = np.float64(sap.variance(ints)); out[f"{tag}/cost"] = np.float64(ints @ w)
        print("solve", tag, L, float(sap.variance(ints)), float(ints @ w), int((ints > 0).sum()))
    np.savez_compressed(os.path.join(OUT, "solve.npz"), **out)


def make_mosap(ns):
    """Multi-output host orchestration of the reference (mosap.py:125-331, 555-610; misc.py:177-311):
    cleanup matrices and cleanup_solution, the brute-force and the randomised integer projection,
    MOSAP.integer_projection and MOSAP.scipy_solve from a fixed x0.  np.random is seeded before every
    call that draws from it (misc.py:203-213)."""
    import contextlib
    import io
    out = {}
    quiet = lambda: contextlib.redirect_stdout(io.StringIO())

    def build(tag, Cs, K, Ks, groups, multi_groups, model_costs):
        flat = [g for gk in groups for g in gk]
        L = len(flat)
        w = np.array([model_costs[list(g)].sum() for g in flat])
        mw = [np.array([model_costs[list(g)].sum() for gk in mg for g in gk]) for mg in multi_groups]
        mosap = ns.mosap.MOSAP([c.copy() for c in Cs], K, list(Ks), [[list(g) for g in gk] for gk in groups],
                               [[[list(g) for g in gk] for gk in mg] for mg in multi_groups], w, mw, verbose=False)
        out[f"{tag}/K"] = np.int64(K); out[f"{tag}/Ks"] = np.array(Ks, dtype=np.int64); out[f"{tag}/w"] = w
        out[f"{tag}/n_outputs"] = np.int64(len(Cs))
        for n, c in enumerate(Cs):
            out[f"{tag}/C{n}"] = c
        for k in range(K):
            out[f"{tag}/groups{k+1}"] = np.array(groups[k], dtype=np.int64).reshape(-1, k + 1)
        for n, mg in enumerate(multi_groups):
            for k in range(Ks[n]):
                out[f"{tag}/multi{n}_groups{k+1}"] = np.array(mg[k], dtype=np.int64).reshape(-1, k + 1)
        return mosap, w, L

    def projections(tag, mosap, w, L, sol, name):
        N = mosap.N
        out[f"{tag}/{name}_sol"] = sol
        psis = [mosap.SAPS[n].psi for n in range(mosap.n_outputs)]
        lb, ub, idx = ns.misc.get_feasible_integer_bounds(sol, N, e=mosap.e)
        out[f"{tag}/{name}_idx"] = idx
        budget = float(w @ np.round(sol)) * 1.01
        vr = np.array(mosap.variances(np.round(sol).astype(int)))
        eps = np.sqrt(vr * 1.05)
        out[f"{tag}/{name}_budget"] = np.float64(budget); out[f"{tag}/{name}_eps"] = eps
        with quiet():
            np.random.seed(1234)
            val, fval = ns.misc.best_closest_integer_solution_BLUE_multi(sol.copy(), psis, w, mosap.e, mosap.mappings, budget=budget)
            np.random.seed(1234)
            val2, fval2 = ns.misc.best_closest_integer_solution_BLUE_multi(sol.copy(), psis, w, mosap.e, mosap.mappings, eps=eps)
            np.random.seed(1234)
            ipb = mosap.integer_projection(sol.copy(), budget=budget)
            np.random.seed(1234)
            ipe = mosap.integer_projection(sol.copy(), eps=eps)
            # max-sample caps that bind: model 1 at most ceil(half of what the budget projection uses)
            caps = np.inf * np.ones(N); caps[1] = max(1.0, np.ceil(0.5 * (np.array(mosap.ES[1]) @ ipb)))
            np.random.seed(1234)
            ipc = mosap.integer_projection(sol.copy(), budget=budget, max_model_samples=caps)
        out[f"{tag}/{name}_budget_val"] = val; out[f"{tag}/{name}_budget_fval"] = np.float64(fval)
        out[f"{tag}/{name}_eps_val"] = val2; out[f"{tag}/{name}_eps_fval"] = np.float64(fval2)
        out[f"{tag}/{name}_projection_budget"] = np.asarray(ipb); out[f"{tag}/{name}_projection_eps"] = np.asarray(ipe)
        out[f"{tag}/{name}_caps"] = caps; out[f"{tag}/{name}_projection_caps"] = np.asarray(ipc)
        print("mosap", tag, name, "LL", len(idx), fval, fval2, int(ipb.sum()), int(ipe.sum()), int(ipc.sum()))

    # ---- case A: two outputs on different coupling graphs (non-trivial mappings), from enumeration.npz
    d = np.load(os.path.join(OUT, "enumeration.npz"))
    tag = "two_outputs"
    K = int(d[f"{tag}/K"]); Ks = d[f"{tag}/Ks"].tolist()
    groups = [d[f"{tag}/groups{k+1}"].tolist() for k in range(K)]
    multi = [[d[f"{tag}/multi{n}_groups{k+1}"].tolist() for k in range(Ks[n])] for n in range(2)]
    N = 6
    mosap, w, L = build(tag, [wishart(N, 61), wishart(N, 62)], K, Ks, groups, multi, 2.0 ** (N - np.arange(N)))
    rng = np.random.RandomState(63)
    sol = np.zeros(L); nz = rng.choice(L, size=N + 2, replace=False)
    sol[nz] = 0.3 + 30 * rng.rand(len(nz)); sol[0] = 3.4
    projections(tag, mosap, w, L, sol, "small")
    m = np.zeros(L); nz = rng.choice(L, size=3 * N, replace=False); m[nz] = 1 + 20 * rng.rand(len(nz)); m[0] = 5.0
    out[f"{tag}/cleanup_m"] = m
    out[f"{tag}/cleanup_X"] = mosap.get_cleanup_matrices(m.copy())
    with quiet():
        out[f"{tag}/cleanup_result"] = mosap.cleanup_solution(m.copy())
    out[f"{tag}/cleanup_variances_before"] = np.array(mosap.variances(m)); out[f"{tag}/cleanup_variances_after"] = np.array(mosap.variances(out[f"{tag}/cleanup_result"]))

    # ---- case B: three outputs sharing all groups of 8 models up to size 3 (BASELINE config 4 style): more than
    #      15 rounding candidates, so the reference randomises (misc.py:196-224)
    tag = "shared_N8K3"
    N, K, No = 8, 3, 3
    groups = all_groups(N, K)
    mosap, w, L = build(tag, [wishart(N, 70 + n) for n in range(No)], K, [K] * No, groups, [groups] * No, 2.0 ** (N - np.arange(N)))
    rng = np.random.RandomState(73)
    flat = [g for gk in groups for g in gk]
    with0 = np.array([i for i, g in enumerate(flat) if 0 in g]); without0 = np.array([i for i, g in enumerate(flat) if 0 not in g])
    sol = np.zeros(L)
    sol[rng.choice(without0, size=11, replace=False)] = 10.3 + 30 * rng.rand(11)      # the 9 largest entries avoid model 0 ...
    sol[rng.choice(with0, size=8, replace=False)] = 1.2 + 3 * rng.rand(8)             # ... and 8 more candidates contain it: LL = 17 > 15
    projections(tag, mosap, w, L, sol, "big")

    # ---- case C: scipy trust-constr driver, three outputs sharing the groups of 5 models up to size 3
    tag = "solve_N5K3"
    N, K, No = 5, 3, 3
    groups = all_groups(N, K)
    mosap, w, L = build(tag, [wishart(N, 80 + n) for n in range(No)], K, [K] * No, groups, [groups] * No, 2.0 ** (N - np.arange(N)))
    x0 = np.ceil(10 * abs(np.random.RandomState(83).randn(L)))
    budget = float(x0 @ w) / 0.9
    x0b = np.concatenate([[max(mosap.variances(x0, delta=1.0e-15))], x0])
    with quiet():
        cont_b = mosap.scipy_solve(budget=budget, x0=x0b.copy())
    eps = np.sqrt(np.array(mosap.variances(x0)) * 1.5)
    with quiet():
        cont_e = mosap.scipy_solve(eps=eps, x0=x0.copy())
        np.random.seed(1234)
        int_b = mosap.integer_projection(cont_b.copy(), budget=budget)
        np.random.seed(1234)
        int_e = mosap.integer_projection(cont_e.copy(), eps=eps)
    out[f"{tag}/x0"] = x0; out[f"{tag}/budget"] = np.float64(budget); out[f"{tag}/eps"] = eps
    out[f"{tag}/continuous_budget"] = cont_b; out[f"{tag}/continuous_eps"] = cont_e
    out[f"{tag}/integer_budget"] = np.asarray(int_b); out[f"{tag}/integer_eps"] = np.asarray(int_e)
    out[f"{tag}/variances_budget"] = np.array(mosap.variances(cont_b)); out[f"{tag}/variances_eps"] = np.array(mosap.variances(cont_e))
    print("mosap solve", max(mosap.variances(cont_b)), float(cont_b @ w), budget, np.array(mosap.variances(cont_e)) / eps ** 2, float(cont_e @ w))
    np.savez_compressed(os.path.join(OUT, "mosap.npz"), **out)


def make_setup(ns):
    """The group bookkeeping of BLUEProblem.setup_solver (blue_models.py:453-509): default clique enumeration,
    user-supplied ``groups`` and per-output ``multi_groups`` (unsorted members, non-cliques, groups outside
    model 0's component), captured at the MOSAP constructor call of the REAL reference class."""
    import contextlib
    import io
    import networkx as nx
    bm = ns.blue_models
    M, No = 7, 2
    rng = np.random.RandomState(91)
    adjs = []
    for n in range(No):
        A = wishart(M, 92 + n)
        A[1, 4] = A[4, 1] = 0.0                      # models 1 and 4 cannot be coupled
        A[2, 5] = A[5, 2] = np.inf                   # uncorrelated, still an edge
        if n == 1:
            A[6, :] = 0.0; A[:, 6] = 0.0; A[6, 6] = 1.3     # model 6 is cut off from model 0's component in output 1
            A[0, 3] = A[3, 0] = 0.0
        adjs.append(A)
    SG = []
    for A in adjs:
        comp = sorted(nx.node_connected_component(nx.from_numpy_array(A), 0))
        SG.append(comp + [-1] * (M - len(comp)))     # rectangular for np.savez; -1 never matches a model id
    costs = 2.0 ** (M - np.arange(M))
    path = os.path.join(OUT, "setup_graph_data.npz")
    np.savez(path, M=M, n_outputs=No, costs=costs, C0=adjs[0], C1=adjs[1], SG=np.array(SG), dV=np.nan * np.ones((No, M, M)))

    captured = {}

    class Capture(Exception):
        pass

    def fake_mosap(C, K, Ks, groups, multi_groups, costs, multi_costs, verbose=True):
        captured.update(K=K, Ks=list(Ks), groups=[[list(g) for g in gk] for gk in groups],
                        multi_groups=[[[list(g) for g in gk] for gk in mg] for mg in multi_groups],
                        costs=np.array(costs), multi_costs=[np.array(c) for c in multi_costs])
        raise Capture()

    real = bm.MOSAP
    bm.MOSAP = fake_mosap
    out = {}
    try:
        user = [[3, 0], [0], [2, 1, 0], [1, 4], [4, 1, 0], [6], [0, 6], [5, 2, 0], [3], [0, 3, 2, 1], [2]]
        multi = [[[0], [1, 0], [2, 0, 1], [4, 3]], [[0], [0, 3], [2, 0], [6, 0], [1, 2], [5]]]
        cases = {"default_K3": dict(K=3), "default_K7": dict(K=7), "user_groups": dict(K=3, groups=user), "user_multi": dict(K=3, multi_groups=multi)}
        for tag, kw in cases.items():
            with contextlib.redirect_stdout(io.StringIO()):
                prob = bm.BLUEProblem(M, datafile=path, n_outputs=No, verbose=False)
            kw = {k: ([list(g) for g in v] if k == "groups" else [[list(g) for g in mg] for mg in v] if k == "multi_groups" else v) for k, v in kw.items()}
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    prob.setup_solver(eps=0.1, solver="scipy", **kw)
            except Capture:
                pass
            out[f"{tag}/K"] = np.int64(captured["K"]); out[f"{tag}/Ks"] = np.array(captured["Ks"], dtype=np.int64)
            out[f"{tag}/costs"] = captured["costs"]
            for k, gk in enumerate(captured["groups"]):
                out[f"{tag}/groups{k+1}"] = np.array(gk, dtype=np.int64).reshape(-1, k + 1)
            for n, mg in enumerate(captured["multi_groups"]):
                out[f"{tag}/n_classes{n}"] = np.int64(len(mg))
                for k, gk in enumerate(mg):
                    out[f"{tag}/multi{n}_groups{k+1}"] = np.array(gk, dtype=np.int64).reshape(-1, k + 1)
                out[f"{tag}/multi_costs{n}"] = captured["multi_costs"][n]
            print("setup", tag, captured["K"], captured["Ks"], [len(g) for g in captured["groups"]])
    finally:
        bm.MOSAP = real
    out["user_groups_input"] = np.array([g + [-1] * (4 - len(g)) for g in user], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "setup.npz"), **out)


def make_pilot():
    """The reference accumulates the pilot sums one sample at a time (blue_fn.py:159-167, N1=1)
    and then forms C_hat = sumsc/N - outer(sumse, sumse)/N^2 (blue_models.py:333)."""
    rng = np.random.RandomState(21)
    N, n = 5, 257
    C = wishart(N, 9)
    Y = rng.standard_normal((n, N)) @ np.linalg.cholesky(C).T + 0.3
    sumse = [0.0 for _ in range(N)]
    sumsc = np.zeros((N, N))
    for s in range(n):
        Ps = Y[s]
        for i in range(N):
            sumse[i] += Ps[i]
        sumsc += np.array([[Ps[i] * Ps[j] for i in range(N)] for j in range(N)])
    sumse = np.array(sumse)
    C_hat = sumsc / n - np.outer(sumse, sumse) / n ** 2
    np.savez_compressed(os.path.join(OUT, "pilot.npz"), Y=Y, sumse=sumse, sumsc=sumsc, C_hat=C_hat)


def make_pilot_mlmc():
    """The reference's own sampling loop ``bluest.blue_fn.blue_fn`` (blue_fn.py:36-227) with
    ``compute_mlmc_differences=True`` on two outputs: column sums, Gram sums and the MLMC difference sums, followed
    by the formulas of ``estimate_missing_covariances`` (blue_models.py:333, :339).  One case sample by sample
    (N1 = 1), one in batches (N1 = 16), a strongly coupled (multilevel-like) model hierarchy in both."""
    import importlib
    bf = importlib.import_module("bluest.blue_fn")
    rng = np.random.RandomState(33)
    M, n, No = 6, 400, 2
    base = rng.standard_normal((n, 1))
    Ys = []
    for o in range(No):                                     # model j = common signal + a perturbation that shrinks with j
        pert = rng.standard_normal((n, M)) * (0.5 ** np.arange(M))[None, :] * (1.0 + o)
        Ys.append((1.0 + 0.2 * o) * base + pert + 0.25 * (o + 1))
    Ys = np.array(Ys)                                       # (No, n, M)

    class Problem:
        def __init__(self):
            self.at = 0

        def evaluate(self, ls, samples):
            if np.ndim(samples[0]) == 0:                    # one sample
                k = self.at; self.at += 1
                return [[Ys[o][k, l] for l in ls] for o in range(No)]
            cnt = len(samples[0])
            k = self.at; self.at += cnt
            return [[Ys[o][k:k + cnt, l] for l in ls] for o in range(No)]

    ls = list(range(M))
    out = {"Y": Ys}
    for tag, N1 in (("one", 1), ("batch", 16)):
        prob = Problem()
        if N1 == 1:
            sampler = lambda ls_: [0.0 for _ in ls_]
        else:
            sampler = lambda ls_, N=1: [np.zeros(N) for _ in ls_]
        sumse, sumsc, cost, sd1, sd2 = bf.blue_fn(ls, n, prob, sampler=sampler, N1=N1, No=No, verbose=False, compute_mlmc_differences=True)
        sumse = np.array([[float(v) for v in row] for row in sumse])
        sd1 = np.array([[[float(sd1[o][i][j]) for j in range(M)] for i in range(M)] for o in range(No)])
        sd2 = np.array([[[float(sd2[o][i][j]) for j in range(M)] for i in range(M)] for o in range(No)])
        C_hat = np.array([sumsc[o] / n - np.outer(sumse[o], sumse[o]) / n ** 2 for o in range(No)])      # blue_models.py:333
        dV = np.full((No, M, M), np.nan)
        for o in range(No):
            for i in range(M):
                for j in range(i + 1, M):
                    dV[o, i, j] = sd2[o, i, j] / n - (sd1[o, i, j] / n) * (sd1[o, i, j] / n)             # blue_models.py:339
        out.update({tag + "/sumse": sumse, tag + "/sumsc": np.array(sumsc), tag + "/sumsd1": sd1, tag + "/sumsd2": sd2,
                    tag + "/C_hat": C_hat, tag + "/dV": dV})
    np.savez_compressed(os.path.join(OUT, "pilot_mlmc.npz"), **out)


if __name__ == "__main__":
    ns = ref_shim.load(with_models=True)
    make_synthetic(ns)
    make_tutorial(ns)
    make_hodgkin(ns)
    make_matern(ns)
    make_enumeration(ns)
    make_estimator(ns)
    make_intproj(ns)
    make_solve(ns)
    make_mosap(ns)
    make_setup(ns)
    make_pilot()
    make_pilot_mlmc()
    print("done")
