"""The N>1 choreography (bluest_b200/dist.py) under gloo with world_size 2 on CPU.

The ShardedEvaluator is engine-agnostic; here a numpy engine built on the ORACLE (test
infrastructure) plays the role of the per-rank blu_ctx, so what is exercised is exactly the
host-side logic that runs on the GPU box: slice balancing, the all-reduce of the partial Phi with
its SUM-reducible support / early-out indicators, the padded all-gather of uneven row slices,
and the row-panel Hessian."""
import contextlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyEngine:
    """Same interface as bluest_b200.dist.GpuEngine, arithmetic from the oracle restatement."""

    def __init__(self, o):
        self.o = o
        self.N = o.N
        self.NP = 4 * ((o.N + 3) // 4)
        L = o.L
        self.phi = torch.zeros(o.N * o.N + 40, dtype=torch.float64)
        self.grad = torch.zeros(L, dtype=torch.float64)
        self.U = torch.zeros(L * self.NP, dtype=torch.float64)
        self.V = torch.zeros(L * self.NP, dtype=torch.float64)
        self.var, self.flags = None, 0

    def stream_context(self):
        return contextlib.nullcontext()

    def set_slice(self, lo, hi):
        self.lo, self.hi = lo, hi

    def shard_phi(self, m):
        o, N = self.o, self.N
        self.m = np.asarray(m, dtype=float)
        ms = np.zeros_like(self.m); ms[self.lo:self.hi] = self.m[self.lo:self.hi]
        phi = o.get_phi(ms)                       # groups outside the slice contribute exact zeros
        buf = np.zeros(N * N + 40)
        buf[:N * N] = np.triu(phi).ravel()        # un-mirrored upper-triangle sums, like the kernel
        flat_groups = [g for gk in o.groups for g in gk]
        for i in range(self.lo, self.hi):
            if abs(self.m[i]) > 1e-6:
                buf[N * N + np.asarray(flat_groups[i])] = 1.0
        buf[N * N + 32] = 1.0 if np.abs(self.m[self.lo:self.hi]).max(initial=0.0) >= 0.05 else 0.0
        self.phi.copy_(torch.from_numpy(buf))
        return self.phi

    def shard_finish(self, delta, want_grad, want_uv):
        o, N = self.o, self.N
        buf = self.phi.numpy()
        up = buf[:N * N].reshape(N, N)
        phi = up + np.triu(up, 1).T + delta * np.eye(N)
        self.flags = 0
        if buf[N * N + 32] <= 0:
            self.flags = 1; self.var = np.inf
            return
        supp = np.where(buf[N * N:N * N + N] > 0)[0]
        P = np.linalg.pinv(phi)
        self.var = np.linalg.pinv(phi[np.ix_(supp, supp)])[0, 0]
        self.P = P
        x = np.ascontiguousarray(P[0])
        Uall = o.ufactor(x).T                      # (L, N)
        g = -(Uall @ x)
        self.grad.zero_(); self.grad[self.lo:self.hi] = torch.from_numpy(g[self.lo:self.hi])
        if want_uv:
            U = np.zeros((o.L, self.NP)); V = np.zeros((o.L, self.NP))
            U[self.lo:self.hi, :N] = Uall[self.lo:self.hi]
            V[self.lo:self.hi, :N] = Uall[self.lo:self.hi] @ (2 * P)
            self.U.copy_(torch.from_numpy(U.ravel())); self.V.copy_(torch.from_numpy(V.ravel()))

    def eval_full(self, m, delta, want_grad, want_uv):
        self.set_slice(0, self.o.L)
        self.shard_phi(m)
        self.shard_finish(delta, want_grad, want_uv)

    def shard_hess(self, rlo, rhi):
        U = self.U.numpy().reshape(-1, self.NP); V = self.V.numpy().reshape(-1, self.NP)
        self.H = U[rlo:rhi] @ V.T

    def hv_partial(self, p):
        U = self.U.numpy().reshape(-1, self.NP)
        t = np.zeros(32)
        t[:self.NP] = np.asarray(p)[self.lo:self.hi] @ U[self.lo:self.hi]
        self.t = torch.from_numpy(t)
        return self.t

    def hv_apply(self, t):
        V = self.V.numpy().reshape(-1, self.NP)
        out = np.zeros(self.o.L)
        out[self.lo:self.hi] = V[self.lo:self.hi] @ t.numpy()[:self.NP]
        self.hv = torch.from_numpy(out)
        return self.hv

    def grad_buffer(self): return self.grad
    def u_buffer(self): return self.U
    def v_buffer(self): return self.V
    def result(self): return self.var, self.flags


def _worker(rank, world, port, N, K, ret, replicate=False):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle as orc
    from bluest_b200.dist import ShardedEvaluator
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        C = orc.wishart_cov(N, 4)
        groups = orc.enumerate_groups(N, K)
        o = orc.SapOracle(C, K, groups)
        ev = ShardedEvaluator(NumpyEngine(o), o.sizes[1:], rank, world, dist=dist, replicate_front=replicate)
        out = {}
        for name, m in (("dense", orc.dense_m(o.L, 2)), ("sparse", orc.sparse_m(o.L, N, 2)), ("tiny", 0.01 * np.ones(o.L))):
            r = ev.evaluate(m, delta=0.0, grad=True, hess=(name != "tiny"))
            e = ev.engine
            out[name] = dict(var=r["var"], flags=r["flags"], lo=r["lo"], hi=r["hi"], rlo=r["rlo"], rhi=r["rhi"],
                             grad=e.grad.numpy().copy(), H=getattr(e, "H", None) if name != "tiny" else None)
        # Hessian operator: factors stay sharded, one 32-double all-reduce per product
        m = orc.dense_m(o.L, 2)
        r = ev.evaluate_factors(m)
        p = np.random.RandomState(7).randn(o.L)
        hp = ev.hess_matvec(p, gather=True).numpy().copy()
        own = ev.hess_matvec(p, gather=False).numpy().copy()
        out["operator"] = dict(var=r["var"], hp=hp, own=own[r["lo"]:r["hi"]], lo=r["lo"], hi=r["hi"])
        ret[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,K,replicate", [(7, 7, False), (8, 4, False), (7, 7, True)])
def test_sharded_evaluation_world2(N, K, replicate):
    import oracle as orc
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, N, K, ret, replicate), nprocs=world, join=True)
    C = orc.wishart_cov(N, 4)
    groups = orc.enumerate_groups(N, K)
    o = orc.SapOracle(C, K, groups)
    for name, m in (("dense", orc.dense_m(o.L, 2)), ("sparse", orc.sparse_m(o.L, N, 2))):
        v, g, H = o.variance_GH(m, hess_mode="factored")
        rows = []
        for r in range(world):
            res = ret[r][name]
            assert abs(res["var"] - v) <= 1e-12 * abs(v)
            tol = 1e-12 if name == "dense" else 1e-9
            assert np.max(np.abs(res["grad"] - g)) <= tol * np.max(np.abs(g))       # gathered: full length on every rank
            rows.append(res["H"])
            assert res["H"].shape == (res["rhi"] - res["rlo"], o.L)
        Hcat = np.vstack(rows)
        assert np.max(np.abs(Hcat - H)) <= (1e-12 if name == "dense" else 1e-9) * np.max(np.abs(H))
    m = orc.dense_m(o.L, 2)
    v, g, H = o.variance_GH(m, hess_mode="factored")
    p = np.random.RandomState(7).randn(o.L)
    for r in range(world):
        res = ret[r]["operator"]
        assert abs(res["var"] - v) <= 1e-12 * abs(v)
        assert np.max(np.abs(res["hp"] - H @ p)) <= 1e-12 * np.max(np.abs(H @ p))        # gathered: all rows on every rank
        assert np.max(np.abs(res["own"] - (H @ p)[res["lo"]:res["hi"]])) <= 1e-12 * np.max(np.abs(H @ p))
    for r in range(world):
        assert ret[r]["tiny"]["flags"] & 1 and np.isinf(ret[r]["tiny"]["var"])
    if not replicate:
        assert ret[0]["dense"]["hi"] == ret[1]["dense"]["lo"]
    assert ret[0]["dense"]["rhi"] == ret[1]["dense"]["rlo"]


# ---------------------------------------------------------------------------------------------
# instance sweep (bluest_b200/sweep.py): budgets split across ranks, no data-path collective
# ---------------------------------------------------------------------------------------------
class _SweepProblem:
    """Oracle-backed stand-in with the members solve_sweep touches (the GPU run uses bluest_b200.SAP)."""

    def __init__(self, N, K):
        import oracle as orc
        from bluest_b200.groups import enumerate_groups, group_costs
        self.o = orc.SapOracle(orc.wishart_cov(N, 0), K, orc.enumerate_groups(N, K))
        self.L, self.N, self.e = self.o.L, N, self.o.e
        self.costs = group_costs(enumerate_groups(N, K), 2.0 ** (N - np.arange(N)))

    def variance(self, m, delta=0):
        return self.o.variance(m, delta)

    def variance_GH(self, m, delta=0, nohess=False):
        return self.o.variance_GH(m, delta, nohess=nohess, hess_mode="factored")

    def get_max_sample_constraints(self, mm):
        return [], []

    def solve(self, budget=None, eps=None, x0=None, continuous_relaxation=False, max_model_samples=None, **kw):
        from bluest_b200.solvers import scipy_solve
        return scipy_solve(self, budget=budget, eps=eps, x0=x0, max_model_samples=max_model_samples, **kw).x


def _sweep_setup():
    rng = np.random.RandomState(0)
    p = _SweepProblem(5, 3)
    x0 = np.ceil(10 * abs(rng.randn(p.L)))
    budgets = float(x0 @ p.costs) * np.array([1.2, 1.5, 2.0, 3.0, 5.0])
    return x0, budgets


def _sweep_worker(rank, world, port, ret):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from bluest_b200.sweep import solve_sweep
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x0, budgets = _sweep_setup()
        made = []
        res = solve_sweep(lambda: made.append(1) or _SweepProblem(5, 3), budgets=budgets, x0=x0, dist=dist, continuous_relaxation=True)
        ret[rank] = dict(res=res, contexts=len(made))
    finally:
        dist.destroy_process_group()


def test_instance_sweep_world2():
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from bluest_b200.sweep import solve_sweep, split_instances
    assert [split_instances(5, 2, r) for r in range(2)] == [(0, 2), (2, 5)]
    assert [split_instances(3, 4, r) for r in range(4)] == [(0, 0), (0, 1), (1, 2), (2, 3)]
    with pytest.raises(ValueError):
        split_instances(3, 2, 2)
    x0, budgets = _sweep_setup()
    serial = solve_sweep(lambda: _SweepProblem(5, 3), budgets=budgets, x0=x0, continuous_relaxation=True)
    world = 2
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_sweep_worker, args=(world, 29900 + (os.getpid() % 90), ret), nprocs=world, join=True)
    for r in range(world):
        res = ret[r]["res"]
        assert ret[r]["contexts"] == 1                                   # one context per rank, reused for its instances
        assert [x["index"] for x in res] == list(range(len(budgets)))       # every rank holds all results, in instance order
        assert [x["rank"] for x in res] == [0, 0, 1, 1, 1]
        for a, b in zip(res, serial):
            assert a["budget"] == b["budget"]
            assert np.array_equal(a["samples"], b["samples"])               # same driver, same inputs: identical allocation
            assert a["cost"] <= a["budget"] * (1 + 1e-9)
    # a larger budget never gives a larger variance
    v = [x["variance"] for x in serial]
    assert all(v[i + 1] <= v[i] * (1 + 1e-6) for i in range(len(v) - 1))
    with pytest.raises(ValueError):
        solve_sweep(lambda: None, budgets=[1.0], epss=[1.0])
