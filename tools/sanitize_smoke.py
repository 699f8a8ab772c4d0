"""Small run through every kernel of the library for compute-sanitizer (memcheck / racecheck):
entry-per-lane and lane-per-group paths, Phi / pinv / gradient / U,V / dense Hessian, the operator,
sliced contexts, cleanup matrix, estimator, integer projection, pilot Gram, Level-1 drop-ins."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import bluest_b200 as blu, oracle as orc
from bluest_b200.dist import GpuEngine

def check(a, b, tol, what):
    err = float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1e-300, np.max(np.abs(np.asarray(b)))))
    print("%-40s maxrel %.2e" % (what, err), flush=True)
    assert err < tol, what

for N, K in ((9, 9), (17, 3), (14, 14)):
    groups = orc.enumerate_groups(N, K)
    L = sum(len(g) for g in groups)
    o = orc.SapOracle(orc.wishart_cov(N, 1), K, groups)
    sap = blu.SAP(orc.wishart_cov(N, 1), K, [[list(g) for g in gk] for gk in groups], np.ones(L), verbose=False)
    m = orc.dense_m(L, 1)
    check(sap.get_phi(m), o.get_phi(m), 1e-12, "N=%d phi" % N)
    vo, go, _ = o.variance_GH(m, nohess=True)
    v, g, _ = sap.variance_GH(m, nohess=True)
    check(g, go, 1e-12, "N=%d gradient" % N)
    v, g, op = sap.variance_GH_operator(m)
    P = np.linalg.pinv(o.get_phi(m)); U = o.ufactor(np.ascontiguousarray(P[0]))
    p = np.random.RandomState(2).randn(L)
    check(op @ p, 2.0 * (U.T @ (P @ (U @ p))), 1e-12, "N=%d operator product" % N)
    if L <= 2100:
        v, g, H = sap.variance_GH(m)
        check(H, 2.0 * (U.T @ P @ U), 1e-12, "N=%d dense Hessian" % N)
        check(sap.get_cleanup_matrix(m), o.cleanup_matrix(m), 1e-12, "N=%d cleanup matrix" % N)
    ms = orc.sparse_m(L, N, 1)
    check(sap.variance(ms), o.variance(ms), 1e-10, "N=%d sparse variance" % N)
    # a sliced context: U rows of the slice only, partial operator product
    e = GpuEngine(sap)
    e.shard_phi(m); sap.sync()
    e.set_slice(L // 3, 2 * L // 3 + 1)
    e.shard_finish(0.0, True, 2); sap.sync()
    import torch
    t = e.hv_partial(torch.from_numpy(p).cuda()); sap.sync()
    check(t.cpu().numpy()[:N], U[:, L // 3:2 * L // 3 + 1] @ p[L // 3:2 * L // 3 + 1], 1e-12, "N=%d sliced partial t" % N)
    e.set_slice(0, L)
    sap.close()

# estimator + integer projection + multi-output on small golden-like problems
N, K = 6, 6
groups = orc.enumerate_groups(N, K); L = sum(len(g) for g in groups)
C = orc.wishart_cov(N, 3)
w = blu.group_costs(groups, 2.0 ** (N - np.arange(N)))
sap = blu.SAP(C, K, [[list(g) for g in gk] for gk in groups], w, verbose=False)
sol = np.zeros(L); sol[[0, 3, 7, 20, 33, 40, 62]] = [2.6, 11.2, 4.7, 9.1, 1.3, 30.5, 6.6]
ip = sap.integer_projection(sol, budget=float(w @ np.round(sol)) * 1.01)
print("integer projection", ip[ip > 0], flush=True)
samples = np.round(sol).astype(int) + 1
sums = [samples[i] * (1.0 + 0.05 * np.arange(len(g))) for i, g in enumerate(g for gk in groups for g in gk)]
print("estimator", sap.compute_BLUE_estimator(sums, samples=samples), flush=True)
mos = blu.MOSAP([C, orc.wishart_cov(N, 4)], K, [K, K], [[list(g) for g in gk] for gk in groups], [[[list(g) for g in gk] for gk in groups]] * 2, w, [w, w], verbose=False)
print("mosap variances", mos.variances(1.0 + sol), flush=True)
Y = np.random.RandomState(0).standard_normal((5000, 7))
s1, S2, Ch = blu.pilot_covariance(Y)
check(S2, Y.T @ Y, 1e-12, "pilot Gram")
# Level 1
k = 3; gk = np.array(groups[k - 1], dtype=np.int64); Lk = len(gk)
inv = np.asarray(sap.invcovs[k - 1]); x = np.random.RandomState(1).rand(N)
g1 = np.zeros(Lk); blu.cmisc.gradK_c(g1, k, Lk, gk.ravel(), inv, x)
print("level-1 gradK_c", float(np.abs(g1).max()), flush=True)
print("sanitize smoke done")
