"""KKT solve (scope-table row f1) at N models, all groups: device time of repeated solves (the first call pays the lazy
module load of its kernels).  Under `ncu --metrics gpu__time_duration.sum -k regex:blu_kkt` it gives the per-kernel split."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import bluest_b200 as blu, oracle as orc
for N in [int(a) for a in sys.argv[1:]] or [15]:
    C = orc.wishart_cov(N, 0)
    ga = blu.enumerate_group_arrays(N)
    L = sum(len(g) for g in ga)
    costs = blu.group_costs(ga, 2.0 ** (N - np.arange(N))); costs = costs / costs.max()
    sap = blu.SAP(C, N, ga, costs, verbose=False)
    M = N + 1
    Gx, scales, has_t = sap.sdp_linear_rows(budget_mode=True)
    n, nlin = L + 1, Gx.shape[0]
    rng = np.random.RandomState(9)
    d = 0.1 + 10.0 * rng.rand(n + nlin)
    r = np.eye(M) + 0.2 * rng.randn(M, M)
    bx = rng.randn(n)
    Z = rng.randn(M, M); Z = Z + Z.T
    bz = np.concatenate([rng.randn(n + nlin), Z.ravel()])
    ms, wall = [], []
    for it in range(6):
        t0 = time.perf_counter()
        ux, uz, t = sap.kkt_solve(has_t, scales, Gx, d, r, bx, bz, return_ms=True)
        wall.append((time.perf_counter() - t0) * 1e3)
        ms.append(t)
    print("N=%d L=%d Q=%d: device ms per solve %s | through the host API %s" % (N, L, M * (M + 1) // 2 + nlin, " ".join("%.3f" % t for t in ms),
                                                                              " ".join("%.2f" % t for t in wall)), flush=True)
    sap.close()
