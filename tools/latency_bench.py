"""tools/latency_bench.py -- host-call latency of the SAP closures on a small problem."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import bluest_b200 as blu, oracle as orc
for N, K in [(10, 3), (10, 10), (5, 5)]:
    C = orc.wishart_cov(N, 0)
    groups = blu.enumerate_groups(N, K)
    L = sum(len(g) for g in groups)
    sap = blu.SAP(C, K, groups, np.ones(L), verbose=False)
    m = orc.dense_m(L, 0)
    for name, fn in (("get_phi", lambda: sap.get_phi(m)), ("variance", lambda: sap.variance(m)),
                     ("variance_GH nohess", lambda: sap.variance_GH(m, nohess=True)), ("variance_GH hess", lambda: sap.variance_GH(m))):
        for _ in range(5): fn()
        t0 = time.perf_counter()
        for _ in range(200): fn()
        print("N=%d K=%d L=%d %-20s %.1f us/call" % (N, K, L, name, (time.perf_counter() - t0) / 200 * 1e6))
    sap.close()
# multi-output: 4 outputs x N=10 (BASELINE config 4), outputs evaluated concurrently
N, No = 10, 4
groups = blu.enumerate_groups(N)
L = sum(len(g) for g in groups)
Cs = [orc.wishart_cov(N, 10 + n) for n in range(No)]
mos = blu.MOSAP(Cs, N, [N] * No, [[list(g) for g in gk] for gk in groups], [[[list(g) for g in gk] for gk in groups] for _ in range(No)],
                np.ones(L), [np.ones(L)] * No, verbose=False)
m = orc.dense_m(L, 0)
for name, fn in (("MOSAP.variances", lambda: mos.variances(m)), ("MOSAP.variance_GH nohess", lambda: mos.variance_GH(m, nohess=True)),
                 ("MOSAP.variance_GH hess", lambda: mos.variance_GH(m))):
    for _ in range(3): fn()
    t0 = time.perf_counter()
    for _ in range(50): fn()
    print("MOSAP 4 x N=10 L=%d %-26s %.1f us/call" % (L, name, (time.perf_counter() - t0) / 50 * 1e6))
