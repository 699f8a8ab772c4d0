"""End-to-end SAP.variance_GH timing (host m -> host var, grad, dense Hessian), plain vs symmetric download."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bluest_b200 as blu
import oracle as orc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 15
groups = blu.enumerate_groups(N)
L = sum(len(g) for g in groups)
sap = blu.SAP(orc.wishart_cov(N, 0), N, groups, np.ones(L), verbose=False)
m = orc.dense_m(L, 0)
for sym in (0, 1, 0, 1):
    sap.set_option("sym_download", sym)
    for _ in range(2):
        v, g, H = sap.variance_GH(m); del H
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        v, g, H = sap.variance_GH(m)
        if _ < n - 1: del H
    dt = (time.perf_counter() - t0) / n
    ok = bool(np.array_equal(H[:2048, -2048:], H[-2048:, :2048].T))
    print("sym_download=%d: %.1f ms per evaluation (%.2f evals/s) symmetric=%s" % (sym, dt * 1e3, 1 / dt, ok))
    del H

# share of the lower triangle that travels by DMA (bottom rows in full) instead of being mirrored by the host
sap.set_option("sym_download", 1)
for pct in (0, 10, 19, 25, 30, 40, 50):
    sap.set_option("sym_full_rows_pct", pct)
    for _ in range(2):
        v, g, H = sap.variance_GH(m); del H
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        v, g, H = sap.variance_GH(m)
        if _ < n - 1: del H
    dt = (time.perf_counter() - t0) / n
    ok = True
    for r0 in range(0, L, 4096):
        ok = ok and bool(np.array_equal(H[r0:r0 + 4096], H[:, r0:r0 + 4096].T))
    print("sym_full_rows_pct=%d: %.1f ms per evaluation (%.2f evals/s) symmetric=%s" % (pct, dt * 1e3, 1 / dt, ok))
    del H
