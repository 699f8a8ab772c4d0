"""Small-problem latency through the host API: MOSAP 4 outputs x 10 models (BASELINE config 4) and a 64-instance
sweep of one 10-model problem (config 2 x sweep), batched kernel vs the per-context path."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import bluest_b200 as blu
import oracle as orc


def main():
    N, No = 10, 4
    groups = blu.enumerate_groups(N); L = sum(len(g) for g in groups)
    cp = lambda gs: [[list(g) for g in gk] for gk in gs]
    mos = blu.MOSAP([orc.wishart_cov(N, 10 + n) for n in range(No)], N, [N] * No, cp(groups), [cp(groups) for _ in range(No)], np.ones(L),
                    [np.ones(L)] * No, verbose=False)
    mh = orc.dense_m(L, 0)
    out = {}
    for name, fn in (("variances", lambda: mos.variances(mh)), ("variance_GH_nohess", lambda: mos.variance_GH(mh, nohess=True))):
        for _ in range(5):
            fn()
        t0 = time.perf_counter()
        for _ in range(200):
            fn()
        out["mosap_4x10_" + name + "_us"] = (time.perf_counter() - t0) / 200 * 1e6
    sap = mos.SAPS[0]
    M = np.array([orc.dense_m(L, j) for j in range(64)])
    for grad in (True, False):
        for _ in range(5):
            blu.evaluate_many(sap, M, grad=grad)
        t0 = time.perf_counter()
        for _ in range(100):
            blu.evaluate_many(sap, M, grad=grad)
        t = (time.perf_counter() - t0) / 100
        out["sweep64_n10_%s_us_per_batch" % ("variance_grad" if grad else "variance")] = t * 1e6
    t0 = time.perf_counter()
    for j in range(64):
        sap.variance_GH(M[j], nohess=True)
    out["sweep64_n10_one_by_one_us_per_batch"] = (time.perf_counter() - t0) * 1e6
    print(out)
    return out


if __name__ == "__main__":
    main()
