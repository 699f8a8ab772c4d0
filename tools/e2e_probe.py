"""tools/e2e_probe.py -- one dense-Hessian evaluation end to end with the download's own timing prints (BLU_DEBUG_TIMING)."""
import os, sys, time
os.environ["BLU_DEBUG_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import bluest_b200 as blu, oracle as orc
N = 15
groups = blu.enumerate_groups(N)
L = sum(len(g) for g in groups)
sap = blu.SAP(orc.wishart_cov(N, 0), N, groups, np.ones(L), verbose=False)
m = orc.dense_m(L, 0)
for opt in sys.argv[1:]:
    k, v = opt.split("="); sap.set_option(k, int(v))
for it in range(int(os.environ.get("E2E_EVALS", "4"))):
    t0 = time.perf_counter()
    v, g, H = sap.variance_GH(m)
    print("evaluation %d: %.1f ms" % (it, (time.perf_counter() - t0) * 1e3), flush=True)
    del H
print("host cores", os.cpu_count())
