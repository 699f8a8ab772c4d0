"""A/B of the symmetric Hessian kernel's staging variants on one box: alternating runs, CUDA events."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import bluest_b200 as blu, oracle as orc
N = int(sys.argv[1]) if len(sys.argv) > 1 else 15
groups = blu.enumerate_groups(N)
L = sum(len(g) for g in groups)
sap = blu.SAP(orc.wishart_cov(N, 0), N, groups, np.ones(L), verbose=False)
m = torch.from_numpy(orc.dense_m(L, 0)).cuda()
for rep in range(3):
    for one in (0, 1):
        sap.set_option("hess_onebuf", one)
        for _ in range(5): sap.eval_device(m, 0.0, grad=True, hess=True)
        sap.timing_log(100)
        for _ in range(100): sap.eval_device(m, 0.0, grad=True, hess=True)
        ph = sap.timing_read()
        print("onebuf=%d: Hessian kernel %.4f ms mean, %.4f median (%.0f GB/s)" % (one, ph[:, 2].mean(), np.median(ph[:, 2]), 8.0 * L * L / np.median(ph[:, 2]) / 1e6), flush=True)
