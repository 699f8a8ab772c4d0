"""MOSAP 4 outputs x 10 models: a few `variances` / `variance_GH(nohess)` calls.  Run with BLU_DEBUG_TIMING=1 to print the
milestones of CTA 0 of the batched kernel (staging, Phi, inverse, gradient) that DESIGN.md quotes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import bluest_b200 as blu, oracle as orc
N, No = 10, 4
groups = blu.enumerate_groups(N); L = sum(len(g) for g in groups)
cp = lambda gs: [[list(g) for g in gk] for gk in gs]
mos = blu.MOSAP([orc.wishart_cov(N, 10 + n) for n in range(No)], N, [N] * No, cp(groups), [cp(groups) for _ in range(No)], np.ones(L), [np.ones(L)] * No, verbose=False)
mh = orc.dense_m(L, 0)
for _ in range(4): mos.variances(mh)
for _ in range(4): mos.variance_GH(mh, nohess=True)
