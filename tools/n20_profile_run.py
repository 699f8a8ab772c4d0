"""Short N=20 run for ncu: evaluation without Hessian (Phi partial, fold, finish, lane-per-group gradient),
factored evaluation (gradient + U factor) and Hessian-operator products."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, ctypes
import bluest_b200 as blu, oracle as orc
from bluest_b200 import _lib
N = 20
groups = blu.enumerate_groups(N)
L = sum(len(g) for g in groups)
sap = blu.SAP(orc.wishart_cov(N, 0), N, groups, np.ones(L), verbose=False)
m = torch.from_numpy(orc.dense_m(L, 0)).cuda()
p = torch.randn(L, dtype=torch.float64, device="cuda"); out = torch.empty_like(p)
for _ in range(3):
    sap.eval_device(m, 0.0, grad=True, hess=False)
    _lib.check(_lib.lib().blu_eval_device(sap._ctx, ctypes.c_void_p(int(m.data_ptr())), 0.0, 1, 3))
    sap.hess_matvec_device(p, out)
sap.sync()
print("ok", sap.last_result())
