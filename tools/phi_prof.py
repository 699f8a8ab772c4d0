"""tools/phi_prof.py -- per-warp time split of the Phi kernel (library built with BLU_NVCC_EXTRA=-DBLU_PHI_PROFILE).
usage: python tools/phi_prof.py N [K]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import bluest_b200 as blu, oracle as orc
from bluest_b200 import _lib
N = int(sys.argv[1]); K = int(sys.argv[2]) if len(sys.argv) > 2 else N
C = orc.wishart_cov(N, 0)
groups = blu.enumerate_groups(N, K)
L = sum(len(g) for g in groups)
sap = blu.SAP(C, K, groups, np.ones(L), verbose=False)
m = torch.from_numpy(orc.dense_m(L, 0)).cuda()
lib = _lib.lib()
st = (ctypes.c_uint64 * 16)()
def stamps():
    _lib.check(lib.blu_ctx_last_stamps(sap._ctx, st)); return np.array(list(st), dtype=np.int64)
for _ in range(3): sap.eval_device(m, 0.0, grad=False, hess=False)
sap.sync(); a = stamps()
sap.eval_device(m, 0.0, grad=False, hess=False); sap.sync(); b = stamps()
d = b - a
w = max(d[12], 1)
print("per warp: prefetch issue %.0f  expand/eff %.0f  barrier wait %.0f  group loop %.0f  (slowest warp total %d)" % (d[13] / w, d[14] / w, d[10] / w, d[15] / w, b[8]))
print("warps %d  mean cycles per warp: total %.0f, waiting on the copy barrier %.0f (%.1f %%)" % (d[12], d[11] / max(d[12], 1), d[10] / max(d[12], 1), 100.0 * d[10] / max(d[11], 1)))
