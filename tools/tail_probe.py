"""tools/tail_probe.py -- %globaltimer milestones of the last CTA of the fused Phi kernel (one GPU, variance-only evaluation).
usage: python tools/tail_probe.py N [K]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import bluest_b200 as blu, oracle as orc
from bluest_b200 import _lib
N = int(sys.argv[1]); K = int(sys.argv[2]) if len(sys.argv) > 2 else N
groups = blu.enumerate_groups(N, K)
L = sum(len(g) for g in groups)
sap = blu.SAP(orc.wishart_cov(N, 0), K, groups, np.ones(L), verbose=False)
m = torch.from_numpy(orc.dense_m(L, 0)).cuda()
lib = _lib.lib(); st = (ctypes.c_uint64 * 16)()
rows = []
for it in range(12):
    sap.eval_device(m, 0.0, grad=False, hess=False); sap.sync()
    _lib.check(lib.blu_ctx_last_stamps(sap._ctx, st))
    a = np.array(list(st), dtype=np.int64)
    if it >= 2:
        rows.append([(a[i] - a[1]) * 1e-3 for i in (2, 3, 4, 6, 10, 7, 9)])
r = np.median(np.array(rows), axis=0)
print("N=%d L=%d  us after the last CTA finished its stream: group fold starts %.2f, final fold starts %.2f, sums complete %.2f, "
      "Phi mirrored + support %.2f, block inverted %.2f, pseudo-inverse written %.2f, done %.2f" % (N, L, *r))
print("sweeps (0 = Gauss-Jordan, > 0 = Jacobi fallback):", float(sap.device_buffer(_lib.BUF_SCAL)[2]))
