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
if hasattr(lib, "blu_prof_warp_read"):
    nW = int(d[12])
    buf = (ctypes.c_longlong * (4 * nW))()
    lib.blu_prof_warp_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
    assert lib.blu_prof_warp_read(buf, nW) == 0
    a = np.array(list(buf), dtype=np.int64).reshape(nW, 4)
    tot = a[:, 0]
    print("per-warp total cycles: min %d  p10 %d  median %d  p90 %d  max %d" % (tot.min(), np.percentile(tot, 10), np.median(tot), np.percentile(tot, 90), tot.max()))
    cta = tot.reshape(-1, 8)
    print("per-CTA max: min %d median %d max %d;  spread inside a CTA (max-min) median %d" % (cta.max(1).min(), np.median(cta.max(1)), cta.max(1).max(), np.median(cta.max(1) - cta.min(1))))
    order = np.argsort(tot)
    for name, idx in (("slowest", order[-8:]), ("fastest", order[:8])):
        print(name, [(int(i), int(tot[i]), int(a[i, 1]), int(a[i, 2]), int(a[i, 3])) for i in idx])
    # correlation with position in the grid
    print("mean total by warp-in-CTA", cta.mean(0).astype(int).tolist())
    print("mean total by CTA octile", [int(x.mean()) for x in np.array_split(cta.mean(1), 8)])
