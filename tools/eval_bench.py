"""tools/eval_bench.py -- time one evaluation mode at a given size (device-resident m).
usage: python tools/eval_bench.py N [hess|nohess|var] [steps] [K]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import bluest_b200 as blu, oracle as orc

N = int(sys.argv[1]); mode = sys.argv[2] if len(sys.argv) > 2 else "nohess"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
K = int(sys.argv[4]) if len(sys.argv) > 4 else N
C = orc.wishart_cov(N, 0)
groups = blu.enumerate_groups(N, K)
L = sum(len(g) for g in groups)
t0 = time.perf_counter()
sap = blu.SAP(C, K, groups, np.ones(L), verbose=False)
if os.environ.get("BLU_PHI_STAGES"):
    sap.set_option("phi_stages", int(os.environ["BLU_PHI_STAGES"]))
sap.sync()
print("N=%d K=%d L=%d setup %.3f s (n_fallback=%d)" % (N, K, L, time.perf_counter() - t0, sap.n_fallback))
m = torch.from_numpy(orc.dense_m(L, 0)).cuda()
want = dict(hess=(True, True), nohess=(True, False), var=(False, False))[mode]
for _ in range(3):
    sap.eval_device(m, 0.0, grad=want[0], hess=want[1])
sap.sync()
sap.timing_log(steps)
t0 = time.perf_counter()
for _ in range(steps):
    sap.eval_device(m, 0.0, grad=want[0], hess=want[1])
sap.sync()
wall = (time.perf_counter() - t0) / steps
ph = sap.timing_read()
S_inv = sum(len(g) * (k + 1) ** 2 for k, g in enumerate(groups))
bytes_nohess = 16.0 * S_inv + 24.0 * L
print("mode=%s  wall/eval %.1f us; phases ms (median): phi+pinv %.4f grad %.4f hess %.4f total %.4f" %
      (mode, wall * 1e6, *np.median(ph, axis=0)))
tot = np.median(ph[:, 3]) * 1e-3
if mode == "nohess":
    print("algorithmic bytes (SURVEY 8d, nohess) %.3f MB -> %.1f GB/s = %.3f of 6555.8" % (bytes_nohess / 1e6, bytes_nohess / tot / 1e9, bytes_nohess / tot / 1e9 / 6555.8))
print("launches per eval", sap.last_launches(), "var", sap.last_result())
