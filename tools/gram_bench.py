"""tools/gram_bench.py -- pilot-covariance Gram kernel (kernel 4) at BASELINE config 5:
1e6 samples x 20 models, device-resident Y.  Prints kernel time, GB/s vs the measured HBM peak
and the parity against the oracle on a 20000-sample prefix."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import bluest_b200 as blu, oracle as orc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10 ** 6
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
C = orc.wishart_cov(N, 0)
rng = np.random.default_rng(0)
Y = rng.standard_normal((n, N)) @ np.linalg.cholesky(C).T
Yd = torch.from_numpy(Y).cuda()
best = 1e9
for _ in range(8):
    s1, S2, Ch, ms = blu.pilot_covariance(Yd, return_ms=True)
    best = min(best, ms)
print("Gram n=%d N=%d: kernel %.1f us -> %.1f GB/s (%.3f of 6555.8); host call incl. alloc/sync ~%.1f us" %
      (n, N, best * 1e3, 8.0 * n * N / (best * 1e-3) / 1e9, 8.0 * n * N / (best * 1e-3) / 1e9 / 6555.8, 0.0))
o1, o2, oc = orc.pilot_covariance(Y[:20000])
g1, g2, gc = blu.pilot_covariance(Y[:20000])
print("parity (20000-sample prefix): s1 %.2e S2 %.2e C_hat %.2e" % (np.abs(g1 - o1).max() / np.abs(o1).max(), np.abs(g2 - o2).max() / np.abs(o2).max(), np.abs(gc - oc).max() / np.abs(oc).max()))
ref = Y.T @ Y
print("full-size vs numpy Y^T Y: %.2e" % (np.abs(S2 - ref).max() / np.abs(ref).max()))
