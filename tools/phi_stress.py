"""tools/phi_stress.py -- randomized parity of Phi(m) against the oracle: dense, sparse and clustered-zero sample vectors,
shuffled group orders (the run-length accumulation of blu_phi.cuh must not depend on the enumeration order)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import bluest_b200 as blu, oracle as orc
worst = 0.0
for N, K, shuffle in ((10, 10, False), (12, 5, False), (9, 9, True), (14, 14, False), (24, 2, False), (8, 3, True)):
    C = orc.wishart_cov(N, N)
    groups = orc.enumerate_groups(N, K)
    rng = np.random.RandomState(N * 7 + K)
    if shuffle:
        groups = [[gk[i] for i in rng.permutation(len(gk))] for gk in groups]
    L = sum(len(g) for g in groups)
    o = orc.SapOracle(C, K, groups)
    sap = blu.SAP(C, K, [[list(g) for g in gk] for gk in groups], np.ones(L), verbose=False)
    if os.environ.get("BLU_PHI_STAGES"):
        sap.set_option("phi_stages", int(os.environ["BLU_PHI_STAGES"]))
    for trial in range(12):
        m = rng.rand(L) * 10 ** rng.uniform(-3, 3, L)
        frac = (0.0, 0.5, 0.9, 0.99)[trial % 4]
        m[rng.rand(L) < frac] = 0.0
        if trial >= 8:                                   # clustered zeros: whole runs of consecutive groups unsampled
            for _ in range(20):
                a = rng.randint(L); m[a:a + rng.randint(1, 70)] = 0.0
        if trial == 11:
            m[:] = 0.0; m[rng.randint(L)] = 3.0
        p, q = sap.get_phi(m), o.get_phi(m)
        err = float(np.max(np.abs(p - q)) / max(np.max(np.abs(q)), 1e-300))
        worst = max(worst, err)
        assert err < 1e-12, (N, K, shuffle, trial, err)
    sap.close()
print("phi stress: worst relative error %.2e" % worst)
