"""Hessian-operator product H p = U (S (U^T p)) with p, H p resident in HBM: time per product."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import bluest_b200 as blu, oracle as orc
for N in [int(a) for a in sys.argv[1:]] or [15, 20]:
    groups = blu.enumerate_groups(N)
    L = sum(len(g) for g in groups)
    sap = blu.SAP(orc.wishart_cov(N, 0), N, groups, np.ones(L), verbose=False)
    v, g, op = sap.variance_GH_operator(orc.dense_m(L, 0))
    p = torch.randn(L, dtype=torch.float64, device="cuda"); out = torch.empty_like(p)
    ext = torch.cuda.ExternalStream(sap.stream())
    for _ in range(5): sap.hess_matvec_device(p, out)
    sap.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(100): sap.hess_matvec_device(p, out)
    e1.record(ext); sap.sync()
    t = e0.elapsed_time(e1) / 100 * 1e-3
    NP = 4 * ((N + 3) // 4)
    ref = op @ p.cpu().numpy()
    print("N=%d L=%d: %.1f us per product, %.0f GB/s over the U factor (2 x %.0f MB); host-API product equals device product: %s"
          % (N, L, t * 1e6, (16.0 * NP * L + 16.0 * L) / t / 1e9, 8.0 * NP * L / 1e6, bool(np.array_equal(ref, out.cpu().numpy()))))
    sap.close()
