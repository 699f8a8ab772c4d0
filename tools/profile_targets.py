"""One short run that launches every kernel family once or twice -- the command line profiled with ncu
(see profiles/README.md): group inversion at 15 models, Gram 1e6 x 20 (plain + telescoped), the 20-model Phi / gradient
streams, the KKT solve at 15 models, the batched small-problem kernel."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch

import bluest_b200 as blu
import oracle as orc


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "n15"):
        section_n15()
    if what in ("all", "gram"):
        section_gram()
    if what in ("all", "n20"):
        section_n20()
    if what in ("all", "batch"):
        section_batch()
    if what in ("kkt_hv",):
        section_kkt_hv()
    print("profile_targets: done")


def section_n15():
    # group inversion, 15 models (one launch per size class) + one dense-Hessian evaluation
    N = 15
    ga = blu.enumerate_group_arrays(N)
    L = sum(len(g) for g in ga)
    costs = blu.group_costs(ga, 2.0 ** (N - np.arange(N))); costs = costs / costs.max()
    sap = blu.SAP(orc.wishart_cov(N, 0), N, ga, costs, verbose=False)
    m = torch.from_numpy(orc.dense_m(L, 0)).cuda()
    for _ in range(2):
        sap.eval_device(m, 0.0, grad=True, hess=True)
    sap.sync()
    # KKT solve
    Gx, scales, has_t = sap.sdp_linear_rows(budget_mode=True)
    rng = np.random.RandomState(0)
    n, nlin, M = L + 1, Gx.shape[0], N + 1
    Z = rng.randn(M, M); Z = Z + Z.T
    sap.kkt_solve(has_t, scales, Gx, 0.5 + rng.rand(n + nlin), np.eye(M) + 0.1 * rng.randn(M, M), rng.randn(n),
                  np.concatenate([rng.randn(n + nlin), Z.ravel()]))
    sap.close()


def section_kkt_hv():
    # KKT solve at 15 models (three solves: the third is the one captured) and the Hessian-operator product at 20 models
    N = 15
    ga = blu.enumerate_group_arrays(N)
    L = sum(len(g) for g in ga)
    costs = blu.group_costs(ga, 2.0 ** (N - np.arange(N))); costs = costs / costs.max()
    sap = blu.SAP(orc.wishart_cov(N, 0), N, ga, costs, verbose=False)
    Gx, scales, has_t = sap.sdp_linear_rows(budget_mode=True)
    rng = np.random.RandomState(0)
    n, nlin, M = L + 1, Gx.shape[0], N + 1
    Z = rng.randn(M, M); Z = Z + Z.T
    for _ in range(3):
        sap.kkt_solve(has_t, scales, Gx, 0.5 + rng.rand(n + nlin), np.eye(M) + 0.1 * rng.randn(M, M), rng.randn(n),
                      np.concatenate([rng.randn(n + nlin), Z.ravel()]))
    sap.close()
    N = 20
    ga = blu.enumerate_group_arrays(N)
    L = sum(len(g) for g in ga)
    sap = blu.SAP(orc.wishart_cov(N, 0), N, ga, np.ones(L), verbose=False)
    v, g, op = sap.variance_GH_operator(orc.dense_m(L, 0))
    p = torch.randn(L, dtype=torch.float64, device="cuda"); out = torch.empty_like(p)
    for _ in range(4):
        sap.hess_matvec_device(p, out)
    sap.sync()
    sap.close()


def section_gram():
    # Gram 1e6 x 20
    Y = torch.randn((10 ** 6, 20), dtype=torch.float64, device="cuda")
    for tele in (False, True):
        for _ in range(2):
            blu.pilot_sums(Y, telescoped=tele)
    del Y


def section_n20():
    # 20 models: Phi + gradient streams
    N = 20
    ga = blu.enumerate_group_arrays(N)
    L = sum(len(g) for g in ga)
    sap = blu.SAP(orc.wishart_cov(N, 0), N, ga, np.ones(L), verbose=False)
    m = torch.from_numpy(orc.dense_m(L, 0)).cuda()
    for _ in range(3):
        sap.eval_device(m, 0.0, grad=True, hess=False)
    sap.sync()
    sap.close()


def section_batch():
    # batched small problems: 64 sample vectors of a 10-model problem
    N = 10
    g10 = blu.enumerate_groups(N); L = sum(len(g) for g in g10)
    sap = blu.SAP(orc.wishart_cov(N, 1), N, [[list(g) for g in gk] for gk in g10], np.ones(L), verbose=False)
    M64 = np.array([orc.dense_m(L, j) for j in range(64)])
    for _ in range(2):
        blu.evaluate_many(sap, M64)
    sap.close()


if __name__ == "__main__":
    main()
