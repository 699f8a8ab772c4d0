"""The toy MLBLUE problem of the reference's tutorials/01_tutorial.ipynb (BASELINE config 1) on the B200:
five synthetic models, pilot covariance on the device, all 31 groups, allocation for a target RMSE with the
scipy driver, integer projection and the BLUE estimate.  Needs a B200 (there is no CPU fallback):

    python examples/tutorial_sap.py

The known answers of the notebook (two stored allocations -> errors 0.01592249 / 0.01592225, costs
91120 / 241606) are pinned by tests/test_gpu_parity.py::test_tutorial_known_answers.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bluest_b200 as blu  # noqa: E402

# covariance of the five models of the tutorial (tests/golden/make_golden.py: make_solve)
C_TUTORIAL = np.array([[2.53500581, 2.3583669, 2.36312599, 1.66331444, 0.72897923],
                       [2.3583669, 2.27345322, 2.07702441, 1.67029559, 0.80520217],
                       [2.36312599, 2.07702441, 2.40127866, 1.37242096, 0.47072499],
                       [1.66331444, 1.67029559, 1.37242096, 1.29027193, 0.67275196],
                       [0.72897923, 0.80520217, 0.47072499, 0.67275196, 1.56637342]])


def main(n_pilot=20000, seed=0):
    N = C_TUTORIAL.shape[0]
    model_costs = 2.0 ** (N - np.arange(N))                               # [32 16 8 4 2]
    rng = np.random.default_rng(seed)
    chol = np.linalg.cholesky(C_TUTORIAL)
    mean = np.linspace(1.0, 0.2, N)

    # 1. pilot samples of all models -> covariance estimate (FP64 Gram contraction on the device)
    Y = mean + rng.standard_normal((n_pilot, N)) @ chol.T
    s1, S2, C_hat = blu.pilot_covariance(Y)
    print("pilot covariance: max |C_hat - C| = %.3e" % np.abs(C_hat - C_TUTORIAL).max())

    # 2. all model groups, their costs, the sample-allocation problem on the device
    groups = blu.enumerate_groups(N)
    costs = blu.group_costs(groups, model_costs)
    sap = blu.SAP(C_hat, N, groups, costs, verbose=False)

    # 3. allocation for a target RMSE: continuous solve (trust-constr on the GPU closures) + integer projection
    eps = 0.01 * np.sqrt(C_hat[0, 0])
    x0 = np.ceil(eps ** -2 * C_hat[0, 0] * np.ones(sap.L) / sap.L)        # feasible start: plenty of samples everywhere
    samples = sap.solve(eps=eps, solver="scipy", x0=x0, hess="operator", sparse_constraints=True)
    used = np.flatnonzero(samples)
    flat = [g for gk in groups for g in gk]
    print("allocation (groups with samples):")
    for i in used:
        print("   models %-18s %6d samples" % ([int(v) for v in flat[i]], samples[i]))
    print("total cost %.0f, RMSE %.6f (target %.6f)" % (sap.tot_cost, np.sqrt(sap.variance(samples)), eps))

    # 4. sample the groups and form the BLUE estimate of E[model 0]
    sums = []
    for i, g in enumerate(flat):
        g = list(g)
        if samples[i] > 0:
            draws = mean[g] + rng.standard_normal((int(samples[i]), N))[:, :] @ chol.T[:, g]
            sums.append(draws.sum(axis=0))
        else:
            sums.append(np.zeros(len(g)))
    mu, var = sap.compute_BLUE_estimator(sums, samples=samples)
    print("BLUE estimate %.5f +- %.5f   (exact mean %.5f)" % (mu, np.sqrt(var), mean[0]))
    sap.close()


if __name__ == "__main__":
    main()
