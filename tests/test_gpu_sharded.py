"""Group-sharded kernels on ONE GPU: two contexts own the two halves of the group enumeration and
the exchange steps are emulated with torch copies (no spin-waits between kernels; the real NCCL
choreography is covered by tests/test_dist_gloo.py on CPU and by bench.py --mode shard)."""
import numpy as np
import pytest

import oracle as orc
from conftest import maxrel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(9, 9), (12, 6)])
def test_two_slices_reproduce_the_full_evaluation(N, K):
    import torch
    import bluest_b200 as blu
    from bluest_b200 import _lib
    from bluest_b200.dist import GpuEngine
    if blu.device_count() <= 0:
        pytest.fail("no CUDA device")
    C = orc.wishart_cov(N, 6)
    groups = orc.enumerate_groups(N, K)
    L = sum(len(g) for g in groups)
    o = orc.SapOracle(C, K, groups)
    sizes = o.sizes[1:]
    slices = blu.balanced_slices(sizes, 2)
    engines = []
    for lo, hi in slices:
        sap = blu.SAP(C, K, [[list(g) for g in gk] for gk in groups], np.ones(L), verbose=False)
        e = GpuEngine(sap)
        e.set_slice(lo, hi)
        engines.append(e)
    for m in (orc.dense_m(L, 3), orc.sparse_m(L, N, 3)):
        full = len(o.support(m)) == N
        tol = 1e-12 if full else 1e-9
        bufs = [e.shard_phi(m) for e in engines]
        for e in engines:
            e.sap.sync()
        total = bufs[0].clone() + bufs[1].clone()              # the all-reduce
        for b in bufs:
            b.copy_(total)
        torch.cuda.synchronize()
        for e in engines:
            e.shard_finish(0.0, True, True)
            e.sap.sync()
        NP = engines[0].NP
        # the all-gather of U, V rows and of the gradient slices
        for name in ("u_buffer", "v_buffer"):
            t0 = getattr(engines[0], name)().view(-1, NP)
            t1 = getattr(engines[1], name)().view(-1, NP)
            (lo0, hi0), (lo1, hi1) = slices
            t0[lo1:hi1] = t1[lo1:hi1]
            t1[lo0:hi0] = t0[lo0:hi0]
        g = torch.cat([engines[0].grad_buffer()[slices[0][0]:slices[0][1]], engines[1].grad_buffer()[slices[1][0]:slices[1][1]]]).cpu().numpy()
        torch.cuda.synchronize()
        panels = []
        rows = [(0, L // 2), (L // 2, L)]                 # row panels are independent of the group slices
        for e, (lo, hi) in zip(engines, rows):
            e.shard_hess(lo, hi)
            e.sap.sync()
            ld = 16 * ((L + 15) // 16)
            panels.append(e.hess_panel()[: (hi - lo) * ld].view(hi - lo, ld)[:, :L].cpu().numpy())
        H = np.vstack(panels)
        vo, go, Ho = o.variance_GH(m, hess_mode="factored")
        for e in engines:
            v, fl = e.result()
            assert abs(v - vo) <= 1e-12 * vo
        assert maxrel(g, go) < tol
        assert maxrel(H, Ho) < tol
    for e in engines:
        e.sap.close()


def test_sliced_soa_gradient_large_class_boundaries():
    """Group slices on a problem large enough for the lane-per-group (SoA tile) kernels: slice
    boundaries fall inside 32-group tiles, the lanes outside the slice must stay silent."""
    import torch
    import bluest_b200 as blu
    from bluest_b200 import _lib
    from bluest_b200.dist import GpuEngine
    N = 14
    C = orc.wishart_cov(N, 7)
    groups = orc.enumerate_groups(N)
    L = sum(len(g) for g in groups)
    o = orc.SapOracle(C, N, groups)
    m = orc.dense_m(L, 4)
    vo, go, _ = o.variance_GH(m, nohess=True)
    slices = [(0, 5001), (5001, 5003), (5003, L)]            # odd boundaries, a 2-group slice
    pieces = []
    for lo, hi in slices:
        sap = blu.SAP(C, N, [[list(g) for g in gk] for gk in groups], np.ones(L), verbose=False)
        e = GpuEngine(sap)
        e.set_slice(lo, hi)
        sap.device_buffer(_lib.BUF_GRAD).fill_(123.0)        # sentinel: nothing outside the slice may be written
        # full Phi on this context first (slice cleared), then the sliced gradient
        e.set_slice(0, L)
        buf = e.shard_phi(m)
        sap.sync()
        e.set_slice(lo, hi)
        e.shard_finish(0.0, True, False)
        sap.sync()
        g = sap.device_buffer(_lib.BUF_GRAD).cpu().numpy()
        assert np.all(g[:lo] == 123.0) and np.all(g[hi:] == 123.0)
        pieces.append(g[lo:hi])
        v, _ = e.result()
        assert abs(v - vo) <= 1e-12 * vo
        sap.close()
    assert maxrel(np.concatenate(pieces), go) < 1e-12


@pytest.mark.parametrize("N,K,cuts", [(9, 9, (0.5,)), (12, 6, (0.3, 0.31)), (14, 14, (0.25, 0.8))])
def test_sliced_hessian_operator(N, K, cuts):
    """Sharded Hessian operator: every slice owner forms its partial t = sum p_i u_i, the partials are
    summed (the 32-double all-reduce), every owner applies V to its own rows.  The factors are never
    gathered.  Checked against the oracle's dense Hessian times p."""
    import torch
    import bluest_b200 as blu
    from bluest_b200.dist import GpuEngine
    C = orc.wishart_cov(N, 8)
    groups = orc.enumerate_groups(N, K)
    L = sum(len(g) for g in groups)
    o = orc.SapOracle(C, K, groups)
    m = orc.dense_m(L, 5)
    P = np.linalg.pinv(o.get_phi(m))
    U = o.ufactor(np.ascontiguousarray(P[0]))
    p = np.random.RandomState(9).randn(L)
    ref = 2.0 * (U.T @ (P @ (U @ p)))
    bounds = [0] + [int(c * L) for c in cuts] + [L]
    slices = list(zip(bounds[:-1], bounds[1:]))
    engines = []
    for lo, hi in slices:
        sap = blu.SAP(C, K, [[list(g) for g in gk] for gk in groups], np.ones(L), verbose=False)
        e = GpuEngine(sap)
        buf = e.shard_phi(m)                                  # full Phi on every context (slice not yet set)
        sap.sync()
        e.set_slice(lo, hi)
        e.shard_finish(0.0, True, 2)                          # gradient and U rows of the slice only (2: no V)
        sap.sync()
        engines.append(e)
    pd = torch.from_numpy(p).cuda()
    ts = []
    for e in engines:
        with e.stream_context():
            ts.append(e.hv_partial(pd).clone())
        e.sap.sync()
    total = torch.stack(ts).sum(0)                            # the all-reduce
    assert float(total[4 * ((N + 3) // 4):].abs().max()) == 0.0 if 4 * ((N + 3) // 4) < 32 else True
    out = np.empty(L)
    for e, (lo, hi) in zip(engines, slices):
        with e.stream_context():
            e._hv_t.copy_(total)
            r = e.hv_apply(e._hv_t)
        e.sap.sync()
        out[lo:hi] = r[lo:hi].cpu().numpy()
    assert maxrel(out, ref) < 1e-12
    # numpy p takes the staging path
    with engines[0].stream_context():
        t2 = engines[0].hv_partial(p).clone()
    engines[0].sap.sync()
    assert torch.equal(t2, ts[0])
    for e in engines:
        e.sap.close()
