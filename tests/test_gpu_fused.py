"""Fused peer-memory all-reduce (blu_shard_eval_fused): world=1 on one GPU must reproduce the plain
evaluation bit for bit; with >= 2 GPUs two processes exchange their partial Phi over NVLink inside
the finish kernel and must agree with the oracle and with each other exactly."""
import os
import sys

import numpy as np
import pytest

import oracle as orc
from conftest import maxrel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_world1_equals_plain():
    import torch
    import bluest_b200 as blu
    from bluest_b200 import _lib
    from bluest_b200.dist import GpuEngine, ShardedEvaluator
    N = 11
    C = orc.wishart_cov(N, 2)
    groups = orc.enumerate_groups(N)
    L = sum(len(g) for g in groups)
    sizes = [len(g) for g in groups]
    sap = blu.SAP(C, N, [[list(g) for g in gk] for gk in groups], np.ones(L), verbose=False)
    ref = blu.SAP(C, N, [[list(g) for g in gk] for gk in groups], np.ones(L), verbose=False)
    ev = ShardedEvaluator(GpuEngine(sap), sizes, 0, 1, fused=True)
    for m in (orc.dense_m(L, 1), orc.sparse_m(L, N, 1), 0.01 * np.ones(L)):
        r = ev.evaluate(m, 0.0, grad=True, hess=False)
        res = ref.variance_GH(m, nohess=True)
        if len(res) == 2:
            assert r["flags"] & 1 and np.isinf(r["var"])
            continue
        assert r["var"] == res[0]
        g = sap.device_buffer(_lib.BUF_GRAD).cpu().numpy()
        assert np.array_equal(g, res[1])
    sap.close(); ref.close()


def _worker(rank, world, port, N, ret):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import bluest_b200 as blu
    from bluest_b200 import _lib
    from bluest_b200.dist import GpuEngine, ShardedEvaluator
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        C = orc.wishart_cov(N, 2)
        groups = orc.enumerate_groups(N)
        L = sum(len(g) for g in groups)
        sap = blu.SAP(C, N, [[list(g) for g in gk] for gk in groups], np.ones(L), verbose=False, device=rank)
        ev = ShardedEvaluator(GpuEngine(sap), [len(g) for g in groups], rank, world, dist=dist, fused=True)
        out = []
        for seed in range(3):
            m = orc.dense_m(L, seed) if seed < 2 else orc.sparse_m(L, N, seed)
            r = ev.evaluate(m, 0.0, grad=True, hess=True, gather_grad=True)
            ld = 16 * ((L + 15) // 16)
            H = sap.device_buffer(_lib.BUF_HESS)[: (r["rhi"] - r["rlo"]) * ld].view(-1, ld)[:, :L].cpu().numpy()
            out.append(dict(var=r["var"], grad=sap.device_buffer(_lib.BUF_GRAD).cpu().numpy(), H=H, rows=(r["rlo"], r["rhi"]),
                            phi=sap.device_buffer(_lib.BUF_PHI)[: N * N].cpu().numpy()))
        # sharded Hessian operator: factors stay on their rank, one 32-double NCCL all-reduce per product
        m = orc.dense_m(L, 0)
        r = ev.evaluate_factors(m)
        p = torch.from_numpy(np.random.RandomState(5).randn(L)).cuda(rank)
        hp = ev.hess_matvec(p, gather=True).cpu().numpy().copy()
        own = ev.hess_matvec(p, gather=False).cpu().numpy()[r["lo"]:r["hi"]].copy()
        out.append(dict(var=r["var"], hp=hp, own=own, lo=r["lo"], hi=r["hi"]))
        ret[rank] = out
        sap.close()
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_clone_lane_and_graph_replay():
    """A clone (second evaluation lane) shares the parent's inverses and must reproduce its results bit for bit,
    also when parent and clone evaluate concurrently on their own streams and when the evaluations are replayed
    from a CUDA graph (the device-side epoch of the fused exchange keeps counting across replays)."""
    import torch
    import bluest_b200 as blu
    from bluest_b200 import _lib
    from bluest_b200.dist import GpuEngine, ShardedEvaluator
    N = 12
    C = orc.wishart_cov(N, 4)
    ga = blu.enumerate_group_arrays(N)
    sizes = [len(g) for g in ga]
    L = sum(sizes)
    sap = blu.SAP(C, N, ga, np.ones(L), verbose=False)
    o = orc.SapOracle(C, N, orc.enumerate_group_arrays(N), invcovs=orc.batched_invcovs(C, orc.enumerate_group_arrays(N)), with_ES=False)
    ev = ShardedEvaluator(GpuEngine(sap), sizes, 0, 1, fused=True)
    lane2 = sap.clone()
    eng2 = GpuEngine(lane2)
    ShardedEvaluator(eng2, sizes, 0, 1, fused=True, set_slice=False)
    with pytest.raises(blu.BluError):
        _lib.check(_lib.lib().blu_ctx_set_slice(lane2._ctx, 0, 10))          # a clone keeps its parent's slice
    ms = [orc.dense_m(L, s) for s in range(4)]
    dms = [torch.from_numpy(m).cuda() for m in ms]
    var = torch.zeros(8, dtype=torch.float64, device="cuda"); grads = torch.zeros((8, L), dtype=torch.float64, device="cuda")
    lanes = [(sap, ev.engine), (lane2, eng2)]

    def enqueue(slot, j, lane):
        s_, e_ = lanes[lane]
        s_.set_grad_output(grads[slot]); e_.shard_eval_fused(dms[j], 0.0, True, 0); s_.save_result(var[slot:slot + 1]); s_.set_grad_output(None)
    for j in range(4):
        enqueue(j, j, 0); enqueue(4 + j, j, 1)                  # both lanes at once, eager
    sap.sync(); lane2.sync()
    for j in range(4):
        vo, go, _ = o.variance_GH(ms[j], nohess=True)
        assert abs(var[j].item() - vo) <= 1e-12 * vo and maxrel(grads[j].cpu().numpy(), go) < 1e-12
        assert var[j].item() == var[4 + j].item() and torch.equal(grads[j], grads[4 + j])     # lanes agree bit for bit
    keep_v, keep_g = var.clone(), grads.clone()
    var.zero_(); grads.zero_()
    gids = []
    for lane, (s_, _) in enumerate(lanes):
        s_.graph_begin()
        for j in range(4):
            enqueue(4 * lane + j, j, lane)
        gids.append(s_.graph_end())
    for rep in range(3):
        sap.graph_launch(gids[0]); lane2.graph_launch(gids[1])
    sap.sync(); lane2.sync()
    assert torch.equal(var, keep_v) and torch.equal(grads, keep_g)
    assert sap.last_result()[1] == 0 and lane2.last_result()[1] == 0
    lane2.close(); sap.close()


def test_fused_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    N, world = 12, 2
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29600 + os.getpid() % 300, N, ret), nprocs=world, join=True)
    C = orc.wishart_cov(N, 2)
    groups = orc.enumerate_groups(N)
    o = orc.SapOracle(C, N, groups)
    for seed in range(3):
        m = orc.dense_m(o.L, seed) if seed < 2 else orc.sparse_m(o.L, N, seed)
        v, g, H = o.variance_GH(m, hess_mode="factored")
        tol = 1e-12 if seed < 2 else 1e-9
        a, b = ret[0][seed], ret[1][seed]
        assert a["var"] == b["var"] and np.array_equal(a["phi"], b["phi"])       # rank-order sum: identical on all ranks
        assert abs(a["var"] - v) <= 1e-12 * v
        assert np.array_equal(a["grad"], b["grad"]) and maxrel(a["grad"], g) < tol
        Hcat = np.vstack([a["H"], b["H"]])
        assert maxrel(Hcat, H) < tol
    # operator: H p with the factors sharded over the two GPUs
    m = orc.dense_m(o.L, 0)
    v, g, H = o.variance_GH(m, hess_mode="factored")
    Hp = H @ np.random.RandomState(5).randn(o.L)
    a, b = ret[0][3], ret[1][3]
    assert np.array_equal(a["hp"], b["hp"]) and maxrel(a["hp"], Hp) < 1e-12
    assert a["hi"] == b["lo"] and maxrel(np.concatenate([a["own"], b["own"]]), Hp) < 1e-12


def _pilot_worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    import bluest_b200 as blu
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    d = np.load(os.path.join(ROOT, "tests", "golden", "pilot_mlmc.npz"))
    Y = d["Y"]                                             # (No, n, M): every rank takes its rows of every output
    n = Y.shape[1]
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    mine = torch.from_numpy(np.ascontiguousarray(Y[:, lo:hi, :])).to("cuda:%d" % rank)
    r = blu.pilot_statistics(mine, dist=dist, telescoped=True)
    ret[rank] = {k: v.copy() for k, v in r.items()}
    dist.barrier()
    dist.destroy_process_group()


def test_pilot_row_split_two_gpus():
    """The sample rows split over two ranks (blue_fn.py:108-110 splits the samples over the MPI ranks, :177-187 adds the
    sums): per-rank Gram kernels, ONE NCCL all-reduce of (N^2+N) n_out doubles, same statistics on every rank."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_pilot_worker, args=(2, 29900 + os.getpid() % 90, ret), nprocs=2, join=True)
    d = np.load(os.path.join(ROOT, "tests", "golden", "pilot_mlmc.npz"))
    M = d["Y"].shape[2]
    iu = np.triu_indices(M, 1)
    for k in ("sumse", "sumsc", "C_hat"):
        assert np.array_equal(ret[0][k], ret[1][k])
        assert maxrel(ret[0][k], d["one/" + k]) < 1e-12
    for o in range(d["Y"].shape[0]):
        assert maxrel(ret[0]["dV"][o][iu], d["one/dV"][o][iu]) < 1e-12
        assert maxrel(ret[0]["sumsd1"][o][iu], d["one/sumsd1"][o][iu]) < 1e-12
