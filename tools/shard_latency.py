"""Per-rank kernel time of the group-sharded evaluation at N models, emulated on ONE GPU: the context
owns slice 0..w-1 of a w-way split in turn; times blu_shard_phi + blu_shard_finish(grad) per slice."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import bluest_b200 as blu, oracle as orc
from bluest_b200.dist import GpuEngine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
groups = blu.enumerate_groups(N)
sizes = [len(g) for g in groups]
L = sum(sizes)
sap = blu.SAP(orc.wishart_cov(N, 0), N, groups, np.ones(L), verbose=False)
eng = GpuEngine(sap)
m = torch.from_numpy(orc.dense_m(L, 0)).cuda()
ext = torch.cuda.ExternalStream(sap.stream())
def timeit(fn, n=30):
    for _ in range(5): fn()
    sap.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(n): fn()
    e1.record(ext); sap.sync()
    return e0.elapsed_time(e1) / n * 1e3
for w in (1, 2, 4, 8):
    sl = blu.balanced_slices(sizes, w)
    for r in sorted(set([0, w - 1])):
        eng.set_slice(*sl[r])
        t_phi = timeit(lambda: eng.shard_phi(m))
        t_fin = timeit(lambda: eng.shard_finish(0.0, False, False))
        t_all = timeit(lambda: (eng.shard_phi(m), eng.shard_finish(0.0, True, False)))
        t_uv = timeit(lambda: (eng.shard_phi(m), eng.shard_finish(0.0, True, True)))
        t_u = timeit(lambda: (eng.shard_phi(m), eng.shard_finish(0.0, True, 2)))
        print("N=%d world=%d rank=%d groups=%d: phi %.1f us, finish %.1f us, phi+finish+grad %.1f us, +U,V %.1f us, +U only %.1f us" % (N, w, r, sl[r][1] - sl[r][0], t_phi, t_fin, t_all, t_uv, t_u), flush=True)
