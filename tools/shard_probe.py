"""Per-slice phase times of the group-sharded 20-model evaluation, measured on ONE GPU: for every rank's slice of a
`world`-way split, the partial-Phi kernel and the finish + gradient kernels are timed separately (CUDA events on the
context's stream).  Shows how well a slicing balances the two streaming kernels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch

import bluest_b200 as blu
import oracle as orc
from bluest_b200 import _lib
from bluest_b200.groups import balanced_slices, stream_cost

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
C = orc.wishart_cov(N, 0)
ga = blu.enumerate_group_arrays(N)
sizes = [len(g) for g in ga]
L = sum(sizes)
sap = blu.SAP(C, N, ga, np.ones(L), verbose=False)
m = torch.from_numpy(orc.dense_m(L, 0)).cuda()
ext = torch.cuda.ExternalStream(sap.stream())
lib = _lib.lib()
import ctypes
mp = ctypes.c_void_p(int(m.data_ptr()))


def timed(fn, n=20):
    for _ in range(3):
        fn()
    sap.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(n):
        fn()
    e1.record(ext)
    sap.sync()
    return e0.elapsed_time(e1) / n * 1e3


weights = {"stream_cost": stream_cost, "k^2": lambda k: float(k * k), "T": lambda k: k * (k + 1) / 2.0, "T+4": lambda k: k * (k + 1) / 2.0 + 4.0,
           "T+12": lambda k: k * (k + 1) / 2.0 + 12.0}
for world in (1, 2, 4, 8):
    for wname, wfn in weights.items():
        if world == 1 and wname != "stream_cost":
            continue
        sl = balanced_slices(sizes, world, weight=wfn)
        rows = []
        for (lo, hi) in sl:
            _lib.check(lib.blu_ctx_set_slice(sap._ctx, lo, hi))
            t_phi = timed(lambda: _lib.check(lib.blu_shard_phi(sap._ctx, mp)))
            t_fin = timed(lambda: _lib.check(lib.blu_shard_finish(sap._ctx, 0.0, 1, 0)))
            rows.append((hi - lo, t_phi, t_fin))
        print("world=%d weight=%-11s max phi %.1f  max fin+grad %.1f  sum-of-max %.1f | " % (world, wname, max(r[1] for r in rows), max(r[2] for r in rows),
              max(r[1] for r in rows) + max(r[2] for r in rows)) + " ".join("[%d: %.0f+%.0f]" % r for r in rows), flush=True)
# fixed costs: a slice of a few groups in the middle of the enumeration
sap2 = sap
for (lo, hi) in ((500000, 500032), (500000, 504096), (500000, 532768)):
    _lib.check(lib.blu_ctx_set_slice(sap._ctx, lo, hi))
    t_phi = timed(lambda: _lib.check(lib.blu_shard_phi(sap._ctx, mp)), 50)
    t_fin = timed(lambda: _lib.check(lib.blu_shard_finish(sap._ctx, 0.0, 0, 0)), 50)
    t_fg = timed(lambda: _lib.check(lib.blu_shard_finish(sap._ctx, 0.0, 1, 0)), 50)
    print("slice of %d groups: phi %.1f us, finish alone %.1f us, finish+grad %.1f us" % (hi - lo, t_phi, t_fin, t_fg), flush=True)
# where the serial tail of the fused Phi kernel goes: globaltimer stamps of its last CTA (world = 1: self push)
_lib.check(lib.blu_ctx_set_slice(sap._ctx, 0, L))
from bluest_b200.dist import GpuEngine
eng = GpuEngine(sap)
eng.connect_peers(0, 1)
for (lo, hi) in ((0, L), (500000, 532768), (0, 130000)):
    _lib.check(lib.blu_ctx_set_slice(sap._ctx, lo, hi))
    for _ in range(3):
        _lib.check(lib.blu_shard_eval_fused(sap._ctx, mp, 0.0, 1, 0))
    st = (ctypes.c_uint64 * 16)()
    _lib.check(lib.blu_ctx_last_stamps(sap._ctx, st))
    s = [int(v) for v in st]
    names = {2: "group-last", 3: "final fold starts", 4: "rank sums", 5: "exchange", 6: "Phi ready", 7: "pinv", 9: "done"}
    print("slice [%d,%d): " % (lo, hi) + ", ".join("%s +%.1f us" % (names[i], (s[i] - s[1]) / 1e3) for i in (2, 3, 4, 5, 6, 7, 9)), flush=True)
sap.close()
