"""Where one group-sharded evaluation spends its time on a real multi-GPU run (launch under torchrun): per rank, the
%globaltimer stamps of the fused Phi kernel's last CTA (fold, peer exchange, pseudo-inverse) and CUDA-event times of
the Phi and gradient kernels of single evaluations."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

import bluest_b200 as blu
import oracle as orc
from bluest_b200 import _lib
from bluest_b200.dist import GpuEngine, ShardedEvaluator

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 20
C = orc.wishart_cov(N, 0)
ga = blu.enumerate_group_arrays(N)
sizes = [len(g) for g in ga]; L = sum(sizes)
sap = blu.SAP(C, N, ga, np.ones(L), verbose=False, device=local)
eng = GpuEngine(sap)
ev = ShardedEvaluator(eng, sizes, rank, world, dist=dist if world > 1 else None, fused=True)
m = torch.from_numpy(orc.dense_m(L, 0)).to("cuda:%d" % local)
ext = torch.cuda.ExternalStream(sap.stream(), device=local)
lib = _lib.lib()
mp = ctypes.c_void_p(int(m.data_ptr()))
for _ in range(5):
    eng.shard_eval_fused(m, 0.0, True, 0)
sap.sync()
if world > 1:
    dist.barrier()
rows = []
for rep in range(6):
    if world > 1:
        dist.barrier()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(ext)
    _lib.check(lib.blu_shard_eval_fused(sap._ctx, mp, 0.0, 0, 0))      # Phi + exchange + pinv only
    e[1].record(ext)
    _lib.check(lib.blu_shard_finish(sap._ctx, 0.0, 1, 0))              # stand-alone finish + gradient
    e[2].record(ext)
    sap.sync()
    st = (ctypes.c_uint64 * 16)()
    _lib.check(lib.blu_ctx_last_stamps(sap._ctx, st))
    s = [int(v) for v in st]
    rows.append((e[0].elapsed_time(e[1]) * 1e3, e[1].elapsed_time(e[2]) * 1e3, [(s[i] - s[1]) / 1e3 for i in (2, 3, 4, 5, 6, 7, 9)]))
r = rows[-1]
print("rank %d slice %s: fused Phi kernel %.1f us, finish+grad %.1f us | tail of the last CTA: group-last +%.1f, final fold +%.1f, rank sums +%.1f, "
      "exchange done +%.1f, Phi ready +%.1f, pinv +%.1f, done +%.1f us" % ((rank, ev.slices[rank], r[0], r[1]) + tuple(r[2])), flush=True)
sap.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
