"""Group enumeration and integer index maps -- host-side, exact.

Mirrors the integer logic of ``BLUEProblem.setup_solver`` (blue_models.py:458-501) and
``MOSAP.__init__`` (mosap.py:41-67): groups are size-major, lexicographic inside a size class;
the flat index of a group is ``cumsizes[k-1] + rank``.
"""
from itertools import combinations

import numpy as np


def enumerate_groups(N, K=None):
    """All subsets of {0..N-1} of size <= K: what nx.enumerate_all_cliques yields on a complete
    model graph after ``groups[k].sort()`` (blue_models.py:462-474, :500-501)."""
    K = N if K is None else min(K, N)
    return [[list(c) for c in combinations(range(N), k)] for k in range(1, K + 1)]


def enumerate_group_arrays(N, K=None):
    """``enumerate_groups`` as one (Lk,k) int64 array per size class (same groups, same order): no Python list
    per group, so the 1 048 575 groups of 20 models take about a second instead of several, and ``SAP`` ingests
    the arrays as they are."""
    K = N if K is None else min(K, N)
    out = []
    for k in range(1, K + 1):
        flat = np.fromiter((v for c in combinations(range(N), k) for v in c), dtype=np.int64)
        out.append(flat.reshape(-1, k))
    return out


def enumerate_cliques(adj, K, component_of=0, nodes=None):
    """Cliques of size <= K of the coupling graph ``adj`` (non-zero = edge), restricted to the
    node set ``nodes`` -- the reference filters with its stored ``SG[n]`` (blue_models.py:462-474, :468) --
    or, when ``nodes`` is None, to the connected component of model ``component_of`` computed from the
    adjacency.  Plain breadth-first extension with increasing vertex ids, which reproduces networkx's
    size-major, lexicographic order."""
    adj = np.asarray(adj) != 0
    N = adj.shape[0]
    if nodes is not None:
        comp = set(int(v) for v in nodes)
    else:
        comp = {component_of}
        frontier = [component_of]
        while frontier:
            nxt = []
            for u in frontier:
                for v in range(N):
                    if v != u and adj[u, v] and v not in comp:
                        comp.add(v); nxt.append(v)
            frontier = nxt
    level = [[v] for v in range(N) if v in comp]
    out = []
    for k in range(1, min(K, N) + 1):
        if not level:
            break
        out.append(level)
        nxt = []
        for c in level:
            for v in range(c[-1] + 1, N):
                if v in comp and all(adj[u, v] for u in c):
                    nxt.append(c + [v])
        level = nxt
    return out


def union_groups(multi_groups):
    """Union over outputs + per-size lexicographic sort (blue_models.py:493-501), O(L log L)."""
    K = max(len(mg) for mg in multi_groups)
    out = []
    for k in range(K):
        seen = set()
        for mg in multi_groups:
            if k < len(mg):
                seen.update(tuple(int(v) for v in g) for g in mg[k])
        out.append([list(g) for g in sorted(seen)])
    return out


def group_costs(groups, model_costs):
    """blue_models.py:137-140: cost of a group = ``sum(model_costs[group])``.  Vectorised per size class,
    adding the members in order (0 + c[g0] + c[g1] + ...) exactly like Python's ``sum``, so the result is
    bit-identical to the reference's (a pairwise ``ndarray.sum`` would round differently from 8 members on)."""
    mc = np.asarray(model_costs)
    out = []
    for gk in groups:
        arr = np.asarray(gk, dtype=np.int64)
        if arr.size == 0:
            continue
        arr = arr.reshape(len(gk), -1)
        acc = np.zeros(arr.shape[0], dtype=np.result_type(mc.dtype, np.int64) if mc.dtype.kind in "iu" else mc.dtype)
        for j in range(arr.shape[1]):
            acc = acc + mc[arr[:, j]]
        out.append(acc)
    return np.concatenate(out) if out else np.zeros(0)


def _class_keys(gk):
    """One integer key per group of a size class: the membership bit mask (models < 63)."""
    arr = np.asarray(gk, dtype=np.int64)
    if arr.size == 0:
        return np.zeros(0, dtype=np.int64)
    arr = arr.reshape(len(gk), -1)
    return np.bitwise_or.reduce(np.left_shift(np.int64(1), arr), axis=1)


def indicator_ES(group_arrays, N):
    """ES[i][g] = int(model i in group g)  (sap.py:89-95, mosap.py:46-52), vectorised."""
    L = sum(len(g) for g in group_arrays)
    ES = np.zeros((N, L), dtype=np.int64)
    off = 0
    for g in group_arrays:
        g = np.asarray(g, dtype=np.int64)
        if g.size:
            cols = off + np.repeat(np.arange(g.shape[0]), g.shape[1])
            ES[g.ravel(), cols] = 1
        off += len(g)
    return [ES[i].copy() for i in range(N)]


def mappings(groups, multi_groups):
    """mappings[n][j] = flat position in ``groups`` of the j-th group of output n
    (mosap.py:54-67); dictionary look-up instead of the reference's O(L^2) scan."""
    sizes = [0] + [len(gk) for gk in groups]
    cum = np.cumsum(sizes)
    # per size class: membership masks of the union, sorted once; every output's groups are located by
    # binary search (the reference scans the union list for every group, mosap.py:59: O(L^2))
    table = {}
    for k, gk in enumerate(groups):
        if len(gk):
            keys = _class_keys(gk)
            order = np.argsort(keys, kind="stable")
            table[len(gk[0])] = (keys[order], order, int(cum[k]))
    out = []
    for mg in multi_groups:
        pos = []
        for gk in mg:
            if len(gk) == 0:
                continue
            size = len(gk[0])
            keys = _class_keys(gk)
            if size not in table:
                raise AssertionError("group %s of an output is missing from the union" % (tuple(int(v) for v in gk[0]),))
            skeys, order, off = table[size]
            at = np.searchsorted(skeys, keys)
            at = np.minimum(at, len(skeys) - 1)
            bad = skeys[at] != keys
            if bad.any():
                raise AssertionError("group %s of an output is missing from the union" % (tuple(int(v) for v in gk[int(np.argmax(bad))]),))
            pos.append(off + order[at])
        out.append(np.concatenate(pos).astype(np.int64) if pos else np.zeros(0, dtype=np.int64))
    return out


def stream_cost(k):
    """Relative cost of one group of size k in the two streaming kernels (Phi: one lane per packed entry, gradient:
    one lane per group): the packed entries T_k = k(k+1)/2 plus a fixed per-group part.  Measured on B200 at 20
    models (tools/shard_probe.py): for 2 / 4 / 8 slices this weight gives the smallest sum of the slowest Phi and the
    slowest gradient slice (8 slices: 41 + 30 us against 43 + 35 us for a per-32-entry step count and 45 + 31 us for
    the reference's k^2)."""
    return k * (k + 1) / 2.0 + 4.0


def balanced_slices(sizes, world, weight=None):
    """Contiguous slices of the flat enumeration with equal work.
    ``sizes`` = [L1..LK]; ``weight(k)`` = cost of one group of size k (default k^2, the reference's
    inner-loop count).  Returns ``world`` (lo, hi) pairs."""
    if weight is None:
        weight = lambda k: float(k * k)
    work = np.concatenate([np.full(int(Lk), weight(k + 1), dtype=np.float64) for k, Lk in enumerate(sizes)]) if sum(sizes) else np.zeros(0)
    cum = np.concatenate([[0.0], np.cumsum(work)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(cum, total * r / world)))
    bounds.append(len(work))
    bounds = [min(max(b, 0), len(work)) for b in bounds]
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return [(bounds[r], bounds[r + 1]) for r in range(world)]
