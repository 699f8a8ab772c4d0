"""Graph-data files and the front half of ``BLUEProblem.setup_solver`` ("next" row f4).

The reference persists a model graph as an ``.npz`` with ``M, n_outputs, costs, C0..C{n-1}, SG, dV``
(``save_graph_data`` / ``load_graph_data``, blue_models.py:265-299), where ``C{n}`` is the graph
adjacency: 0 = models cannot be coupled, inf = uncorrelated, anything else = the covariance.
``get_covariance`` (blue_models.py:166-179) turns that into the covariance matrix the allocation
problem sees: 0 -> NaN (never indexed by a clique), inf -> 0.
"""
import numpy as np

from .groups import enumerate_cliques, group_costs, union_groups
from .mosap import MOSAP


def load_graph_data(filename, n_outputs=None):
    """Read a reference-format graph file.  Returns a dict with M, n_outputs, costs, adjacency (list),
    C (list, the ``get_covariance`` view), SG (list of connected-component node lists), dV."""
    data = dict(np.load(filename))
    M = int(data["M"])
    No = int(data["n_outputs"]) if n_outputs is None else int(n_outputs)
    if No > int(data["n_outputs"]):
        raise ValueError("Loaded data number of models and/or number of outputs mismatch with the user-given values")
    adj = [np.array(data["C%d" % n], dtype=np.float64) for n in range(No)]
    C = []
    for A in adj:
        Cn = A.copy()
        mask0 = Cn == 0
        maskinf = np.isinf(Cn)
        Cn[mask0] = np.nan
        Cn[maskinf] = 0
        C.append(Cn)
    dV = data.get("dV", None)
    return {"M": M, "n_outputs": No, "costs": np.array(data["costs"], dtype=np.float64), "adjacency": adj, "C": C,
            "SG": data["SG"].tolist()[:No], "dV": None if dV is None else [dV[n] for n in range(No)]}


def save_graph_data(filename, costs, adjacency, SG, dV=None):
    """Write a reference-format graph file (blue_models.py:265-270)."""
    M = len(costs)
    C_dict = {"C%d" % n: np.asarray(A) for n, A in enumerate(adjacency)}
    if dV is None:
        dV = np.nan * np.ones((len(adjacency), M, M))
    np.savez(filename, M=M, n_outputs=len(adjacency), costs=np.asarray(costs), **C_dict, SG=np.asarray(SG), dV=np.asarray(dV))


def sample_file_name(filename, ls):
    """blue_fn.py:97-100: the sample file of the coupled models ``ls`` is ``<base><l0><l1>...<ext>``."""
    ext = "." + filename.split(".")[-1]
    base = ".".join(filename.split(".")[:-1]) + "".join(str(l) for l in ls)
    return base + ext


def save_sample_file(filename, ls, values, inputs, outputs_to_save=None):
    """The sample snapshot ``blue_fn`` keeps when it is given ``filename`` (blue_fn.py:97-104, 132-145, 189-222), same
    keys and same append semantics, for model outputs that were evaluated in batches:

      values   (n_outputs, n, len(ls)) array (device tensors are copied to the host): ``values[o, s, i]`` = output o of
               model ``ls[i]`` for sample s -> keys ``values_<o>_<i>`` (only the outputs in ``outputs_to_save``)
      inputs   (n, len(ls)) or list of per-model arrays: the samples the models were evaluated on -> ``inputs_<i>``
      plus ``models``, ``n_samples``, ``n_outputs``.

    If the file exists its lists are extended and ``n_samples`` is increased (blue_fn.py:205-218); the models and the
    number of outputs must match.  Returns the path written."""
    import os
    if hasattr(values, "cpu"):
        values = values.cpu().numpy()
    values = np.asarray(values)
    No, n, Lm = values.shape
    if Lm != len(ls):
        raise ValueError("values has %d model columns, expected %d" % (Lm, len(ls)))
    if outputs_to_save is None:
        outputs_to_save = list(range(No))
    path = sample_file_name(filename, ls)
    out = {"values_%d_%d" % (o, i): [values[o, s, i] for s in range(n)] if o in outputs_to_save else [] for o in range(No) for i in range(Lm)}
    for i in range(Lm):
        col = inputs[i] if isinstance(inputs, (list, tuple)) else np.asarray(inputs)[:, i]
        # the reference appends a sample's input once per SAVED OUTPUT (the append sits inside its loop over the
        # outputs, blue_fn.py:135-140): reproduced, so that files written here and there can extend one another
        out["inputs_%d" % i] = [col[s] for s in range(n) for _ in outputs_to_save]
    out["models"] = np.array([ls])
    out["n_samples"] = np.array([n])
    out["n_outputs"] = np.array([No])
    if os.path.isfile(path):
        old = dict(np.load(path, allow_pickle=True))
        if [int(v) for v in np.ravel(old["models"])] != [int(v) for v in ls]:
            raise AssertionError("sample file %s holds other models" % path)
        if int(np.ravel(old["n_outputs"])[0]) != No:
            raise AssertionError("sample file %s holds another number of outputs" % path)
        for key in list(old.keys()):
            if "values" in key or "inputs" in key:
                out[key] = [item for item in old[key]] + out[key]
        out["n_samples"] = np.array([int(np.ravel(old["n_samples"])[0]) + n])
    np.savez_compressed(path, **out)
    return path


def load_sample_file(filename, ls=None):
    """Read a sample snapshot back: ``(values (n_outputs, n, len(ls)) with NaN for outputs that were not saved, inputs
    (n, len(ls)), n_samples)``.  ``ls`` given: ``filename`` is the base name handed to ``blue_fn``; None: the full path."""
    path = sample_file_name(filename, ls) if ls is not None else filename
    d = np.load(path, allow_pickle=True)
    models = [int(v) for v in np.ravel(d["models"])]
    No = int(np.ravel(d["n_outputs"])[0]); n = int(np.ravel(d["n_samples"])[0])
    vals = np.full((No, n, len(models)), np.nan)
    for o in range(No):
        for i in range(len(models)):
            a = np.asarray(d["values_%d_%d" % (o, i)], dtype=np.float64)
            if a.size:
                vals[o, :, i] = a
    ins = np.array([np.asarray(d["inputs_%d" % i], dtype=np.float64) for i in range(len(models))]).T
    rep = ins.shape[0] // n if n else 1                    # one copy of every input per saved output (see save_sample_file)
    return vals, ins[::max(rep, 1)], n


def is_subclique(adj, nodes):
    """``is_subclique`` of blue_models.py:33-36 on an adjacency matrix: every pair of ``nodes`` -- a node with
    itself included, the model graphs carry self-loops -- is joined by an edge (non-zero adjacency)."""
    A = np.asarray(adj)
    nodes = [int(v) for v in nodes]
    return all(A[i, j] != 0 for a, i in enumerate(nodes) for j in nodes[a:])


def prepare_groups(graph, K=4, groups=None, multi_groups=None):
    """The group bookkeeping of ``BLUEProblem.setup_solver`` (blue_models.py:453-501) -> (K, Ks, groups, multi_groups).

    Without user input: all cliques of size <= K of every output's model graph inside model 0's component.
    With ``groups`` (used for every output) or ``multi_groups`` (one list per output): the user's groups are
    sorted in place, those that are not cliques of the output's graph or leave model 0's component are dropped,
    the rest are binned by size -- size classes may stay empty -- and ``Ks[n]`` is the largest size that survived.
    Then the union over the outputs, each size class sorted lexicographically."""
    M, No = graph["M"], graph["n_outputs"]
    if multi_groups is not None and len(multi_groups) != No:
        raise ValueError("multi_groups must be a list of groupings of the same length as the number of outputs.")
    if groups is not None and multi_groups is None:
        multi_groups = [groups for _ in range(No)]
    if multi_groups is None:
        K = min(K, M)
        multi_groups, Ks = [], []
        for n in range(No):
            SG = graph.get("SG") if hasattr(graph, "get") else None
            mg = enumerate_cliques(graph["adjacency"][n], K, component_of=0, nodes=None if SG is None else SG[n])   # blue_models.py:468
            multi_groups.append(mg)
            Ks.append(min(K, len(mg)))
    else:
        first_Ks = [min(max(len(g) for g in user), M) for user in multi_groups]
        for n in range(No):
            component = set(int(v) for v in graph["SG"][n])
            binned = [[] for _ in range(first_Ks[n])]
            for g in multi_groups[n]:
                g.sort()                                                   # in place, like the reference
                if is_subclique(graph["adjacency"][n], g) and all(int(v) in component for v in g):
                    binned[len(g) - 1].append(g)
            multi_groups[n] = binned
        Ks = [min(max(len(g) for gk in mg for g in gk), M) for mg in multi_groups]
    Kmax = max(Ks)
    union = [[] for _ in range(Kmax)]
    seen = [set() for _ in range(Kmax)]
    for n in range(No):
        for k in range(Ks[n]):
            for g in multi_groups[n][k]:
                key = tuple(int(v) for v in g)
                if key not in seen[k]:
                    seen[k].add(key)
                    union[k].append(g)
    for k in range(Kmax):
        union[k].sort()
    return Kmax, Ks, union, multi_groups


def setup_mosap(graph, K=4, device=0, verbose=False, groups=None, multi_groups=None):
    """The construction half of ``BLUEProblem.setup_solver`` (blue_models.py:453-509) for a loaded graph: group
    bookkeeping (``prepare_groups``), group costs, and the device-backed MOSAP.  ``graph`` is what
    ``load_graph_data`` returns."""
    Kmax, Ks, union, multi_groups = prepare_groups(graph, K=K, groups=groups, multi_groups=multi_groups)
    costs = group_costs(union, graph["costs"])
    multi_costs = [group_costs(mg, graph["costs"]) for mg in multi_groups]
    return MOSAP(graph["C"], Kmax, Ks, union, multi_groups, costs, multi_costs, verbose=verbose, device=device)


def setup_solver(graph, K=4, budget=None, eps=None, groups=None, multi_groups=None, solver="scipy", continuous_relaxation=False,
                 max_model_samples=None, x0=None, device=0, verbose=False, **solve_kwargs):
    """``BLUEProblem.setup_solver`` (blue_models.py:448-538) on a loaded graph: builds the MOSAP, solves the
    allocation and returns ``(mosap, blue_data)`` with the reference's ``blue_data`` dictionary (``models``: the
    groups that received samples, ``samples``, ``errors`` = sqrt of the output variances, ``total_cost``)."""
    from .mosap import BLUESTError
    if budget is None and eps is None:
        raise ValueError("Need to specify either budget or RMSE tolerance")
    if budget is not None and eps is not None:
        eps = None                                                         # blue_models.py:450
    if eps is not None and np.ndim(eps) == 0:
        eps = [eps for _ in range(graph["n_outputs"])]
    mosap = setup_mosap(graph, K=K, device=device, verbose=verbose, groups=groups, multi_groups=multi_groups)
    mosap.solve(eps=eps, budget=budget, solver=solver, x0=x0, continuous_relaxation=continuous_relaxation,
                max_model_samples=max_model_samples, **solve_kwargs)
    if mosap.samples is None:
        raise BLUESTError("MOSAP solution failed!")
    Vs = mosap.variances(mosap.samples)
    picked = np.argwhere(mosap.samples > 0).flatten()
    blue_data = {"models": [mosap.flattened_groups[i] for i in picked], "samples": mosap.samples[mosap.samples > 0].copy(),
                 "errors": np.sqrt(Vs), "total_cost": mosap.tot_cost}
    return mosap, blue_data
