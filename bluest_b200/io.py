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


def setup_mosap(graph, K=4, device=0, verbose=False):
    """The group-enumeration half of ``BLUEProblem.setup_solver`` (blue_models.py:458-509) for a loaded
    graph: cliques of size <= K per output inside model 0's component, union + sort, group costs, and
    the device-backed MOSAP.  ``graph`` is what ``load_graph_data`` returns."""
    M, No = graph["M"], graph["n_outputs"]
    K = min(K, M)
    multi_groups, Ks = [], []
    for n in range(No):
        A = graph["adjacency"][n]
        mg = enumerate_cliques(A, K, component_of=0)
        multi_groups.append(mg)
        Ks.append(min(K, len(mg)))
    Kmax = max(Ks)
    groups = union_groups(multi_groups)
    costs = group_costs(groups, graph["costs"])
    multi_costs = [group_costs(mg, graph["costs"]) for mg in multi_groups]
    return MOSAP(graph["C"], Kmax, Ks, groups, multi_groups, costs, multi_costs, verbose=verbose, device=device)
