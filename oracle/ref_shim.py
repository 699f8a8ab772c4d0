"""oracle/ref_shim.py -- TEST INFRASTRUCTURE ONLY.

Imports the *actual* reference Python (bluest.misc / bluest.sap / bluest.mosap /
bluest.blue_models) from a read-only checkout, in the build container only.  The
reference's ``bluest/__init__.py`` pulls mpi4py, cvxpy and cvxopt, none of which are
installed, so the package is registered as a bare namespace and those three are
replaced by inert stubs.  The compiled ``_cmisc_bluest`` comes
from ``oracle/_ref`` (built by oracle/Makefile from the reference's own cmisc.cpp).

Used by tests/golden/make_golden.py to generate the committed golden vectors, and by
``-m "not gpu"`` tests to pin oracle/oracle.py against the real thing when the
checkout is present.  Nothing on the GPU box may depend on it.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DEFAULT = os.environ.get("BLUEST_REF", "/root/reference")


def available(ref=REF_DEFAULT):
    return os.path.isfile(os.path.join(ref, "bluest", "misc.py")) and _ref_so() is not None


def _ref_so():
    d = os.path.join(HERE, "_ref")
    if not os.path.isdir(d):
        return None
    for f in os.listdir(d):
        if f.startswith("_cmisc_bluest") and f.endswith(".so"):
            return os.path.join(d, f)
    return None


class _OneRankComm:
    """1-rank stand-in for mpi4py's COMM_WORLD (only what blue_models/blue_fn touch)."""
    def Get_rank(self): return 0
    def Get_size(self): return 1
    def bcast(self, obj, root=0): return obj
    def allreduce(self, obj, op=None): return obj
    def barrier(self): return None
    Barrier = barrier
    def Split(self, *a, **k): return self


def load(ref=REF_DEFAULT, with_models=False):
    """Return a namespace with .misc, .sap, .mosap (and .blue_models if asked)."""
    if not available(ref):
        raise RuntimeError("reference checkout or oracle/_ref build not available")
    refdir = os.path.dirname(_ref_so())
    if refdir not in sys.path:
        sys.path.insert(0, refdir)
    if "bluest" not in sys.modules or not getattr(sys.modules["bluest"], "_shim", False):
        pkg = types.ModuleType("bluest")
        pkg.__path__ = [os.path.join(ref, "bluest")]
        pkg._shim = True
        sys.modules["bluest"] = pkg
        for name in ("cvxpy", "cvxopt"):
            if name not in sys.modules:
                stub = types.ModuleType(name)
                stub.matrix = stub.spmatrix = stub.solvers = None
                stub._stub = True
                sys.modules[name] = stub
        if "mpi4py" not in sys.modules:
            mpi = types.ModuleType("mpi4py")
            MPI = types.ModuleType("mpi4py.MPI")
            MPI.COMM_WORLD = _OneRankComm()
            MPI.SUM = "SUM"
            MPI.Comm = _OneRankComm
            mpi.MPI = MPI
            mpi._stub = True
            sys.modules["mpi4py"] = mpi
            sys.modules["mpi4py.MPI"] = MPI
    ns = types.SimpleNamespace()
    ns.cmisc = importlib.import_module("_cmisc_bluest")
    ns.misc = importlib.import_module("bluest.misc")
    ns.sap = importlib.import_module("bluest.sap")
    ns.mosap = importlib.import_module("bluest.mosap")
    if with_models:
        ns.blue_models = importlib.import_module("bluest.blue_models")
    return ns
