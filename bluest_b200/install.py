"""install() -- swap the reference's SAP for the B200-backed one inside an imported ``bluest``.

What a BLUEST maintainer would do by hand done at run time:

* Level 2: ``bluest.sap.SAP`` / ``bluest.mosap.SAP`` become a hybrid class that keeps every
  host-side method of the reference's SAP (``solve``, ``cvxopt_solve``, ``cvxpy_solve``,
  ``ipopt_solve``, ``scipy_solve``, ``get_max_sample_constraints`` ...) and takes ``__init__``,
  ``get_variance_functions``, ``psi``, ``invcovs``, ``integer_projection`` and
  ``compute_BLUE_estimator`` from ``bluest_b200.SAP`` -- so ``BLUEProblem.setup_solver`` / ``MOSAP`` run unmodified on the GPU closures.
* Level 1: the five ``_cmisc_bluest`` routines bound in ``bluest.misc`` are replaced by
  ``bluest_b200.cmisc``.

``uninstall()`` restores the originals.  Nothing here falls back to the CPU: with the hybrid
installed and no GPU, constructing a SAP raises BluError.
"""
import sys

from . import cmisc
from .sap import SAP as B200SAP

_saved = {}
_L1 = ("assemble_psi_c", "objectiveK_c", "gradK_c", "hessKQ_c", "cleanupK_c")
_L2_METHODS = ("__init__", "get_variance_functions", "_m", "eval_device", "upload_m", "sync", "last_result", "last_timing",
               "timing_log", "timing_read", "last_launches", "device_ptr", "device_buffer", "stream", "close", "__del__", "compute_BLUE_estimator", "integer_projection",
               "variance_GH_begin", "variance_GH_end", "hess_matvec", "hess_operator", "hess_matvec_device", "set_option", "get_option",
               "graph_begin", "graph_end", "graph_launch", "save_result", "set_grad_output")


def make_hybrid(ref_sap_cls):
    """Reference SAP (host orchestration, solvers) + B200 SAP (setup, closures, psi, invcovs)."""
    ns = {name: getattr(B200SAP, name) for name in _L2_METHODS}
    ns["psi"] = B200SAP.psi
    ns["invcovs"] = B200SAP.invcovs
    for lazy in ("ES", "e", "flattened_groups"):             # built on first use (sap.py:66-71, 89-95)
        ns[lazy] = getattr(B200SAP, lazy)
    ns["__doc__"] = "bluest.sap.SAP with the sample-allocation hot path on the B200 (bluest_b200)."
    return type("SAP", (ref_sap_cls,), ns)


def install(level1=True, level2=True):
    """Patch the already-importable ``bluest`` package in place.  Returns the hybrid SAP class."""
    import importlib
    sap_mod = importlib.import_module("bluest.sap")
    mosap_mod = importlib.import_module("bluest.mosap")
    misc_mod = importlib.import_module("bluest.misc")
    hybrid = None
    if level2 and "sap.SAP" not in _saved:
        _saved["sap.SAP"] = sap_mod.SAP
        _saved["mosap.SAP"] = mosap_mod.SAP
        hybrid = make_hybrid(sap_mod.SAP)
        sap_mod.SAP = hybrid
        mosap_mod.SAP = hybrid
        pkg = sys.modules.get("bluest")
        if pkg is not None and getattr(pkg, "SAP", None) is _saved["sap.SAP"]:
            _saved["pkg.SAP"] = pkg.SAP
            pkg.SAP = hybrid
    if level1 and "misc" not in _saved:
        _saved["misc"] = {n: getattr(misc_mod, n) for n in _L1}
        for n in _L1:
            setattr(misc_mod, n, getattr(cmisc, n))
    return hybrid if hybrid is not None else sap_mod.SAP


def uninstall():
    import importlib
    if "sap.SAP" in _saved:
        importlib.import_module("bluest.sap").SAP = _saved.pop("sap.SAP")
        importlib.import_module("bluest.mosap").SAP = _saved.pop("mosap.SAP")
        if "pkg.SAP" in _saved:
            sys.modules["bluest"].SAP = _saved.pop("pkg.SAP")
    if "misc" in _saved:
        misc_mod = importlib.import_module("bluest.misc")
        for n, f in _saved.pop("misc").items():
            setattr(misc_mod, n, f)
