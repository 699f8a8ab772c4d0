"""bluest_b200 -- B200-native (sm_100a, FP64) sample-allocation hot path of BLUEST.

Drop-in for the reference's ``SAP`` / ``MOSAP`` objective closures and its ``_cmisc_bluest``
native routines; everything numerical runs in hand-written CUDA behind libbluest_b200.so
(include/bluest_b200.h).  No CPU fallback: importing works anywhere, calling needs a GPU.
"""
from ._lib import BluError, device_count, lib          # noqa: F401
from .groups import (balanced_slices, stream_cost, enumerate_cliques, enumerate_groups, enumerate_group_arrays, group_costs,     # noqa: F401
                     indicator_ES, mappings, union_groups)
from .sap import SAP                                    # noqa: F401
from .mosap import MOSAP, BLUESTError                   # noqa: F401
from .pilot import (pilot_covariance, pilot_sums, pilot_statistics, finalize_sums, PilotAccumulator, fill_missing_covariances,   # noqa: F401
                    estimate_missing_covariances)
from . import cmisc, intproj, io                        # noqa: F401
from .dist import ShardedEvaluator, GpuEngine          # noqa: F401
from .install import install, uninstall                # noqa: F401
from .sweep import solve_sweep, split_instances        # noqa: F401
from .batch import Batch, evaluate_many                 # noqa: F401

__all__ = ["SAP", "MOSAP", "BLUESTError", "BluError", "pilot_covariance", "pilot_sums", "pilot_statistics", "finalize_sums", "PilotAccumulator", "cmisc", "enumerate_groups", "enumerate_group_arrays",
           "enumerate_cliques", "union_groups", "group_costs", "indicator_ES", "mappings", "balanced_slices",
           "device_count", "lib", "ShardedEvaluator", "GpuEngine", "install", "uninstall", "solve_sweep", "split_instances", "Batch", "evaluate_many"]
