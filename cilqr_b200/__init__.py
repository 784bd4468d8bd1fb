"""cilqr_b200 -- B200-native batched constrained-iLQR solver behind the interface of
mpt0816/Cilqr's ``IlqrOptimizer::Plan`` (reference ``algorithm/ilqr/ilqr_optimizer.h:41-48``).

The product is ``lib/libcilqr_b200.so`` (hand-written sm_100a CUDA + an extern-"C" ABI declared in
``include/cilqr_b200.h``).  This package is the thin Python mirror used by the tests and the
benchmark; it never computes anything itself and raises if the CUDA library is missing.
"""
from .solver import (CilqrError, Params, Solver, STATUS_NAMES, default_params, lib_path,  # noqa: F401
                     load_library)
from . import scenarios  # noqa: F401

__all__ = ["Solver", "Params", "CilqrError", "default_params", "load_library", "lib_path",
           "STATUS_NAMES", "scenarios"]
