"""The CPU oracle built on the PORTABLE libm (oracle/libcilqr_oracle_pm.so = cilqr_oracle.c with -DCILQR_PM_LIBM).

TEST INFRASTRUCTURE ONLY.  Same API as oracle/binding.py (this module re-executes that file against the other
library).  It is the bit-for-bit target of the STRICT build of the CUDA kernel (libcilqr_b200_strict.so), which
uses the same sin / cos / tan / log / hypot source (cilqr_b200/csrc/pm_math.h) on the device.  The default oracle
(glibc libm) is the one pinned against the compiled reference; the two differ only through the last bits of those
five functions.
"""
import importlib.util as _ilu
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
_spec = _ilu.spec_from_file_location("oracle._binding_pm_impl", _os.path.join(_here, "binding.py"))
_m = _ilu.module_from_spec(_spec)
_spec.loader.exec_module(_m)
_m._LIB_PATH = _os.path.join(_here, "libcilqr_oracle_pm.so")
_m._MAKE_TARGET = "libcilqr_oracle_pm.so"
globals().update({k: getattr(_m, k) for k in dir(_m) if not k.startswith("__")})
