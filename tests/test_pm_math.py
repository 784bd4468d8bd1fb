"""The portable libm (cilqr_b200/csrc/pm_math.h) is an ordinary libm: measured against glibc on the ranges the solver
uses.  (Its purpose is bit-identical results on host and device, tests/test_gpu_strict.py; this test only shows that
exchanging glibc for it perturbs results by the last bits, like CUDA's libm does.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")


@pytest.fixture(scope="module")
def pm():
    subprocess.check_call(["make", "-C", HERE, "libpm_math.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(HERE, "libpm_math.so"))
    for f in (L.pm_export_batch, L.pm_export_libm_batch):
        f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    for n in ("sin", "cos", "tan", "log"):
        getattr(L, "pm_export_" + n).restype = C.c_double
        getattr(L, "pm_export_" + n).argtypes = [C.c_double]
    L.pm_export_hypot.restype = C.c_double
    L.pm_export_hypot.argtypes = [C.c_double, C.c_double]
    return L


def _ulps(got, ref):
    u = np.abs(np.nextafter(ref, np.inf) - ref)
    return np.abs(got - ref) / u


def test_accuracy_against_glibc(pm):
    rng = np.random.default_rng(1)
    n = 400_000
    cases = {0: (rng.uniform(-7.0, 7.0, n), 1.0), 1: (rng.uniform(-7.0, 7.0, n), 1.0), 2: (rng.uniform(-1.3, 1.3, n), 3.0),
             3: (np.exp(rng.uniform(-30.0, 30.0, n)), 1.0), 4: (rng.uniform(-300.0, 300.0, n), 1.0)}
    names = ["sin", "cos", "tan", "log", "hypot"]
    y = np.ascontiguousarray(rng.uniform(-300.0, 300.0, n) * rng.choice([1.0, 1e-3], n))
    worst = {}
    for which, (x, bound) in cases.items():
        x = np.ascontiguousarray(x)
        got, ref = np.zeros(n), np.zeros(n)
        pm.pm_export_batch(which, n, x.ctypes.data, y.ctypes.data, got.ctypes.data)
        pm.pm_export_libm_batch(which, n, x.ctypes.data, y.ctypes.data, ref.ctypes.data)
        e = _ulps(got, ref)
        worst[names[which]] = float(e.max())
        assert e.max() <= bound + 1.0, (names[which], e.max())  # glibc itself is within 1 ulp of the true value
    print("\n[pm_math] max difference from glibc in ulps:", worst)


def test_special_values(pm):
    assert pm.pm_export_log(0.0) == -np.inf and np.isnan(pm.pm_export_log(-1.0)) and pm.pm_export_log(np.inf) == np.inf
    assert pm.pm_export_log(1.0) == 0.0 and abs(pm.pm_export_log(5e-324) - np.log(5e-324)) < 1e-12
    assert np.isnan(pm.pm_export_sin(np.inf)) and np.isnan(pm.pm_export_tan(np.nan))
    assert pm.pm_export_sin(0.0) == 0.0 and pm.pm_export_cos(0.0) == 1.0
    assert pm.pm_export_hypot(3.0, 4.0) == 5.0 and pm.pm_export_hypot(np.inf, np.nan) == np.inf
    assert abs(pm.pm_export_hypot(3e200, 4e200) / 5e200 - 1.0) < 1e-15 and pm.pm_export_hypot(0.0, 0.0) == 0.0
    # large arguments: deterministic (the rollout is rejected whatever the value), finite
    assert np.isfinite(pm.pm_export_sin(1e9)) and abs(pm.pm_export_sin(1e9)) <= 1.0
