"""ctypes binding of the tracker oracle (oracle/tracker_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtracker_oracle.so")
REF_LIB_PATH = os.path.join(_HERE, "_ref", "libcilqr_ref_tracker.so")


class Config(C.Structure):
    _fields_ = [("sumulation_dt", C.c_double), ("dt", C.c_double), ("tolerance", C.c_double), ("max_num_iteration", C.c_int)] + [
        (n, C.c_double) for n in ("lat_weight_l", "lat_weight_theta", "lat_weight_delta", "lat_weight_delta_rate",
                                  "lat_preview_time", "lon_weight_s", "lon_weight_v", "lon_weight_a", "lon_weight_j",
                                  "wheel_base", "delta_min", "delta_max", "min_acceleration", "max_acceleration",
                                  "delta_rate_min", "delta_rate_max", "jerk_min", "jerk_max")]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tracker_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libtracker_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.tracker_oracle_default_config.argtypes = [C.POINTER(Config)]
        L.tracker_oracle_plan.argtypes = [C.POINTER(Config), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.tracker_oracle_solve_lqr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_uint,
                                               C.c_void_p, C.POINTER(C.c_int)]
        L.tracker_oracle_init_guess.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def ref_lib():
    """The reference's own Tracker compiled against the Eigen stand-in (oracle/_ref); None when it was not built."""
    global _ref
    if _ref is None and os.path.exists(REF_LIB_PATH):
        L = C.CDLL(REF_LIB_PATH)
        L.ref_tracker_plan.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _ref = L
    return _ref


def default_config() -> Config:
    c = Config()
    lib().tracker_oracle_default_config(C.byref(c))
    return c


def start_record(start4) -> np.ndarray:
    """TrajectoryPoint handed to Tracker::Plan: x, y, theta, velocity set (trajectory_planner.cpp:73-75)."""
    r = np.zeros(13)
    r[2], r[3], r[4], r[6] = start4[0], start4[1], start4[2], start4[3]
    return r


def plan(start13, coarse, cfg: Config | None = None):
    """-> (ok, trajectory [K,13], total DARE iterations)"""
    cfg = cfg or default_config()
    coarse = np.ascontiguousarray(coarse, np.float64)
    start13 = np.ascontiguousarray(start13, np.float64)
    K = coarse.shape[0]
    out = np.zeros((K, 13))
    it = C.c_int()
    ok = lib().tracker_oracle_plan(C.byref(cfg), start13.ctypes.data, coarse.ctypes.data, K, out.ctypes.data, C.byref(it))
    return bool(ok), out, it.value


def init_guess(traj):
    traj = np.ascontiguousarray(traj, np.float64)
    K = traj.shape[0]
    X, U = np.zeros((K, 6)), np.zeros((K - 1, 2))
    lib().tracker_oracle_init_guess(traj.ctypes.data, K, X.ctypes.data, U.ctypes.data)
    return X, U


def solve_lqr(A, B, Q, R, tol=0.01, max_iter=150):
    A, B, Q = (np.ascontiguousarray(x, np.float64) for x in (A, B, Q))
    K = np.zeros(3)
    it = C.c_int()
    lib().tracker_oracle_solve_lqr(A.ctypes.data, B.ctypes.data, Q.ctypes.data, R, tol, max_iter, K.ctypes.data, C.byref(it))
    return K, it.value


def ref_plan(start13, coarse):
    coarse = np.ascontiguousarray(coarse, np.float64)
    start13 = np.ascontiguousarray(start13, np.float64)
    out = np.zeros((coarse.shape[0], 13))
    rc = ref_lib().ref_tracker_plan(start13.ctypes.data, coarse.ctypes.data, coarse.shape[0], out.ctypes.data)
    return rc, out
