"""ctypes binding of the DP planner oracle (oracle/dp_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdp_oracle.so")
_lib = None
NT, NS, NL = 5, 7, 10


class Config(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "tf", "delta_t", "dp_nominal_velocity", "dp_w_obstacle", "dp_w_lateral", "dp_w_lateral_change",
        "dp_w_lateral_velocity_change", "dp_w_longitudinal_velocity_bias", "dp_w_longitudinal_velocity_change",
        "max_velocity", "width", "wheel_base", "front_hang_length", "rear_hang_length")]


class Env(C.Structure):
    _fields_ = [("R", C.c_int), ("ref", C.c_void_p), ("NB", C.c_int), ("barrier", C.c_void_p), ("V", C.c_int),
                ("n_static", C.c_int), ("static_poly", C.c_void_p), ("static_nv", C.c_void_p),
                ("n_dyn", C.c_int), ("T", C.c_int), ("dyn_time", C.c_void_p), ("dyn_samples", C.c_void_p),
                ("dyn_poly", C.c_void_p), ("dyn_nv", C.c_void_p)]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("dp_oracle.c", "dp_oracle.h")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(map(os.path.getmtime, srcs)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libdp_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.dp_default_config.argtypes = [C.POINTER(Config)]
        L.dp_default_config.restype = None
        L.dp_build_barrier.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.dp_evaluate_station.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_void_p]
        L.dp_evaluate_station.restype = None
        L.dp_get_projection.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        L.dp_get_projection.restype = None
        L.dp_check_optimization_collision.argtypes = [C.POINTER(Config), C.POINTER(Env), C.c_double, C.c_double,
                                                      C.c_double, C.c_double]
        L.dp_num_knots.argtypes = [C.POINTER(Config)]
        L.dp_plan.argtypes = [C.POINTER(Config), C.POINTER(Env), C.c_double, C.c_double, C.c_double, C.c_void_p,
                              C.POINTER(C.c_double), C.c_void_p]
        _lib = L
    return _lib


def default_config() -> Config:
    c = Config()
    lib().dp_default_config(C.byref(c))
    return c


class Scene:
    """One Environment as contiguous arrays (keeps them alive for the ctypes struct)."""

    def __init__(self, ref, barrier=None, static_poly=None, static_nv=None, dyn_time=None, dyn_samples=None,
                 dyn_poly=None, dyn_nv=None):
        self.ref = np.ascontiguousarray(ref, np.float64)
        self.barrier = build_barrier(self.ref) if barrier is None else np.ascontiguousarray(barrier, np.float64)
        self.static_poly = np.ascontiguousarray(static_poly if static_poly is not None else np.zeros((0, 4, 2)), np.float64)
        ns, V = self.static_poly.shape[:2]
        self.static_nv = np.ascontiguousarray(static_nv if static_nv is not None else np.full(ns, V), np.int32)
        self.dyn_poly = np.ascontiguousarray(dyn_poly if dyn_poly is not None else np.zeros((0, 1, V, 2)), np.float64)
        nd, T = self.dyn_poly.shape[:2]
        assert self.dyn_poly.shape[2] == V or nd == 0
        self.dyn_time = np.ascontiguousarray(dyn_time if dyn_time is not None else np.zeros((nd, T)), np.float64)
        self.dyn_samples = np.ascontiguousarray(dyn_samples if dyn_samples is not None else np.full(nd, T), np.int32)
        self.dyn_nv = np.ascontiguousarray(dyn_nv if dyn_nv is not None else np.full(nd, V), np.int32)
        self.env = Env(len(self.ref), self.ref.ctypes.data, len(self.barrier), self.barrier.ctypes.data, V, ns,
                       self.static_poly.ctypes.data, self.static_nv.ctypes.data, nd, T, self.dyn_time.ctypes.data,
                       self.dyn_samples.ctypes.data, self.dyn_poly.ctypes.data, self.dyn_nv.ctypes.data)


def build_barrier(ref) -> np.ndarray:
    ref = np.ascontiguousarray(ref, np.float64)
    cap = 2 * (int((ref[-1, 0] - ref[0, 0]) / 0.1) + 2)
    out = np.zeros((cap, 2))
    n = lib().dp_build_barrier(len(ref), ref.ctypes.data, out.ctypes.data, cap)
    assert n >= 0
    return np.ascontiguousarray(out[:n])


def evaluate_station(ref, s: float) -> np.ndarray:
    ref = np.ascontiguousarray(ref, np.float64)
    out = np.zeros(7)
    lib().dp_evaluate_station(len(ref), ref.ctypes.data, s, out.ctypes.data)
    return out


def get_projection(ref, x: float, y: float) -> np.ndarray:
    ref = np.ascontiguousarray(ref, np.float64)
    out = np.zeros(2)
    lib().dp_get_projection(len(ref), ref.ctypes.data, x, y, out.ctypes.data)
    return out


def check_collision(scene: Scene, time: float, x: float, y: float, theta: float, cfg: Config | None = None) -> bool:
    cfg = cfg or default_config()
    return bool(lib().dp_check_optimization_collision(C.byref(cfg), C.byref(scene.env), time, x, y, theta))


def plan(scene: Scene, x: float, y: float, theta: float, cfg: Config | None = None):
    """-> (ok, trajectory [K,13], min_cost, waypoints [NT,3])"""
    cfg = cfg or default_config()
    K = lib().dp_num_knots(C.byref(cfg))
    traj = np.zeros((K, 13))
    wp = np.zeros((NT, 3))
    mc = C.c_double()
    ok = lib().dp_plan(C.byref(cfg), C.byref(scene.env), x, y, theta, traj.ctypes.data, C.byref(mc), wp.ctypes.data)
    return bool(ok), traj, mc.value, wp
