"""ctypes binding of the corridor oracle (oracle/corridor_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcorridor_oracle.so")
_lib = None

CODE_NAMES = ["ok", "no_points", "few_points", "origin_ub", "capacity", "point_capacity"]


class Config(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("max_diff_x", "max_diff_y", "radius", "max_axis_x", "max_axis_y",
                                          "lane_segment_length")]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("corridor_oracle.c", "corridor_oracle.h")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(map(os.path.getmtime, srcs)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libcorridor_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.corr_default_config.argtypes = [C.POINTER(Config)]
        L.corr_default_config.restype = None
        L.corr_convex_hull_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.corr_convex_hull_f32.restype = C.c_int
        L.corr_add_corridor_points.argtypes = [C.POINTER(Config), C.c_double, C.c_double, C.c_double, C.c_void_p,
                                               C.POINTER(C.c_int)]
        L.corr_add_corridor_points.restype = None
        L.corr_build_corridor.argtypes = [C.POINTER(Config), C.c_double, C.c_double, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.corr_build_corridor.restype = C.c_int
        L.corr_plan.argtypes = [C.POINTER(Config), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.corr_plan.restype = C.c_int
        L.corr_lane_constraints.argtypes = [C.POINTER(Config), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.corr_lane_constraints.restype = C.c_int
        _lib = L
    return _lib


def default_config() -> Config:
    c = Config()
    lib().corr_default_config(C.byref(c))
    return c


def convex_hull(points, clockwise: bool = False) -> np.ndarray:
    p = np.ascontiguousarray(points, np.float32).reshape(-1, 2)
    out = np.zeros(max(len(p), 1), np.int32)
    n = lib().corr_convex_hull_f32(p.ctypes.data, len(p), int(clockwise), out.ctypes.data)
    return out[:n]


def add_corridor_points(x, y, theta, cfg: Config | None = None) -> np.ndarray:
    cfg = cfg or default_config()
    out = np.zeros((8, 2))
    n = C.c_int(0)
    lib().corr_add_corridor_points(C.byref(cfg), x, y, theta, out.ctypes.data, C.byref(n))
    return out[:n.value]


def build_corridor(origin_x, origin_y, points, cap: int = 256, cfg: Config | None = None):
    cfg = cfg or default_config()
    p = np.ascontiguousarray(points, np.float64).reshape(-1, 2)
    cons = np.zeros((cap, 3))
    poly = np.zeros((cap, 2))
    m = C.c_int(0)
    rc = lib().corr_build_corridor(C.byref(cfg), origin_x, origin_y, p.ctypes.data, len(p), cons.ctypes.data,
                                   poly.ctypes.data, cap, C.byref(m))
    return rc, cons[:m.value], poly[:m.value]


def plan_batch(traj, obs_points, obs_cnt, M_max: int, cfg: Config | None = None):
    """Corridor::BuildCorridorConstraints for every trajectory of the batch -> corridor [B,K,M_max,3],
    cnt [B,K], polygon [B,K,M_max,2], code [B,K]."""
    cfg = cfg or default_config()
    traj = np.ascontiguousarray(traj, np.float64)
    pts = np.ascontiguousarray(obs_points, np.float64)
    cnt = np.ascontiguousarray(obs_cnt, np.int32)
    B, K, P = pts.shape[:3]
    cor = np.zeros((B, K, M_max, 3))
    ccnt = np.zeros((B, K), np.int32)
    poly = np.zeros((B, K, M_max, 2))
    code = np.zeros((B, K), np.int32)
    for b in range(B):
        lib().corr_plan(C.byref(cfg), K, traj[b].ctypes.data, pts[b].ctypes.data, cnt[b].ctypes.data, P, M_max,
                        cor[b].ctypes.data, ccnt[b].ctypes.data, poly[b].ctypes.data, code[b].ctypes.data)
    return cor, ccnt, poly, code


def lane_constraints(boundary, is_left: bool, cap: int = 1024, cfg: Config | None = None):
    cfg = cfg or default_config()
    p = np.ascontiguousarray(boundary, np.float64).reshape(-1, 2)
    out = np.zeros((cap, 7))
    n = lib().corr_lane_constraints(C.byref(cfg), p.ctypes.data, len(p), int(is_left), out.ctypes.data, cap)
    return n, out[:max(n, 0)]
