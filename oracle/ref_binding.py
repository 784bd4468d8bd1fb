"""ctypes binding of oracle/_ref/libcilqr_ref_geom.so: the REFERENCE'S OWN geometry / reference-line code
(algorithm/math/*.cpp, utils/discretized_trajectory.cpp, utils/discrete_points_math.cc), compiled from
/root/reference by `make -C oracle _ref`.  TEST INFRASTRUCTURE ONLY -- it pins the restatements."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libcilqr_ref_geom.so")
REFERENCE_ROOT = "/root/reference"
_lib = None


def available() -> bool:
    return all(os.path.exists(os.path.join(_HERE, "_ref", n)) for n in
               ("libcilqr_ref_geom.so", "libcilqr_ref_dp.so", "libcilqr_ref_solver.so", "libcilqr_ref_corridor.so")) or os.path.isdir(REFERENCE_ROOT)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.check_call(["make", "-C", _HERE, "_ref"], stdout=subprocess.DEVNULL)
        L = C.CDLL(LIB_PATH)
        d = C.c_double
        L.ref_normalize_angle.argtypes = [d]
        L.ref_normalize_angle.restype = d
        L.ref_slerp.argtypes = [d] * 5
        L.ref_slerp.restype = d
        L.ref_segment_distance.argtypes = [d] * 6
        L.ref_segment_distance.restype = d
        L.ref_evaluate_stations.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_get_projections.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_get_cartesians.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_polygon_overlaps_disc_box.argtypes = [C.c_int, C.c_void_p, d, d, d]
        L.ref_polygon_is_point_in.argtypes = [C.c_int, C.c_void_p, d, d]
        L.ref_disc_box_is_point_in.argtypes = [d] * 5
        L.ref_compute_path_profile.argtypes = [d, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _f(a):
    return np.ascontiguousarray(a, np.float64)


def evaluate_stations(ref, stations):
    ref, st = _f(ref), _f(stations)
    out = np.zeros((len(st), 7))
    lib().ref_evaluate_stations(len(ref), ref.ctypes.data, len(st), st.ctypes.data, out.ctypes.data)
    return out


def get_projections(ref, xy):
    ref, xy = _f(ref), _f(xy)
    out = np.zeros((len(xy), 2))
    lib().ref_get_projections(len(ref), ref.ctypes.data, len(xy), xy.ctypes.data, out.ctypes.data)
    return out


def get_cartesians(ref, sl):
    ref, sl = _f(ref), _f(sl)
    out = np.zeros((len(sl), 2))
    lib().ref_get_cartesians(len(ref), ref.ctypes.data, len(sl), sl.ctypes.data, out.ctypes.data)
    return out


def polygon_overlaps_disc_box(poly, cx, cy, radius) -> bool:
    p = _f(poly)
    return bool(lib().ref_polygon_overlaps_disc_box(len(p), p.ctypes.data, cx, cy, radius))


def polygon_is_point_in(poly, x, y) -> bool:
    p = _f(poly)
    return bool(lib().ref_polygon_is_point_in(len(p), p.ctypes.data, x, y))


def compute_path_profile(dt, xy):
    xy = _f(xy)
    n = len(xy)
    v, a, k = np.zeros(n), np.zeros(n), np.zeros(n)
    ok = lib().ref_compute_path_profile(dt, n, xy.ctypes.data, v.ctypes.data, a.ctypes.data, k.ctypes.data)
    return bool(ok), v, a, k


# ---- the reference's own DpPlanner / Environment (oracle/_ref/libcilqr_ref_dp.so) --------------------------------
DP_LIB_PATH = os.path.join(_HERE, "_ref", "libcilqr_ref_dp.so")
_dplib = None


def dp_lib():
    global _dplib
    if _dplib is None:
        if not os.path.exists(DP_LIB_PATH):
            subprocess.check_call(["make", "-C", _HERE, "_ref"], stdout=subprocess.DEVNULL)
        L = C.CDLL(DP_LIB_PATH)
        d, i, p = C.c_double, C.c_int, C.c_void_p
        L.ref_dp_plan.argtypes = [i, p, i, i, p, p, i, i, p, p, p, p, d, d, d, p, i, p]
        L.ref_check_optimization_collision.argtypes = [i, p, i, i, p, p, i, i, p, p, p, p, i, p, p]
        L.ref_road_barrier.argtypes = [i, p, p, i]
        _dplib = L
    return _dplib


def _scene_args(ref, static_poly, static_nv, dyn_time, dyn_samples, dyn_poly, dyn_nv):
    ref = _f(ref)
    sp, dt_, dp_ = _f(static_poly), _f(dyn_time), _f(dyn_poly)
    snv = np.ascontiguousarray(static_nv, np.int32)
    dsm = np.ascontiguousarray(dyn_samples, np.int32)
    dnv = np.ascontiguousarray(dyn_nv, np.int32)
    V = sp.shape[1] if sp.ndim == 3 and sp.shape[0] else (dp_.shape[2] if dp_.ndim == 4 and dp_.shape[0] else 4)
    T = dp_.shape[1] if dp_.ndim == 4 else 0
    keep = (ref, sp, snv, dt_, dsm, dp_, dnv)
    args = [len(ref), ref.ctypes.data, V, sp.shape[0], sp.ctypes.data, snv.ctypes.data, dp_.shape[0], T,
            dt_.ctypes.data, dsm.ctypes.data, dp_.ctypes.data, dnv.ctypes.data]
    return args, keep


def dp_plan(ref, static_poly, static_nv, dyn_time, dyn_samples, dyn_poly, dyn_nv, x, y, theta):
    """The reference's DpPlanner::Plan -> (ok, trajectory [K,11])."""
    args, keep = _scene_args(ref, static_poly, static_nv, dyn_time, dyn_samples, dyn_poly, dyn_nv)
    traj = np.zeros((512, 11))
    ok = C.c_int(0)
    K = dp_lib().ref_dp_plan(*args, float(x), float(y), float(theta), traj.ctypes.data, 512, C.byref(ok))
    return bool(ok.value), traj[:K].copy()


def check_optimization_collision(ref, static_poly, static_nv, dyn_time, dyn_samples, dyn_poly, dyn_nv, queries):
    args, keep = _scene_args(ref, static_poly, static_nv, dyn_time, dyn_samples, dyn_poly, dyn_nv)
    q = _f(queries)
    out = np.zeros(len(q), np.int32)
    dp_lib().ref_check_optimization_collision(*args, len(q), q.ctypes.data, out.ctypes.data)
    return out.astype(bool)


def road_barrier(ref):
    ref = _f(ref)
    cap = 2 * (int((ref[-1, 0] - ref[0, 0]) / 0.1) + 4)
    out = np.zeros((cap, 2))
    n = dp_lib().ref_road_barrier(len(ref), ref.ctypes.data, out.ctypes.data, cap)
    assert n >= 0
    return out[:n]


# ---- the reference's own CILQR solver (oracle/_ref/libcilqr_ref_solver.so, compiled against Eigen-lite) ------------
SOLVER_LIB_PATH = os.path.join(_HERE, "_ref", "libcilqr_ref_solver.so")
_slib = None


def solver_lib():
    global _slib
    if _slib is None:
        if not os.path.exists(SOLVER_LIB_PATH):
            subprocess.check_call(["make", "-C", _HERE, "_ref"], stdout=subprocess.DEVNULL)
        L = C.CDLL(SOLVER_LIB_PATH)
        d, i, p = C.c_double, C.c_int, C.c_void_p
        L.ref_ilqr_solve.argtypes = [d, i, i, i, i, p, p, p, p, p, p, p, p, p, p, p, i, p, p]
        L.ref_ilqr_solve_cfg.argtypes = [d, i, i, i, i, p, p, p, p, p, p, p, p, p, p, p, i, p, p, p]
        L.ref_dynamics.argtypes = [d, p, p, p]
        L.ref_dynamics.restype = None
        L.ref_dynamics_jacobian.argtypes = [d, p, p, p, p]
        L.ref_dynamics_jacobian.restype = None
        L.ref_barrier_value.argtypes = [d]
        L.ref_barrier_value.restype = d
        _slib = L
    return _slib


def ilqr_solve(batch, b: int, dt: float = 0.1, overrides=None):
    """The reference's IlqrOptimizer on scenario b of a ScenarioBatch -> dict(states, controls, init_states,
    init_controls, cost_hist [n,5], n_iter_trajs).  overrides = (max_iter_num, abs_cost_tol, rel_cost_tol)."""
    N, K = batch.N, batch.N + 1
    X, U, X0, U0 = np.zeros((K, 6)), np.zeros((N, 2)), np.zeros((K, 6)), np.zeros((N, 2))
    ch = np.zeros((1024, 5))
    nc, ni = C.c_int(), C.c_int()
    arrs = [_f(batch.start[b]), _f(batch.coarse[b]), _f(batch.corridor[b]),
            np.ascontiguousarray(batch.corridor_cnt[b], np.int32), _f(batch.lane_left[b]), _f(batch.lane_right[b])]
    ov = _f(overrides) if overrides is not None else None
    rc = solver_lib().ref_ilqr_solve_cfg(dt, N, batch.M_max, batch.lane_left.shape[1], batch.lane_right.shape[1],
                                         *[a.ctypes.data for a in arrs], X.ctypes.data, U.ctypes.data, X0.ctypes.data,
                                         U0.ctypes.data, ch.ctypes.data, len(ch), C.byref(nc), C.byref(ni),
                                         ov.ctypes.data if ov is not None else None)
    if rc != 0:
        raise RuntimeError(f"ref_ilqr_solve failed: {rc}")
    return dict(states=X, controls=U, init_states=X0, init_controls=U0, cost_hist=ch[:nc.value].copy(),
                n_iter_trajs=ni.value)


def dynamics(x, u, dt: float = 0.1):
    x, u, out = _f(x), _f(u), np.zeros(6)
    solver_lib().ref_dynamics(dt, x.ctypes.data, u.ctypes.data, out.ctypes.data)
    return out


def dynamics_jacobian(x, u, dt: float = 0.1):
    x, u, A, B = _f(x), _f(u), np.zeros((6, 6)), np.zeros((6, 2))
    solver_lib().ref_dynamics_jacobian(dt, x.ctypes.data, u.ctypes.data, A.ctypes.data, B.ctypes.data)
    return A, B


# ---- the reference's own Corridor (oracle/_ref/libcilqr_ref_corridor.so) -----------------------------------------
CORRIDOR_LIB_PATH = os.path.join(_HERE, "_ref", "libcilqr_ref_corridor.so")
_clib = None


def corridor_lib():
    global _clib
    if _clib is None:
        if not os.path.exists(CORRIDOR_LIB_PATH):
            subprocess.check_call(["make", "-C", _HERE, "_ref"], stdout=subprocess.DEVNULL)
        L = C.CDLL(CORRIDOR_LIB_PATH)
        d, i, p = C.c_double, C.c_int, C.c_void_p
        L.ref_build_corridor.argtypes = [d, d, d, i, p, p, p, i]
        L.ref_lane_constraints.argtypes = [i, p, i, p, i]
        _clib = L
    return _clib


def build_corridor(x, y, theta, obstacle_points, cap: int = 128):
    """The reference's AddCorridorPoints + BuildCorridor for one knot -> (m, constraints [m,3], polygon [m,2]);
    m = -1 when BuildCorridor returns false."""
    pts = _f(obstacle_points).reshape(-1, 2)
    cons, poly = np.zeros((cap, 3)), np.zeros((cap, 2))
    m = corridor_lib().ref_build_corridor(float(x), float(y), float(theta), len(pts), pts.ctypes.data,
                                          cons.ctypes.data, poly.ctypes.data, cap)
    return m, cons[:max(m, 0)].copy(), poly[:max(m, 0)].copy()


def lane_constraints(boundary, is_left: bool, cap: int = 1024):
    bd = _f(boundary).reshape(-1, 2)
    out = np.zeros((cap, 7))
    n = corridor_lib().ref_lane_constraints(len(bd), bd.ctypes.data, int(is_left), out.ctypes.data, cap)
    return n, out[:max(n, 0)].copy()
