"""ctypes binding of the CPU oracle (oracle/cilqr_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
``--impl reference`` legs of bench.py.  Nothing under cilqr_b200/ imports this module.
Parity: pinned bit for bit against the reference's own solver source compiled with an Eigen stand-in (see the
header of cilqr_oracle.h and tests/test_reference_pins.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcilqr_oracle.so")
_MAKE_TARGET = "libcilqr_oracle.so"  # (oracle/binding_pm.py re-executes this module for libcilqr_oracle_pm.so)

STATUS_NAMES = ["converged_abs", "converged_rel", "converged_grad", "lambda_overflow", "max_iter"]


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "front_hang_length", "wheel_base", "rear_hang_length", "width",
        "max_velocity", "min_acceleration", "max_acceleration",
        "jerk_min", "jerk_max", "delta_min", "delta_max", "delta_rate_min", "delta_rate_max",
        "safe_margin",
        "w_jerk", "w_delta_rate", "w_x_target", "w_y_target", "w_theta", "w_v", "w_a", "w_delta",
        "abs_cost_tol", "rel_cost_tol", "barrier_t", "barrier_eps", "delta_t")] + [
        ("num_of_disc", C.c_int), ("max_iter_num", C.c_int)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class Problem(C.Structure):
    _fields_ = [("N", C.c_int), ("M_max", C.c_int), ("S_left", C.c_int), ("S_right", C.c_int),
                ("start", _dp), ("coarse", _dp), ("corridor", _dp), ("corridor_cnt", _ip),
                ("lane_left", _dp), ("lane_right", _dp), ("init_mode", C.c_int), ("init_states", _dp),
                ("init_controls", _dp)]


class Result(C.Structure):
    _fields_ = [("states", _dp), ("controls", _dp), ("init_states", _dp), ("init_controls", _dp),
                ("status", C.c_int), ("iters", C.c_int), ("accepted", C.c_int),
                ("alpha_hash", C.c_uint), ("cost", C.c_double * 5), ("cost_init", C.c_double * 5),
                ("lambda_", C.c_double), ("trace", _dp), ("trace_cap", C.c_int),
                ("trace_len", C.c_int), ("cost_hist", _dp), ("cost_hist_cap", C.c_int),
                ("cost_hist_len", C.c_int)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cilqr_oracle.c")
    deps = [src, os.path.join(_HERE, "cilqr_oracle.h"), os.path.join(_HERE, "..", "cilqr_b200", "csrc", "pm_math.h")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["make", "-C", _HERE, "-B", _MAKE_TARGET], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.cilqr_oracle_default_params.argtypes = [C.POINTER(Params)]
        L.cilqr_oracle_solve.argtypes = [C.POINTER(Params), C.POINTER(Problem), C.POINTER(Result)]
        L.cilqr_oracle_solve.restype = C.c_int
        L.cilqr_oracle_solve_batch.argtypes = [C.POINTER(Params)] + [C.c_int] * 5 + [
            _dp, _dp, _dp, _ip, _dp, _dp, _dp, _dp, _dp, C.c_int]
        L.cilqr_oracle_solve_batch.restype = C.c_int
        L.cilqr_oracle_solve_batch_init.argtypes = [C.POINTER(Params)] + [C.c_int] * 5 + [
            _dp, _dp, _dp, _ip, _dp, _dp, C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_int]
        L.cilqr_oracle_solve_batch_init.restype = C.c_int
        L.cilqr_oracle_normalize_angle.argtypes = [C.c_double]
        L.cilqr_oracle_normalize_angle.restype = C.c_double
        L.cilqr_oracle_dynamics.argtypes = [C.POINTER(Params), _dp, _dp, _dp]
        L.cilqr_oracle_dynamics_jacobian.argtypes = [C.POINTER(Params), _dp, _dp, _dp, _dp]
        for f in ("barrier_value", "barrier_dcoef"):
            fn = getattr(L, "cilqr_oracle_" + f)
            fn.argtypes = [C.POINTER(Params), C.c_double]
            fn.restype = C.c_double
        L.cilqr_oracle_barrier_hcoef.argtypes = [C.POINTER(Params), C.c_double, _dp, _dp]
        L.cilqr_oracle_segment_distance.argtypes = [_dp, C.c_double, C.c_double]
        L.cilqr_oracle_segment_distance.restype = C.c_double
        L.cilqr_oracle_disc_radius.argtypes = [C.POINTER(Params)]
        L.cilqr_oracle_disc_radius.restype = C.c_double
        L.cilqr_oracle_ctx_create.argtypes = [C.POINTER(Params), C.POINTER(Problem)]
        L.cilqr_oracle_ctx_create.restype = C.c_void_p
        L.cilqr_oracle_ctx_destroy.argtypes = [C.c_void_p]
        L.cilqr_oracle_ctx_constraints.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.cilqr_oracle_ctx_iqr.argtypes = [C.c_void_p, _dp, _dp]
        L.cilqr_oracle_ctx_total_cost.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.cilqr_oracle_ctx_total_cost.restype = C.c_double
        L.cilqr_oracle_ctx_linearize.argtypes = [C.c_void_p] + [_dp] * 8
        L.cilqr_oracle_ctx_backward.argtypes = [C.c_void_p, C.c_double, _dp, _dp, _dp]
        L.cilqr_oracle_ctx_forward.argtypes = [C.c_void_p, C.c_double, _dp, _dp, _dp, _dp]
        L.cilqr_oracle_ctx_nearest.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.cilqr_oracle_ctx_nearest.restype = C.c_int
        _lib = L
    return _lib


def default_params() -> Params:
    p = Params()
    lib().cilqr_oracle_default_params(C.byref(p))
    return p


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def make_problem(batch, b: int):
    """-> (Problem, keepalive) for scenario b of a ScenarioBatch-like object."""
    arrs = dict(start=_f64(batch.start[b]), coarse=_f64(batch.coarse[b]),
                corridor=_f64(batch.corridor[b]),
                cnt=np.ascontiguousarray(batch.corridor_cnt[b], dtype=np.int32),
                ll=_f64(batch.lane_left[b]), lr=_f64(batch.lane_right[b]))
    pb = Problem(batch.N, batch.M_max, arrs["ll"].shape[0], arrs["lr"].shape[0], _d(arrs["start"]),
                 _d(arrs["coarse"]), _d(arrs["corridor"]), arrs["cnt"].ctypes.data_as(_ip),
                 _d(arrs["ll"]), _d(arrs["lr"]))
    return pb, arrs


def solve(batch, b: int, params: Params | None = None, trace: bool = False, hist: bool = False):
    """Full solve of one scenario; returns a dict of numpy results."""
    p = params or default_params()
    pb, keep = make_problem(batch, b)
    K, N = batch.N + 1, batch.N
    out = dict(states=np.zeros((K, 6)), controls=np.zeros((N, 2)), init_states=np.zeros((K, 6)),
               init_controls=np.zeros((N, 2)))
    r = Result()
    r.states, r.controls = _d(out["states"]), _d(out["controls"])
    r.init_states, r.init_controls = _d(out["init_states"]), _d(out["init_controls"])
    cap = p.max_iter_num + 2
    tr = np.zeros((cap, 8))
    ch = np.zeros((cap, 5))
    if trace:
        r.trace, r.trace_cap = _d(tr), cap
    if hist:
        r.cost_hist, r.cost_hist_cap = _d(ch), cap
    rc = lib().cilqr_oracle_solve(C.byref(p), C.byref(pb), C.byref(r))
    out.update(rc=rc, status=r.status, iters=r.iters, accepted=r.accepted, alpha_hash=r.alpha_hash,
               cost=np.array(r.cost[:]), cost_init=np.array(r.cost_init[:]), lam=r.lambda_)
    if trace:
        out["trace"] = tr[:r.trace_len].copy()
    if hist:
        out["cost_hist"] = ch[:r.cost_hist_len].copy()
    del keep
    return out


def solve_batch(batch, params: Params | None = None, nthreads: int = 1, init_mode: int = 0, init_states=None,
                init_controls=None):
    """-> states[B,K,6], controls[B,N,2], status[B,8] (status, iters, cost5, alpha_hash), n_converged.
    init_mode 1: open-loop rollout of init_controls; 2: (init_states, init_controls) as the initial guess."""
    p = params or default_params()
    B, K, N = batch.B, batch.N + 1, batch.N
    states = np.zeros((B, K, 6))
    controls = np.zeros((B, N, 2))
    status = np.zeros((B, 8))
    start, coarse, corridor = _f64(batch.start), _f64(batch.coarse), _f64(batch.corridor)
    cnt = np.ascontiguousarray(batch.corridor_cnt, dtype=np.int32)
    ll, lr = _f64(batch.lane_left), _f64(batch.lane_right)
    ix = _f64(init_states) if init_states is not None else None
    iu = _f64(init_controls) if init_controls is not None else None
    conv = lib().cilqr_oracle_solve_batch_init(
        C.byref(p), B, N, batch.M_max, ll.shape[1], lr.shape[1], _d(start), _d(coarse), _d(corridor),
        cnt.ctypes.data_as(_ip), _d(ll), _d(lr), init_mode, _d(ix) if ix is not None else None,
        _d(iu) if iu is not None else None, _d(states), _d(controls), _d(status), nthreads)
    return states, controls, status, conv


class Ctx:
    """Stage-level access (mirrors the private members of IlqrOptimizer)."""

    def __init__(self, batch, b: int, params: Params | None = None):
        self.p = params or default_params()
        self.pb, self._keep = make_problem(batch, b)
        self.N, self.K, self.M_max = batch.N, batch.N + 1, batch.M_max
        self.S = (self.pb.S_left, self.pb.S_right)
        self.h = lib().cilqr_oracle_ctx_create(C.byref(self.p), C.byref(self.pb))

    def close(self):
        if self.h:
            lib().cilqr_oracle_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def constraints(self):
        cor = np.zeros((self.K, self.M_max, 3))
        ll = np.zeros((self.S[0], 3))
        lr = np.zeros((self.S[1], 3))
        lib().cilqr_oracle_ctx_constraints(self.h, _d(cor), _d(ll), _d(lr))
        return cor, ll, lr

    def iqr(self):
        X, U = np.zeros((self.K, 6)), np.zeros((self.N, 2))
        lib().cilqr_oracle_ctx_iqr(self.h, _d(X), _d(U))
        return X, U

    def total_cost(self, X, U):
        c5 = np.zeros(5)
        X, U = _f64(X), _f64(U)
        lib().cilqr_oracle_ctx_total_cost(self.h, _d(X), _d(U), _d(c5))
        return c5

    def linearize(self, X, U):
        N, K = self.N, self.K
        X, U = _f64(X), _f64(U)
        o = dict(A=np.zeros((N, 6, 6)), B=np.zeros((N, 6, 2)), Jx=np.zeros((K, 6)),
                 Ju=np.zeros((N, 2)), Hx=np.zeros((K, 6, 6)), Hu=np.zeros((N, 2, 2)))
        lib().cilqr_oracle_ctx_linearize(self.h, _d(X), _d(U), _d(o["A"]), _d(o["B"]), _d(o["Jx"]),
                                         _d(o["Ju"]), _d(o["Hx"]), _d(o["Hu"]))
        return o

    def backward(self, lam: float):
        Ks, ks, dV = np.zeros((self.N, 2, 6)), np.zeros((self.N, 2)), np.zeros(2)
        lib().cilqr_oracle_ctx_backward(self.h, lam, _d(Ks), _d(ks), _d(dV))
        return Ks, ks, dV

    def forward(self, alpha: float, X, U):
        X, U = _f64(X), _f64(U)
        Xn, Un = np.zeros_like(X), np.zeros_like(U)
        lib().cilqr_oracle_ctx_forward(self.h, alpha, _d(X), _d(U), _d(Xn), _d(Un))
        return Xn, Un

    def nearest(self, side: int, x: float, y: float) -> int:
        return lib().cilqr_oracle_ctx_nearest(self.h, side, x, y)
