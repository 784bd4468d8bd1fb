"""ctypes mirror of include/cilqr_b200.h.

``Solver.plan_batch`` corresponds to ``IlqrOptimizer::Plan`` (reference
``algorithm/ilqr/ilqr_optimizer.cc:53-95``) for a batch of scenarios in host memory;
``Solver.plan_batch_device`` takes raw device pointers (e.g. ``torch.Tensor.data_ptr()``).
No computation happens in Python and there is no CPU fallback: if the shared library cannot be
loaded or no sm_100 GPU is present the calls raise ``CilqrError``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

STATUS_NAMES = ["converged_abs", "converged_rel", "converged_grad", "lambda_overflow", "max_iter"]
E_INVALID, E_CUDA, E_NO_DEVICE, E_CAPACITY, E_SMEM, E_TIMEOUT, E_NCCL = -1, -2, -3, -4, -5, -6, -7

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class CilqrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cilqr_b200 error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    """POD mirror of CilqrParams (VehicleParam / IlqrConfig / Weights / barrier constants)."""
    _fields_ = [(n, C.c_double) for n in (
        "front_hang_length", "wheel_base", "rear_hang_length", "width",
        "max_velocity", "min_acceleration", "max_acceleration",
        "jerk_min", "jerk_max", "delta_min", "delta_max", "delta_rate_min", "delta_rate_max",
        "safe_margin",
        "w_jerk", "w_delta_rate", "w_x_target", "w_y_target", "w_theta", "w_v", "w_a", "w_delta",
        "abs_cost_tol", "rel_cost_tol", "barrier_t", "barrier_eps", "delta_t")] + [
        ("num_of_disc", C.c_int32), ("max_iter_num", C.c_int32)]


class BatchIn(C.Structure):
    _fields_ = [("B", C.c_int32), ("N", C.c_int32), ("M_max", C.c_int32), ("S_left", C.c_int32),
                ("S_right", C.c_int32), ("start", C.c_void_p), ("coarse", C.c_void_p),
                ("corridor", C.c_void_p), ("corridor_cnt", C.c_void_p), ("lane_left", C.c_void_p),
                ("lane_right", C.c_void_p), ("init_mode", C.c_int32), ("init_states", C.c_void_p),
                ("init_controls", C.c_void_p)]


class BatchOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "states", "controls", "status", "trajectory", "init_states", "init_controls", "cost_hist",
        "iter_states", "iter_controls", "hist_len")] + [("hist_cap", C.c_int32), ("result", C.c_void_p)]


class DebugOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "corridor", "lanes", "X0", "U0", "cost0", "A11", "Jx", "Ju", "Hx", "Hu", "Kg", "kg", "dV",
        "Xn", "Un", "costn", "nearest", "gnorm")]


class CorridorConfig(C.Structure):
    """POD mirror of CilqrCorridorConfig (CorridorConfig, planner_config.h:75-86)."""
    _fields_ = [(n, C.c_double) for n in ("max_diff_x", "max_diff_y", "radius", "max_axis_x", "max_axis_y",
                                          "lane_segment_length")] + [("point_cap", C.c_int32)]


class CorridorIn(C.Structure):
    _fields_ = [("B", C.c_int32), ("K", C.c_int32), ("P_max", C.c_int32), ("M_max", C.c_int32),
                ("traj", C.c_void_p), ("obs_points", C.c_void_p), ("obs_cnt", C.c_void_p)]


class CorridorOut(C.Structure):
    _fields_ = [("corridor", C.c_void_p), ("corridor_cnt", C.c_void_p), ("polygon", C.c_void_p),
                ("code", C.c_void_p)]


class DpConfig(C.Structure):
    """POD mirror of CilqrDpConfig (PlannerConfig planner_config.h:88-141 + VehicleParam fields)."""
    _fields_ = [(n, C.c_double) for n in (
        "tf", "delta_t", "dp_nominal_velocity", "dp_w_obstacle", "dp_w_lateral", "dp_w_lateral_change",
        "dp_w_lateral_velocity_change", "dp_w_longitudinal_velocity_bias", "dp_w_longitudinal_velocity_change",
        "max_velocity", "width", "wheel_base", "front_hang_length", "rear_hang_length")]


class TrackerConfig(C.Structure):
    """POD mirror of CilqrTrackerConfig (TrackerConfig planner_config.h:18-43 + VehicleParam fields)."""
    _fields_ = [(n, C.c_double) for n in (
        "sumulation_dt", "dt", "tolerance", "lat_weight_l", "lat_weight_theta", "lat_weight_delta", "lat_weight_delta_rate",
        "lat_preview_time", "lon_weight_s", "lon_weight_v", "lon_weight_a", "lon_weight_j", "wheel_base", "delta_min",
        "delta_max", "min_acceleration", "max_acceleration", "delta_rate_min", "delta_rate_max", "jerk_min", "jerk_max")] + [
        ("max_num_iteration", C.c_int32)]


class DpIn(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "R", "NB", "V", "n_static", "n_dyn", "T")] + [
        (n, C.c_void_p) for n in ("ref", "barrier", "start", "static_poly", "static_nv", "dyn_time", "dyn_samples",
                                  "dyn_poly", "dyn_nv")]


class DpOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("trajectory", "coarse", "xytheta", "ok", "cost", "waypoints")]


CORRIDOR_CODE_NAMES = ["ok", "no_points", "few_points", "origin", "capacity", "point_capacity"]

EXPORTS = ["cilqr_abi_version", "cilqr_default_params", "cilqr_create", "cilqr_destroy",
           "cilqr_plan_batch", "cilqr_plan_batch_device", "cilqr_synchronize",
           "cilqr_kernel_launches", "cilqr_last_kernel_ms", "cilqr_occupancy", "cilqr_strerror",
           "cilqr_last_cuda_error", "cilqr_debug_first_iteration", "cilqr_debug_stats",
           "cilqr_debug_completion_histogram", "cilqr_debug_host_path", "cilqr_corridor_default_config", "cilqr_corridor_batch",
           "cilqr_corridor_batch_device", "cilqr_lane_constraints", "cilqr_lane_constraints_device",
           "cilqr_corridor_last_kernel_ms", "cilqr_dp_default_config", "cilqr_dp_num_knots",
           "cilqr_dp_plan_batch", "cilqr_dp_plan_batch_device", "cilqr_dp_last_kernel_ms",
           "cilqr_tracker_default_config", "cilqr_tracker_batch", "cilqr_tracker_batch_device",
           "cilqr_tracker_last_kernel_ms", "cilqr_multi_create", "cilqr_multi_destroy", "cilqr_multi_devices", "cilqr_multi_shard_size",
           "cilqr_multi_last_error", "cilqr_plan_sharded"]

_libs = {}


def lib_path(variant: str = "") -> str:
    return _build.STRICT_LIB_PATH if variant == "strict" else _build.LIB_PATH


def load_library(build_if_missing: bool = True, variant: str = ""):
    """Loads libcilqr_b200.so (building it with nvcc when absent).  Raises if that fails.
    variant="strict" loads the parity instrument libcilqr_b200_strict.so instead (tests / bench parity legs only)."""
    if variant in _libs:
        return _libs[variant]
    path = lib_path(variant)
    if build_if_missing and _build.is_stale(path):
        _build.build_library(strict=variant == "strict")
    if not os.path.exists(path):
        raise CilqrError(E_NO_DEVICE, f"{path} is missing and could not be built; "
                         "the CUDA extension is required (no CPU fallback)")
    L = C.CDLL(path)
    L.cilqr_abi_version.restype = C.c_int
    L.cilqr_default_params.argtypes = [C.POINTER(Params)]
    L.cilqr_default_params.restype = None
    L.cilqr_create.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.POINTER(C.c_void_p)]
    L.cilqr_destroy.argtypes = [C.c_void_p]
    L.cilqr_destroy.restype = None
    L.cilqr_plan_batch.argtypes = [C.c_void_p, C.POINTER(BatchIn), C.POINTER(BatchOut)]
    L.cilqr_plan_batch_device.argtypes = [C.c_void_p, C.POINTER(BatchIn), C.POINTER(BatchOut), C.c_void_p]
    L.cilqr_synchronize.argtypes = [C.c_void_p]
    L.cilqr_kernel_launches.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.cilqr_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.cilqr_occupancy.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.cilqr_strerror.argtypes = [C.c_int]
    L.cilqr_strerror.restype = C.c_char_p
    L.cilqr_last_cuda_error.argtypes = [C.c_void_p]
    L.cilqr_last_cuda_error.restype = C.c_char_p
    L.cilqr_debug_first_iteration.argtypes = [C.c_void_p, C.POINTER(BatchIn), C.POINTER(DebugOut)]
    L.cilqr_debug_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.cilqr_debug_completion_histogram.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.cilqr_debug_host_path.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.cilqr_corridor_default_config.argtypes = [C.POINTER(CorridorConfig)]
    L.cilqr_corridor_default_config.restype = None
    L.cilqr_corridor_batch.argtypes = [C.c_void_p, C.POINTER(CorridorConfig), C.POINTER(CorridorIn),
                                       C.POINTER(CorridorOut)]
    L.cilqr_corridor_batch_device.argtypes = [C.c_void_p, C.POINTER(CorridorConfig), C.POINTER(CorridorIn),
                                              C.POINTER(CorridorOut), C.c_void_p]
    L.cilqr_lane_constraints.argtypes = [C.c_void_p, C.POINTER(CorridorConfig), C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
    L.cilqr_lane_constraints_device.argtypes = [C.c_void_p, C.POINTER(CorridorConfig), C.c_int, C.c_int, C.c_int,
                                                C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.cilqr_corridor_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.cilqr_dp_default_config.argtypes = [C.POINTER(DpConfig)]
    L.cilqr_dp_default_config.restype = None
    L.cilqr_dp_num_knots.argtypes = [C.POINTER(DpConfig)]
    L.cilqr_dp_plan_batch.argtypes = [C.c_void_p, C.POINTER(DpConfig), C.POINTER(DpIn), C.POINTER(DpOut)]
    L.cilqr_dp_plan_batch_device.argtypes = [C.c_void_p, C.POINTER(DpConfig), C.POINTER(DpIn), C.POINTER(DpOut),
                                             C.c_void_p]
    L.cilqr_dp_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.cilqr_tracker_default_config.argtypes = [C.POINTER(TrackerConfig)]
    L.cilqr_tracker_default_config.restype = None
    L.cilqr_tracker_batch.argtypes = [C.c_void_p, C.POINTER(TrackerConfig), C.c_int, C.c_int] + [C.c_void_p] * 6
    L.cilqr_tracker_batch_device.argtypes = [C.c_void_p, C.POINTER(TrackerConfig), C.c_int, C.c_int] + [C.c_void_p] * 7
    L.cilqr_tracker_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.cilqr_multi_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_void_p)]
    L.cilqr_multi_destroy.argtypes = [C.c_void_p]
    L.cilqr_multi_destroy.restype = None
    L.cilqr_multi_devices.argtypes = [C.c_void_p]
    L.cilqr_multi_shard_size.argtypes = [C.c_void_p, C.c_int]
    L.cilqr_multi_last_error.argtypes = [C.c_void_p]
    L.cilqr_multi_last_error.restype = C.c_char_p
    L.cilqr_plan_sharded.argtypes = [C.c_void_p, C.POINTER(BatchIn), C.POINTER(BatchOut), C.POINTER(C.c_void_p)]
    _libs[variant] = L
    return L


def default_params() -> Params:
    p = Params()
    load_library().cilqr_default_params(C.byref(p))
    return p


def default_corridor_config(point_cap: int = 0) -> CorridorConfig:
    c = CorridorConfig()
    load_library().cilqr_corridor_default_config(C.byref(c))
    c.point_cap = point_cap
    return c


def default_dp_config() -> DpConfig:
    c = DpConfig()
    load_library().cilqr_dp_default_config(C.byref(c))
    return c


def dp_num_knots(cfg: DpConfig | None = None) -> int:
    cfg = cfg or default_dp_config()
    return load_library().cilqr_dp_num_knots(C.byref(cfg))


def _ptr(x):
    """numpy array / torch tensor / int / None -> address."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()  # torch.Tensor


class Solver:
    """Owns one cilqr_handle (device buffers + streams) on ``device``."""

    def __init__(self, params: Params | None = None, device: int = 0, N_max: int = 200,
                 M_max: int = 32, S_max: int = 64, B_max: int = 1 << 21, variant: str = ""):
        self._L = load_library(variant=variant)
        self.params = params or default_params()
        h = C.c_void_p()
        rc = self._L.cilqr_create(C.byref(self.params), device, N_max, M_max, S_max, B_max, C.byref(h))
        if rc != 0:
            raise CilqrError(rc, self._L.cilqr_strerror(rc).decode())
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.cilqr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            msg = self._L.cilqr_strerror(rc).decode()
            if rc in (E_CUDA, E_TIMEOUT):
                msg += " -- " + self._L.cilqr_last_cuda_error(self._h).decode()
            raise CilqrError(rc, msg)

    @staticmethod
    def _make_in(B, N, M_max, S_left, S_right, start, coarse, corridor, corridor_cnt, lane_left, lane_right):
        return BatchIn(B, N, M_max, S_left, S_right, _ptr(start), _ptr(coarse), _ptr(corridor),
                       _ptr(corridor_cnt), _ptr(lane_left), _ptr(lane_right))

    def plan_batch(self, batch, trajectory: bool = False, init_guess: bool = False, hist_cap: int = 0,
                   out: dict | None = None, result: bool = False, init_mode: int = 0, init_states=None,
                   init_controls=None) -> dict:
        """Host path.  ``batch`` is a ScenarioBatch-like object with numpy arrays (pinned memory
        gives asynchronous copies).  Returns numpy outputs (or fills the arrays given in ``out``).
        init_mode 1: open-loop rollout of ``init_controls`` [B,N,2]; 2: (``init_states`` [B,K,6], ``init_controls``)
        as the initial guess instead of iqr (ilqr_optimizer.cc:168-169)."""
        B, N, K = batch.B, batch.N, batch.N + 1
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        arrs = [f64(batch.start), f64(batch.coarse), f64(batch.corridor),
                np.ascontiguousarray(batch.corridor_cnt, dtype=np.int32), f64(batch.lane_left), f64(batch.lane_right)]
        bi = self._make_in(B, N, batch.M_max, arrs[4].shape[1], arrs[5].shape[1], *arrs)
        if init_mode:
            gx = f64(init_states) if init_states is not None else None
            gu = f64(init_controls)
            arrs += [gx, gu]  # keep alive
            bi.init_mode, bi.init_states, bi.init_controls = init_mode, _ptr(gx), _ptr(gu)
        o = out if out is not None else {}
        o.setdefault("states", np.empty((B, K, 6)))
        o.setdefault("controls", np.empty((B, N, 2)))
        o.setdefault("status", np.empty((B, 8)))
        if trajectory:
            o.setdefault("trajectory", np.empty((B, K, 13)))
        if result:
            o.setdefault("result", np.empty((B, K, 13)))
        if init_guess:
            o.setdefault("init_states", np.empty((B, K, 6)))
            o.setdefault("init_controls", np.empty((B, N, 2)))
        if hist_cap > 0:
            o.setdefault("cost_hist", np.zeros((B, hist_cap, 5)))
            o.setdefault("iter_states", np.zeros((B, hist_cap, K, 6)))
            o.setdefault("iter_controls", np.zeros((B, hist_cap, N, 2)))
            o.setdefault("hist_len", np.zeros((B, 2), dtype=np.int32))
        bo = BatchOut(*[_ptr(o.get(n)) for n in ("states", "controls", "status", "trajectory", "init_states",
                                                  "init_controls", "cost_hist", "iter_states", "iter_controls",
                                                  "hist_len")], hist_cap, _ptr(o.get("result")))
        self._check(self._L.cilqr_plan_batch(self._h, C.byref(bi), C.byref(bo)))
        return o

    def plan_batch_device(self, B, N, M_max, S_left, S_right, start, coarse, corridor, corridor_cnt,
                          lane_left, lane_right, states, controls, status, trajectory=None,
                          init_states=None, init_controls=None, stream: int | None = None, result=None):
        """Device path: every array argument is a device pointer (int) or a CUDA torch tensor.
        Enqueues the solve kernel and returns immediately."""
        bi = self._make_in(B, N, M_max, S_left, S_right, start, coarse, corridor, corridor_cnt, lane_left, lane_right)
        bo = BatchOut(_ptr(states), _ptr(controls), _ptr(status), _ptr(trajectory), _ptr(init_states),
                      _ptr(init_controls), None, None, None, None, 0, _ptr(result))
        self._check(self._L.cilqr_plan_batch_device(self._h, C.byref(bi), C.byref(bo), stream))

    def debug_first_iteration(self, B, N, M_max, S_left, S_right, start, coarse, corridor, corridor_cnt,
                              lane_left, lane_right, outs: dict):
        bi = self._make_in(B, N, M_max, S_left, S_right, start, coarse, corridor, corridor_cnt, lane_left, lane_right)
        d = DebugOut(*[_ptr(outs.get(n)) for n, _ in DebugOut._fields_])
        self._check(self._L.cilqr_debug_first_iteration(self._h, C.byref(bi), C.byref(d)))

    def synchronize(self):
        self._check(self._L.cilqr_synchronize(self._h))

    def debug_host_path(self, watchdog_ms: int = 0, starve_after: int = -1):
        """Test hook: shorten the input watchdog / withhold the input watermark beyond ``starve_after``."""
        self._check(self._L.cilqr_debug_host_path(self._h, watchdog_ms, starve_after))

    def kernel_launches(self) -> int:
        n = C.c_int64()
        self._check(self._L.cilqr_kernel_launches(self._h, C.byref(n)))
        return n.value

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        self._check(self._L.cilqr_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def debug_stats(self) -> dict:
        """Scheduler counters of the last launch (summed over warps)."""
        a = (C.c_uint64 * 8)()
        self._check(self._L.cilqr_debug_stats(self._h, a))
        k = ["passes", "idle_polls", "failed_claims", "init", "back", "roll", "eval", "type_switches"]
        return dict(zip(k, [int(x) for x in a]))

    def completion_histogram(self):
        """Scenarios completed per 2 ms bucket since the start of the last launch."""
        a = (C.c_uint64 * 256)()
        self._check(self._L.cilqr_debug_completion_histogram(self._h, a))
        return [int(x) for x in a]

    def occupancy(self, N: int, S_left: int, S_right: int):
        w, s = C.c_int(), C.c_int()
        self._check(self._L.cilqr_occupancy(self._h, N, S_left, S_right, C.byref(w), C.byref(s)))
        return w.value, s.value

    # ---- corridor builder (Corridor::Plan, algorithm/ilqr/corridor.cc:17-54) --------------------------
    def corridor_batch(self, traj, obs_points, obs_cnt, M_max: int, polygon: bool = False,
                       cfg: CorridorConfig | None = None) -> dict:
        """Host path of BuildCorridorConstraints for a batch: numpy in, numpy out."""
        cfg = cfg or default_corridor_config()
        traj = np.ascontiguousarray(traj, np.float64)
        pts = np.ascontiguousarray(obs_points, np.float64)
        cnt = np.ascontiguousarray(obs_cnt, np.int32)
        B, K, P = pts.shape[:3]
        o = {"corridor": np.zeros((B, K, M_max, 3)), "corridor_cnt": np.zeros((B, K), np.int32),
             "code": np.zeros((B, K), np.int32)}
        if polygon:
            o["polygon"] = np.zeros((B, K, M_max, 2))
        ci = CorridorIn(B, K, P, M_max, _ptr(traj), _ptr(pts), _ptr(cnt))
        co = CorridorOut(_ptr(o["corridor"]), _ptr(o["corridor_cnt"]), _ptr(o.get("polygon")), _ptr(o["code"]))
        self._check(self._L.cilqr_corridor_batch(self._h, C.byref(cfg), C.byref(ci), C.byref(co)))
        return o

    def corridor_batch_device(self, B, K, P_max, M_max, traj, obs_points, obs_cnt, corridor, corridor_cnt, code,
                              polygon=None, cfg: CorridorConfig | None = None, stream: int | None = None):
        """Device path: pointers / CUDA tensors; enqueues the build kernel and returns."""
        cfg = cfg or default_corridor_config()
        ci = CorridorIn(B, K, P_max, M_max, _ptr(traj), _ptr(obs_points), _ptr(obs_cnt))
        co = CorridorOut(_ptr(corridor), _ptr(corridor_cnt), _ptr(polygon), _ptr(code))
        self._check(self._L.cilqr_corridor_batch_device(self._h, C.byref(cfg), C.byref(ci), C.byref(co), stream))

    def lane_constraints(self, boundary, is_left: bool, S_max: int, cfg: CorridorConfig | None = None):
        """CalLeft/RightLaneConstraints for B boundary polylines [B,n,2] -> ([B,S_max,7], count [B])."""
        cfg = cfg or default_corridor_config()
        bd = np.ascontiguousarray(boundary, np.float64)
        B, n = bd.shape[:2]
        out = np.zeros((B, S_max, 7))
        cnt = np.zeros(B, np.int32)
        self._check(self._L.cilqr_lane_constraints(self._h, C.byref(cfg), B, n, S_max, int(is_left), _ptr(bd),
                                                   _ptr(out), _ptr(cnt)))
        return out, cnt

    def lane_constraints_device(self, B, n, S_max, is_left, boundary, out, count,
                                cfg: CorridorConfig | None = None, stream: int | None = None):
        cfg = cfg or default_corridor_config()
        self._check(self._L.cilqr_lane_constraints_device(self._h, C.byref(cfg), B, n, S_max, int(is_left),
                                                          _ptr(boundary), _ptr(out), _ptr(count), stream))

    def corridor_last_kernel_ms(self) -> float:
        ms = C.c_float()
        self._check(self._L.cilqr_corridor_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    # ---- coarse DP planner (DpPlanner::Plan, algorithm/planner/dp_planner.cpp:135-281) -----------------
    def dp_plan_batch(self, dpb, barrier, cfg: DpConfig | None = None, waypoints: bool = False) -> dict:
        """Host path: ``dpb`` is a scenarios.DpBatch-like object, ``barrier`` the road barrier [NB,2] sorted by x."""
        cfg = cfg or default_dp_config()
        K = dp_num_knots(cfg)
        if K < 2:
            raise CilqrError(E_INVALID, "bad DP configuration (tf / delta_t)")
        B = dpb.start.shape[0]
        f64 = lambda a: np.ascontiguousarray(a, np.float64)  # noqa: E731
        i32 = lambda a: np.ascontiguousarray(a, np.int32)  # noqa: E731
        arrs = [f64(dpb.ref), f64(barrier), f64(dpb.start), f64(dpb.static_poly), i32(dpb.static_nv), f64(dpb.dyn_time),
                i32(dpb.dyn_samples), f64(dpb.dyn_poly), i32(dpb.dyn_nv)]
        V = arrs[3].shape[2] if arrs[3].ndim == 4 and arrs[3].shape[1] > 0 else (arrs[7].shape[3] if arrs[7].ndim == 5 else 4)
        di = DpIn(B, arrs[0].shape[0], arrs[1].shape[0], V, arrs[3].shape[1], arrs[7].shape[1],
                  arrs[7].shape[2] if arrs[7].ndim == 5 else 0, *[_ptr(a) for a in arrs])
        o = {"trajectory": np.zeros((B, K, 13)), "coarse": np.zeros((B, K, 6)), "xytheta": np.zeros((B, K, 3)),
             "ok": np.zeros(B, np.int32), "cost": np.zeros(B)}
        if waypoints:
            o["waypoints"] = np.zeros((B, 5, 3))
        do = DpOut(*[_ptr(o.get(n)) for n in ("trajectory", "coarse", "xytheta", "ok", "cost", "waypoints")])
        self._check(self._L.cilqr_dp_plan_batch(self._h, C.byref(cfg), C.byref(di), C.byref(do)))
        return o

    def dp_plan_batch_device(self, B, R, NB, V, n_static, n_dyn, T, ref, barrier, start, static_poly, static_nv,
                             dyn_time, dyn_samples, dyn_poly, dyn_nv, ok, trajectory=None, coarse=None, xytheta=None,
                             cost=None, waypoints=None, cfg: DpConfig | None = None, stream: int | None = None):
        """Device path: pointers / CUDA tensors; enqueues the planner kernel and returns."""
        cfg = cfg or default_dp_config()
        di = DpIn(B, R, NB, V, n_static, n_dyn, T, *[_ptr(a) for a in (ref, barrier, start, static_poly, static_nv,
                                                                        dyn_time, dyn_samples, dyn_poly, dyn_nv)])
        do = DpOut(_ptr(trajectory), _ptr(coarse), _ptr(xytheta), _ptr(ok), _ptr(cost), _ptr(waypoints))
        self._check(self._L.cilqr_dp_plan_batch_device(self._h, C.byref(cfg), C.byref(di), C.byref(do), stream))

    # ---- tracker initial guess (Tracker::Plan, algorithm/ilqr/tracker.cc:169-215 + InitGuess) ----------------------
    def tracker_batch(self, start4, coarse_traj, cfg: "TrackerConfig | None" = None) -> dict:
        """Host path: start4 [B,4], coarse_traj [B,K,13] (TrajectoryPoint records) -> trajectory [B,K,13],
        guess_states [B,K,6], guess_controls [B,K-1,2], ok [B]."""
        if cfg is None:
            cfg = TrackerConfig()
            self._L.cilqr_tracker_default_config(C.byref(cfg))
        st = np.ascontiguousarray(start4, np.float64)
        co = np.ascontiguousarray(coarse_traj, np.float64)
        B, K = co.shape[:2]
        o = {"trajectory": np.zeros((B, K, 13)), "guess_states": np.zeros((B, K, 6)), "guess_controls": np.zeros((B, K - 1, 2)),
             "ok": np.zeros(B, np.int32)}
        self._check(self._L.cilqr_tracker_batch(self._h, C.byref(cfg), B, K, _ptr(st), _ptr(co), _ptr(o["trajectory"]),
                                                _ptr(o["guess_states"]), _ptr(o["guess_controls"]), _ptr(o["ok"])))
        return o

    def tracker_batch_device(self, B, K, start4, coarse_traj, ok, traj=None, guess_states=None, guess_controls=None,
                             cfg: "TrackerConfig | None" = None, stream: int | None = None):
        if cfg is None:
            cfg = TrackerConfig()
            self._L.cilqr_tracker_default_config(C.byref(cfg))
        self._check(self._L.cilqr_tracker_batch_device(self._h, C.byref(cfg), B, K, _ptr(start4), _ptr(coarse_traj), _ptr(traj),
                                                       _ptr(guess_states), _ptr(guess_controls), _ptr(ok), stream))

    def tracker_last_kernel_ms(self) -> float:
        ms = C.c_float()
        self._check(self._L.cilqr_tracker_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def dp_last_kernel_ms(self) -> float:
        ms = C.c_float()
        self._check(self._L.cilqr_dp_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value


class MultiSolver:
    """Owns one cilqr_multi (one solver handle per GPU of this box, one host process): ``plan_sharded`` solves a host
    batch as contiguous shards, optionally leaving every shard's result block on every GPU (one NCCL all-gather)."""

    def __init__(self, n_devices: int, params: Params | None = None, devices=None, N_max: int = 200, M_max: int = 32,
                 S_max: int = 64, B_max_per_device: int = 1 << 20):
        self._L = load_library()
        self.params = params or default_params()
        h = C.c_void_p()
        devs = (C.c_int * n_devices)(*devices) if devices is not None else None
        rc = self._L.cilqr_multi_create(C.byref(self.params), n_devices, devs, N_max, M_max, S_max, B_max_per_device,
                                        C.byref(h))
        if rc != 0:
            raise CilqrError(rc, self._L.cilqr_strerror(rc).decode())
        self._h = h
        self.n_devices = n_devices

    def close(self):
        if getattr(self, "_h", None):
            self._L.cilqr_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shard_size(self, B: int) -> int:
        return self._L.cilqr_multi_shard_size(self._h, B)

    def plan_sharded(self, batch, gathered=None, result: bool = False) -> dict:
        """``gathered``: optional list of n_devices device pointers / CUDA tensors (one per GPU, each
        n_devices * shard_size * (6K + 2N + 8) doubles) that receive all result blocks."""
        B, N, K = batch.B, batch.N, batch.N + 1
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        arrs = [f64(batch.start), f64(batch.coarse), f64(batch.corridor),
                np.ascontiguousarray(batch.corridor_cnt, dtype=np.int32), f64(batch.lane_left), f64(batch.lane_right)]
        bi = BatchIn(B, N, batch.M_max, arrs[4].shape[1], arrs[5].shape[1], *[_ptr(a) for a in arrs])
        o = {"states": np.empty((B, K, 6)), "controls": np.empty((B, N, 2)), "status": np.empty((B, 8))}
        if result:
            o["result"] = np.empty((B, K, 13))
        bo = BatchOut(_ptr(o["states"]), _ptr(o["controls"]), _ptr(o["status"]), None, None, None, None, None, None, None, 0,
                      _ptr(o.get("result")))
        g = None
        if gathered is not None:
            g = (C.c_void_p * self.n_devices)(*[_ptr(x) for x in gathered])
        rc = self._L.cilqr_plan_sharded(self._h, C.byref(bi), C.byref(bo), g)
        if rc != 0:
            raise CilqrError(rc, self._L.cilqr_strerror(rc).decode() + " -- " +
                             self._L.cilqr_multi_last_error(self._h).decode())
        return o
