"""The DEVICE code of the corridor and DP-planner kernels (cilqr_b200/csrc/corridor_kernel.cuh, dp_kernel.cuh),
compiled for the HOST by the development harnesses tools/corridor_host_emul.cc and tools/dp_host_emul.cc (CUDA
qualifiers and intrinsics shimmed, one "thread" per CTA) and compared bit for bit with the oracles.

This is not a CPU path of the product -- the library has none -- but a way to catch a logic regression in the
kernels' source in a container without a GPU; the GPU tests (tests/test_gpu_corridor.py, tests/test_gpu_dp.py) remain
the parity tests proper."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from cilqr_b200 import scenarios

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, name):
    so = str(tmp_path / f"lib{name}.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tools", f"{name}.cc")])
    return C.CDLL(so)


def _p(a):
    return C.c_void_p(a.ctypes.data)


def test_corridor_device_code_equals_the_oracle(tmp_path):
    from oracle import corridor_binding as cb
    L = _build(tmp_path, "corridor_host_emul")
    for seed, B, N, n_obs, cap in ((23, 6, 40, 11, 65), (22, 4, 60, 20, 105)):
        _, ci = scenarios.generate_with_obstacles(seed, 0, B, N=N, n_obs=n_obs)
        M = 24
        cor, cnt, poly, code = cb.plan_batch(ci.traj, ci.obs_points, ci.obs_cnt, M)
        K, P = ci.K, ci.P_max
        cor2, cnt2 = np.zeros((B, K, M, 3)), np.zeros((B, K), np.int32)
        poly2, code2 = np.zeros((B, K, M, 2)), np.zeros((B, K), np.int32)
        L.emul_corridor(B, K, P, M, cap, _p(ci.traj), _p(ci.obs_points), _p(ci.obs_cnt), _p(cor2), _p(cnt2), _p(poly2),
                        _p(code2))
        assert not code2.any() and np.array_equal(cnt, cnt2)
        m = np.arange(M)[None, None, :] < cnt[..., None]
        assert np.array_equal(cor[m], cor2[m]) and np.array_equal(poly[m], poly2[m])


def test_dp_device_code_equals_the_oracle(tmp_path):
    from oracle import dp_binding as dp
    L = _build(tmp_path, "dp_host_emul")
    B = 3
    db = scenarios.generate_dp(5, B)
    barrier = dp.build_barrier(db.ref)
    cfg = dp.default_config()
    K = dp.lib().dp_num_knots(cfg)
    cf = np.array([getattr(cfg, n) for n, _ in dp.Config._fields_])
    dims = np.array([B, len(db.ref), len(barrier), 4, db.static_poly.shape[1], db.dyn_poly.shape[1], db.dyn_poly.shape[2]],
                    np.int32)
    traj, coarse, xyt = np.zeros((B, K, 13)), np.zeros((B, K, 6)), np.zeros((B, K, 3))
    ok, cost, wp = np.zeros(B, np.int32), np.zeros(B), np.zeros((B, 5, 3))
    L.emul_dp(_p(cf), _p(dims), _p(db.ref), _p(barrier), _p(db.start), _p(db.static_poly), _p(db.static_nv),
              _p(db.dyn_time), _p(db.dyn_samples), _p(db.dyn_poly), _p(db.dyn_nv), _p(traj), _p(coarse), _p(xyt), _p(ok),
              _p(cost), _p(wp))
    for b in range(B):
        sc = dp.Scene(db.ref, barrier, db.static_poly[b], db.static_nv[b], db.dyn_time[b], db.dyn_samples[b],
                      db.dyn_poly[b], db.dyn_nv[b])
        o_ok, o_traj, o_cost, o_wp = dp.plan(sc, *db.start[b], cfg)
        assert bool(ok[b]) == o_ok and cost[b] == o_cost and np.array_equal(wp[b], o_wp)
        assert np.array_equal(traj[b], o_traj, equal_nan=True)
        assert np.array_equal(coarse[b][:, :3], o_traj[:, 2:5], equal_nan=True)


def test_tracker_device_code_equals_the_oracle(tmp_path):
    """tracker_kernel (cilqr_b200/csrc/tracker_kernel.cuh) compiled for the host against oracle/tracker_oracle.c, which is
    itself bit-identical to the reference's own Tracker (tests/test_reference_pins.py): trajectory, InitGuess copy, ok."""
    from oracle import dp_binding as dp
    from oracle import tracker_binding as tb
    L = _build(tmp_path, "tracker_host_emul")
    db = scenarios.generate_dp(5, 4, n_obs=6)
    barrier = dp.build_barrier(db.ref)
    cfg = tb.default_config()
    names = ["sumulation_dt", "dt", "tolerance", "lat_weight_l", "lat_weight_theta", "lat_weight_delta", "lat_weight_delta_rate",
             "lat_preview_time", "lon_weight_s", "lon_weight_v", "lon_weight_a", "lon_weight_j", "wheel_base", "delta_min",
             "delta_max", "min_acceleration", "max_acceleration", "delta_rate_min", "delta_rate_max", "jerk_min", "jerk_max"]
    cf = np.array([getattr(cfg, n) for n in names])
    coarse, starts = [], []
    for b in range(db.B):
        sc = dp.Scene(db.ref, barrier, db.static_poly[b], db.static_nv[b], db.dyn_time[b], db.dyn_samples[b], db.dyn_poly[b],
                      db.dyn_nv[b])
        ok, traj, _, _ = dp.plan(sc, *db.start[b])
        if ok and not np.isnan(traj).any():
            coarse.append(traj)
            starts.append([db.start[b][0] + 0.2, db.start[b][1] - 0.1, db.start[b][2] + 0.03, 9.0])
    coarse, starts = np.ascontiguousarray(coarse), np.ascontiguousarray(starts)
    B, K = coarse.shape[:2]
    assert B >= 2
    traj, gx, gu, ok = np.zeros((B, K, 13)), np.zeros((B, K, 6)), np.zeros((B, K - 1, 2)), np.zeros(B, np.int32)
    L.emul_tracker(_p(cf), cfg.max_num_iteration, B, K, _p(starts), _p(coarse), _p(traj), _p(gx), _p(gu), _p(ok))
    for b in range(B):
        oko, to, _ = tb.plan(tb.start_record(starts[b]), coarse[b])
        X, U = tb.init_guess(to)
        assert ok[b] == int(oko) == 1
        assert np.array_equal(traj[b], to) and np.array_equal(gx[b], X) and np.array_equal(gu[b], U)
