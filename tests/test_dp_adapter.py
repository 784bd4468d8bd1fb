"""The reference's C++ call site of the coarse planner (TrajectoryPlanner -> DpPlanner::Plan,
trajectory_planner.cpp:24,32) driven through the header-compatible planning::DpPlanner of
include/cilqr/dp_planner_b200.h, compiled against the stand-ins in tests/adapter/stubs."""
import subprocess

import numpy as np
import pytest

from cilqr_b200 import build as cbuild
from cilqr_b200 import scenarios


def _scene(tmp_path, seed=91):
    from oracle import dp_binding as dp
    db = scenarios.generate_dp(seed, 1)
    barrier = dp.build_barrier(db.ref)
    T = db.dyn_poly.shape[2]
    parts = [np.array([len(db.ref), db.static_poly.shape[1], db.dyn_poly.shape[1], T], dtype=np.float64),
             db.start[0], db.ref.ravel(),
             np.array([len(barrier[::2])], dtype=np.float64), barrier[::2].ravel(),
             np.array([len(barrier[1::2])], dtype=np.float64), barrier[1::2].ravel(),
             db.static_poly[0].ravel()]
    for j in range(db.dyn_poly.shape[1]):
        parts.append(np.concatenate([db.dyn_time[0, j][:, None], db.dyn_poly[0, j].reshape(T, 8)], axis=1).ravel())
    np.concatenate(parts).astype(np.float64).tofile(tmp_path / "scene.bin")
    return db, barrier


def test_dp_adapter_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = cbuild.build_dp_demo()
    if torch.cuda.is_available():
        pytest.skip("covered by the gpu test")
    _scene(tmp_path)
    r = subprocess.run([exe, str(tmp_path / "scene.bin"), str(tmp_path / "r.bin")], capture_output=True, text=True)
    assert r.returncode == 1, r.stdout + r.stderr
    assert "cilqr_create failed" in r.stderr and "no CPU fallback" in r.stderr
    d = np.fromfile(tmp_path / "r.bin")
    assert d[0] == 0 and d[1] == 0  # Plan() returned false, result untouched


@pytest.mark.gpu
def test_dp_adapter_matches_oracle(tmp_path):
    from oracle import dp_binding as dp
    exe = cbuild.build_dp_demo()
    db, barrier = _scene(tmp_path)
    r = subprocess.run([exe, str(tmp_path / "scene.bin"), str(tmp_path / "r.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    d = np.fromfile(tmp_path / "r.bin")
    sc = dp.Scene(db.ref, barrier, db.static_poly[0], db.static_nv[0], db.dyn_time[0], db.dyn_samples[0],
                  db.dyn_poly[0], db.dyn_nv[0])
    ok, traj, cost, _ = dp.plan(sc, *db.start[0])
    K = int(d[1])
    assert K == len(traj) and bool(d[0]) == ok and abs(d[2] - cost) <= 1e-9 * (abs(cost) + 1)
    got = d[3:].reshape(K, 11)
    assert np.array_equal(np.isnan(got), np.isnan(traj[:, :11]))
    assert np.nanmax(np.abs(got - traj[:, :11]) / (np.abs(traj[:, :11]) + 1.0)) < 1e-9
