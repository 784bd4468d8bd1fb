"""The reference's C++ call site of the corridor (TrajectoryPlanner -> Corridor::Plan,
trajectory_planner.cpp:25,49-57, planning_node.cc:87-103) driven through the header-compatible planning::Corridor of
include/cilqr/corridor_b200.h, compiled against the stand-ins in tests/adapter/stubs."""
import subprocess

import numpy as np
import pytest

from cilqr_b200 import build as cbuild
from cilqr_b200 import scenarios

N_OBS, DT = 11, 0.1


def _scene(tmp_path, seed=61, N=40):
    _, ci = scenarios.generate_with_obstacles(seed, 0, 1, N=N, n_obs=N_OBS, drop_far=False)
    K = ci.K
    n_ped, n_mov = int(round(N_OBS * 6 / 11.0)), int(round(N_OBS * 3 / 11.0))
    n_sta = N_OBS - n_ped - n_mov
    pts = ci.obs_points[0]  # [K, 4*N_OBS, 2]: static obstacles first, then the dynamic ones
    rd = scenarios.road("gentle")
    s = np.arange(0.0, 120.0, 0.1)
    left = np.stack(rd.frenet_to_xy(s, 2.5), axis=1)
    right = np.stack(rd.frenet_to_xy(s, -6.0), axis=1)
    t = np.arange(K) * DT
    parts = [np.array([K, n_sta * 4, N_OBS - n_sta, len(left), len(right)], dtype=np.float64),
             np.concatenate([ci.traj[0], t[:, None]], axis=1).ravel(), pts[0, :n_sta * 4].ravel()]
    for j in range(n_sta, N_OBS):
        parts.append(np.array([K], dtype=np.float64))
        parts.append(np.concatenate([t[:, None], pts[:, 4 * j:4 * j + 4].reshape(K, 8)], axis=1).ravel())
    parts += [left.ravel(), right.ravel()]
    np.concatenate(parts).astype(np.float64).tofile(tmp_path / "scene.bin")
    return ci, left, right


def test_corridor_adapter_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = cbuild.build_corridor_demo()
    if torch.cuda.is_available():
        pytest.skip("covered by the gpu test")
    _scene(tmp_path)
    r = subprocess.run([exe, str(tmp_path / "scene.bin"), str(tmp_path / "r.bin")], capture_output=True, text=True)
    assert r.returncode == 1, r.stdout + r.stderr
    assert "cilqr_create failed" in r.stderr and "no CPU fallback" in r.stderr
    d = np.fromfile(tmp_path / "r.bin")
    assert d[0] == 0  # Plan() returned false


@pytest.mark.gpu
def test_corridor_adapter_matches_oracle(tmp_path):
    from oracle import corridor_binding as cb
    exe = cbuild.build_corridor_demo()
    ci, left, right = _scene(tmp_path)
    r = subprocess.run([exe, str(tmp_path / "scene.bin"), str(tmp_path / "r.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    d = np.fromfile(tmp_path / "r.bin")
    assert d[0] == 1 and int(d[1]) == ci.K
    M = 64
    cor, cnt, poly, code = cb.plan_batch(ci.traj, ci.obs_points, ci.obs_cnt, M)
    assert not code.any()
    o = 2
    for k in range(ci.K):
        m = int(d[o]); o += 1
        assert m == cnt[0, k]
        assert np.array_equal(d[o:o + 3 * m].reshape(m, 3), cor[0, k, :m]); o += 3 * m
        assert np.array_equal(d[o:o + 2 * m].reshape(m, 2), poly[0, k, :m]); o += 2 * m
        assert int(d[o]) == 4 * N_OBS + 8; o += 1  # points_for_corridors: obstacle points + 8 box points
    for bd, is_left in ((left, True), (right, False)):
        n, seg = cb.lane_constraints(bd, is_left)
        assert int(d[o]) == n; o += 1
        assert np.array_equal(d[o:o + 7 * n].reshape(n, 7), seg); o += 7 * n
    assert o == len(d)
