"""The reference's C++ call site (TrajectoryPlanner -> IlqrOptimizer::Plan, trajectory_planner.cpp:26,80-97)
driven through the header-compatible planning::IlqrOptimizer of include/cilqr/ilqr_optimizer_b200.h.
Eigen / ROS are not installed here, so the adapter is compiled against the type stand-ins in
tests/adapter/stubs (same member names as the reference headers)."""
import os
import subprocess

import numpy as np
import pytest

from cilqr_b200 import build as cbuild
from cilqr_b200 import scenarios


def _write_scenario(path, batch, b):
    K = batch.N + 1
    parts = [np.array([batch.N, batch.M_max, batch.lane_left.shape[1], batch.lane_right.shape[1]], dtype=np.float64),
             batch.start[b].ravel(), batch.coarse[b].ravel(), batch.corridor_cnt[b].astype(np.float64).ravel(),
             batch.corridor[b].ravel(), batch.lane_left[b].ravel(), batch.lane_right[b].ravel()]
    assert parts[3].size == K
    np.concatenate(parts).astype(np.float64).tofile(path)


def _read_result(path, K):
    d = np.fromfile(path, dtype=np.float64)
    ok, status, iters, n_cost, n_iter = d[:5]
    o = 5
    opt = d[o:o + K * 13].reshape(K, 13)
    o += K * 13
    it0 = d[o:o + K * 13].reshape(K, 13)
    o += K * 13
    cost = d[o:o + int(n_cost) * 5].reshape(int(n_cost), 5)
    return dict(ok=ok, status=int(status), iters=int(iters), n_iter=int(n_iter), opt=opt, iter0=it0, cost=cost)


def test_adapter_compiles_and_fails_loudly_without_gpu(tmp_path):
    """Compile check of the drop-in header + the 'no CPU fallback' contract: without a device Plan()
    returns false and leaves opt_trajectory empty (the caller's failure signal,
    trajectory_planner.cpp:91-94)."""
    import torch
    exe = cbuild.build_adapter_demo()
    if torch.cuda.is_available():
        pytest.skip("covered by the gpu test")
    batch = scenarios.generate(5, 0, 1, N=30)
    _write_scenario(tmp_path / "s.bin", batch, 0)
    r = subprocess.run([exe, str(tmp_path / "s.bin"), str(tmp_path / "r.bin")], capture_output=True, text=True)
    assert r.returncode == 1, r.stdout + r.stderr
    assert "cilqr_create failed" in r.stderr and "no CPU fallback" in r.stderr
    res = _read_result(tmp_path / "r.bin", batch.N + 1)
    assert res["ok"] == 0 and np.all(res["opt"] == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("N,seed", [(80, 20260101), (30, 5)])
def test_adapter_matches_oracle(tmp_path, oracle, N, seed):
    """B = 1 through the C++ adapter: opt_trajectory, iter_trajs[0] and cost() against the oracle.
    N = 80 is the shipped horizon (tf = 8 s, dt = 0.1 s, planner_config.h:94,99)."""
    exe = cbuild.build_adapter_demo()
    batch = scenarios.generate(seed, 0, 1, N=N, n_obs=11 if N == 80 else 20)
    K = N + 1
    _write_scenario(tmp_path / "s.bin", batch, 0)
    r = subprocess.run([exe, str(tmp_path / "s.bin"), str(tmp_path / "r.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    res = _read_result(tmp_path / "r.bin", K)
    o = oracle.solve(batch, 0, hist=True)
    assert res["ok"] == 1
    assert (res["status"], res["iters"]) == (o["status"], o["iters"])
    X, U = o["states"], o["controls"]
    opt = res["opt"]
    # TransformToTrajectory (ilqr_optimizer.cc:771-791): time, s, x, y, theta, kappa, v, a, jerk, delta, delta_rate
    np.testing.assert_allclose(opt[:, 0], np.arange(K) * 0.1, rtol=0, atol=1e-12)
    assert np.all(opt[:, 1] == 0) and np.all(opt[:, 11:] == 0)
    tol = dict(rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(opt[:, [2, 3, 4, 6, 7, 9]], X, **tol)
    np.testing.assert_allclose(opt[:, 5], np.tan(X[:, 5]) / 1.0, **tol)
    np.testing.assert_allclose(opt[:-1, [8, 10]], U, **tol)
    assert np.all(opt[-1, [8, 10]] == 0)  # jerk / delta_rate stay 0 at the last knot
    # iter_trajs[0] is the initial guess (:170); cost() = initial + every accepted iterate (:173,283,296)
    np.testing.assert_allclose(res["iter0"][:, [2, 3, 4, 6, 7, 9]], o["init_states"], **tol)
    assert res["cost"].shape == o["cost_hist"].shape
    np.testing.assert_allclose(res["cost"], o["cost_hist"], rtol=1e-7)
    assert res["n_iter"] == len(o["cost_hist"]) - (1 if o["status"] <= 1 else 0)
