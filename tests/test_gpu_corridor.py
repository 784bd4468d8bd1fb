"""GPU parity of the batched corridor builder (cilqr_corridor_batch / _device, cilqr_lane_constraints)
against the CPU oracle (oracle/corridor_oracle.c) and the committed cv2-backed fixture, through the C ABI.

Bar: the hull work is float32 / index work -- bit-exact.  The kernel evaluates every float and double
expression with the reference's types and without FMA contraction; the one input that is not IEEE-exact is
cos/sin of the knot heading (CUDA's sincos vs glibc's differ in the last ulp on rare arguments), which can
move a box corner by one ulp.  So: plane counts and codes identical on >= 99.9 % of knots, planes
bit-identical on >= 99 % of knots and within 1e-6 relative on all knots with the same count.
"""
import os

import numpy as np
import pytest

from cilqr_b200 import scenarios
from cilqr_b200.solver import default_corridor_config

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def corr_oracle():
    from oracle import corridor_binding as cb
    cb.build()
    return cb


def _check(got, cor, cnt, code, min_cnt_same=0.999, min_exact=0.99):
    M = cor.shape[2]
    same = (got["corridor_cnt"] == cnt) & (got["code"] == code)
    m = (np.arange(M)[None, None, :] < cnt[..., None]) & same[..., None]
    exact = np.array([np.array_equal(g[mm], c[mm]) for g, c, mm in
                      zip(got["corridor"].reshape(-1, M, 3), cor.reshape(-1, M, 3), m.reshape(-1, M))])
    err = np.abs(got["corridor"] - cor) / (np.abs(cor) + 1.0)
    err = np.where(m[..., None], err, 0.0)
    print(f"\n[corridor parity] knots {cnt.size}: same count+code {same.mean():.5f}, bit-exact planes "
          f"{exact[same.ravel()].mean():.5f}, max rel err {err.max():.2e}, planes/knot {cnt.mean():.2f}")
    assert same.mean() >= min_cnt_same
    assert exact[same.ravel()].mean() >= min_exact
    assert err.max() < 1e-6


def test_corridor_golden_fixture(solver):
    z = np.load(os.path.join(GOLD, "corridor_golden_v1.npz"))
    M = z["corridor"].shape[2]
    got = solver.corridor_batch(z["traj"], z["obs_points"], z["obs_cnt"], M, polygon=True)
    assert not got["code"].any()
    _check(got, z["corridor"], z["corridor_cnt"], np.zeros_like(z["corridor_cnt"]))
    m = np.arange(M)[None, None, :] < z["corridor_cnt"][..., None]
    assert np.abs(got["polygon"] - z["polygon"])[m].max() < 1e-6


@pytest.mark.parametrize("N,B,seed,n_obs", [(50, 256, 21, 20), (100, 64, 22, 20), (80, 32, 23, 11)])
def test_corridor_parity_with_oracle(solver, corr_oracle, N, B, seed, n_obs):
    _, ci = scenarios.generate_with_obstacles(seed, 0, B, N=N, n_obs=n_obs)
    M = 24
    cor, cnt, poly, code = corr_oracle.plan_batch(ci.traj, ci.obs_points, ci.obs_cnt, M)
    got = solver.corridor_batch(ci.traj, ci.obs_points, ci.obs_cnt, M, polygon=True,
                                cfg=default_corridor_config(point_cap=4 * n_obs + 16))
    _check(got, cor, cnt, code)
    assert solver.corridor_last_kernel_ms() > 0


def test_corridor_edge_cases(solver, corr_oracle):
    _, ci = scenarios.generate_with_obstacles(31, 0, 4, N=20)
    M = 24
    # no obstacle points at all: the corridor is the +-10 m box (4 planes)
    zero = np.zeros_like(ci.obs_cnt)
    got = solver.corridor_batch(ci.traj, ci.obs_points, zero, M)
    assert not got["code"].any() and (got["corridor_cnt"] == 4).all()
    cor, cnt, _, code = corr_oracle.plan_batch(ci.traj, ci.obs_points, zero, M)
    _check(got, cor, cnt, code)
    # P_max = 0 (no obstacle array)
    got0 = solver.corridor_batch(ci.traj, np.zeros((ci.B, ci.K, 0, 2)), zero, M)
    assert np.array_equal(got0["corridor"], got["corridor"])
    # plane capacity: M_max = 3 < 4 planes -> code 4 at every knot, count 0 (loud, never truncated)
    small = solver.corridor_batch(ci.traj, ci.obs_points, zero, 3)
    assert (small["code"] == 4).all() and not small["corridor_cnt"].any()
    # point capacity: more points inside the window than point_cap -> code 5
    tight = solver.corridor_batch(ci.traj, ci.obs_points, ci.obs_cnt, M, cfg=default_corridor_config(point_cap=12))
    full = solver.corridor_batch(ci.traj, ci.obs_points, ci.obs_cnt, M, cfg=default_corridor_config(point_cap=96))
    d = ci.obs_points - ci.traj[:, :, None, :2]
    valid = np.arange(ci.P_max)[None, None, :] < ci.obs_cnt[..., None]
    inside = ((np.abs(d[..., 0]) <= 25) & (np.abs(d[..., 1]) <= 25) & valid).sum(axis=2) + 8
    assert ((tight["code"] == 5) == (inside > 12)).all()
    ok = tight["code"] == 0
    assert np.array_equal(tight["corridor"][ok], full["corridor"][ok])
    # slots beyond obs_cnt are never read (they hold NaN in the generator's output)
    assert np.isfinite(full["corridor"][np.arange(M)[None, None, :] < full["corridor_cnt"][..., None]]).all()
    # B = 0 and argument validation
    e = solver.corridor_batch(np.zeros((0, 5, 3)), np.zeros((0, 5, 4, 2)), np.zeros((0, 5), np.int32), M)
    assert e["corridor"].shape == (0, 5, M, 3)
    import cilqr_b200
    with pytest.raises(cilqr_b200.CilqrError):
        solver.corridor_batch(ci.traj, np.zeros((ci.B, ci.K, 300, 2)), zero, M)


def test_lane_constraints_parity(solver, corr_oracle):
    rd = scenarios.road("gentle")
    s = np.arange(rd.s_min, rd.s_max + 1e-9, 0.1)
    polys = []
    for lat in (2.5, -6.0, 1.0, -3.0):
        bx, by = rd.frenet_to_xy(s, lat)
        polys.append(np.stack([bx, by], axis=1))
    polys = np.stack(polys)
    for left in (True, False):
        out, cnt = solver.lane_constraints(polys, left, S_max=128)
        for b in range(len(polys)):
            n, seg = corr_oracle.lane_constraints(polys[b], left)
            assert cnt[b] == n
            assert np.array_equal(out[b, :n], seg)
    out, cnt = solver.lane_constraints(np.zeros((2, 7, 2)), True, S_max=8)
    assert (cnt == -1).all()
    out, cnt = solver.lane_constraints(polys[:1], True, S_max=5)
    assert cnt[0] == -2


def test_corridor_feeds_the_solver_on_the_device(solver, corr_oracle, oracle):
    """Corridor::Plan -> IlqrOptimizer::Plan chained on the device (trajectory_planner.cpp:63-86): the build
    kernel writes corridor / corridor_cnt straight into the buffers the solve kernel reads."""
    import torch
    dev = torch.device("cuda:0")
    B, N, M = 48, 60, 20
    batch, ci = scenarios.generate_with_obstacles(41, 0, B, N=N, M_max=M)
    K = N + 1
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    traj, pts, pcnt = t(ci.traj), t(ci.obs_points), t(ci.obs_cnt)
    cor = torch.zeros(B, K, M, 3, dtype=torch.float64, device=dev)
    ccnt = torch.zeros(B, K, dtype=torch.int32, device=dev)
    code = torch.zeros(B, K, dtype=torch.int32, device=dev)
    X = torch.zeros(B, K, 6, dtype=torch.float64, device=dev)
    U = torch.zeros(B, N, 2, dtype=torch.float64, device=dev)
    S = torch.zeros(B, 8, dtype=torch.float64, device=dev)
    cfg = default_corridor_config(point_cap=96)
    solver.corridor_batch_device(B, K, ci.P_max, M, traj, pts, pcnt, cor, ccnt, code, cfg=cfg)
    solver.plan_batch_device(B, N, M, batch.S, batch.S, t(batch.start), t(batch.coarse), cor, ccnt,
                             t(batch.lane_left), t(batch.lane_right), X, U, S)
    solver.synchronize()
    assert not code.cpu().numpy().any()
    # oracle pipeline on the host
    ocor, ocnt, _, ocode = corr_oracle.plan_batch(ci.traj, ci.obs_points, ci.obs_cnt, M)
    assert not ocode.any()
    assert np.array_equal(ccnt.cpu().numpy(), ocnt)
    ob = scenarios.ScenarioBatch(N, M, batch.S, batch.start, batch.coarse, ocor, ocnt, batch.lane_left, batch.lane_right)
    Xo, Uo, So, _ = oracle.solve_batch(ob, nthreads=os.cpu_count() or 1)
    Sg, Xg = S.cpu().numpy(), X.cpu().numpy()
    same = (Sg[:, 0] == So[:, 0]) & (Sg[:, 1] == So[:, 1]) & (Sg[:, 7] == So[:, 7])
    e = (np.abs(Xg - Xo) / (np.abs(Xo) + 1.0)).reshape(B, -1).max(axis=1)
    print(f"\n[corridor->solve] identical decision path {same.sum()}/{B}, max rel err on those {e[same].max():.2e}, "
          f"converged {(Sg[:, 0] <= 2).sum()}/{B}, planes/knot {ocnt.mean():.2f}")
    assert same.mean() >= 0.9
    assert (e[same] < 1e-4).mean() >= 0.97


def test_corridor_properties_at_scale(solver):
    """Size-independent properties on a batch too large for the CPU checker: the knot is strictly inside its
    corridor, no in-range obstacle point is inside, results are deterministic and shard-invariant."""
    B, N, M = 2048, 100, 24
    _, ci = scenarios.generate_with_obstacles(51, 0, B, N=N)
    cfg = default_corridor_config(point_cap=96)
    got = solver.corridor_batch(ci.traj, ci.obs_points, ci.obs_cnt, M, cfg=cfg)
    assert not got["code"].any()
    cor, cnt = got["corridor"], got["corridor_cnt"]
    valid = np.arange(M)[None, None, :] < cnt[..., None]
    nrm = np.hypot(cor[..., 0], cor[..., 1])
    x, y = ci.traj[..., 0, None], ci.traj[..., 1, None]
    with np.errstate(invalid="ignore", divide="ignore"):
        g = (cor[..., 0] * x + cor[..., 1] * y - cor[..., 2]) / nrm
    assert (g[valid] < -1e-6).all()
    again = solver.corridor_batch(ci.traj, ci.obs_points, ci.obs_cnt, M, cfg=cfg)
    assert np.array_equal(again["corridor"][valid], cor[valid]) and np.array_equal(again["corridor_cnt"], cnt)
    half = solver.corridor_batch(ci.traj[B // 2:], ci.obs_points[B // 2:], ci.obs_cnt[B // 2:], M, cfg=cfg)
    assert np.array_equal(half["corridor"][valid[B // 2:]], cor[B // 2:][valid[B // 2:]])
    # every obstacle point inside the +-25 m window lies outside (or within 2 mm of) the polygon
    sub = slice(0, 64)
    pts = ci.obs_points[sub]
    pv = np.arange(ci.P_max)[None, None, :] < ci.obs_cnt[sub][..., None]
    d = pts - ci.traj[sub][:, :, None, :2]
    near = pv & (np.abs(d[..., 0]) <= 25) & (np.abs(d[..., 1]) <= 25)
    pts = np.nan_to_num(pts)
    with np.errstate(invalid="ignore", divide="ignore"):
        viol = (pts[..., None, 0] * cor[sub][:, :, None, :, 0] + pts[..., None, 1] * cor[sub][:, :, None, :, 1]
                - cor[sub][:, :, None, :, 2]) / nrm[sub][:, :, None, :]
    viol = np.where(valid[sub][:, :, None, :], viol, -np.inf).max(axis=3)
    assert (viol[near] > -2e-3).all()
