"""GPU parity of the batched DP planner (cilqr_dp_plan_batch / _device) against the CPU oracle
(oracle/dp_oracle.c), through the C ABI.

The kernel evaluates every double expression in the reference's order without FMA contraction (the TU is
compiled with -fmad=false), so lattice decisions match the oracle's unless a libm result (sin, cos, atan, fmod,
hypot: CUDA vs glibc, last ulp) flips a comparison.  Bar: ok flag, optimum waypoints and cost identical on
>= 97 % of scenarios; on those the trajectory within 1e-9 relative; mismatches are counted and printed.
"""
import numpy as np
import pytest

from cilqr_b200 import scenarios
from cilqr_b200.solver import default_dp_config, dp_num_knots

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dp_oracle():
    from oracle import dp_binding as dp
    dp.build()
    return dp


def _oracle_batch(dp, db, barrier):
    """The oracle on every scene, one host thread per core (ctypes releases the GIL inside dp_plan)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    cfg = dp.default_config()

    def one(b):
        sc = dp.Scene(db.ref, barrier, db.static_poly[b], db.static_nv[b], db.dyn_time[b], db.dyn_samples[b],
                      db.dyn_poly[b], db.dyn_nv[b])
        return dp.plan(sc, *db.start[b], cfg)

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        return list(ex.map(one, range(db.B)))


def test_dp_parity_on_2048_scenes(solver, dp_oracle):
    """24 scenes are too thin for a kernel whose sin / cos / atan differ from glibc in the last ulp and
    feed strict-'<' lattice decisions.  2 048 scenes: every field of every trajectory point, the optimum's way-points,
    the cost and the ok flag against the oracle (itself bit-identical to the reference's own DpPlanner::Plan,
    tests/test_reference_pins.py).  Every mismatch is counted and printed."""
    B = 2048
    db = scenarios.generate_dp(20260107, B)
    barrier = dp_oracle.build_barrier(db.ref)
    ref = _oracle_batch(dp_oracle, db, barrier)
    got = solver.dp_plan_batch(db, barrier, waypoints=True)
    ok_same = np.array([got["ok"][b] == r[0] for b, r in enumerate(ref)])
    same = np.array([got["ok"][b] == r[0] and np.array_equal(got["waypoints"][b][:, :2], r[3][:, :2])
                     for b, r in enumerate(ref)])
    nan_same = np.array([np.array_equal(np.isnan(got["trajectory"][b]), np.isnan(r[1])) for b, r in enumerate(ref)])
    err = np.array([np.nanmax(np.abs(got["trajectory"][b] - r[1]) / (np.abs(r[1]) + 1.0)) if nan_same[b] else np.inf
                    for b, r in enumerate(ref)])
    cerr = np.array([abs(got["cost"][b] - r[2]) / (abs(r[2]) + 1.0) for b, r in enumerate(ref)])
    bitwise = np.array([np.array_equal(got["trajectory"][b], r[1], equal_nan=True) for b, r in enumerate(ref)])
    print(f"\n[dp parity 2048] ok flag equal {ok_same.sum()}/{B}; identical optimum {same.sum()}/{B}; on those: NaN pattern "
          f"equal {nan_same[same].sum()}, max rel err trajectory {err[same].max():.2e}, cost {cerr[same].max():.2e}, "
          f"bit-identical trajectories {bitwise[same].sum()}; planned ok {int(got['ok'].sum())}/{B}; different optimum: "
          f"{np.where(~same)[0][:16].tolist()}; kernel {solver.dp_last_kernel_ms():.1f} ms")
    assert ok_same.mean() >= 0.999
    assert same.mean() >= 0.995
    assert nan_same[same].all()
    assert err[same].max() < 1e-9 and cerr[same].max() < 1e-9


def test_dp_parity_with_oracle(solver, dp_oracle):
    B = 24
    db = scenarios.generate_dp(71, B)
    barrier = dp_oracle.build_barrier(db.ref)
    ref = _oracle_batch(dp_oracle, db, barrier)
    got = solver.dp_plan_batch(db, barrier, waypoints=True)
    assert dp_num_knots() == 81 and got["trajectory"].shape == (B, 81, 13)
    same = np.array([got["ok"][b] == r[0] and np.array_equal(got["waypoints"][b][:, :2], r[3][:, :2])
                     for b, r in enumerate(ref)])
    # a plan that stands still (station index 0) repeats a point: ComputePathProfile divides 0 by 0 there
    # (discrete_points_math.cc:118-150) and the reference returns NaN curvature -- so do both restatements
    nan_same = np.array([np.array_equal(np.isnan(got["trajectory"][b]), np.isnan(r[1])) for b, r in enumerate(ref)])
    assert nan_same.all()
    err = np.array([np.nanmax(np.abs(got["trajectory"][b] - r[1]) / (np.abs(r[1]) + 1.0)) for b, r in enumerate(ref)])
    cerr = np.array([abs(got["cost"][b] - r[2]) / (abs(r[2]) + 1.0) for b, r in enumerate(ref)])
    print(f"\n[dp parity] B={B}: identical optimum {same.sum()}/{B}; on those: max rel err trajectory "
          f"{err[same].max():.2e}, cost {cerr[same].max():.2e}; planned ok {int(got['ok'].sum())}/{B}; "
          f"kernel {solver.dp_last_kernel_ms():.1f} ms")
    assert same.mean() >= 0.97
    assert err[same].max() < 1e-9 and cerr[same].max() < 1e-9
    # the solver / corridor views of the same result
    assert np.array_equal(got["coarse"][..., :3], got["trajectory"][..., 2:5], equal_nan=True)
    assert np.array_equal(got["coarse"][..., 3:5], got["trajectory"][..., 6:8], equal_nan=True)
    assert np.array_equal(got["coarse"][..., 5], got["trajectory"][..., 9], equal_nan=True)
    assert np.array_equal(got["xytheta"], got["trajectory"][..., 2:5], equal_nan=True)


def test_dp_known_answers_and_edges(solver, dp_oracle):
    import cilqr_b200
    s = np.arange(0, 300.05, 0.1)
    z = np.zeros_like(s)
    ref = np.stack([s, s, z, z, z, np.full_like(s, 2.5), np.full_like(s, 6.0)], axis=1)
    barrier = dp_oracle.build_barrier(ref)
    T = 3
    free = scenarios.DpBatch(ref, np.zeros((2, 3)), np.zeros((2, 0, 4, 2)), np.zeros((2, 0), np.int32),
                             np.zeros((2, 0, T)), np.zeros((2, 0), np.int32), np.zeros((2, 0, T, 4, 2)),
                             np.zeros((2, 0), np.int32))
    got = solver.dp_plan_batch(free, barrier, waypoints=True)
    # free straight road: station index 3 (16 m per 1.6 s = the nominal 10 m/s), lateral index 9 (the centre line)
    assert got["ok"].all() and (got["waypoints"][..., 0] == 3).all() and (got["waypoints"][..., 1] == 9).all()
    assert np.allclose(got["cost"], 10.0, atol=1e-9)
    assert np.allclose(got["trajectory"][:, 17:, 6], 10.0)
    # an obstacle on the centre line: avoided, and identical to the oracle
    box = np.array([[38.0, -1.0], [38.0, 1.0], [42.0, 1.0], [42.0, -1.0]])
    blocked = scenarios.DpBatch(ref, np.zeros((1, 3)), box[None, None], np.full((1, 1), 4, np.int32),
                                np.zeros((1, 0, T)), np.zeros((1, 0), np.int32), np.zeros((1, 0, T, 4, 2)),
                                np.zeros((1, 0), np.int32))
    g2 = solver.dp_plan_batch(blocked, barrier, waypoints=True)
    ok, traj, cost, wp = dp_oracle.plan(dp_oracle.Scene(ref, barrier, box[None]), 0.0, 0.0, 0.0)
    assert g2["ok"][0] == ok and np.array_equal(g2["waypoints"][0], wp) and abs(g2["cost"][0] - cost) < 1e-9
    assert np.abs(g2["trajectory"][0] - traj).max() < 1e-9 and np.abs(traj[:, 3]).max() > 1.0
    # B = 0 and argument validation
    empty = scenarios.DpBatch(ref, np.zeros((0, 3)), np.zeros((0, 0, 4, 2)), np.zeros((0, 0), np.int32),
                              np.zeros((0, 0, T)), np.zeros((0, 0), np.int32), np.zeros((0, 0, T, 4, 2)),
                              np.zeros((0, 0), np.int32))
    assert solver.dp_plan_batch(empty, barrier)["ok"].shape == (0,)
    bad = default_dp_config()
    bad.delta_t = 0.0
    with pytest.raises(cilqr_b200.CilqrError):
        solver.dp_plan_batch(free, barrier, cfg=bad)


def test_dp_feeds_corridor_and_solver_on_the_device(solver):
    """DpPlanner::Plan -> Corridor::Plan -> IlqrOptimizer::Plan (trajectory_planner.cpp:32-86) chained on the
    device: the planner writes `coarse` / `xytheta` in the layouts the next two kernels read."""
    import torch
    dev = torch.device("cuda:0")
    B, M = 16, 24
    db = scenarios.generate_dp(81, B)
    from oracle import dp_binding as dp
    barrier = dp.build_barrier(db.ref)
    K = dp_num_knots()
    N = K - 1
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    ok = torch.zeros(B, dtype=torch.int32, device=dev)
    coarse = torch.zeros(B, K, 6, dtype=torch.float64, device=dev)
    xyt = torch.zeros(B, K, 3, dtype=torch.float64, device=dev)
    solver.dp_plan_batch_device(B, len(db.ref), len(barrier), 4, db.static_poly.shape[1], db.dyn_poly.shape[1],
                                db.dyn_poly.shape[2], t(db.ref), t(barrier), t(db.start), t(db.static_poly),
                                t(db.static_nv), t(db.dyn_time), t(db.dyn_samples), t(db.dyn_poly), t(db.dyn_nv), ok,
                                coarse=coarse, xytheta=xyt)
    # obstacle points per knot = the corners of every obstacle at that knot's time (Environment's queries)
    pts = np.concatenate([np.broadcast_to(db.static_poly[:, None], (B, K) + db.static_poly.shape[1:]).reshape(B, K, -1, 2),
                          np.transpose(db.dyn_poly, (0, 2, 1, 3, 4)).reshape(B, K, -1, 2)], axis=2)
    cnt = np.full((B, K), pts.shape[2], np.int32)
    cor = torch.zeros(B, K, M, 3, dtype=torch.float64, device=dev)
    ccnt = torch.zeros(B, K, dtype=torch.int32, device=dev)
    code = torch.zeros(B, K, dtype=torch.int32, device=dev)
    solver.corridor_batch_device(B, K, pts.shape[2], M, xyt, t(pts), t(cnt), cor, ccnt, code)
    # lanes: a window of 40 segments per side starting ~15 m behind each start (like scenarios.generate)
    rd = scenarios.road("gentle")
    s0 = np.array([dp.get_projection(db.ref, x, y)[0] for x, y, _ in db.start])
    lanes = []
    for side, left in ((0, True), (1, False)):
        stn = rd.lane_station[side]
        first = np.clip(np.searchsorted(stn, s0 - 15.0, side="right") - 1, 0, len(stn) - 41)
        idx = first[:, None] + np.arange(41)[None, :]
        lanes.append(np.ascontiguousarray(scenarios._lane_constraints(rd.lane_pts[side][idx], left=left)))
    start4 = np.concatenate([db.start, np.full((B, 1), 10.0)], axis=1)  # planning_node.cc:24-30: v = 10
    X = torch.zeros(B, K, 6, dtype=torch.float64, device=dev)
    U = torch.zeros(B, N, 2, dtype=torch.float64, device=dev)
    S = torch.zeros(B, 8, dtype=torch.float64, device=dev)
    solver.synchronize()  # the solve below runs on another handle's stream
    big = __import__("cilqr_b200").Solver(device=0, N_max=N, M_max=M, S_max=lanes[0].shape[1], B_max=B)
    big.plan_batch_device(B, N, M, lanes[0].shape[1], lanes[1].shape[1], t(start4), coarse, cor, ccnt, t(lanes[0]),
                          t(lanes[1]), X, U, S)
    solver.synchronize()
    big.synchronize()
    okh, codeh, Sh = ok.cpu().numpy(), code.cpu().numpy(), S.cpu().numpy()
    print(f"\n[dp->corridor->solve] planned {int(okh.sum())}/{B}, corridor failures {int((codeh != 0).sum())}, "
          f"solver exits {np.bincount(Sh[:, 0].astype(int), minlength=5).tolist()}, mean iterations {Sh[:, 1].mean():.1f}")
    assert not (codeh != 0).any()
    assert np.isfinite(X.cpu().numpy()).all() and (Sh[:, 0] >= 0).all()
    big.close()


def test_dp_golden_fixture(solver):
    """The committed oracle outputs (tests/golden/dp_golden_v1.npz) as the target."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dp_golden_v1.npz"))
    db = scenarios.DpBatch(z["ref"], z["start"], z["static_poly"], z["static_nv"], z["dyn_time"], z["dyn_samples"],
                           z["dyn_poly"], z["dyn_nv"])
    got = solver.dp_plan_batch(db, z["barrier"], waypoints=True)
    same = np.array([got["ok"][b] == z["ok"][b] and np.array_equal(got["waypoints"][b][:, :2], z["waypoints"][b][:, :2])
                     for b in range(db.B)])
    assert same.sum() >= db.B - 1  # a last-ulp libm difference may flip one lattice decision
    for b in np.where(same)[0]:
        t = z["trajectory"][b]
        assert np.array_equal(np.isnan(got["trajectory"][b]), np.isnan(t))
        assert np.nanmax(np.abs(got["trajectory"][b] - t) / (np.abs(t) + 1.0)) < 1e-9
        assert abs(got["cost"][b] - z["cost"][b]) <= 1e-9 * (abs(z["cost"][b]) + 1.0)
