"""PINS AGAINST THE REFERENCE ITSELF.  The pieces of the path that build without ROS / Eigen / OpenCV -- the
reference's own geometry (algorithm/math/*.cpp), reference-line queries (utils/discretized_trajectory.cpp) and
path profile (utils/discrete_points_math.cc) -- are compiled from /root/reference into oracle/_ref (recipe:
`make -C oracle _ref`, wrapper oracle/ref_geom_wrapper.cc) and the restatements in oracle/ are compared with
them bit for bit:

  solver oracle : NormalizeAngle, LineSegment2d::DistanceTo (the distance FindNeastLaneSegment minimises)
  DP oracle     : slerp, EvaluateStation, GetCartesian, GetProjection, Polygon2d::IsPointIn,
                  Polygon2d::HasOverlap(Box2d) on the planner's disc boxes, ComputePathProfile
"""
import ctypes as C

import numpy as np
import pytest

from cilqr_b200 import scenarios
from oracle import dp_binding as dp
from oracle import ref_binding as ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="neither oracle/_ref nor /root/reference is present")


def _lines():
    s = np.arange(0, 120.05, 0.1)
    z = np.zeros_like(s)
    straight = np.stack([s, s, z, z, z, np.full_like(s, 2.5), np.full_like(s, 6.0)], axis=1)
    return [straight, scenarios.reference_line("gentle"), scenarios.reference_line("shipped")]


def test_normalize_angle_and_slerp(oracle):
    L, R = dp.lib(), ref.lib()
    L.dp_slerp.restype = C.c_double
    L.dp_slerp.argtypes = [C.c_double] * 5
    O = oracle.lib()
    rng = np.random.default_rng(0)
    angles = np.r_[rng.uniform(-20, 20, 4000), [0.0, np.pi, -np.pi, 3 * np.pi, -3 * np.pi, 2 * np.pi, 1e-300, -1e-300]]
    for a in angles:
        assert O.cilqr_oracle_normalize_angle(float(a)) == R.ref_normalize_angle(float(a))
    for _ in range(4000):
        a0, a1 = rng.uniform(-7, 7, 2)
        t0 = rng.uniform(0, 10)
        t1 = t0 + rng.choice([0.0, 1e-11, 0.1, 1.0])
        t = rng.uniform(t0 - 0.1, t1 + 0.1)
        assert L.dp_slerp(a0, t0, a1, t1, t) == R.ref_slerp(a0, t0, a1, t1, t)


def test_segment_distance(oracle):
    O, R = oracle.lib(), ref.lib()
    rng = np.random.default_rng(1)
    for i in range(5000):
        x0, y0, x1, y1, px, py = rng.uniform(-50, 50, 6)
        if i % 10 == 0:
            x1, y1 = x0, y0  # degenerate segment (length <= kMathEpsilon): distance to the start point
        if i % 7 == 0:
            px, py = x0 + 0.3 * (x1 - x0), y0 + 0.3 * (y1 - y0)  # on the segment
        seg = (C.c_double * 7)(0.0, 0.0, 0.0, x0, y0, x1, y1)
        assert O.cilqr_oracle_segment_distance(seg, px, py) == R.ref_segment_distance(x0, y0, x1, y1, px, py)


def test_reference_line_queries():
    rng = np.random.default_rng(2)
    for line in _lines():
        st = np.r_[rng.uniform(line[0, 0] - 5, line[-1, 0] + 5, 800), line[[0, 1, -1, 7], 0]]
        want = ref.evaluate_stations(line, st)
        got = np.stack([dp.evaluate_station(line, float(s)) for s in st])
        assert np.array_equal(got, want)
        sl = np.stack([rng.uniform(line[0, 0] - 2, line[-1, 0] + 2, 500), rng.uniform(-7, 4, 500)], axis=1)
        want = ref.get_cartesians(line, sl)
        out = np.zeros(2)
        for (s, l), w in zip(sl, want):
            dp.lib().dp_get_cartesian(len(line), line.ctypes.data, C.c_double(s), C.c_double(l),
                                      out.ctypes.data_as(C.c_void_p))
            assert np.array_equal(out, w)
        idx = rng.integers(0, len(line), 150)
        xy = line[idx, 1:3] + rng.normal(0, 3.0, size=(150, 2))
        xy = np.r_[xy, line[[0, -1], 1:3] + [[-4.0, 1.0], [3.0, -2.0]]]  # beyond both ends
        want = ref.get_projections(line, xy)
        got = np.stack([dp.get_projection(line, float(x), float(y)) for x, y in xy])
        assert np.array_equal(got, want)


def test_polygon_box_overlap_and_point_in_polygon():
    L = dp.lib()
    rng = np.random.default_rng(3)
    radius = 1.2965
    hits = 0
    for i in range(6000):
        nv = int(rng.integers(3, 7))
        if i % 3 == 0:  # oriented rectangle (vehicles, pedestrians)
            c, th, hx, hy = rng.uniform(-4, 4, 2), rng.uniform(0, np.pi), rng.uniform(0.2, 3), rng.uniform(0.2, 1.5)
            k = np.array([[-hx, -hy], [-hx, hy], [hx, hy], [hx, -hy]])
            rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
            poly = c + k @ rot.T
        else:  # any simple-ish polygon: convex position around a centre, either orientation
            ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
            poly = rng.uniform(-4, 4, 2) + np.stack([np.cos(ang), np.sin(ang)], axis=1) * rng.uniform(0.3, 3.0, (nv, 1))
            if i % 2:
                poly = poly[::-1]
        poly = np.ascontiguousarray(poly)
        cx, cy = rng.uniform(-5, 5, 2)
        want = ref.polygon_overlaps_disc_box(poly, cx, cy, radius)
        got = bool(L.dp_polygon_overlaps_box(poly.ctypes.data_as(C.c_void_p), len(poly), C.c_double(cx), C.c_double(cy),
                                             C.c_double(radius)))
        assert got == want
        hits += want
        x, y = rng.uniform(-5, 5, 2)
        assert bool(L.dp_polygon_is_point_in(poly.ctypes.data_as(C.c_void_p), len(poly), C.c_double(x),
                                             C.c_double(y))) == ref.polygon_is_point_in(poly, x, y)
    assert 500 < hits < 5500  # both outcomes are exercised


def test_path_profile():
    L = dp.lib()
    rng = np.random.default_rng(4)
    for n in (2, 3, 17, 81):
        for _ in range(20):
            xy = np.cumsum(rng.uniform(0.2, 1.5, (n, 2)), axis=0) + rng.normal(0, 0.05, (n, 2))
            ok, v, a, k = ref.compute_path_profile(0.1, xy)
            assert ok
            x, y = np.ascontiguousarray(xy[:, 0]), np.ascontiguousarray(xy[:, 1])
            v2, a2, k2 = np.zeros(n), np.zeros(n), np.zeros(n)
            L.dp_path_profile(C.c_double(0.1), n, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p),
                              v2.ctypes.data_as(C.c_void_p), a2.ctypes.data_as(C.c_void_p), k2.ctypes.data_as(C.c_void_p))
            assert np.array_equal(v, v2) and np.array_equal(a, a2) and np.array_equal(k, k2)


def test_planner_results_are_consistent_with_the_reference_geometry():
    """End to end on a planned scene: every knot of the oracle's plan lies where the REFERENCE's GetCartesian puts
    its (s, l), and the reference's own ComputePathProfile reproduces the velocity / acceleration / curvature columns."""
    db = scenarios.generate_dp(31, 1, n_obs=5)
    barrier = dp.build_barrier(db.ref)
    sc = dp.Scene(db.ref, barrier, db.static_poly[0], db.static_nv[0], db.dyn_time[0], db.dyn_samples[0],
                  db.dyn_poly[0], db.dyn_nv[0])
    ok, traj, cost, wp = dp.plan(sc, *db.start[0])
    okp, v, a, k = ref.compute_path_profile(0.1, traj[:, 2:4])
    assert okp and np.array_equal(traj[:, 6], v) and np.array_equal(traj[:, 7], a) and np.array_equal(traj[:, 5], k, equal_nan=True)
    sl = ref.get_projections(db.ref, traj[:, 2:4])
    assert np.abs(sl[:, 0] - traj[:, 1]).max() < 0.05  # projecting the knots back gives their stations


# ---- the reference's own Environment and DpPlanner (environment.cpp, dp_planner.cpp compiled unmodified) ----------
def _scene(db, b):
    return (db.ref, db.static_poly[b], db.static_nv[b], db.dyn_time[b], db.dyn_samples[b], db.dyn_poly[b], db.dyn_nv[b])


def test_road_barrier_equals_set_reference():
    for line in _lines()[1:]:
        mine = dp.build_barrier(line)
        theirs = ref.road_barrier(line)  # left points then right points (road_barrier_ itself is private)
        assert len(mine) == len(theirs)
        assert np.array_equal(np.array(sorted(map(tuple, mine))), np.array(sorted(map(tuple, theirs))))
        assert (np.diff(mine[:, 0]) >= 0).all()


def test_check_optimization_collision_equals_the_reference():
    db = scenarios.generate_dp(17, 3)
    barrier = dp.build_barrier(db.ref)
    rng = np.random.default_rng(5)
    total = 0
    for b in range(db.B):
        sc = dp.Scene(db.ref, barrier, *_scene(db, b)[1:])
        # poses along and around the road near the scene's start, at times inside and outside the obstacle tracks
        s0 = dp.get_projection(db.ref, *db.start[b, :2])[0]
        q = []
        for _ in range(1500):
            r = dp.evaluate_station(db.ref, s0 + rng.uniform(-5, 110))
            l = rng.uniform(-7.5, 4.0)
            q.append([rng.choice([rng.uniform(0, 8.5), 0.0, 8.0, 1.6]), r[1] - l * np.sin(r[3]), r[2] + l * np.cos(r[3]),
                      r[3] + rng.normal(0, 0.3)])
        q = np.array(q)
        want = ref.check_optimization_collision(*_scene(db, b), q)
        got = np.array([dp.check_collision(sc, *row) for row in q])
        # query time == last sample time dereferences end() in the reference (undefined): not compared
        defined = q[:, 0] != 8.0
        assert np.array_equal(got[defined], want[defined])
        total += int(want.sum())
    assert total > 200  # collisions do occur in the sample


def test_whole_plans_equal_the_reference_bit_for_bit():
    """DpPlanner::Plan of the reference itself against the oracle on whole scenes: return value and every field of
    every trajectory point."""
    db = scenarios.generate_dp(5, 6)
    barrier = dp.build_barrier(db.ref)
    oks = []
    for b in range(db.B):
        okr, tr = ref.dp_plan(*_scene(db, b), *db.start[b])
        ok, traj, cost, wp = dp.plan(dp.Scene(db.ref, barrier, *_scene(db, b)[1:]), *db.start[b])
        assert okr == ok and len(tr) == len(traj) == 81
        assert np.array_equal(tr, traj[:, :11], equal_nan=True)
        oks.append(ok)
    assert any(oks) and not all(oks)  # both a successful and a failed plan are covered


def test_committed_fixture_equals_the_reference():
    """tests/golden/dp_golden_v1.npz (the GPU tests' target) holds exactly what the reference's planner returns."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dp_golden_v1.npz"))
    for b in range(len(z["start"])):
        okr, tr = ref.dp_plan(z["ref"], z["static_poly"][b], z["static_nv"][b], z["dyn_time"][b], z["dyn_samples"][b],
                              z["dyn_poly"][b], z["dyn_nv"][b], *z["start"][b])
        assert okr == bool(z["ok"][b])
        assert np.array_equal(tr, z["trajectory"][b][:, :11], equal_nan=True)


def test_known_answer_scenes_on_the_reference():
    s = np.arange(0, 300.05, 0.1)
    zz = np.zeros_like(s)
    line = np.stack([s, s, zz, zz, zz, np.full_like(s, 2.5), np.full_like(s, 6.0)], axis=1)
    none = (np.zeros((0, 4, 2)), np.zeros(0, np.int32), np.zeros((0, 1)), np.zeros(0, np.int32), np.zeros((0, 1, 4, 2)),
            np.zeros(0, np.int32))
    ok, tr = ref.dp_plan(line, *none, 0.0, 0.0, 0.0)
    assert ok and np.allclose(tr[17:, 6], 10.0) and np.allclose(tr[:, 3], 0.0)
    ok2, traj, cost, wp = dp.plan(dp.Scene(line), 0.0, 0.0, 0.0)
    assert ok2 and np.array_equal(tr, traj[:, :11])


# ---- the reference's own CILQR solver (ilqr_optimizer.cc, vehicle_model.cc, barrier_function.h compiled unmodified
# ---- against the Eigen-lite stand-in of oracle/ref_stubs) ---------------------------------------------------------
def test_dynamics_and_jacobian_equal_the_reference(oracle):
    L, p = oracle.lib(), oracle.default_params()
    rng = np.random.default_rng(6)
    for _ in range(300):
        x = np.array([rng.normal(0, 20), rng.normal(0, 20), rng.uniform(-7, 7), rng.uniform(0, 20), rng.uniform(-5, 5),
                      rng.uniform(-0.9, 0.9)])
        u = np.array([rng.uniform(-10, 10), rng.uniform(-0.3, 0.3)])
        out, A, B = np.zeros(6), np.zeros(36), np.zeros(12)
        dp_ = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        L.cilqr_oracle_dynamics(C.byref(p), dp_(x), dp_(u), dp_(out))
        assert np.array_equal(out, ref.dynamics(x, u))
        L.cilqr_oracle_dynamics_jacobian(C.byref(p), dp_(x), dp_(u), dp_(A), dp_(B))
        Ar, Br = ref.dynamics_jacobian(x, u)
        assert np.array_equal(A.reshape(6, 6), Ar) and np.array_equal(B.reshape(6, 2), Br)


def test_barrier_value_equals_the_reference(oracle):
    L, p = oracle.lib(), oracle.default_params()
    R = ref.solver_lib()
    for g in np.r_[np.linspace(-5, 1, 400), [-0.01, -0.0100001, -0.0099999, 0.0, 1e-12, -1e-12]]:
        assert L.cilqr_oracle_barrier_value(C.byref(p), float(g)) == R.ref_barrier_value(float(g))


def _solve_both(oracle, batch, b):
    r = ref.ilqr_solve(batch, b)
    o = oracle.solve(batch, b, hist=True)
    return r, o


@pytest.mark.parametrize("seed,B,N,road", [(11, 24, 50, "gentle"), (12, 8, 100, "gentle"), (13, 12, 80, "shipped"),
                                           (14, 6, 200, "gentle"), (15, 16, 30, "shipped")])
def test_whole_solves_equal_the_reference_bit_for_bit(oracle, seed, B, N, road):
    """IlqrOptimizer::Optimize of the reference itself against the oracle: the iqr initial guess, the returned
    states and controls, and the five-component cost of every accepted iterate (cost()), all bit-identical."""
    batch = scenarios.generate(seed, 0, B, N=N, road_name=road)
    exits = set()
    for b in range(B):
        r, o = _solve_both(oracle, batch, b)
        assert np.array_equal(r["init_states"], o["init_states"]) and np.array_equal(r["init_controls"], o["init_controls"])
        assert len(r["cost_hist"]) == len(o["cost_hist"]) and np.array_equal(r["cost_hist"], np.array(o["cost_hist"]))
        assert np.array_equal(r["states"], o["states"]) and np.array_equal(r["controls"], o["controls"])
        exits.add(o["status"])
    print(f"\\n[solver pin] seed {seed} N={N} {road}: {B} solves bit-identical; exits seen {sorted(exits)}")


def test_solver_fixture_equals_the_reference():
    """tests/golden/cilqr_golden_v1.npz (a GPU-test target) holds exactly what the reference's solver returns."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cilqr_golden_v1.npz"))
    batch = scenarios.ScenarioBatch(int(z["N"]), int(z["M_max"]), int(z["S"]), z["start"], z["coarse"], z["corridor"],
                                    z["corridor_cnt"], z["lane_left"], z["lane_right"])
    for b in range(batch.B):
        r = ref.ilqr_solve(batch, b)
        assert np.array_equal(r["states"], z["states"][b]) and np.array_equal(r["controls"], z["controls"][b])
        assert np.array_equal(r["init_states"], z["init_states"][b])


# ---- the reference's own Corridor (corridor.cc compiled unmodified; cv::convexHull = the cv2-pinned restatement) ---
def test_build_corridor_equals_the_reference():
    from oracle import corridor_binding as cb
    M = 64
    knots = 0
    for seed, B, N, n_obs in ((5, 12, 50, 20), (6, 8, 80, 11), (7, 6, 30, 5)):
        _, ci = scenarios.generate_with_obstacles(seed, 0, B, N=N, n_obs=n_obs)
        cor, cnt, poly, code = cb.plan_batch(ci.traj, ci.obs_points, ci.obs_cnt, M)
        for b in range(ci.B):
            for k in range(ci.K):
                n = int(ci.obs_cnt[b, k])
                m, cons, pl = ref.build_corridor(*ci.traj[b, k], ci.obs_points[b, k, :n], cap=M)
                assert m == cnt[b, k] and code[b, k] == 0
                assert np.array_equal(cons, cor[b, k, :m]) and np.array_equal(pl, poly[b, k, :m])
                knots += 1
    assert knots > 1400
    # an empty obstacle set gives the box of AddCorridorPoints
    m, cons, _ = ref.build_corridor(3.0, -1.0, 0.4, np.zeros((0, 2)))
    rc, cons2, _ = cb.build_corridor(3.0, -1.0, cb.add_corridor_points(3.0, -1.0, 0.4))
    assert m == 4 and rc == 0 and np.array_equal(cons, cons2)


def test_corridor_fixture_equals_the_reference():
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "corridor_golden_v1.npz"))
    for b in range(z["traj"].shape[0]):
        for k in range(z["traj"].shape[1]):
            n = int(z["obs_cnt"][b, k])
            m, cons, pl = ref.build_corridor(*z["traj"][b, k], z["obs_points"][b, k, :n])
            assert m == z["corridor_cnt"][b, k]
            assert np.array_equal(cons, z["corridor"][b, k, :m]) and np.array_equal(pl, z["polygon"][b, k, :m])


def test_lane_constraints_equal_the_reference():
    from oracle import corridor_binding as cb
    for name in ("gentle", "shipped"):
        rd = scenarios.road(name)
        s = np.arange(rd.s_min, rd.s_max + 1e-9, 0.1)
        for lat, left in ((2.5, True), (-6.0, False), (0.7, True)):
            bd = np.stack(rd.frenet_to_xy(s, lat), axis=1)
            n, seg = ref.lane_constraints(bd, left)
            n2, seg2 = cb.lane_constraints(bd, left)
            assert n == n2 and np.array_equal(seg, seg2)
    assert ref.lane_constraints(np.zeros((5, 2)), True)[0] == -1 == cb.lane_constraints(np.zeros((5, 2)), True)[0]


def test_solver_exits_equal_the_reference(oracle):
    """The rarer exits of Optimize -- absolute-tolerance convergence (:281,287) and the lambda-overflow "unsolved"
    exit (:302-307) -- pinned on scenarios selected for them (the relative-tolerance exit dominates the rest)."""
    batch = scenarios.generate(32, 0, 2048, N=30, road_name="shipped")
    picks = {0: [247, 959, 1022, 1550], 3: [34, 513, 1313, 1420]}
    for status, ids in picks.items():
        for b in ids:
            r, o = _solve_both(oracle, batch, b)
            assert o["status"] == status
            assert len(r["cost_hist"]) == len(o["cost_hist"]) and np.array_equal(r["cost_hist"], np.array(o["cost_hist"]))
            assert np.array_equal(r["states"], o["states"]) and np.array_equal(r["controls"], o["controls"])


def test_max_iteration_exit_and_long_solves_equal_the_reference(oracle):
    """With the cost tolerances switched off the solve runs to `max_iter_num` (:312-316) or to the lambda-overflow
    exit -- up to 200 iterations of accumulated rounding, still bit-identical.  (The gradient-norm exit :236-241 is
    pinned by test_gradient_norm_exit_equals_the_reference below.)"""
    batch = scenarios.generate(41, 0, 10, N=30)
    seen = set()
    for mi, at, rt in ((3, 1e-2, 1e-2), (200, 0.0, 0.0)):
        p = oracle.default_params()
        p.max_iter_num, p.abs_cost_tol, p.rel_cost_tol = mi, at, rt
        for b in range(batch.B):
            o = oracle.solve(batch, b, params=p, hist=True)
            r = ref.ilqr_solve(batch, b, overrides=(mi, at, rt))
            assert len(r["cost_hist"]) == len(o["cost_hist"]) and np.array_equal(r["cost_hist"], np.array(o["cost_hist"]))
            assert np.array_equal(r["states"], o["states"]) and np.array_equal(r["controls"], o["controls"])
            seen.add(o["status"])
    assert 4 in seen and 3 in seen


from picks import GRAD_EXIT_PICKS  # noqa: E402  (tests/picks.py; tests/test_gpu_parity.py solves the same ones on the GPU)


def test_gradient_norm_exit_equals_the_reference(oracle):
    """The gradient-norm exit (ilqr_optimizer.cc:235-241, CalGradientNorm :322-332): `gnorm < 1e-6 && lambda < 1e-5`.
    With the default tolerances (1e-2) the cost tests always fire first; with both set to 0 the iteration runs on
    until the feed-forward gains vanish while lambda is still small.  Reference and oracle must leave through the
    same exit after the same accepts, bit for bit."""
    seen = 0
    for (seed, N), ids in GRAD_EXIT_PICKS.items():
        batch = scenarios.generate(seed, 0, max(ids) + 1, N=N)
        p = oracle.default_params()
        p.abs_cost_tol = p.rel_cost_tol = 0.0
        for b in ids:
            o = oracle.solve(batch, b, params=p, hist=True)
            r = ref.ilqr_solve(batch, b, overrides=(200, 0.0, 0.0))
            assert o["status"] == 2, (seed, N, b, o["status"])
            assert len(r["cost_hist"]) == len(o["cost_hist"]) and np.array_equal(r["cost_hist"], np.array(o["cost_hist"]))
            assert np.array_equal(r["states"], o["states"]) and np.array_equal(r["controls"], o["controls"])
            # the reference pushes no cost / iterate on this exit: the last cost_ entry is the last accept's
            assert len(o["cost_hist"]) == o["accepted"] + 1
            seen += 1
    assert seen == 8


# ---- the reference's own Tracker (tracker.cc + linear_quadratic_regulator.cc compiled unmodified) -----------------
def test_tracker_equals_the_reference_bit_for_bit():
    """Tracker::Plan (algorithm/ilqr/tracker.cc:169-215: 800 simulation steps, two DARE fixed-point solves each, RK4)
    of the reference itself against oracle/tracker_oracle.c on DP-planned coarse trajectories: every field of every
    trajectory point bit-identical (about 50 000 DARE iterations per plan)."""
    from oracle import dp_binding as dpo
    from oracle import tracker_binding as tb
    dpo.build()
    assert tb.ref_lib() is not None, "oracle/_ref/libcilqr_ref_tracker.so missing: run `make -C oracle _ref`"
    n = 0
    for seed, B in ((3, 6), (4, 6)):
        db = scenarios.generate_dp(seed, B, n_obs=6)
        barrier = dpo.build_barrier(db.ref)
        for b in range(B):
            sc = dpo.Scene(db.ref, barrier, db.static_poly[b], db.static_nv[b], db.dyn_time[b], db.dyn_samples[b],
                           db.dyn_poly[b], db.dyn_nv[b])
            ok, traj, _, _ = dpo.plan(sc, *db.start[b])
            if not ok or np.isnan(traj).any():
                continue
            for v0 in (10.0, 4.0):
                s13 = tb.start_record([db.start[b][0] + 0.3, db.start[b][1] - 0.2, db.start[b][2] + 0.05, v0])
                oko, to, its = tb.plan(s13, traj)
                rc, tr = tb.ref_plan(s13, traj)
                assert rc == int(oko) == 1 and its > 10000
                assert np.array_equal(to, tr)
                n += 1
    assert n >= 12
    # a follow trajectory whose time stamps the simulation clock cannot reach: both report failure
    bad = traj.copy()
    bad[:, 0] *= 0.999
    s13 = tb.start_record([*db.start[b], 10.0])
    assert tb.ref_plan(s13, bad)[0] in (0, -1) and not tb.plan(s13, bad)[0]
