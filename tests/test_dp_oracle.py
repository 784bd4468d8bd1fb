"""Known-answer and property tests of the DP planner oracle (oracle/dp_oracle.c; reference
algorithm/planner/dp_planner.cpp).  The reference ships no fixtures for this path; the pin against the reference's own compiled planner is
tests/test_reference_pins.py."""
import numpy as np
import pytest

from oracle import dp_binding as dp


def straight_ref(length=300.0, left=2.5, right=6.0):
    s = np.arange(0, length + 0.05, 0.1)
    z = np.zeros_like(s)
    return np.stack([s, s, z, z, z, np.full_like(s, left), np.full_like(s, right)], axis=1)


def arc_ref(radius=40.0, length=200.0):
    s = np.arange(0, length + 0.05, 0.1)
    th = s / radius
    return np.stack([s, radius * np.sin(th), radius * (1 - np.cos(th)), th, np.full_like(s, 1 / radius),
                     np.full_like(s, 2.5), np.full_like(s, 6.0)], axis=1)


def box(cx, cy, hx, hy):
    return np.array([[cx - hx, cy - hy], [cx - hx, cy + hy], [cx + hx, cy + hy], [cx + hx, cy - hy]])


def test_lattice_and_segment_counts():
    # tf = 8, dt = 0.1: layer 0 covers t in [0, 1.6] (17 knots), the others 16 each (dp_planner.cpp:287-298)
    assert dp.lib().dp_num_knots(dp.default_config()) == 81


def test_evaluate_station_and_projection():
    ref = arc_ref()
    r = dp.evaluate_station(ref, 33.33)
    assert r[0] == 33.33 and abs(r[3] - 33.33 / 40.0) < 1e-12
    assert abs(r[1] - 40 * np.sin(33.33 / 40)) < 1e-4 and abs(r[2] - 40 * (1 - np.cos(33.33 / 40))) < 1e-4
    # clamping at both ends (QueryLowerBoundStationPoint, discretized_trajectory.cpp:34-46)
    assert np.allclose(dp.evaluate_station(ref, -5.0)[1:3], ref[0, 1:3] - 5.0 * np.array([1, 0]), atol=1e-2)  # linear extrapolation of the first chord
    # a point 1.5 m left of station 50 projects to (50, +1.5); right is negative (copysign, :184-186)
    th = 50 / 40.0
    px, py = 40 * np.sin(th) - 1.5 * np.sin(th), 40 * (1 - np.cos(th)) + 1.5 * np.cos(th)
    sl = dp.get_projection(ref, px, py)
    assert abs(sl[0] - 50.0) < 0.06 and abs(sl[1] - 1.5) < 1e-3
    sl = dp.get_projection(ref, 40 * np.sin(th) + 2.0 * np.sin(th), 40 * (1 - np.cos(th)) - 2.0 * np.cos(th))
    assert abs(sl[1] + 2.0) < 1e-3


def test_barrier_is_sorted_and_on_the_bounds():
    ref = straight_ref(100.0)
    b = dp.build_barrier(ref)
    assert len(b) == 2 * (int(100.0 / 0.1) + 1)
    assert (np.diff(b[:, 0]) >= 0).all()
    assert set(np.round(b[:, 1], 9)) == {2.5, -6.0}


def test_box_polygon_overlap_rules():
    """Polygon2d::HasOverlap(Box2d) (polygon2d.cpp:150-165) is vertex-in-box OR corner-in-polygon -- it misses a
    thin polygon that crosses the box without either; the oracle keeps that."""
    ref = straight_ref(100.0, left=50.0, right=50.0)  # barrier far away
    cfg = dp.default_config()
    length = cfg.wheel_base + cfg.rear_hang_length + cfg.front_hang_length
    radius = np.hypot(0.25 * length, 0.5 * cfg.width)
    f2x, r2x = 0.75 * length - cfg.rear_hang_length, 0.25 * length - cfg.rear_hang_length
    x, y = 20.0, 0.0
    scene = lambda poly: dp.Scene(ref, static_poly=np.asarray(poly, float)[None])  # noqa: E731
    assert not dp.check_collision(dp.Scene(ref), 0.0, x, y, 0.0)
    assert dp.check_collision(scene(box(x + f2x, y, 0.2, 0.2)), 0.0, x, y, 0.0)          # polygon inside the front box
    assert dp.check_collision(scene(box(x + r2x, y, 5.0, 5.0)), 0.0, x, y, 0.0)          # rear box inside the polygon
    assert not dp.check_collision(scene(box(x + f2x + radius + 0.5, y, 0.2, 0.2)), 0.0, x, y, 0.0)  # disjoint
    thin = [[x + f2x - 5, y - 0.05], [x + f2x - 5, y + 0.05], [x + f2x + 5, y + 0.05], [x + f2x + 5, y - 0.05]]
    assert not dp.check_collision(scene(thin), 0.0, x, y, 0.0)                           # crosses both boxes, missed
    # heading moves the discs: at theta = pi/2 the front disc is above the rear axle
    assert dp.check_collision(scene(box(x, y + f2x, 0.2, 0.2)), 0.0, x, y, np.pi / 2)
    # road barrier points count as obstacles (environment.cpp:62-85)
    tight = dp.Scene(straight_ref(100.0, left=1.0, right=6.0))
    assert dp.check_collision(tight, 0.0, x, 0.5, 0.0) and not dp.check_collision(tight, 0.0, x, -2.0, 0.0)


def test_dynamic_obstacle_sample_selection():
    """CheckDynamicCollision (environment.cpp:124-141): skipped outside [first, last] sample time; otherwise the
    first sample whose time is greater than the query."""
    ref = straight_ref(100.0, left=50.0, right=50.0)
    cfg = dp.default_config()
    f2x = 0.75 * (cfg.wheel_base + cfg.rear_hang_length + cfg.front_hang_length) - cfg.rear_hang_length
    t = np.array([[0.0, 1.0, 2.0]])
    polys = np.stack([box(20 + f2x, 0, 0.2, 0.2), box(60, 0, 0.2, 0.2), box(20 + f2x, 0, 0.2, 0.2)])[None]
    sc = dp.Scene(ref, static_poly=np.zeros((0, 4, 2)), dyn_time=t, dyn_poly=polys)
    assert not dp.check_collision(sc, 0.5, 20.0, 0.0, 0.0)   # upper_bound(0.5) -> sample at t = 1 (far away)
    assert dp.check_collision(sc, 1.5, 20.0, 0.0, 0.0)       # -> sample at t = 2
    assert not dp.check_collision(sc, 2.5, 20.0, 0.0, 0.0)   # after the last sample: skipped
    assert not dp.check_collision(sc, 0.0, 20.0, 0.0, 0.0)   # upper_bound(0) -> t = 1


def test_free_road_keeps_the_centre_at_nominal_speed():
    ok, traj, cost, wp = dp.plan(dp.Scene(straight_ref()), 0.0, 0.0, 0.0)
    assert ok
    # station step 16 m per 1.6 s layer = the nominal 10 m/s (index 3), lateral index NL-1 = the centre line
    assert (wp[:, 0] == 3).all() and (wp[:, 1] == 9).all()
    assert np.allclose(wp[:, 2], 16.0 * np.arange(1, 6))
    # the only cost is the first layer's speed change from the start state's ds0 = 0: |16/1.6| * w = 10
    assert abs(cost - 10.0) < 1e-9
    assert np.allclose(traj[:, 0], 0.1 * np.arange(81)) and np.allclose(traj[:, 3], 0.0)
    assert np.allclose(traj[17:, 6], 10.0) and np.allclose(traj[:16, 6], 16.0 / 17 / 0.1)  # 17 knots in layer 0


def test_obstacle_on_the_centre_line_is_avoided():
    ref = straight_ref()
    cfg = dp.default_config()
    sc = dp.Scene(ref, static_poly=box(40.0, 0.0, 2.0, 1.0)[None])
    ok, traj, cost, wp = dp.plan(sc, 0.0, 0.0, 0.0, cfg)
    assert ok and cost < cfg.dp_w_obstacle
    assert np.abs(traj[:, 3]).max() > 1.0  # it left the centre line
    for k in range(81):  # no knot of the result collides
        assert not dp.check_collision(sc, traj[k, 0], traj[k, 2], traj[k, 3], traj[k, 4])
    # and the result is the arg-min of a brute-force walk over the same lattice decisions: costs are additive, so
    # re-planning from the same start is deterministic
    ok2, traj2, cost2, _ = dp.plan(sc, 0.0, 0.0, 0.0, cfg)
    assert cost2 == cost and np.array_equal(traj, traj2)


def test_blocked_road_reports_failure():
    ref = straight_ref()
    wall = np.stack([box(30.0, y, 0.5, 0.6) for y in np.arange(-7.0, 4.0, 1.0)])
    ok, traj, cost, _ = dp.plan(dp.Scene(ref, static_poly=wall), 0.0, 0.0, 0.0)
    # every lattice path beyond 30 m crosses the wall; stopping short (station index 0..1) stays free
    cfg = dp.default_config()
    assert ok == (cost < cfg.dp_w_obstacle)
    assert traj[-1, 1] < 30.0 or not ok


def test_curved_road_follows_the_arc():
    ref = arc_ref()
    ok, traj, cost, wp = dp.plan(dp.Scene(ref), 0.0, 0.0, 0.0)
    assert ok and (wp[:, 1] == 9).all()
    r = np.hypot(traj[:, 2], traj[:, 3] - 40.0)
    assert np.abs(r - 40.0).max() < 1e-3
    assert np.abs(traj[20:-2, 5] - 1 / 40.0).max() < 2e-3  # kappa from ComputePathProfile
    assert np.allclose(traj[:, 9], np.arctan(traj[:, 5] * 1.0))  # delta = atan(kappa * wheel_base), :270


def test_oracle_reproduces_the_committed_fixture():
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dp_golden_v1.npz"))
    assert np.array_equal(dp.build_barrier(z["ref"]), z["barrier"])
    for b in range(len(z["start"])):
        sc = dp.Scene(z["ref"], z["barrier"], z["static_poly"][b], z["static_nv"][b], z["dyn_time"][b],
                      z["dyn_samples"][b], z["dyn_poly"][b], z["dyn_nv"][b])
        ok, traj, cost, wp = dp.plan(sc, *z["start"][b])
        assert ok == bool(z["ok"][b]) and cost == z["cost"][b]
        assert np.array_equal(wp, z["waypoints"][b]) and np.array_equal(traj, z["trajectory"][b], equal_nan=True)
