"""Known-answer tests of the CPU oracle's primitives (the reference ships none: SURVEY.md section 4).
Each check is an analytic property of the reference formula the oracle restates."""
import ctypes as C
import math

import numpy as np
import pytest


def _arr(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def test_default_params_match_reference_headers(oracle):
    p = oracle.default_params()
    # vehicle_param.h:26-64
    assert (p.front_hang_length, p.wheel_base, p.rear_hang_length, p.width) == (0.96, 1.0, 0.929, 1.942)
    assert (p.max_velocity, p.min_acceleration, p.max_acceleration) == (20.0, -5.0, 5.0)
    assert p.delta_max == pytest.approx(40.0 / 180 * math.pi, abs=0) and p.delta_min == -p.delta_max
    assert p.delta_rate_max == p.delta_max / 3.0
    # planner_config.h:45-66, barrier_function.h:143-146 (IlqrConfig::t = 100 is never read)
    assert (p.w_x_target, p.w_y_target, p.w_theta, p.w_jerk, p.w_delta_rate) == (0.5, 0.5, 1e-3, 1.0, 1.0)
    assert (p.barrier_t, p.barrier_eps, p.num_of_disc, p.max_iter_num) == (5.0, 0.01, 5, 200)


def test_disc_radius(oracle):
    # ilqr_optimizer.cc:97-104: hypot(1.942/2, 2.889/2/5)
    r = oracle.lib().cilqr_oracle_disc_radius(C.byref(oracle.default_params()))
    assert r == pytest.approx(1.01307, abs=1e-5)
    assert r == math.hypot(1.942 / 2.0, (0.96 + 1.0 + 0.929) / 2.0 / 5)


@pytest.mark.parametrize("a,expect", [
    (0.0, 0.0), (math.pi, -math.pi), (-math.pi, -math.pi), (3 * math.pi, -math.pi),
    (math.pi - 1e-9, math.pi - 1e-9), (-math.pi - 1e-9, math.pi - 1e-9), (7.0, 7.0 - 2 * math.pi),
    (-7.0, -7.0 + 2 * math.pi), (100.0, 100.0 - 32 * math.pi)])
def test_normalize_angle(oracle, a, expect):
    # math_utils.cpp:53-59 -> [-pi, pi)
    got = oracle.lib().cilqr_oracle_normalize_angle(a)
    assert got == pytest.approx(expect, abs=1e-9)
    assert -math.pi <= got < math.pi


def test_barrier_is_c1_at_minus_eps_and_matches_formulas(oracle):
    L, p = oracle.lib(), oracle.default_params()
    eps, rt = p.barrier_eps, 1.0 / p.barrier_t
    v = lambda g: L.cilqr_oracle_barrier_value(C.byref(p), g)  # noqa: E731
    d = lambda g: L.cilqr_oracle_barrier_dcoef(C.byref(p), g)  # noqa: E731
    # log branch (barrier_function.h:107-108,118-119)
    for g in (-5.0, -1.0, -0.5, -0.011):
        assert v(g) == pytest.approx(-rt * math.log(-g), rel=1e-15)
        assert d(g) == pytest.approx(-rt / g, rel=1e-15)
    # C0 / C1 continuity at g = -eps (value and first derivative of the quadratic extension)
    assert v(-eps - 1e-12) == pytest.approx(v(-eps), abs=1e-9)
    assert d(-eps - 1e-12) == pytest.approx(d(-eps), rel=1e-9)
    # derivative coefficient is d(value)/dg on both branches
    for g in (-2.0, -0.02, -0.005, 0.0, 0.3):
        h = 1e-7
        assert (v(g + h) - v(g - h)) / (2 * h) == pytest.approx(d(g), rel=1e-5)


def test_barrier_hessian_quirk_q6(oracle):
    """Relaxed branch: the reference reuses the gradient coefficient for the outer product and
    drops ddx (barrier_function.h:137-139), so the 'Hessian' there is not the second derivative."""
    L, p = oracle.lib(), oracle.default_params()
    co, cd = C.c_double(), C.c_double()
    L.cilqr_oracle_barrier_hcoef(C.byref(p), -0.5, C.byref(co), C.byref(cd))
    assert co.value == pytest.approx(0.2 / 0.25) and cd.value == pytest.approx(0.2 / -0.5)
    L.cilqr_oracle_barrier_hcoef(C.byref(p), 0.1, C.byref(co), C.byref(cd))
    assert co.value == pytest.approx(0.2 * (0.1 + 0.02) / 1e-4) and cd.value == 0.0
    assert co.value == pytest.approx(L.cilqr_oracle_barrier_dcoef(C.byref(p), 0.1))


def test_segment_distance_cases(oracle):
    L = oracle.lib()
    seg, ps = _arr([0, 0, 0, 1.0, 2.0, 4.0, 6.0])  # 3-4-5 direction, length 5
    dist = lambda x, y: L.cilqr_oracle_segment_distance(ps, x, y)  # noqa: E731
    assert dist(1.0, 2.0) == 0.0
    assert dist(-2.0, -2.0) == pytest.approx(5.0)            # before start: proj <= 0
    assert dist(4.0 + 3.0, 6.0 + 4.0) == pytest.approx(5.0)  # past end: proj >= length
    assert dist(1.0 + 1.5 - 0.8 * 2, 2.0 + 2.0 + 0.6 * 2) == pytest.approx(2.0)  # interior: |cross|
    deg, pd = _arr([0, 0, 0, 1.0, 1.0, 1.0, 1.0])  # degenerate: length <= 1e-10
    assert L.cilqr_oracle_segment_distance(pd, 4.0, 5.0) == pytest.approx(5.0)


def _dyn(oracle, x, u):
    L, p = oracle.lib(), oracle.default_params()
    x, px = _arr(x)
    u, pu = _arr(u)
    out, po = _arr(np.zeros(6))
    L.cilqr_oracle_dynamics(C.byref(p), px, pu, po)
    return out


def test_dynamics_straight_line_and_wrap(oracle):
    # v = 10 along +x, no steering: x advances by v*dt + a*dt^2/2 (midpoint rule is exact here)
    n = _dyn(oracle, [0, 0, 0, 10.0, 2.0, 0.0], [1.0, 0.0])
    assert n[0] == pytest.approx(10.0 * 0.1 + 0.5 * 2.0 * 0.01)
    assert n[1] == pytest.approx(0.0, abs=1e-15)
    assert n[3] == pytest.approx(10.0 + 0.1 * (2.0 + 0.05 * 1.0))
    assert n[4] == pytest.approx(2.1)
    # theta is wrapped into [-pi, pi) after the step (vehicle_model.cc:116)
    n = _dyn(oracle, [0, 0, math.pi - 0.01, 10.0, 0.0, 0.3], [0.0, 0.0])
    assert -math.pi <= n[2] < math.pi and n[2] < 0


def test_dynamics_jacobian_vs_finite_differences(oracle):
    """A, B against central differences of Dynamics.  The reference's hand derivative
    (vehicle_model.cc:61-85) is the exact Jacobian of the midpoint step *except* where it drops
    second-order dt terms (quirk Q20): A(2,3..5) and B(2,1) use tan(delta + dt/2*delta_rate) and
    v in place of the mid-stage quantities.  Exact entries are held to 1e-6, the approximated
    ones to their documented O(dt^2) gap."""
    L, p = oracle.lib(), oracle.default_params()
    rng = np.random.default_rng(1)
    for _ in range(20):
        x = np.array([rng.normal(), rng.normal(), rng.uniform(-3, 3), rng.uniform(1, 15), rng.uniform(-2, 2),
                      rng.uniform(-0.5, 0.5)])
        u = np.array([rng.uniform(-5, 5), rng.uniform(-0.2, 0.2)])
        A, pA = _arr(np.zeros(36))
        Bm, pB = _arr(np.zeros(12))
        xx, px = _arr(x)
        uu, pu = _arr(u)
        L.cilqr_oracle_dynamics_jacobian(C.byref(p), px, pu, pA, pB)
        A, Bm = A.reshape(6, 6), Bm.reshape(6, 2)
        h = 1e-6
        fdA = np.stack([(_dyn(oracle, x + h * e, u) - _dyn(oracle, x - h * e, u)) / (2 * h) for e in np.eye(6)], axis=1)
        fdB = np.stack([(_dyn(oracle, x, u + h * e) - _dyn(oracle, x, u - h * e)) / (2 * h) for e in np.eye(2)], axis=1)
        # rows 3..5 and the identity/structural entries are exact
        np.testing.assert_allclose(A[3:], fdA[3:], atol=1e-6)
        np.testing.assert_allclose(Bm[3:], fdB[3:], atol=1e-6)
        np.testing.assert_allclose(A[:2, :3], fdA[:2, :3], atol=1e-6)
        # structure: zero pattern of the reference matrices
        assert A[2, 0] == A[2, 1] == 0.0 and A[2, 2] == 1.0 and A[3, 4] == 0.1
        assert np.count_nonzero(Bm) == 4
        # hand-derived entries: first order in dt agrees
        np.testing.assert_allclose(A[:3, 3:], fdA[:3, 3:], atol=5e-2 * 0.1 * (1 + abs(x[3])))
        np.testing.assert_allclose(Bm[:3], fdB[:3], atol=5e-2 * 0.1 * (1 + abs(x[3])))


def test_nearest_segment_first_minimum_wins(oracle, small_batch):
    """FindNeastLaneSegment uses strict '<' (ilqr_optimizer.cc:612): among equidistant segments the
    lowest index is returned.  Two identical consecutive segments are an exact tie."""
    import copy
    batch = copy.deepcopy(small_batch.slice(0, 1))
    for lane in (batch.lane_left, batch.lane_right):
        lane[0, 4] = lane[0, 3]  # segment 4 := copy of segment 3
    ctx = oracle.Ctx(batch, 0)
    for side, lane in ((0, batch.lane_left[0]), (1, batch.lane_right[0])):
        mx, my = 0.5 * (lane[3, 3] + lane[3, 5]), 0.5 * (lane[3, 4] + lane[3, 6])
        assert ctx.nearest(side, mx + 0.3, my - 0.2) == 3
    ctx.close()


def test_initial_guess_modes_are_consistent_with_iqr(oracle):
    """init_mode (oracle cilqr_oracle_problem): handing the solver iqr's own result back as the caller's guess
    (mode 2: InitGuess, ilqr_optimizer.cc:107-139 / the commented-out line :168) or as controls to roll out open loop
    (mode 1: OpenLoopRollout, slover/ilqr.h:362-370 -- iqr's rollout IS Dynamics applied to its clamped controls,
    :830-841) must reproduce the default solve bit for bit; a different guess must change the solve."""
    import numpy as np
    from cilqr_b200 import scenarios
    batch = scenarios.generate(19, 0, 16, N=40)
    X, U, S, _ = oracle.solve_batch(batch, nthreads=4)
    X0 = np.stack([oracle.solve(batch, b)["init_states"] for b in range(batch.B)])
    U0 = np.stack([oracle.solve(batch, b)["init_controls"] for b in range(batch.B)])
    X2, U2, S2, _ = oracle.solve_batch(batch, nthreads=4, init_mode=2, init_states=X0, init_controls=U0)
    assert np.array_equal(X2, X) and np.array_equal(U2, U) and np.array_equal(S2, S)
    X1, U1, S1, _ = oracle.solve_batch(batch, nthreads=4, init_mode=1, init_controls=U0)
    assert np.array_equal(X1, X) and np.array_equal(U1, U) and np.array_equal(S1, S)
    Xz, Uz, Sz, _ = oracle.solve_batch(batch, nthreads=4, init_mode=1, init_controls=np.zeros_like(U0))
    assert not np.array_equal(Xz, X) and np.isfinite(Xz).all()
    # the open-loop guess with zero controls coasts: its first cost entry differs from iqr's, the solve still improves it
    assert (Sz[:, 0] <= 4).all()
