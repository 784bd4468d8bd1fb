"""The corridor oracle (oracle/corridor_oracle.c) against what pins it:

* the REAL OpenCV hull -- committed cv2 outputs (tests/golden/corridor_hull_v1.npz) and, when cv2 is
  importable, live fuzzing (cv::convexHull is what the reference calls, corridor.cc:184,218,242);
* an independent NumPy restatement of BuildCorridor whose hulls are cv2's (tests/golden/corridor_golden_v1.npz
  and live);
* analytic properties of the construction (the knot is strictly inside, no obstacle point is inside).
"""
import os

import numpy as np
import pytest

from cilqr_b200 import scenarios
from oracle import corridor_binding as cb

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _same_up_to_rotation(a, b):
    return len(a) == len(b) and (len(a) == 0 or any(np.array_equal(np.roll(b, s, axis=0), a) for s in range(len(a))))


def test_hull_matches_committed_opencv_outputs():
    z = np.load(os.path.join(GOLD, "corridor_hull_v1.npz"))
    po = np.r_[0, np.cumsum(z["n_points"])]
    ho = np.r_[0, np.cumsum(z["n_hull"])]
    assert len(z["n_points"]) == 480
    for i in range(len(z["n_points"])):
        p = z["points"][po[i]:po[i + 1]]
        ref = z["hull"][ho[i]:ho[i + 1]]
        got = cb.convex_hull(p, bool(z["clockwise"][i]))
        assert np.array_equal(got, ref), (i, ref, got)


def test_hull_live_against_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(99)
    for it in range(3000):
        kind = it % 4
        n = int(rng.integers(1, 90))
        if kind == 0:
            p = rng.normal(size=(n, 2)) * 30
        elif kind == 1:
            q = rng.permutation(100)[:min(n, 100)]
            p = np.c_[q % 10, q // 10].astype(float)  # distinct, heavily collinear
        elif kind == 2:
            t = rng.uniform(0, 2 * np.pi, n)
            p = np.c_[np.cos(t), np.sin(t)] * 290
        else:  # duplicates everywhere: OpenCV's choice AMONG coincident points is not reproduced
            p = rng.integers(-3, 4, size=(n, 2)).astype(float)
        p = np.ascontiguousarray(p, np.float32)
        for cw in (False, True):
            ref = cv2.convexHull(p, clockwise=cw, returnPoints=False).ravel()
            got = cb.convex_hull(p, cw)
            if kind < 3:
                assert np.array_equal(got, ref), (kind, n, cw)
            else:  # same hull polygon, vertex for vertex (coordinates), possibly another copy of a point
                assert _same_up_to_rotation(p[ref], p[got]), (n, cw)


def test_build_corridor_matches_committed_cv2_backed_restatement():
    z = np.load(os.path.join(GOLD, "corridor_golden_v1.npz"))
    M = z["corridor"].shape[2]
    cor, cnt, poly, code = cb.plan_batch(z["traj"], z["obs_points"], z["obs_cnt"], M)
    assert not code.any()
    assert np.array_equal(cnt, z["corridor_cnt"])
    m = np.arange(M)[None, None, :] < cnt[..., None]
    assert np.array_equal(cor[m], z["corridor"][m])  # bit-exact: same expression types, same hull
    assert np.array_equal(poly[m], z["polygon"][m])


def test_build_corridor_live_against_cv2_backed_restatement():
    pytest.importorskip("cv2")
    from oracle import corridor_numpy as cn
    _, ci = scenarios.generate_with_obstacles(11, 0, 8, N=30)
    for b in range(ci.B):
        for k in range(0, ci.K, 3):
            n = ci.obs_cnt[b, k]
            pts = np.r_[ci.obs_points[b, k, :n], cn.add_corridor_points(*ci.traj[b, k])]
            assert np.array_equal(pts[n:], cb.add_corridor_points(*ci.traj[b, k]))
            rc, cons, pl = cn.build_corridor(ci.traj[b, k, 0], ci.traj[b, k, 1], pts)
            rc2, cons2, pl2 = cb.build_corridor(ci.traj[b, k, 0], ci.traj[b, k, 1], pts)
            assert rc == rc2 == 0
            assert np.array_equal(cons, cons2) and np.array_equal(pl, pl2)


def test_corridor_contains_the_knot_and_excludes_every_obstacle_point():
    _, ci = scenarios.generate_with_obstacles(3, 0, 16, N=40)
    M = 32
    cor, cnt, poly, code = cb.plan_batch(ci.traj, ci.obs_points, ci.obs_cnt, M)
    assert not code.any() and cnt.min() >= 3
    for b in range(ci.B):
        for k in range(ci.K):
            pl = cor[b, k, :cnt[b, k]]
            nrm = np.hypot(pl[:, 0], pl[:, 1])
            x, y = ci.traj[b, k, :2]
            assert ((pl[:, 0] * x + pl[:, 1] * y - pl[:, 2]) / nrm < -1e-6).all()  # a x + b y < c at the knot
            pts = ci.obs_points[b, k, :ci.obs_cnt[b, k]]
            near = (np.abs(pts[:, 0] - x) <= 25) & (np.abs(pts[:, 1] - y) <= 25)
            viol = (pts[near, None, 0] * pl[None, :, 0] + pts[near, None, 1] * pl[None, :, 1] - pl[None, :, 2]) / nrm
            assert (viol.max(axis=1) > -2e-3).all()  # float32 hull: points may sit a millimetre inside a face


def test_box_only_corridor_is_the_box():
    # no obstacle in range: the corridor is the +-10 m heading-aligned box of AddCorridorPoints
    th = 0.3
    pts = cb.add_corridor_points(5.0, -2.0, th)
    rc, cons, poly = cb.build_corridor(5.0, -2.0, pts)
    assert rc == 0 and len(cons) == 4
    nrm = np.hypot(cons[:, 0], cons[:, 1])
    dist = (cons[:, 2] - cons[:, 0] * 5.0 - cons[:, 1] * -2.0) / nrm
    assert np.allclose(dist, 10.0, atol=1e-3)
    ang = np.sort(np.mod(np.arctan2(cons[:, 1], cons[:, 0]) - th, 2 * np.pi))
    assert np.allclose(ang, [0, np.pi / 2, np.pi, 3 * np.pi / 2], atol=1e-4) or np.allclose(
        np.sort(np.mod(ang + 1e-3, 2 * np.pi)), np.array([0, np.pi / 2, np.pi, 3 * np.pi / 2]) + 1e-3, atol=1e-4)


def test_failure_codes():
    cfg = cb.default_config()
    rc, _, _ = cb.build_corridor(0.0, 0.0, np.zeros((0, 2)))
    assert rc == 1  # corridor.cc:127-130
    rc, _, _ = cb.build_corridor(0.0, 0.0, np.array([[1.0, 0], [0, 1.0], [30.0, 0]]))
    assert rc == 2  # fewer than four points pass the +-25 m filter, corridor.cc:179-182
    pts = cb.add_corridor_points(0.0, 0.0, 0.0, cfg)
    cons = np.zeros((2, 3))
    poly = np.zeros((2, 2))
    import ctypes as C
    m = C.c_int(0)
    rc = cb.lib().corr_build_corridor(C.byref(cfg), 0.0, 0.0, pts.ctypes.data, 8, cons.ctypes.data, poly.ctypes.data,
                                      2, C.byref(m))
    assert rc == 4 and m.value == 0


def test_lane_constraints_match_the_generator_and_the_reference_conventions():
    rd = scenarios.road("gentle")
    s = np.arange(rd.s_min, rd.s_max + 1e-9, 0.1)
    for side, lat in enumerate((2.5, -6.0)):
        bx, by = rd.frenet_to_xy(s, lat)
        n, seg = cb.lane_constraints(np.stack([bx, by], axis=1), is_left=(side == 0))
        want = scenarios._lane_constraints(rd.lane_pts[side][None], left=(side == 0))[0]
        assert n == len(want)
        assert np.array_equal(seg, want)
        # left segments run pt[i] -> pt[i-1], right pt[i-1] -> pt[i] (corridor.cc:279,300); the road centre
        # satisfies a x + b y < c for both
        cx, cy = rd.frenet_to_xy(s[::50], 0.0)
        j = n // 2
        mid = 0.5 * (seg[j, 3:5] + seg[j, 5:7])
        i = np.argmin((cx - mid[0]) ** 2 + (cy - mid[1]) ** 2)
        assert seg[j, 0] * cx[i] + seg[j, 1] * cy[i] < seg[j, 2]
    assert cb.lane_constraints(np.zeros((5, 2)), True)[0] == -1  # fewer than two sampled points
    line = np.stack([np.arange(0, 100, 0.1), np.zeros(1000)], axis=1)
    assert cb.lane_constraints(line, False, cap=3)[0] == -2
