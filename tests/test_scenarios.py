"""Synthetic scenario generator: determinism, shapes, and the feasibility guarantees the
reference's own corridor builder gives by construction."""
import math

import numpy as np

from cilqr_b200 import scenarios


def test_deterministic_and_chunk_addressable():
    a = scenarios.generate(5, 0, 2048, N=20, workers=1)
    b = scenarios.generate(5, 1024, 1024, N=20, workers=2)
    for x, y in ((a.start, b.start), (a.coarse, b.coarse), (a.corridor, b.corridor), (a.corridor_cnt, b.corridor_cnt),
                 (a.lane_left, b.lane_left), (a.lane_right, b.lane_right)):
        assert np.array_equal(x[1024:], y)
    c = scenarios.generate(6, 0, 8, N=20)
    assert not np.array_equal(c.start, a.start[:8])


def test_shapes_and_counts():
    b = scenarios.generate(1, 0, 16, N=50)
    assert b.start.shape == (16, 4) and b.coarse.shape == (16, 51, 6) and b.corridor.shape == (16, 51, 20, 3)
    assert b.corridor_cnt.dtype == np.int32 and b.corridor_cnt.min() >= 4 and b.corridor_cnt.max() <= 20
    assert b.lane_left.shape == (16, 40, 7) and b.lane_right.shape == (16, 40, 7)
    assert scenarios.algorithmic_bytes(100, 20, 40) == 8 * (4 + 606 + 6060 + 560) + 404 + 8 * (606 + 200 + 8)


def test_coarse_point_strictly_inside_every_shrunk_plane():
    """Corridor::Plan builds each polygon around the coarse point; the generator must leave the point
    feasible after the (r_disc + margin) shrink of ilqr_optimizer.cc:438-473."""
    b = scenarios.generate(2, 0, 64, N=60)
    r = math.hypot(1.942 / 2.0, (0.96 + 1.0 + 0.929) / 2.0 / 5) + 0.2
    a, bb, c = b.corridor[..., 0], b.corridor[..., 1], b.corridor[..., 2]
    nrm = np.hypot(a, bb)
    used = np.arange(b.M_max)[None, None, :] < b.corridor_cnt[..., None]
    g = a * b.coarse[:, :, None, 0] + bb * b.coarse[:, :, None, 1] - (c - r * nrm)
    assert np.all(g[used] <= 1e-9)
    assert np.all(nrm[used] > 0)


def test_lane_half_planes_contain_the_road_interior():
    """HalfPlaneConstraint (corridor.cc:322-331): segment start/end as stored; the coarse path lies on
    the feasible side (a x + b y < c) of the nearest left and right segments."""
    b = scenarios.generate(3, 0, 8, N=40)
    for lane in (b.lane_left, b.lane_right):
        st, en = lane[..., 3:5], lane[..., 5:7]
        seglen = np.hypot(*(en - st).transpose(2, 0, 1))
        assert np.all(seglen > 4.0) and np.all(seglen < 6.5)
        mid = 0.5 * (st + en)
        for i in range(b.B):
            p = b.coarse[i, 0, :2]
            j = np.argmin(np.hypot(*(mid[i] - p).T))
            assert lane[i, j, 0] * p[0] + lane[i, j, 1] * p[1] < lane[i, j, 2]
