"""Independent NumPy restatement of Corridor::BuildCorridor (algorithm/ilqr/corridor.cc:122-263) that calls
the REAL OpenCV hull (cv2.convexHull) where the reference calls cv::convexHull.

TEST INFRASTRUCTURE ONLY.  Used to cross-check oracle/corridor_oracle.c (whose hull is a restatement of
OpenCV's) and to generate tests/golden/corridor_golden_v1.npz.  float32 / float64 types follow the
reference's C++ expression types (cv::Point2f, Eigen::Vector2f/3f are float; everything else double).
"""
from __future__ import annotations

import numpy as np

K_MATH_EPSILON = 1e-10
f32 = np.float32

DEFAULT_CFG = dict(max_diff_x=25.0, max_diff_y=25.0, radius=150.0, max_axis_x=10.0, max_axis_y=10.0,
                   lane_segment_length=5.0)


def _hull(points_f32: np.ndarray, clockwise: bool) -> np.ndarray:
    import cv2  # imported lazily: only the tests that pin against OpenCV need it

    return cv2.convexHull(np.ascontiguousarray(points_f32, np.float32), clockwise=clockwise,
                          returnPoints=False).ravel()


def add_corridor_points(x: float, y: float, theta: float, cfg=DEFAULT_CFG) -> np.ndarray:
    """corridor.cc:89-120 with is_multiple_sample = false: every corner twice."""
    c, s = np.cos(theta), np.sin(theta)
    dx1, dy1 = c * cfg["max_axis_x"], s * cfg["max_axis_x"]
    dx2, dy2 = s * cfg["max_axis_y"], -c * cfg["max_axis_y"]
    corners = [(x + dx1 + dx2, y + dy1 + dy2), (x + dx1 - dx2, y + dy1 - dy2),
               (x - dx1 - dx2, y - dy1 - dy2), (x - dx1 + dx2, y - dy1 + dy2)]
    out = []
    for i in range(4):
        nx = (i + 1) % 4
        for ratio in (0.0, 1.0):
            out.append((corners[i][0] * (1 - ratio) + corners[nx][0] * ratio,
                        corners[i][1] * (1 - ratio) + corners[nx][1] * ratio))
    return np.array(out, np.float64)


def build_corridor(origin_x: float, origin_y: float, points: np.ndarray, cfg=DEFAULT_CFG):
    """Returns (code, constraints [m][3], polygon [m][2]); code 0 = ok, 2 = fewer than 4 points."""
    points = np.asarray(points, np.float64).reshape(-1, 2)
    n = len(points)
    if n == 0:
        return 1, None, None
    dx = points[:, 0] - origin_x
    dy = points[:, 1] - origin_y
    norm = np.sqrt(dx * dx + dy * dy)
    keep = ~((np.abs(dx) > cfg["max_diff_x"]) | (np.abs(dy) > cfg["max_diff_y"])) & ~(np.abs(norm) < K_MATH_EPSILON)
    filt = points[keep]
    nf = len(filt)
    dx, dy, norm = dx[keep], dy[keep], norm[keep]
    R = cfg["radius"]
    flip = np.zeros((n + 1, 2), np.float32)
    flip[:nf, 0] = (dx + 2 * (R - norm) * dx / norm).astype(np.float32)
    flip[:nf, 1] = (dy + 2 * (R - norm) * dy / norm).astype(np.float32)
    if nf < 4:
        return 2, None, None
    vidx = _hull(flip, False)
    if (vidx >= nf).any():
        raise NotImplementedError("origin on the flipped hull (never happens with the corridor box points)")
    vdata = filt[vidx].astype(np.float32)  # cv::Point2f(filterd_points[v].x(), ...)
    interior_x, interior_y = float(origin_x), float(origin_y)
    vidx2 = _hull(vdata, False)
    nv, nh2 = len(vdata), len(vidx2)
    tcons = []
    for j in range(nh2):
        j1 = (j + 1) % nh2
        ray = vdata[vidx2[j1]] - vdata[vidx2[j]]  # float32
        nrm = np.array([ray[1], -ray[0]], np.float32)
        z = f32(nrm[0] * nrm[0]) + f32(nrm[1] * nrm[1])
        if z > 0:
            nrm = nrm / np.sqrt(f32(z))
        idx = int(vidx2[j])
        while idx != int(vidx2[j1]):
            c = (float(vdata[idx, 0]) - interior_x) * float(nrm[0]) + (float(vdata[idx, 1]) - interior_y) * float(nrm[1])
            tcons.append((nrm[0], nrm[1], f32(c)))
            idx = (idx + 1) % nv
    tcons = np.array(tcons, np.float32).reshape(-1, 3)
    dual = np.stack([tcons[:, 0] / tcons[:, 2], tcons[:, 1] / tcons[:, 2]], axis=1).astype(np.float32)
    dv = dual[_hull(dual, True)]
    nd = len(dv)
    poly = np.zeros((nd, 2))
    for i in range(nd):
        i1 = (i + 1) % nd
        ray = dv[i1] - dv[i]
        c = float(f32(ray[1] * dv[i, 0]) - f32(ray[0] * dv[i, 1]))  # float expression, then widened
        poly[i] = (interior_x + float(ray[1]) / c, interior_y - float(ray[0]) / c)
    cons = np.zeros((nd, 3))
    for i in range(nd):
        i1 = (i + 1) % nd
        r = poly[i1] - poly[i]
        cons[i] = (-r[1], r[0], -r[1] * poly[i, 0] + r[0] * poly[i, 1])
    return 0, cons, poly
