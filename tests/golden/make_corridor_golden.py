"""Generates the corridor fixtures:

* tests/golden/corridor_hull_v1.npz -- point sets together with the hull indices returned by the REAL
  OpenCV (cv2.convexHull, python cv2 4.13.0 in this image).  These ARE outputs of the reference's
  third-party dependency (cv::convexHull, called at algorithm/ilqr/corridor.cc:184,218,242) and pin the
  hull restatement in oracle/corridor_oracle.c.
* tests/golden/corridor_golden_v1.npz -- a small batch of Corridor::Plan inputs with the outputs of the
  NumPy restatement (oracle/corridor_numpy.py), whose three hulls per knot are computed by cv2 itself.
  The arithmetic around the hulls is a restatement (the reference needs ROS/Eigen/OpenCV C++ to build).

Re-run only deliberately:  python tests/golden/make_corridor_golden.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cilqr_b200 import scenarios  # noqa: E402
from oracle import corridor_numpy as cn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def hull_cases(rng):
    cases = []
    for it in range(240):
        kind = it % 6
        n = int(rng.integers(1, 70))
        if kind == 0:
            p = rng.normal(size=(n, 2)) * 50
        elif kind == 1:  # distinct grid points: many collinear triples
            q = rng.permutation(81)[:min(n, 81)]
            p = np.c_[q % 9, q // 9].astype(float)
        elif kind == 2:  # circle (all on the hull)
            t = rng.uniform(0, 2 * np.pi, n)
            p = np.c_[np.cos(t), np.sin(t)] * 290
        elif kind == 3:  # one line, distinct
            x = rng.permutation(n).astype(float)
            p = np.c_[x, 2 * x + 1]
        elif kind == 4:  # convex position on a parabola
            x = rng.permutation(n).astype(float)
            p = np.c_[x, 0.25 * x * x]
        else:  # the corridor's own pattern: flipped obstacle points + box corners twice + zero slots
            pts = rng.normal(size=(n, 2)) * 9
            th = rng.uniform(0, 2 * np.pi)
            c, s = np.cos(th) * 10, np.sin(th) * 10
            cs = np.array([[c + s, s - c], [c - s, s + c], [-c - s, -s + c], [-c + s, -s - c]])
            allp = np.r_[pts, cs[[0, 1, 1, 2, 2, 3, 3, 0]]]
            d = np.hypot(allp[:, 0], allp[:, 1])[:, None]
            p = np.r_[allp + 2 * (150 - d) * allp / d, np.zeros((int(rng.integers(1, 6)), 2))]
        p = np.ascontiguousarray(p, np.float32)
        for cw in (False, True):
            cases.append((p, cw, cv2.convexHull(p, clockwise=cw, returnPoints=False).ravel().astype(np.int32)))
    return cases


if __name__ == "__main__":
    rng = np.random.default_rng(20261017)
    cases = hull_cases(rng)
    np.savez_compressed(os.path.join(HERE, "corridor_hull_v1.npz"),
                        points=np.concatenate([c[0] for c in cases]),
                        n_points=np.array([len(c[0]) for c in cases], np.int32),
                        clockwise=np.array([c[1] for c in cases], np.int8),
                        hull=np.concatenate([c[2] for c in cases]),
                        n_hull=np.array([len(c[2]) for c in cases], np.int32),
                        cv2_version=cv2.__version__)

    _, ci = scenarios.generate_with_obstacles(515151, 0, 6, N=24, n_obs=11)
    B, K, P = ci.obs_points.shape[:3]
    M = 24
    cor = np.zeros((B, K, M, 3))
    poly = np.zeros((B, K, M, 2))
    cnt = np.zeros((B, K), np.int32)
    for b in range(B):
        for k in range(K):
            n = ci.obs_cnt[b, k]
            pts = np.r_[ci.obs_points[b, k, :n], cn.add_corridor_points(*ci.traj[b, k])]
            rc, cons, pl = cn.build_corridor(ci.traj[b, k, 0], ci.traj[b, k, 1], pts)
            assert rc == 0
            cnt[b, k] = len(cons)
            cor[b, k, :len(cons)] = cons
            poly[b, k, :len(cons)] = pl
    np.savez_compressed(os.path.join(HERE, "corridor_golden_v1.npz"), traj=ci.traj, obs_points=ci.obs_points,
                        obs_cnt=ci.obs_cnt, corridor=cor, corridor_cnt=cnt, polygon=poly)
    for f in ("corridor_hull_v1.npz", "corridor_golden_v1.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
    print("hull cases", len(cases), "corridor knots", B * K, "planes per knot", cnt.min(), cnt.mean(), cnt.max())
