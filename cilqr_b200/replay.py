"""Batched replay of recorded scenarios: pickle -> tensors -> DpPlanner -> Corridor -> IlqrOptimizer -> result.

The reference records a scenario with ``script/reference_publisher.py ... serialize`` (``:232-236``: a pickle of
``{"center": CenterLine, "static": Obstacles, "dynamic": DynamicObstacles}``, ROS messages of ``msg/*.msg``) and replays
ONE of them through ROS topics with ``script/pickle_publisher.py:21-40``; ``PlanningNode`` turns the messages into an
``Environment`` (``planning_node.cc:33-80``) and plans on a click (``:82-88``).  This module does the same for MANY
recordings at once without ROS: it reads the pickles (the message classes are not needed -- ``genpy`` messages pickle
as their slot values), builds the flat arrays of ``include/cilqr_b200.h`` exactly as ``PlanningNode`` /
``Environment`` / ``Corridor`` would (centre line, road barrier, obstacle polygons, per-knot obstacle points, lane
boundaries), runs the three batched stages on the GPU through the C ABI, and returns the planner's published result
records (``TrajectoryPlanner::Plan``'s post-processing, ``trajectory_planner.cpp:103-125``).

    python -m cilqr_b200.replay a.pickle b.pickle ... [--out results.npz]

Host-side preparation only; all planning arithmetic happens in libcilqr_b200.so.
"""
from __future__ import annotations

import io
import math
import pickle
import sys
from dataclasses import dataclass

import numpy as np

# slot names of the reference's messages (msg/*.msg) and of the geometry_msgs / std_msgs they embed, in order
_SLOTS = {
    "CenterLine": ["points"],
    "CenterLinePoint": ["s", "x", "y", "theta", "kappa", "left_bound", "right_bound"],
    "Obstacles": ["obstacles"],
    "DynamicObstacles": ["obstacles"],
    "DynamicObstacle": ["polygon", "trajectory"],
    "DynamicTrajectoryPoint": ["time", "x", "y", "theta"],
    "Polygon": ["points"],
    "Point32": ["x", "y", "z"],
    "Point": ["x", "y", "z"],
    "Header": ["seq", "stamp", "frame_id"],
    "Time": ["secs", "nsecs"],
}


class _Msg:
    """Stand-in for any pickled ROS message: genpy.Message pickles as the list of its slot values."""
    _name = "Msg"

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        elif isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):  # (dict, slots dict)
            self.__dict__.update(state[0] or {})
            self.__dict__.update(state[1])
        else:
            for k, v in zip(_SLOTS.get(self._name, []), state):
                setattr(self, k, v)


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if name in _SLOTS or module.startswith(("planning", "geometry_msgs", "std_msgs", "genpy", "rospy")):
            return type(name, (_Msg,), {"_name": name})
        return super().find_class(module, name)


@dataclass
class Scene:
    """One recorded Environment (planning_node.cc:33-80)."""
    center: np.ndarray    # [R,7] s, x, y, theta, kappa, left_bound, right_bound
    static: list          # polygons [V,2]
    dynamic: list         # (times [T], polygons [T,V,2]) per obstacle, already transformed to map coordinates


def _polygon_points(poly) -> np.ndarray:
    return np.array([[p.x, p.y] for p in poly.points], dtype=np.float64).reshape(-1, 2)


def load_pickle(path_or_bytes) -> Scene:
    """reference_publisher.py:232-236 -> Scene.  Python-2 pickles (text protocol) are read with latin-1 strings."""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    rec = _Unpickler(io.BytesIO(data), encoding="latin1").load()
    center = np.array([[getattr(p, k) for k in _SLOTS["CenterLinePoint"]] for p in rec["center"].points], dtype=np.float64)
    static = [_polygon_points(o) for o in rec["static"].obstacles] if rec.get("static") is not None else []
    dynamic = []
    if rec.get("dynamic") is not None:
        for ob in rec["dynamic"].obstacles:  # DynamicObstaclesCallback, planning_node.cc:62-80
            shape = _polygon_points(ob.polygon)
            t = np.array([tp.time for tp in ob.trajectory], dtype=np.float64)
            x = np.array([tp.x for tp in ob.trajectory])
            y = np.array([tp.y for tp in ob.trajectory])
            th = np.array([tp.theta for tp in ob.trajectory])
            c, s = np.cos(th)[:, None], np.sin(th)[:, None]  # math::Pose::transform, pose.h:40-46
            poly = np.stack([x[:, None] + shape[None, :, 0] * c - shape[None, :, 1] * s,
                             y[:, None] + shape[None, :, 0] * s + shape[None, :, 1] * c], axis=-1)
            dynamic.append((t, poly))
    return Scene(center, static, dynamic)


# ---- Environment::set_reference, environment.cpp:20-49 ------------------------------------------------------------
def _normalize_angle(a):
    r = np.fmod(a + math.pi, 2.0 * math.pi)
    return np.where(r < 0.0, r + 2.0 * math.pi, r) - math.pi


def evaluate_station(center: np.ndarray, s: np.ndarray):
    """DiscretizedTrajectory::EvaluateStation (discretized_trajectory.cpp:110-121, LinearInterpolateTrajectory :62-84,
    slerp math_utils.h:208-225), vectorised -> x, y, theta, left_bound, right_bound."""
    st = center[:, 0]
    it = np.searchsorted(st, s, side="left")          # std::lower_bound
    it = np.clip(it, 1, len(st) - 1)
    p0, p1 = center[it - 1], center[it]
    ds = p1[:, 0] - p0[:, 0]
    same = np.abs(ds) < 1e-10
    w = np.where(same, 0.0, (s - p0[:, 0]) / np.where(same, 1.0, ds))
    lerp = lambda a, b: (1 - w) * a + w * b  # noqa: E731
    a0, a1 = _normalize_angle(p0[:, 3]), _normalize_angle(p1[:, 3])
    d = a1 - a0
    d = np.where(d > math.pi, d - 2 * math.pi, np.where(d < -math.pi, d + 2 * math.pi, d))
    theta = np.where(same, p0[:, 3], _normalize_angle(a0 + d * w))
    return lerp(p0[:, 1], p1[:, 1]), lerp(p0[:, 2], p1[:, 2]), theta, lerp(p0[:, 5], p1[:, 5]), lerp(p0[:, 6], p1[:, 6])


def road_barriers(center: np.ndarray):
    """-> (left [n,2], right [n,2], sorted barrier [2n,2]) as Environment::set_reference builds them."""
    s0, s1 = center[0, 0], center[-1, 0]
    n = int((s1 - s0) / 0.1)
    s = s0 + np.arange(n + 1) * 0.1
    x, y, th, lb, rb = evaluate_station(center, s)
    left = np.stack([x - lb * np.sin(th), y + lb * np.cos(th)], axis=1)      # GetCartesian(s, left_bound)
    right = np.stack([x + rb * np.sin(th), y - rb * np.cos(th)], axis=1)     # GetCartesian(s, -right_bound)
    both = np.concatenate([left, right])
    return left, right, np.ascontiguousarray(both[np.argsort(both[:, 0], kind="stable")])


def _dynamic_points_at(times: np.ndarray, polys: np.ndarray, t: float):
    """Environment::QueryDynamicObstacles (environment.cpp:131-149): the obstacle's polygon at time t, or None."""
    if times[0] > t + 1e-10 or times[-1] < t - 1e-10:
        return None
    i = int(np.searchsorted(times + 1e-10, t, side="right"))  # first sample with t < time + eps
    return polys[min(i, len(times) - 1)]


def replay(scenes, starts=None, solver=None, tf: float = 8.0, dt: float = 0.1, M_max: int = 32, v0: float = 10.0) -> dict:
    """Plans every scene from its start pose (default: PlanningNode's initial state (0, 0, 0), v = 10,
    planning_node.cc:24-30).  Scenes are grouped by centre line (the DP planner's reference is shared by a batch).
    -> dict(ok [B] bool, dp_ok, corridor_ok, status [B,8], result [B,K,13], coarse [B,K,6])."""
    import cilqr_b200
    from .scenarios import DpBatch
    from .solver import dp_num_knots
    scenes = list(scenes)
    B = len(scenes)
    starts = np.zeros((B, 3)) if starts is None else np.asarray(starts, dtype=np.float64).reshape(B, 3)
    own = solver is None
    solver = solver or cilqr_b200.Solver(device=0)
    K = dp_num_knots()
    N = K - 1
    out = {"ok": np.zeros(B, bool), "dp_ok": np.zeros(B, bool), "corridor_ok": np.zeros(B, bool),
           "status": np.full((B, 8), np.nan), "result": np.full((B, K, 13), np.nan), "coarse": np.full((B, K, 6), np.nan)}
    groups = {}
    for i, sc in enumerate(scenes):
        groups.setdefault(sc.center.tobytes(), []).append(i)
    for ids in groups.values():
        center = scenes[ids[0]].center
        left, right, barrier = road_barriers(center)
        n = len(ids)
        ns = max(len(scenes[i].static) for i in ids)
        nd = max(len(scenes[i].dynamic) for i in ids)
        V = max([4] + [len(p) for i in ids for p in scenes[i].static] + [pl.shape[1] for i in ids for _, pl in scenes[i].dynamic])
        T = max([1] + [len(t) for i in ids for t, _ in scenes[i].dynamic])
        sp, snv = np.zeros((n, ns, V, 2)), np.zeros((n, ns), np.int32)
        dtm, dsm = np.zeros((n, nd, T)), np.zeros((n, nd), np.int32)
        dpl, dnv = np.zeros((n, nd, T, V, 2)), np.zeros((n, nd), np.int32)
        for j, i in enumerate(ids):
            for o, p in enumerate(scenes[i].static):
                sp[j, o, :len(p)], snv[j, o] = p, len(p)
            for o, (t, pl) in enumerate(scenes[i].dynamic):
                dtm[j, o, :len(t)], dsm[j, o] = t, len(t)
                dpl[j, o, :len(t), :pl.shape[1]], dnv[j, o] = pl, pl.shape[1]
        db = DpBatch(center, np.ascontiguousarray(starts[ids]), sp, snv, dtm, dsm, dpl, dnv)
        dp = solver.dp_plan_batch(db, barrier)
        dp_ok = dp["ok"].astype(bool)
        # per-knot obstacle points: static corners first, then the dynamic obstacles' at the knot's time
        # (Corridor::BuildCorridorConstraints, corridor.cc:56-87 with Environment::Query*ObstaclesPoints :163-194)
        pts_list = []
        for i in ids:
            per_knot = []
            for k in range(K):
                p = [q for q in scenes[i].static]
                for t, pl in scenes[i].dynamic:
                    q = _dynamic_points_at(t, pl, k * dt)
                    if q is not None:
                        p.append(q)
                per_knot.append(np.concatenate(p) if p else np.zeros((0, 2)))
            pts_list.append(per_knot)
        P = max([1] + [len(q) for pk in pts_list for q in pk])
        pts, cnt = np.full((n, K, P, 2), np.nan), np.zeros((n, K), np.int32)
        for j, pk in enumerate(pts_list):
            for k, q in enumerate(pk):
                pts[j, k, :len(q)], cnt[j, k] = q, len(q)
        xyt = np.nan_to_num(dp["xytheta"])
        cor = solver.corridor_batch(xyt, pts, cnt, M_max=M_max, cfg=cilqr_b200.solver.default_corridor_config(point_cap=min(250, max(64, P + 16))))
        cor_ok = (cor["code"] == 0).all(axis=1)
        S_cap = 255  # the solver caches nearest-segment indices as bytes
        ll, nl = solver.lane_constraints(left[None], True, S_cap)
        lr, nr = solver.lane_constraints(right[None], False, S_cap)
        if nl[0] < 1 or nr[0] < 1:
            raise ValueError("lane boundaries: fewer than two sampled points, or more than 255 segments per side")
        good = np.where(dp_ok & cor_ok)[0]
        for j, i in enumerate(ids):
            out["dp_ok"][i], out["corridor_ok"][i] = dp_ok[j], cor_ok[j]
            out["coarse"][i] = dp["coarse"][j]
        if len(good) == 0:
            continue
        g = len(good)
        from .scenarios import ScenarioBatch
        start4 = np.concatenate([starts[ids][good], np.full((g, 1), v0)], axis=1)  # trajectory_planner.cpp:73-75
        sb = ScenarioBatch(N, M_max, 0, np.ascontiguousarray(start4), np.ascontiguousarray(dp["coarse"][good]),
                           np.ascontiguousarray(cor["corridor"][good]), np.ascontiguousarray(cor["corridor_cnt"][good]),
                           np.ascontiguousarray(np.broadcast_to(ll[:, :nl[0]], (g, nl[0], 7))),
                           np.ascontiguousarray(np.broadcast_to(lr[:, :nr[0]], (g, nr[0], 7))))
        big = cilqr_b200.Solver(device=solver.device, N_max=N, M_max=M_max, S_max=int(max(nl[0], nr[0])), B_max=g)
        res = big.plan_batch(sb, result=True)
        big.close()
        for j, gi in enumerate(good):
            i = ids[gi]
            out["ok"][i] = True
            out["status"][i], out["result"][i] = res["status"][j], res["result"][j]
    if own:
        solver.close()
    return out


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("pickles", nargs="+")
    ap.add_argument("--out", default=None)
    a = ap.parse_args(argv)
    scenes = [load_pickle(p) for p in a.pickles]
    r = replay(scenes)
    for p, ok, st in zip(a.pickles, r["ok"], r["status"]):
        print(f"{p}: {'planned' if ok else 'failed'}" + (f", exit {int(st[0])} after {int(st[1])} iterations, cost {st[2]:.3f}" if ok else ""))
    if a.out:
        np.savez_compressed(a.out, **r)


if __name__ == "__main__":
    main(sys.argv[1:])
