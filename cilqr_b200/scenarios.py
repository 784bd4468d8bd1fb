"""Synthetic `random_pedestrian`-style scenario generator.

Produces the *inputs* of ``IlqrOptimizer::Plan`` (reference ``algorithm/ilqr/ilqr_optimizer.h:41-48``)
in the scenario-major double wire format of ``include/cilqr_b200.h``:

* ``start[B,4]``            x, y, theta, v              (``trajectory_planner.cpp:73-75``)
* ``coarse[B,K,6]``         x, y, theta, v, a, delta    (what ``TransformGoals`` reads, ``ilqr_optimizer.cc:141-152``)
* ``corridor[B,K,M,3]``     raw half-planes a*x+b*y<c   (``corridor.h:20``), ``corridor_cnt[B,K]``
* ``lane_left/right[B,S,7]`` a,b,c,x0,y0,x1,y1          (``corridor.cc:265-305,322-331``)

The reference's own scenario source (``script/reference_publisher.py``) is unseeded, needs ROS, and
feeds a DP planner + OpenCV corridor builder that are out of scope; this module
re-creates the same road (``reference_publisher.py:200-209``, last straight lengthened), the same
obstacle mix and size/speed distributions (``:116-194``), a DP-lattice-like piece-wise linear
coarse trajectory with the ``ComputePathProfile`` finite differences
(``algorithm/utils/discrete_points_math.cc:27-176``), and per-knot half-plane sets of the shape
``Corridor::Plan`` emits (4 faces of the +-10 m box of ``AddCorridorPoints`` ``corridor.cc:89-120``
plus one separating plane per nearby obstacle, un-normalised like polygon edges).

Determinism: scenarios are produced in chunks of ``CHUNK`` ids; chunk ``c`` of seed ``s`` uses
``numpy.random.Philox(key=[s, c])``, so any rank can build any id range independently.
"""
from __future__ import annotations

import math
import os
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np

CHUNK = 1024
WHEEL_BASE = 1.0  # vehicle_param.h:31
_LEFT_BOUND = 2.5  # reference_publisher.py:25
_RIGHT_BOUND = 6.0  # reference_publisher.py:26
_LANE_SEG = 5.0  # planner_config.h:85
_DISC_RADIUS = math.hypot(1.942 / 2.0, (0.96 + 1.0 + 0.929) / 2.0 / 5)  # ilqr_optimizer.cc:97-104
_SAFE_MARGIN = 0.2  # planner_config.h:59
_BOX_HALF = 10.0  # planner_config.h:81-82
_MAX_DIFF = 25.0  # planner_config.h:77-78
_N_LAYERS = 5  # dp_planner.h:27 (NT)
_A_LAT = 8.0  # m/s^2, curve speed cap of the synthetic coarse trajectory
_A_DEC = 2.5  # m/s^2, braking used to anticipate the cap
_A_ACC = 2.0  # m/s^2
_D_CLEAR = 0.0  # m, minimum clearance of the coarse point beyond the shrunk plane
_DV_LAYER = 1.0  # m/s, per-layer deviation from the scenario's nominal speed

# "shipped": reference_publisher.py:200-209 with the last straight lengthened.
# "gentle" (default for the synthetic batches): same layout with the arc radii scaled by 2.5.
# On the shipped radii (5-12 m) the reference's own LQR initial guess (iqr, ilqr_optimizer.cc:793-842)
# leaves the +-10 m corridor box in 20-40 % of random placements and the solve then stops on the
# relative-cost test after 0-2 iterations with a cost of 1e6-1e8 -- valid reference behaviour but a
# degenerate benchmark; with 12.5-30 m radii that happens in ~1 % of scenarios.
_ROAD_CONFIGS = {
    "shipped": [30.0, (-90.0, 10.0), 10.0, (180.0, 5.0), 36.0, (-180.0, 12.0), 250.0],
    "gentle": [30.0, (-90.0, 25.0), 10.0, (180.0, 12.5), 36.0, (-180.0, 30.0), 250.0],
}
_ROAD_BACK = 40.0  # straight extension behind s=0 so lane windows can start behind the ego


@dataclass
class ScenarioBatch:
    """One batch in wire format.  All arrays are C-contiguous."""

    N: int
    M_max: int
    S: int
    start: np.ndarray
    coarse: np.ndarray
    corridor: np.ndarray
    corridor_cnt: np.ndarray
    lane_left: np.ndarray
    lane_right: np.ndarray

    @property
    def B(self) -> int:
        return self.start.shape[0]

    @property
    def K(self) -> int:
        return self.N + 1

    def slice(self, lo: int, hi: int) -> "ScenarioBatch":
        return ScenarioBatch(
            self.N, self.M_max, self.S,
            np.ascontiguousarray(self.start[lo:hi]), np.ascontiguousarray(self.coarse[lo:hi]),
            np.ascontiguousarray(self.corridor[lo:hi]), np.ascontiguousarray(self.corridor_cnt[lo:hi]),
            np.ascontiguousarray(self.lane_left[lo:hi]), np.ascontiguousarray(self.lane_right[lo:hi]))

    def input_bytes(self) -> int:
        return sum(a.nbytes for a in (self.start, self.coarse, self.corridor, self.corridor_cnt,
                                      self.lane_left, self.lane_right))

    def save(self, path: str) -> None:
        np.savez_compressed(path, N=self.N, M_max=self.M_max, S=self.S, start=self.start,
                            coarse=self.coarse, corridor=self.corridor,
                            corridor_cnt=self.corridor_cnt, lane_left=self.lane_left,
                            lane_right=self.lane_right)

    @staticmethod
    def load(path: str) -> "ScenarioBatch":
        z = np.load(path)
        return ScenarioBatch(int(z["N"]), int(z["M_max"]), int(z["S"]), z["start"], z["coarse"],
                             z["corridor"], z["corridor_cnt"], z["lane_left"], z["lane_right"])


def algorithmic_bytes(N: int, M: int, S: int) -> int:
    """Compulsory HBM bytes per trajectory for double I/O (DESIGN.md, section "Roofline").

    in : start 4 + coarse 6K + corridor 3MK doubles + cnt K int32 + lanes 2*7S doubles
    out: states 6K + controls 2N + status 8 doubles
    This is SURVEY 8(d)'s fp32 formula with 8-byte elements (count field stays int32).
    """
    K = N + 1
    return 8 * (4 + 6 * K + 3 * M * K + 14 * S) + 4 * K + 8 * (6 * K + 2 * N + 8)


# ----------------------------------------------------------------------------------------------
class _Road:
    """Centre line as an analytic function of station (straights and arcs)."""

    def __init__(self, name: str = "gentle"):
        x, y, yaw, s = -_ROAD_BACK, 0.0, 0.0, -_ROAD_BACK
        pieces = []  # (s0, s1, x0, y0, yaw0, kappa)
        for seg in [_ROAD_BACK] + _ROAD_CONFIGS[name]:
            if isinstance(seg, tuple):
                deg, radius = seg
                ang = math.radians(deg)
                length = abs(ang) * radius
                kappa = math.copysign(1.0 / radius, ang)
            else:
                length, kappa = float(seg), 0.0
            pieces.append((s, s + length, x, y, yaw, kappa))
            if kappa == 0.0:
                x += length * math.cos(yaw)
                y += length * math.sin(yaw)
            else:
                x += (math.sin(yaw + kappa * length) - math.sin(yaw)) / kappa
                y += (-math.cos(yaw + kappa * length) + math.cos(yaw)) / kappa
                yaw += kappa * length
            s += length
        self.pieces = np.array(pieces)
        self.s_min = -_ROAD_BACK
        self.s_max = s
        self._build_lanes()

    def eval(self, s):
        """-> x, y, theta, kappa (arrays shaped like s)."""
        s = np.asarray(s, dtype=np.float64)
        idx = np.clip(np.searchsorted(self.pieces[:, 1], s, side="right"), 0, len(self.pieces) - 1)
        p = self.pieces[idx]
        ds = s - p[..., 0]
        x0, y0, yaw0, kap = p[..., 2], p[..., 3], p[..., 4], p[..., 5]
        straight = kap == 0.0
        ksafe = np.where(straight, 1.0, kap)
        yaw = yaw0 + kap * ds
        xa = x0 + (np.sin(yaw) - np.sin(yaw0)) / ksafe
        ya = y0 + (-np.cos(yaw) + np.cos(yaw0)) / ksafe
        xs = x0 + ds * np.cos(yaw0)
        ys = y0 + ds * np.sin(yaw0)
        return np.where(straight, xs, xa), np.where(straight, ys, ya), yaw, kap

    def frenet_to_xy(self, s, l):
        x, y, th, _ = self.eval(s)
        return x - l * np.sin(th), y + l * np.cos(th)

    def _build_lanes(self):
        # boundary polylines at 0.1 m centre-line resolution, then LaneBoundarySample
        # (corridor.cc:307-320): keep a point once it is >= 5 m (minus 1e-10) from the last kept one
        s = np.arange(self.s_min, self.s_max + 1e-9, 0.1)
        self.lane_pts, self.lane_station = [], []
        for lat in (_LEFT_BOUND, -_RIGHT_BOUND):
            bx, by = self.frenet_to_xy(s, lat)
            keep = [0]
            lx, ly = bx[0], by[0]
            for i in range(1, len(s)):
                if math.hypot(bx[i] - lx, by[i] - ly) >= _LANE_SEG - 1e-10:
                    keep.append(i)
                    lx, ly = bx[i], by[i]
            keep = np.array(keep)
            self.lane_pts.append(np.stack([bx[keep], by[keep]], axis=1))
            self.lane_station.append(s[keep])


_ROADS: dict = {}


def road(name: str = "gentle") -> _Road:
    if name not in _ROADS:
        _ROADS[name] = _Road(name)
    return _ROADS[name]


def _lane_constraints(pts, left: bool):
    """pts [n,S+1,2] -> [n,S,7] (a,b,c,x0,y0,x1,y1).  Left segments run pt[i]->pt[i-1], right
    pt[i-1]->pt[i] (corridor.cc:279,300); HalfPlaneConstraint (corridor.cc:322-331)."""
    p0, p1 = pts[:, :-1], pts[:, 1:]
    st, en = (p1, p0) if left else (p0, p1)
    nx = en[..., 0] - st[..., 0]
    ny = en[..., 1] - st[..., 1]
    a = ny
    b = -nx
    c = a * st[..., 0] + b * st[..., 1]
    return np.stack([a, b, c, st[..., 0], st[..., 1], en[..., 0], en[..., 1]], axis=-1)


def _path_profile(x, y, dt):
    """ComputePathProfile (discrete_points_math.cc:27-176), vectorised over the leading axis.
    -> speeds, accelerations, kappas, each [n,K]."""
    seg = np.sqrt((x[:, 1:] - x[:, :-1]) ** 2 + (y[:, 1:] - y[:, :-1]) ** 2)
    acc_s = np.concatenate([np.zeros_like(x[:, :1]), np.cumsum(seg, axis=1)], axis=1)
    speeds = (acc_s[:, 1:] - acc_s[:, :-1]) / dt
    speeds = np.concatenate([speeds, speeds[:, -1:]], axis=1)
    accel = (speeds[:, 1:] - speeds[:, :-1]) / dt
    accel = np.concatenate([accel, accel[:, -1:]], axis=1)

    def d_ds(f):
        out = np.empty_like(f)
        out[:, 0] = (f[:, 1] - f[:, 0]) / (acc_s[:, 1] - acc_s[:, 0])
        out[:, -1] = (f[:, -1] - f[:, -2]) / (acc_s[:, -1] - acc_s[:, -2])
        out[:, 1:-1] = (f[:, 2:] - f[:, :-2]) / (acc_s[:, 2:] - acc_s[:, :-2])
        return out

    xds, yds = d_ds(x), d_ds(y)
    xdds, ydds = d_ds(xds), d_ds(yds)
    n2 = xds * xds + yds * yds
    kappa = (xds * ydds - yds * xdds) / (np.sqrt(n2) * n2 + 1e-6)
    return speeds, accel, kappa


def _gen_chunk(seed: int, chunk: int, n: int, N: int, n_obs: int, M_max: int, S: int, dt: float,
               road_name: str = "gentle", want_points: bool = False):
    rd = road(road_name)
    rng = np.random.Generator(np.random.Philox(key=[seed, chunk]))
    K = N + 1
    T = N * dt
    t = np.arange(K) * dt

    # ---- coarse trajectory: 5 time layers, piece-wise constant speed, piece-wise linear lateral
    vscale = min(1.0, 150.0 / (12.0 * T))
    v_layer = (rng.uniform(5.0, 10.0, size=(n, 1)) + rng.uniform(-_DV_LAYER, _DV_LAYER, size=(n, _N_LAYERS))) * vscale
    travel_max = 12.0 * vscale * T
    s0 = rng.uniform(0.0, rd.s_max - travel_max - 30.0, size=n)
    l_knots = np.empty((n, _N_LAYERS + 1))
    l_knots[:, 0] = rng.uniform(-1.0, 0.5, size=n)
    for i in range(_N_LAYERS):
        l_knots[:, i + 1] = np.clip(l_knots[:, i] + rng.uniform(-1.8, 1.8, size=n), -4.2, 0.9)
    t_layer = T / _N_LAYERS
    lay = np.minimum((t / t_layer).astype(np.int64), _N_LAYERS - 1)  # [K]
    frac = t / t_layer - lay  # [K]
    # station: the layer speed is a target; it is capped in curves to a lateral acceleration of
    # _A_LAT (anticipating the cap with a _A_DEC braking look-ahead) and rate-limited, so that the
    # coarse trajectory stays trackable by the reference's weak LQR initial guess (iqr)
    look = np.arange(0.0, 33.0, 3.0)
    s_path = np.empty((n, K))
    s_path[:, 0] = s0
    v_now = None
    for k in range(N):
        kap = np.abs(rd.eval(s_path[:, k, None] + look[None, :])[3])  # [n, len(look)]
        vcap2 = _A_LAT / np.maximum(kap, 1e-6) + 2.0 * _A_DEC * look[None, :]
        target = np.minimum(v_layer[:, lay[k]], np.sqrt(vcap2.min(axis=1)))
        v_now = target if v_now is None else np.clip(target, v_now - _A_DEC * dt, v_now + _A_ACC * dt)
        s_path[:, k + 1] = s_path[:, k] + dt * v_now
    s_rel = s_path - s0[:, None]
    l_path = l_knots[:, lay] + (l_knots[:, lay + 1] - l_knots[:, lay]) * frac
    cx, cy, cth, ckap = rd.eval(s_path)
    px = cx - l_path * np.sin(cth)
    py = cy + l_path * np.cos(cth)
    # heading as in dp_planner.cpp:243: road heading + atan((dl/ds)/(1-kappa*l)), backward differences
    dl = np.diff(l_path, axis=1, prepend=l_path[:, :1])
    ds = np.maximum(np.diff(s_path, axis=1, prepend=s_path[:, :1]), 1e-3)
    ptheta = cth + np.arctan((dl / ds) / (1.0 - ckap * l_path))
    speeds, accel, kappa = _path_profile(px, py, dt)
    pdelta = np.arctan(kappa * WHEEL_BASE)  # dp_planner.cpp:270
    coarse = np.stack([px, py, ptheta, speeds, accel, pdelta], axis=-1)

    start = np.stack([px[:, 0] + rng.normal(0.0, 0.1, size=n), py[:, 0] + rng.normal(0.0, 0.1, size=n),
                      ptheta[:, 0] + rng.normal(0.0, 0.02, size=n),
                      np.clip(speeds[:, 0] + rng.uniform(-1.0, 1.0, size=n), 0.5, 19.0)], axis=-1)

    # ---- obstacles (reference_publisher.py:116-194): 55 % pedestrians, 27 % moving, 18 % static
    n_ped = int(round(n_obs * 6 / 11.0))
    n_mov = int(round(n_obs * 3 / 11.0))
    n_sta = n_obs - n_ped - n_mov
    travel = s_rel[:, -1]  # [n]
    half = np.empty((n_obs, 2))
    half[:n_ped] = (0.5, 0.5)
    half[n_ped:] = (2.0, 1.0)
    # station / lateral of every obstacle at every knot time: [n, n_obs, K]
    o_s = np.empty((n, n_obs, K))
    o_l = np.empty((n, n_obs, K))
    # pedestrians cross from one road edge to the other at 0.4-1.4 m/s, starting at s/20 s
    ped_s = s0[:, None] + rng.uniform(5.0, 10.0 + travel[:, None], size=(n, n_ped))
    ped_v = 0.4 + rng.random(size=(n, n_ped))
    ped_dir = np.where(rng.random(size=(n, n_ped)) > 0.5, 1.0, -1.0)
    road_ub, road_lb = _LEFT_BOUND + 1.0, -_RIGHT_BOUND - 1.0
    ped_t0 = (ped_s - s0[:, None]) / 20.0
    prog = np.clip((t[None, None, :] - ped_t0[..., None]) * ped_v[..., None], 0.0, road_ub - road_lb)
    o_s[:, :n_ped] = ped_s[..., None]
    o_l[:, :n_ped] = np.where(ped_dir[..., None] > 0, road_ub - prog, road_lb + prog)
    # moving vehicles at 4-6 m/s on lateral {0,-4}
    mov_s = s0[:, None] + rng.uniform(-10.0, travel[:, None], size=(n, n_mov))
    mov_v = 4.0 + 2.0 * rng.random(size=(n, n_mov))
    mov_l = np.where(rng.random(size=(n, n_mov)) > 0.5, 0.0, -4.0)
    o_s[:, n_ped:n_ped + n_mov] = mov_s[..., None] + mov_v[..., None] * t[None, None, :]
    o_l[:, n_ped:n_ped + n_mov] = mov_l[..., None]
    # static vehicles on lateral {1,0,-4}
    sta_s = s0[:, None] + rng.uniform(10.0, 20.0 + travel[:, None], size=(n, n_sta))
    sta_l = np.array([1.0, 0.0, -4.0])[rng.integers(0, 3, size=(n, n_sta))]
    o_s[:, n_ped + n_mov:] = sta_s[..., None]
    o_l[:, n_ped + n_mov:] = sta_l[..., None]
    o_s = np.minimum(o_s, rd.s_max - 1.0)

    ox, oy, oth, _ = rd.eval(o_s)
    ocx = ox - o_l * np.sin(oth)
    ocy = oy + o_l * np.cos(oth)
    oth[:, :n_ped] = 0.0  # pedestrians keep theta = 0 (reference_publisher.py:188)
    cos_o, sin_o = np.cos(oth), np.sin(oth)
    # nearest corner of every obstacle to the coarse point of the same knot
    best_d2 = np.full((n, n_obs, K), np.inf)
    best_dx = np.zeros((n, n_obs, K))
    best_dy = np.zeros((n, n_obs, K))
    hx = half[None, :, 0, None]
    hy = half[None, :, 1, None]
    corners = []
    for sx, sy in ((-1, -1), (-1, 1), (1, 1), (1, -1)):  # transform_footprint order
        kx = ocx + sx * hx * cos_o - sy * hy * sin_o - px[:, None, :]
        ky = ocy + sx * hx * sin_o + sy * hy * cos_o - py[:, None, :]
        if want_points:
            corners.append(np.stack([kx + px[:, None, :], ky + py[:, None, :]], axis=-1))  # [n,n_obs,K,2]
        d2 = kx * kx + ky * ky
        better = d2 < best_d2
        best_d2 = np.where(better, d2, best_d2)
        best_dx = np.where(better, kx, best_dx)
        best_dy = np.where(better, ky, best_dy)
    dist = np.sqrt(best_d2)
    in_range = (np.abs(best_dx) <= _MAX_DIFF) & (np.abs(best_dy) <= _MAX_DIFF) & (dist > 1e-9)
    n_plane_obs = min(M_max - 4, n_obs)
    order = np.argsort(np.where(in_range, dist, np.inf), axis=1, kind="stable")[:, :n_plane_obs]  # [n,P,K]
    sel_ok = np.take_along_axis(in_range, order, axis=1)
    sel_d = np.take_along_axis(dist, order, axis=1)
    sel_dx = np.take_along_axis(best_dx, order, axis=1)
    sel_dy = np.take_along_axis(best_dy, order, axis=1)
    d_safe = _DISC_RADIUS + _SAFE_MARGIN + _D_CLEAR
    nxn = sel_dx / np.where(sel_ok, sel_d, 1.0)
    nyn = sel_dy / np.where(sel_ok, sel_d, 1.0)
    c_pl = nxn * px[:, None, :] + nyn * py[:, None, :] + np.maximum(sel_d, d_safe)
    edge = rng.uniform(1.0, 12.0, size=(n, n_plane_obs, K))  # polygon-edge magnitude (corridor.cc:249-260)
    cnt_obs = sel_ok.sum(axis=1)  # [n,K]; selected planes are the first cnt_obs of `order`

    corridor = np.zeros((n, K, M_max, 3))
    cth_, sth_ = np.cos(ptheta), np.sin(ptheta)
    box_n = ((cth_, sth_), (-sth_, cth_), (-cth_, -sth_), (sth_, -cth_))  # counter-clockwise faces
    for f, (bx, by) in enumerate(box_n):
        corridor[:, :, f, 0] = 2 * _BOX_HALF * bx
        corridor[:, :, f, 1] = 2 * _BOX_HALF * by
        corridor[:, :, f, 2] = 2 * _BOX_HALF * (bx * px + by * py + _BOX_HALF)
    obs_planes = np.stack([nxn * edge, nyn * edge, c_pl * edge], axis=-1)  # [n,P,K,3]
    obs_planes = np.where(sel_ok[..., None], obs_planes, 0.0)
    corridor[:, :, 4:4 + n_plane_obs, :] = np.transpose(obs_planes, (0, 2, 1, 3))
    corridor_cnt = (4 + cnt_obs).astype(np.int32)

    # ---- lanes: window of S segments per side starting ~15 m behind the ego
    lanes = []
    for side in range(2):
        st = rd.lane_station[side]
        first = np.searchsorted(st, s0 - 15.0, side="right") - 1
        first = np.clip(first, 0, len(st) - (S + 1))
        idx = first[:, None] + np.arange(S + 1)[None, :]
        lanes.append(_lane_constraints(rd.lane_pts[side][idx], left=(side == 0)))
    # ---- ego-centred frame: the reference always plans with the ego near the map origin
    # (planning_node.cc:24-30); the 3-norm normalisation of (a,b,c) (ilqr_optimizer.cc:475-495)
    # makes the barrier depend on that, so translate everything by the first coarse point.
    ox, oy = px[:, 0].copy(), py[:, 0].copy()
    coarse[:, :, 0] -= ox[:, None]
    coarse[:, :, 1] -= oy[:, None]
    start[:, 0] -= ox
    start[:, 1] -= oy
    corridor[..., 2] -= corridor[..., 0] * ox[:, None, None] + corridor[..., 1] * oy[:, None, None]
    for ln in lanes:
        ln[..., 2] -= ln[..., 0] * ox[:, None] + ln[..., 1] * oy[:, None]
        ln[..., 3] -= ox[:, None]
        ln[..., 5] -= ox[:, None]
        ln[..., 4] -= oy[:, None]
        ln[..., 6] -= oy[:, None]
    if want_points:
        # obstacle points per knot as Environment::QueryStatic/DynamicObstaclesPoints hand them to
        # Corridor::BuildCorridor (environment.cpp:163-194): static obstacles first, then the dynamic ones at
        # the knot's time, four corners each -> [n, K, 4*n_obs, 2], ego-centred like everything else
        pts = np.stack(corners, axis=2)  # [n, n_obs, 4, K, 2]
        perm = list(range(n_ped + n_mov, n_obs)) + list(range(n_ped + n_mov))
        pts = np.transpose(pts[:, perm], (0, 3, 1, 2, 4)).reshape(n, K, 4 * n_obs, 2).copy()
        pts[..., 0] -= ox[:, None, None]
        pts[..., 1] -= oy[:, None, None]
        return start, coarse, corridor, corridor_cnt, lanes[0], lanes[1], pts
    return start, coarse, corridor, corridor_cnt, lanes[0], lanes[1]


def generate(seed: int, first_id: int, count: int, N: int = 100, n_obs: int = 20, M_max: int = 20,
             S: int = 40, dt: float = 0.1, workers: int | None = None,
             road_name: str = "gentle") -> ScenarioBatch:
    """Scenarios ``first_id .. first_id+count-1`` of stream ``seed``.  ``first_id`` must be a
    multiple of CHUNK unless the whole request lies inside one chunk."""
    if M_max < 5:
        raise ValueError("M_max must be >= 5 (4 box faces + at least one obstacle plane)")
    c0, c1 = first_id // CHUNK, (first_id + count - 1) // CHUNK
    jobs = list(range(c0, c1 + 1))
    if workers is None:
        workers = max(1, min(len(jobs), (os.cpu_count() or 2) // 2, 16))

    def run(c):
        return _gen_chunk(seed, c, CHUNK, N, n_obs, M_max, S, dt, road_name)

    if workers > 1:
        with ThreadPoolExecutor(workers) as ex:
            parts = list(ex.map(run, jobs))
    else:
        parts = [run(c) for c in jobs]
    lo = first_id - c0 * CHUNK
    arrs = [np.concatenate([p[i] for p in parts], axis=0)[lo:lo + count] for i in range(6)]
    arrs = [np.ascontiguousarray(a) for a in arrs]
    return ScenarioBatch(N, M_max, S, *arrs)


@dataclass
class CorridorInputs:
    """Inputs of Corridor::Plan for a batch (include/cilqr_b200.h, CilqrCorridorIn)."""

    traj: np.ndarray        # [B,K,3] x, y, theta of the coarse trajectory
    obs_points: np.ndarray  # [B,K,P_max,2]
    obs_cnt: np.ndarray     # [B,K] int32

    @property
    def B(self) -> int:
        return self.traj.shape[0]

    @property
    def K(self) -> int:
        return self.traj.shape[1]

    @property
    def P_max(self) -> int:
        return self.obs_points.shape[2]


def generate_with_obstacles(seed: int, first_id: int, count: int, N: int = 100, n_obs: int = 20,
                            M_max: int = 20, S: int = 40, dt: float = 0.1, road_name: str = "gentle",
                            drop_far: bool = True):
    """Like generate(), and additionally the obstacle point clouds of the same scenarios: returns
    (ScenarioBatch, CorridorInputs).  With ``drop_far`` the points of obstacles that are further than
    60 m from the knot are dropped and the rest compacted (ragged obs_cnt), as a caller with a spatial
    index would do; BuildCorridor's own +-25 m filter makes the result independent of that."""
    c0, c1 = first_id // CHUNK, (first_id + count - 1) // CHUNK
    parts = [_gen_chunk(seed, c, CHUNK, N, n_obs, M_max, S, dt, road_name, want_points=True) for c in range(c0, c1 + 1)]
    lo = first_id - c0 * CHUNK
    arrs = [np.ascontiguousarray(np.concatenate([p[i] for p in parts], axis=0)[lo:lo + count]) for i in range(7)]
    batch = ScenarioBatch(N, M_max, S, *arrs[:6])
    pts = arrs[6]
    B, K, P, _ = pts.shape
    cnt = np.full((B, K), P, dtype=np.int32)
    if drop_far:
        d = pts - batch.coarse[:, :, None, :2]
        keep = (np.abs(d[..., 0]) <= 60.0) & (np.abs(d[..., 1]) <= 60.0)
        order = np.argsort(~keep, axis=2, kind="stable")  # kept points first, original order preserved
        pts = np.take_along_axis(pts, order[..., None], axis=2)
        cnt = keep.sum(axis=2).astype(np.int32)
        pts[np.arange(P)[None, None, :] >= cnt[..., None]] = np.nan  # slots >= cnt must never be read
    traj = np.ascontiguousarray(batch.coarse[:, :, :3])
    return batch, CorridorInputs(traj, np.ascontiguousarray(pts), cnt)


@dataclass
class DpBatch:
    """Inputs of DpPlanner::Plan for a batch (include/cilqr_b200.h, CilqrDpIn).  The centre line and the road
    barrier are shared by the batch (one road), obstacles and the start state are per scenario."""

    ref: np.ndarray          # [R,7] s, x, y, theta, kappa, left_bound, right_bound
    start: np.ndarray        # [B,3] x, y, theta
    static_poly: np.ndarray  # [B,n_static,4,2]
    static_nv: np.ndarray    # [B,n_static] int32
    dyn_time: np.ndarray     # [B,n_dyn,T]
    dyn_samples: np.ndarray  # [B,n_dyn] int32
    dyn_poly: np.ndarray     # [B,n_dyn,T,4,2]
    dyn_nv: np.ndarray       # [B,n_dyn] int32

    @property
    def B(self) -> int:
        return self.start.shape[0]


def reference_line(road_name: str = "gentle", step: float = 0.1) -> np.ndarray:
    """The centre line as the reference publishes it (script/reference_publisher.py:200-226: station, pose,
    curvature and the two bounds every 0.1 m) -> [R,7]."""
    rd = road(road_name)
    s = np.arange(0.0, rd.s_max + 1e-9, step)
    x, y, th, kap = rd.eval(s)
    return np.ascontiguousarray(np.stack([s, x, y, th, kap, np.full_like(s, _LEFT_BOUND), np.full_like(s, _RIGHT_BOUND)],
                                         axis=1))


def generate_dp(seed: int, count: int, n_obs: int = 11, tf: float = 8.0, dt: float = 0.1,
                road_name: str = "gentle") -> DpBatch:
    """random_pedestrian-style scenes for the DP planner: the obstacle mix, sizes and speeds of
    script/reference_publisher.py:116-194 (pedestrians crossing, moving and static vehicles), each dynamic obstacle
    as a polygon trajectory sampled every dt (DynamicObstacles.msg), the ego start near the centre line."""
    rd = road(road_name)
    rng = np.random.Generator(np.random.Philox(key=[seed, 977]))
    n = count
    T = int(round(tf / dt)) + 1
    t = np.arange(T) * dt
    n_ped = int(round(n_obs * 6 / 11.0))
    n_mov = int(round(n_obs * 3 / 11.0))
    n_sta = n_obs - n_ped - n_mov
    s0 = rng.uniform(5.0, rd.s_max - 130.0, size=n)
    l0 = rng.uniform(-1.0, 0.5, size=n)
    cx, cy, cth, _ = rd.eval(s0)
    start = np.stack([cx - l0 * np.sin(cth), cy + l0 * np.cos(cth), cth + rng.normal(0.0, 0.02, size=n)], axis=-1)

    def corners(s, l, th_fixed, hx, hy):
        ox, oy, oth, _ = rd.eval(s)
        px, py = ox - l * np.sin(oth), oy + l * np.cos(oth)
        th = oth if th_fixed is None else np.full_like(oth, th_fixed)
        c, sn = np.cos(th), np.sin(th)
        out = []
        for sx, sy in ((-1, -1), (-1, 1), (1, 1), (1, -1)):
            out.append(np.stack([px + sx * hx * c - sy * hy * sn, py + sx * hx * sn + sy * hy * c], axis=-1))
        return np.stack(out, axis=-2)

    ped_s = s0[:, None] + rng.uniform(8.0, 90.0, size=(n, n_ped))
    ped_v = 0.4 + rng.random(size=(n, n_ped))
    ped_dir = np.where(rng.random(size=(n, n_ped)) > 0.5, 1.0, -1.0)
    ub, lb = _LEFT_BOUND + 1.0, -_RIGHT_BOUND - 1.0
    prog = np.clip((t[None, None, :] - ((ped_s - s0[:, None]) / 20.0)[..., None]) * ped_v[..., None], 0.0, ub - lb)
    ped_l = np.where(ped_dir[..., None] > 0, ub - prog, lb + prog)
    ped = corners(np.broadcast_to(ped_s[..., None], ped_l.shape), ped_l, 0.0, 0.5, 0.5)
    mov_s = s0[:, None, None] + rng.uniform(-10.0, 80.0, size=(n, n_mov, 1)) + (4.0 + 2.0 * rng.random(size=(n, n_mov, 1))) * t
    mov_l = np.broadcast_to(np.where(rng.random(size=(n, n_mov, 1)) > 0.5, 0.0, -4.0), mov_s.shape)
    mov = corners(np.minimum(mov_s, rd.s_max - 1.0), mov_l, None, 2.0, 1.0)
    sta_s = s0[:, None] + rng.uniform(15.0, 95.0, size=(n, n_sta))
    sta_l = np.array([1.0, 0.0, -4.0])[rng.integers(0, 3, size=(n, n_sta))]
    sta = corners(sta_s, sta_l, None, 2.0, 1.0)
    dyn_poly = np.ascontiguousarray(np.concatenate([ped, mov], axis=1))
    nd = n_ped + n_mov
    return DpBatch(reference_line(road_name), np.ascontiguousarray(start), np.ascontiguousarray(sta),
                   np.full((n, n_sta), 4, np.int32), np.ascontiguousarray(np.broadcast_to(t, (n, nd, T))),
                   np.full((n, nd), T, np.int32), dyn_poly, np.full((n, nd), 4, np.int32))


def config_batch(index: int, count: int | None = None, first_id: int = 0) -> ScenarioBatch:
    """BASELINE.json configs; seed = 20260101 + index."""
    table = {
        0: dict(N=80, n_obs=11, B=1),       # shipped scenario, emulated
        1: dict(N=50, n_obs=20, B=1024),
        2: dict(N=100, n_obs=20, B=65536),
        3: dict(N=100, n_obs=20, B=1048576),
        4: dict(N=100, n_obs=20, B=32768),  # horizon sweep: pass N explicitly via generate()
    }[index]
    return generate(20260101 + index, first_id, count or table["B"], N=table["N"],
                    n_obs=table["n_obs"])
