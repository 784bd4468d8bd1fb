"""Independent pure-Python restatement of DpPlanner::Plan (algorithm/planner/dp_planner.cpp) written from the
reference source, class by class, to cross-check oracle/dp_oracle.c (tests/test_dp_oracle_cross.py).

TEST INFRASTRUCTURE ONLY.  Python floats are IEEE doubles and math.* calls glibc, so when both restatements
follow the reference's operation order the results agree bit for bit.  Slow (tens of seconds per scene).
"""
from __future__ import annotations

import bisect
import math

K_MATH_EPS = 1e-10  # math::kMathEpsilon
DP_EPS = 1e-3       # constexpr double kMathEpsilon = 1e-3 in dp_planner.cpp:29 (shadows the former there)
NT, NS, NL = 5, 7, 10


def normalize_angle(angle):  # math_utils.cpp:53-59
    a = math.fmod(angle + math.pi, 2.0 * math.pi)
    if a < 0.0:
        a += 2.0 * math.pi
    return a - math.pi


def slerp(a0, t0, a1, t1, t):  # math_utils.h:208-225
    if abs(t1 - t0) <= K_MATH_EPS:
        return normalize_angle(a0)
    a0_n = normalize_angle(a0)
    a1_n = normalize_angle(a1)
    d = a1_n - a0_n
    if d > math.pi:
        d = d - 2 * math.pi
    elif d < -math.pi:
        d = d + 2 * math.pi
    r = (t - t0) / (t1 - t0)
    return normalize_angle(a0_n + d * r)


class Reference:
    """DiscretizedTrajectory over the centre line (utils/discretized_trajectory.cpp)."""

    def __init__(self, rows):
        self.pts = [tuple(float(v) for v in r) for r in rows]  # s, x, y, theta, kappa, left_bound, right_bound
        self.s = [p[0] for p in self.pts]

    def _lower_bound_station(self, station):  # :34-46
        if station >= self.s[-1]:
            return len(self.pts) - 1
        if station < self.s[0]:
            return 0
        return bisect.bisect_left(self.s, station)

    @staticmethod
    def _interp(p0, p1, s):  # LinearInterpolateTrajectory :62-84
        s0, s1 = p0[0], p1[0]
        if abs(s1 - s0) < K_MATH_EPS:
            return p0
        w = (s - s0) / (s1 - s0)
        return (s, (1 - w) * p0[1] + w * p1[1], (1 - w) * p0[2] + w * p1[2], slerp(p0[3], p0[0], p1[3], p1[0], s),
                (1 - w) * p0[4] + w * p1[4], (1 - w) * p0[5] + w * p1[5], (1 - w) * p0[6] + w * p1[6])

    def evaluate_station(self, station):  # :110-121
        it = self._lower_bound_station(station)
        if it == 0:
            it = 1
        return self._interp(self.pts[it - 1], self.pts[it], station)

    def get_cartesian(self, station, lateral):  # :192-196
        r = self.evaluate_station(station)
        return r[1] - lateral * math.sin(r[3]), r[2] + lateral * math.cos(r[3])

    def get_projection(self, x, y):  # :156-190
        best, idx = float("inf"), 0
        for i, p in enumerate(self.pts):  # QueryNearestPoint :136-154
            dx, dy = p[1] - x, p[2] - y
            d = dx * dx + dy * dy
            if d < best:
                best, idx = d, i
        pp = self.pts[idx]
        i0, i1 = max(0, idx - 1), min(len(self.pts) - 1, idx + 1)
        if i0 < i1:
            a, b = self.pts[i0], self.pts[i1]
            v0x, v0y = x - a[1], y - a[2]
            v1x, v1y = b[1] - a[1], b[2] - a[2]
            v1_norm = math.sqrt(v1x * v1x + v1y * v1y)
            dot = v0x * v1x + v0y * v1y
            pp = self._interp(a, b, a[0] + dot / v1_norm)
        nr_x, nr_y = x - pp[1], y - pp[2]
        return pp[0], math.copysign(math.hypot(nr_x, nr_y), nr_y * math.cos(pp[3]) - nr_x * math.sin(pp[3]))


class Polygon:
    """Polygon2d: bounds + IsPointIn + HasOverlap(Box2d) (math/polygon2d.cpp:120-165,246-256)."""

    def __init__(self, pts):
        self.pts = [(float(x), float(y)) for x, y in pts]
        xs, ys = [p[0] for p in self.pts], [p[1] for p in self.pts]
        self.min_x, self.max_x, self.min_y, self.max_y = min(xs), max(xs), min(ys), max(ys)

    def is_point_in(self, x, y):
        if x < self.min_x or x > self.max_x or y < self.min_y or y > self.max_y:
            return False
        j, c = len(self.pts) - 1, 0
        for i, (xi, yi) in enumerate(self.pts):
            xj, yj = self.pts[j]
            if (yi > y) != (yj > y):
                side = (xi - x) * (yj - y) - (yi - y) * (xj - x)
                if (side > 0.0) if yi < yj else (side < 0.0):
                    c += 1
            j = i
        return bool(c & 1)

    def has_overlap(self, box):
        if box.max_x < self.min_x or box.min_x > self.max_x or box.max_y < self.min_y or box.min_y > self.max_y:
            return False
        if any(box.is_point_in(x, y) for x, y in self.pts):
            return True
        return any(self.is_point_in(x, y) for x, y in box.corners())


class Box:
    """Box2d(AABox2d) after Shift (math/box2d.cpp:93-105,123-129; math/aabox2d.cpp:63-71)."""

    def __init__(self, cx, cy, half):
        self.cx, self.cy, self.half = cx, cy, half
        self.min_x, self.max_x, self.min_y, self.max_y = cx - half, cx + half, cy - half, cy + half

    def is_point_in(self, x, y):
        x0, y0 = x - self.cx, y - self.cy
        dx = abs(x0 * 1.0 + y0 * 0.0)
        dy = abs(-x0 * 0.0 + y0 * 1.0)
        return dx <= self.half + K_MATH_EPS and dy <= self.half + K_MATH_EPS

    def corners(self):
        return [(self.cx + self.half, self.cy - self.half), (self.cx + self.half, self.cy + self.half),
                (self.cx - self.half, self.cy + self.half), (self.cx - self.half, self.cy - self.half)]


class Environment:
    """utils/environment.cpp: collision checks only."""

    def __init__(self, cfg, barrier, statics, dynamics):
        self.cfg = cfg
        self.barrier = [(float(x), float(y)) for x, y in barrier]  # sorted by x
        self.bx = [p[0] for p in self.barrier]
        self.statics = [Polygon(p) for p in statics]
        self.dynamics = [[(float(t), Polygon(p)) for t, p in ob] for ob in dynamics]
        length = cfg["wheel_base"] + cfg["rear_hang_length"] + cfg["front_hang_length"]  # vehicle_param.h:80-85
        self.radius = math.hypot(0.25 * length, 0.5 * cfg["width"])
        self.r2x = 0.25 * length - cfg["rear_hang_length"]
        self.f2x = 0.75 * length - cfg["rear_hang_length"]

    def check_static(self, box):  # :51-87
        for ob in self.statics:
            if ob.has_overlap(box):
                return True
        if not self.barrier:
            return False
        if box.max_x < self.bx[0] or box.min_x > self.bx[-1]:
            return False
        start = bisect.bisect_right(self.bx, box.min_x)
        end = bisect.bisect_right(self.bx, box.max_x)
        if start > 0:
            start -= 1
        return any(box.is_point_in(*self.barrier[i]) for i in range(start, end))

    def check_dynamic(self, time, box):  # :124-141
        for ob in self.dynamics:
            if not ob or ob[0][0] > time or ob[-1][0] < time:
                continue
            i = bisect.bisect_right([t for t, _ in ob], time)
            if i >= len(ob):
                i = len(ob) - 1  # the reference dereferences end() here
            if ob[i][1].has_overlap(box):
                return True
        return False

    def check_optimization_collision(self, time, x, y, theta):  # :99-122
        r = self.radius
        half = (r - (-r)) / 2.0
        c0 = (-r + r) / 2.0
        xf, xr = x + self.f2x * math.cos(theta), x + self.r2x * math.cos(theta)
        yf, yr = y + self.f2x * math.sin(theta), y + self.r2x * math.sin(theta)
        f_box, r_box = Box(c0 + xf, c0 + yf, half), Box(c0 + xr, c0 + yr, half)
        return (self.check_static(f_box) or self.check_static(r_box) or self.check_dynamic(time, f_box)
                or self.check_dynamic(time, r_box))


class DpPlanner:
    def __init__(self, cfg, ref: Reference, env: Environment):
        self.cfg, self.ref, self.env = cfg, ref, env
        self.unit_time = cfg["tf"] / NT
        lin = lambda n, a, b: [a + (b - a) / (n - 1) * i for i in range(n)]  # noqa: E731  math::LinSpaced
        self.time_ = lin(NT, self.unit_time, cfg["tf"])
        self.station_ = lin(NS, 0, self.unit_time * cfg["max_velocity"])
        self.lateral_ = lin(NL - 1, 0, 1)
        self.safe_margin = cfg["width"] / 2 * 1.5

    def lateral_offset(self, s, l_ind):  # dp_planner.h:83-92
        if l_ind == NL - 1:
            return 0.0
        r = self.ref.evaluate_station(s)
        lb = -r[6] + self.safe_margin
        ub = r[5] - self.safe_margin
        return lb + (ub - lb) * self.lateral_[l_ind]

    def interpolate_linearly(self, parent_s, parent_l_ind, cur_t, cur_s_ind, cur_l_ind):  # :283-320
        cfg = self.cfg
        nseg, t = 0, 0.0
        while t < cfg["tf"] + cfg["delta_t"] - K_MATH_EPS:
            if cur_t == 0:
                if 0.0 - DP_EPS < t < self.unit_time + DP_EPS:
                    nseg += 1
            elif self.time_[cur_t] - self.unit_time + K_MATH_EPS < t < self.time_[cur_t] + K_MATH_EPS:
                nseg += 1
            t += cfg["delta_t"]
        p_l, p_s = self.start_l, self.start_s
        if parent_l_ind >= 0:
            p_s = parent_s
            p_l = self.lateral_offset(p_s, parent_l_ind)
        cur_s = p_s + self.station_[cur_s_ind]
        cur_l = self.lateral_offset(cur_s, cur_l_ind)
        s_step = self.station_[cur_s_ind] / nseg
        l_step = (cur_l - p_l) / nseg
        return [(p_s + i * s_step, p_l + i * l_step) for i in range(nseg)]

    def collision_cost(self, parent, cur):  # :40-85
        pt, ps_i, pl_i = parent
        parent_s = grandparent_s = self.start_s
        last_l, last_s = self.start_l, self.start_s
        if pt >= 0:
            cell = self.space[pt][ps_i][pl_i]
            parent_s = cell[1]
            if pt > 0:
                grandparent_s = self.space[pt - 1][cell[2]][cell[3]][1]
            prev = self.interpolate_linearly(grandparent_s, cell[3], pt, ps_i, pl_i)
            last_s, last_l = prev[-1]
        path = self.interpolate_linearly(parent_s, pl_i, cur[0], cur[1], cur[2])
        nseg = len(path)
        for i, (s, l) in enumerate(path):
            dl = l - last_l
            ds = max(s - last_s, DP_EPS)
            last_l, last_s = l, s
            cx, cy = self.ref.get_cartesian(s, l)
            r = self.ref.evaluate_station(s)
            lb = min(0.0, -r[6] + self.safe_margin)
            ub = max(0.0, r[5] - self.safe_margin)
            if l < lb - DP_EPS or l > ub + DP_EPS:
                return self.cfg["dp_w_obstacle"]
            heading = r[3] + math.atan((dl / ds) / (1 - r[4] * l))
            parent_time = 0.0 if pt < 0 else self.time_[pt]
            time = parent_time + i * (self.unit_time / nseg)
            if self.env.check_optimization_collision(time, cx, cy, heading):
                return self.cfg["dp_w_obstacle"]
        return 0.0

    def get_cost(self, parent, cur):  # :87-133
        cfg = self.cfg
        pt, ps_i, pl_i = parent
        parent_s = grandparent_s = self.start_s
        parent_l = grandparent_l = self.start_l
        if pt >= 0:
            cell = self.space[pt][ps_i][pl_i]
            parent_s = cell[1]
            parent_l = self.lateral_offset(parent_s, pl_i)
            if pt >= 1:
                grandparent_s = self.space[pt - 1][cell[2]][cell[3]][1]
                grandparent_l = self.lateral_offset(grandparent_s, cell[3])
        cur_s = parent_s + self.station_[cur[1]]
        cur_l = self.lateral_offset(cur_s, cur[2])
        ds1, dl1 = cur_s - parent_s, cur_l - parent_l
        ds0, dl0 = parent_s - grandparent_s, parent_l - grandparent_l
        if self.collision_cost(parent, cur) >= cfg["dp_w_obstacle"]:
            return cur_s, cfg["dp_w_obstacle"]
        cost_lateral = abs(cur_l)
        cost_lateral_change = abs(parent_l - cur_l) / (self.station_[cur[1]] + DP_EPS)
        cost_lateral_change_t = abs(dl1 - dl0) / self.unit_time
        cost_v = abs(ds1 / self.unit_time - cfg["dp_nominal_velocity"])
        cost_dv = abs((ds1 - ds0) / self.unit_time)
        return cur_s, (cfg["dp_w_lateral"] * cost_lateral + cfg["dp_w_lateral_change"] * cost_lateral_change
                       + cfg["dp_w_lateral_velocity_change"] * cost_lateral_change_t
                       + cfg["dp_w_longitudinal_velocity_bias"] * cost_v
                       + cfg["dp_w_longitudinal_velocity_change"] * cost_dv)

    def plan(self, start_x, start_y, start_theta):  # :135-281 -> (ok, rows [K][11], min_cost, waypoints)
        cfg = self.cfg
        self.start_s, self.start_l = self.ref.get_projection(start_x, start_y)
        big = 1.7976931348623157e308
        self.space = [[[[big, 2.2250738585072014e-308, -1, -1] for _ in range(NL)] for _ in range(NS)] for _ in range(NT)]
        for i in range(NS):
            for j in range(NL):
                cur_s, c = self.get_cost((-1, -1, -1), (0, i, j))
                self.space[0][i][j][1] = cur_s
                self.space[0][i][j][0] = c
        for i in range(NT - 1):
            for j in range(NS):
                for k in range(NL):
                    for m in range(NS):
                        for n in range(NL):
                            cur_s, delta = self.get_cost((i, j, k), (i + 1, m, n))
                            cur_cost = self.space[i][j][k][0] + delta
                            if cur_cost < self.space[i + 1][m][n][0]:
                                self.space[i + 1][m][n] = [cur_cost, cur_s, j, k]
        min_cost, ms, ml = big, 0, 0
        for i in range(NS):
            for j in range(NL):
                if self.space[NT - 1][i][j][0] < min_cost:
                    ms, ml, min_cost = i, j, self.space[NT - 1][i][j][0]
        wps = [None] * NT
        for i in range(NT - 1, -1, -1):
            cell = list(self.space[i][ms][ml])
            wps[i] = ((i, ms, ml), cell)
            ms, ml = cell[2], cell[3]
        rows, xy = [], []
        last_l, last_s = self.start_l, self.start_s
        for i in range(NT):
            parent_s = wps[i - 1][1][1] if i > 0 else self.start_s
            seg = self.interpolate_linearly(parent_s, wps[i][1][3], i, wps[i][0][1], wps[i][0][2])
            for s, l in seg:
                dl = l - last_l
                ds = max(s - last_s, DP_EPS)
                last_l, last_s = l, s
                x, y = self.ref.get_cartesian(s, l)
                tp = self.ref.evaluate_station(s)
                n = len(rows)
                rows.append([cfg["delta_t"] * n, s, x, y, tp[3] + math.atan((dl / ds) / (1 - tp[4] * l)),
                             0.0, 0.0, 0.0, 0.0, 0.0, 0.0])
                xy.append((x, y))
        speeds, accel, kappas = path_profile(cfg["delta_t"], xy)
        for i, row in enumerate(rows):
            row[5] = kappas[i]
            row[9] = math.atan(kappas[i] * cfg["wheel_base"])
            row[6] = speeds[i]
            row[7] = accel[i]
        waypoints = [(w[0][1], w[0][2], w[1][1]) for w in wps]
        return min_cost < cfg["dp_w_obstacle"], rows, min_cost, waypoints


def path_profile(dt, xy):  # discrete_points_math.cc:27-176
    n = len(xy)
    acc = [0.0]
    distance, (fx, fy) = 0.0, xy[0]
    for i in range(1, n):
        nx, ny = xy[i]
        seg = math.sqrt((fx - nx) * (fx - nx) + (fy - ny) * (fy - ny))
        acc.append(seg + distance)
        distance += seg
        fx, fy = nx, ny
    speeds = [(acc[i] - acc[i - 1]) / dt for i in range(1, n)]
    speeds.append(speeds[-1])
    accel = [(speeds[i] - speeds[i - 1]) / dt for i in range(1, n)]
    accel.append(accel[-1])

    def div(a, b):  # IEEE semantics for a standing-still plan (0/0 -> NaN), like the C++
        try:
            return a / b
        except ZeroDivisionError:
            return float("nan") if a == 0 or a != a else math.copysign(float("inf"), a) * math.copysign(1.0, b)

    def d_ds(f):
        out = []
        for i in range(n):
            lo, hi = (0, 1) if i == 0 else ((n - 2, n - 1) if i == n - 1 else (i - 1, i + 1))
            out.append(div(f[hi] - f[lo], acc[hi] - acc[lo]))
        return out

    xds, yds = d_ds([p[0] for p in xy]), d_ds([p[1] for p in xy])
    xdds, ydds = d_ds(xds), d_ds(yds)
    kappas = []
    for i in range(n):
        n2 = xds[i] * xds[i] + yds[i] * yds[i]
        kappas.append(div(xds[i] * ydds[i] - yds[i] * xdds[i], math.sqrt(n2) * n2 + 1e-6) if n2 == n2 else float("nan"))
    return speeds, accel, kappas


DEFAULT_CFG = dict(tf=8.0, delta_t=0.1, dp_nominal_velocity=10.0, dp_w_obstacle=1000.0, dp_w_lateral=0.1,
                   dp_w_lateral_change=0.5, dp_w_lateral_velocity_change=1.0, dp_w_longitudinal_velocity_bias=10.0,
                   dp_w_longitudinal_velocity_change=1.0, max_velocity=20.0, width=1.942, wheel_base=1.0,
                   front_hang_length=0.96, rear_hang_length=0.929)
