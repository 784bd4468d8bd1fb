"""Second, independent CPU restatement of mpt0816/Cilqr's CILQR solve in NumPy float64.

TEST INFRASTRUCTURE ONLY (see oracle/cilqr_oracle.h).  The reference ships no tests or golden vectors;
this file exists so that two restatements written in different styles (the structured C in
cilqr_oracle.c, dense 6x6 linear algebra here) check each other; the C one is additionally pinned bit
for bit against the reference's own solver source (tests/test_reference_pins.py).  Pure-Python loops: use it for small horizons / a handful of scenarios only.

Written from the reference sources, dense like the Eigen code (no sparsity shortcuts):
  algorithm/ilqr/ilqr_optimizer.cc   Plan :53-95, Optimize :154-320, Backward :334-390,
                                     Forward :392-415, costs :417-603, derivatives :620-769, iqr :793-842
  algorithm/ilqr/vehicle_model.cc    Dynamics :88-121, DynamicsJacbian :21-86
  algorithm/ilqr/barrier_function.h  RelaxBarrierFunction :81-147
  algorithm/math/math_utils.cpp      NormalizeAngle :53-59
  algorithm/math/line_segment2d.cpp  ctor :40-49, DistanceTo :61-75
"""
from __future__ import annotations

import math

import numpy as np

ALPHAS = [1.0000, 0.5012, 0.2512, 0.1259, 0.0631, 0.0316, 0.0158, 0.0079, 0.0040, 0.0020, 0.0010]


class P:
    """vehicle_param.h:26-64, planner_config.h:45-73,94, barrier_function.h:143-146."""
    front_hang_length, wheel_base, rear_hang_length, width = 0.96, 1.0, 0.929, 1.942
    max_velocity, min_acceleration, max_acceleration = 20.0, -5.0, 5.0
    jerk_min, jerk_max = -10.0, 10.0
    delta_min, delta_max = -40.0 / 180 * math.pi, 40.0 / 180 * math.pi
    delta_rate_min, delta_rate_max = delta_min / 3.0, delta_max / 3.0
    safe_margin = 0.2
    w_jerk, w_delta_rate, w_x, w_y, w_theta, w_v, w_a, w_delta = 1.0, 1.0, 0.5, 0.5, 1e-3, 0.0, 0.0, 0.0
    abs_cost_tol, rel_cost_tol = 1e-2, 1e-2
    t, eps = 5.0, 0.01
    dt = 0.1
    num_of_disc = 5
    max_iter_num = 200


def normalize_angle(a):
    r = math.fmod(a + math.pi, 2.0 * math.pi)
    if r < 0.0:
        r += 2.0 * math.pi
    return r - math.pi


def f_cont(s, u):
    th, v, a, de = normalize_angle(s[2]), s[3], s[4], normalize_angle(s[5])
    return np.array([v * math.cos(th), v * math.sin(th), v * math.tan(de) / P.wheel_base, a, u[0], u[1]])


def dynamics(s, u):
    k1 = f_cont(s, u)
    k2 = f_cont(s + 0.5 * P.dt * k1, u)
    n = s + P.dt * k2
    n[2] = normalize_angle(n[2])
    n[5] = normalize_angle(n[5])
    return n


def dynamics_jacobian(s, u):
    L, dt = P.wheel_base, P.dt
    v, theta, delta, a, dr = s[3], normalize_angle(s[2]), normalize_angle(s[5]), s[4], u[1]
    theta_mid = theta + 0.5 * dt * v * math.tan(delta) / L
    td, tdr = math.tan(delta), math.tan(delta + 0.5 * dt * dr)
    c, sn = math.cos(theta_mid), math.sin(theta_mid)
    A = np.eye(6)
    B = np.zeros((6, 2))
    vm = 0.5 * a * dt + v
    A[0, 2] = -dt * vm * sn
    A[0, 3] = dt * c - 0.5 * dt * dt * vm * sn * td / L
    A[0, 4] = 0.5 * dt * dt * c
    A[0, 5] = -0.5 * dt * dt * v * vm * (td * td + 1) * sn / L
    A[1, 2] = dt * vm * c
    A[1, 3] = dt * sn + 0.5 * dt * dt * vm * c * td / L
    A[1, 4] = 0.5 * dt * dt * sn
    A[1, 5] = 0.5 * dt * dt * v * vm * (td * td + 1) * c / L
    A[2, 3] = dt * tdr / L
    A[2, 4] = 0.5 * dt * dt * tdr / L
    A[2, 5] = dt * (v * (tdr * tdr + 1)) / L
    A[3, 4] = dt
    B[2, 1] = 0.5 * dt * dt * v * (tdr * tdr + 1) / L
    B[3, 0] = 0.5 * dt * dt
    B[4, 0] = dt
    B[5, 1] = dt
    return A, B


def bar_value(x):
    rt = 1.0 / P.t
    if x < -P.eps:
        return -rt * math.log(-x)
    return 0.5 * rt * (((-x - 2.0 * P.eps) / P.eps) ** 2 - 1) - rt * math.log(P.eps)


def bar_jac(x, dx):
    rt = 1.0 / P.t
    if x < -P.eps:
        return -rt / x * dx
    return rt * (x + 2.0 * P.eps) / P.eps / P.eps * dx


def bar_hess(x, dx, ddx=None):
    rt = 1.0 / P.t
    n = dx.shape[0]
    if ddx is None:
        ddx = np.zeros((n, n))
    if x < -P.eps:
        return np.outer(rt / x / x * dx, dx) - rt / x * ddx
    return np.outer(rt * (x + 2.0 * P.eps) / P.eps / P.eps * dx, dx)


class Segment:
    def __init__(self, x0, y0, x1, y1):
        self.s = (x0, y0)
        self.e = (x1, y1)
        dx, dy = x1 - x0, y1 - y0
        self.length = math.hypot(dx, dy)
        self.u = (0.0, 0.0) if self.length <= 1e-10 else (dx / self.length, dy / self.length)

    def distance_to(self, px, py):
        if self.length <= 1e-10:
            return math.hypot(px - self.s[0], py - self.s[1])
        x0, y0 = px - self.s[0], py - self.s[1]
        proj = x0 * self.u[0] + y0 * self.u[1]
        if proj <= 0.0:
            return math.hypot(x0, y0)
        if proj >= self.length:
            return math.hypot(px - self.e[0], py - self.e[1])
        return abs(x0 * self.u[1] - y0 * self.u[0])


class Solver:
    """State of one IlqrOptimizer::Plan call."""

    def __init__(self, start, coarse, corridor, cnt, lane_left, lane_right):
        self.N = coarse.shape[0] - 1
        self.K = self.N + 1
        self.goals = np.array(coarse, dtype=np.float64)
        self.goals[0] = [start[0], start[1], start[2], start[3], 0.0, 0.0]
        length = P.front_hang_length + P.wheel_base + P.rear_hang_length
        r = math.hypot(P.width / 2.0, length / 2.0 / P.num_of_disc)
        self.disc_radius = r

        def shrink(e, d):
            e = np.array(e[:3], dtype=np.float64)
            e[2] = e[2] - d * (e[0] * e[0] + e[1] * e[1]) / math.hypot(e[0], e[1])
            return e / math.hypot(math.hypot(e[0], e[1]), e[2])

        self.corridor = [[shrink(corridor[k][m], r + P.safe_margin) for m in range(int(cnt[k]))]
                         for k in range(self.K)]
        self.lanes = []
        for lane in (lane_left, lane_right):
            self.lanes.append([(shrink(row, r), Segment(row[3], row[4], row[5], row[6])) for row in lane])
        Ld = length / P.num_of_disc
        self.off = [Ld * (j - 0.5) - P.rear_hang_length for j in range(P.num_of_disc)]

    # ---- costs
    def nearest(self, side, x, y):
        best, bi = float("inf"), -1
        for i, (_, seg) in enumerate(self.lanes[side]):
            d = seg.distance_to(x, y)
            if d < best:
                best, bi = d, i
        return self.lanes[side][bi][0]

    def total_cost(self, X, U):
        j = 0.0
        for i in range(self.K):
            d = X[i] - self.goals[i]
            j += P.w_x * d[0] ** 2 + P.w_y * d[1] ** 2 + P.w_theta * d[2] ** 2
        for i in range(self.N):
            j += P.w_jerk * U[i, 0] ** 2 + P.w_delta_rate * U[i, 1] ** 2
        xc = 0.0
        for i in range(self.K):
            s = X[i]
            for g in (-s[3], s[3] - P.max_velocity, s[4] - P.max_acceleration, P.min_acceleration - s[4],
                      s[5] - P.delta_max, P.delta_min - s[5]):
                xc += bar_value(g)
        uc = 0.0
        for i in range(self.N):
            u = U[i]
            for g in (u[0] - P.jerk_max, P.jerk_min - u[0], u[1] - P.delta_rate_max, P.delta_rate_min - u[1]):
                uc += bar_value(g)
        dyn = xc + uc
        cor = 0.0
        lan = 0.0
        for i in range(self.K):
            for off in self.off:
                x = X[i, 0] + off * math.cos(X[i, 2])
                y = X[i, 1] + off * math.sin(X[i, 2])
                for c in self.corridor[i]:
                    cor += bar_value(c[0] * x + c[1] * y - c[2])
        for i in range(self.K):
            for off in self.off:
                x = X[i, 0] + off * math.cos(X[i, 2])
                y = X[i, 1] + off * math.sin(X[i, 2])
                for side in (0, 1):
                    c = self.nearest(side, x, y)
                    lan += bar_value(c[0] * x + c[1] * y - c[2])
        return np.array([j + dyn + cor + lan, j, dyn, cor, lan])

    # ---- derivatives (dense 6x6 like the Eigen code)
    def cost_derivs(self, idx, s, u):
        g = self.goals[idx]
        Jx = np.array([2 * P.w_x * (s[0] - g[0]), 2 * P.w_y * (s[1] - g[1]), 2 * P.w_theta * (s[2] - g[2]), 0, 0, 0.0])
        Ju = np.array([2 * P.w_jerk * u[0], 2 * P.w_delta_rate * u[1]])
        Hx = np.diag([2 * P.w_x, 2 * P.w_y, 2 * P.w_theta, 2 * P.w_v, 2 * P.w_a, 2 * P.w_delta]).astype(np.float64)
        Hu = np.diag([2 * P.w_jerk, 2 * P.w_delta_rate]).astype(np.float64)
        e = np.eye(6)
        terms = [(0.0 - s[3], -e[3]), (s[3] - P.max_velocity, e[3]), (P.min_acceleration - s[4], -e[4]),
                 (s[4] - P.max_acceleration, e[4]), (P.delta_min - s[5], -e[5]), (s[5] - P.delta_max, e[5])]
        Jx = Jx + sum(bar_jac(gv, d) for gv, d in terms)
        Hx = Hx + sum(bar_hess(gv, d) for gv, d in terms)
        e2 = np.eye(2)
        uterms = [(P.jerk_min - u[0], -e2[0]), (u[0] - P.jerk_max, e2[0]), (P.delta_rate_min - u[1], -e2[1]),
                  (u[1] - P.delta_rate_max, e2[1])]
        Ju = Ju + sum(bar_jac(gv, d) for gv, d in uterms)
        Hu = Hu + sum(bar_hess(gv, d) for gv, d in uterms)
        ddx = np.zeros((6, 6))
        for off in self.off:
            lc, ls = off * math.cos(s[2]), off * math.sin(s[2])
            x, y = s[0] + lc, s[1] + ls
            for c in self.corridor[idx]:
                dx = np.array([c[0], c[1], -c[0] * ls + c[1] * lc, 0, 0, 0.0])
                gv = c[0] * x + c[1] * y - c[2]
                Jx = Jx + bar_jac(gv, dx)
                ddx[2, 2] = -c[0] * lc - c[1] * ls
                Hx = Hx + bar_hess(gv, dx, ddx)
        for off in self.off:
            lc, ls = off * math.cos(s[2]), off * math.sin(s[2])
            x, y = s[0] + lc, s[1] + ls
            for side in (0, 1):
                c = self.nearest(side, x, y)
                dx = np.array([c[0], c[1], -c[0] * ls + c[1] * lc, 0, 0, 0.0])
                gv = c[0] * x + c[1] * y - c[2]
                Jx = Jx + bar_jac(gv, dx)
                ddx[2, 2] = -c[0] * lc - c[1] * ls
                Hx = Hx + bar_hess(gv, dx, ddx)
        return Jx, Ju, Hx, Hu

    def linearize(self, X, U):
        N = self.N
        self.As, self.Bs, self.Jx, self.Ju, self.Hx, self.Hu = [], [], [], [], [], []
        for i in range(N):
            A, B = dynamics_jacobian(X[i], U[i])
            jx, ju, hx, hu = self.cost_derivs(i, X[i], U[i])
            self.As.append(A); self.Bs.append(B); self.Jx.append(jx); self.Ju.append(ju)
            self.Hx.append(hx); self.Hu.append(hu)
        jx, _, hx, _ = self.cost_derivs(N, X[N], np.zeros(2))
        self.Jx.append(jx)
        self.Hx.append(hx)

    def backward(self, lam):
        N = self.N
        Vx, Vxx = self.Jx[N].copy(), self.Hx[N].copy()
        self.Ks, self.ks = [None] * N, [None] * N
        dV = [0.0, 0.0]
        for i in range(N - 1, -1, -1):
            A, B = self.As[i], self.Bs[i]
            Qx = self.Jx[i] + A.T @ Vx
            Qu = self.Ju[i] + B.T @ Vx
            Qxx = self.Hx[i] + A.T @ Vxx @ A
            Quu = self.Hu[i] + B.T @ Vxx @ B
            Qux = B.T @ Vxx @ A
            T = Quu + lam * np.eye(2)
            det = T[0, 0] * T[1, 1] - T[1, 0] * T[0, 1]
            inv = np.array([[T[1, 1], -T[0, 1]], [-T[1, 0], T[0, 0]]]) * (1.0 / det)
            K = -inv @ Qux
            k = -inv @ Qu
            self.Ks[i], self.ks[i] = K, k
            Vx = Qx + K.T @ Quu @ k + K.T @ Qu + Qux.T @ k
            Vxx = Qxx + K.T @ Quu @ K + K.T @ Qux + Qux.T @ K
            # `Vxx = 0.5 * (Vxx + Vxx.transpose())` (:381) is assigned coefficient by coefficient, column-major,
            # without a temporary (no product in it): the upper triangle reads already-overwritten entries (Q22)
            Vxx = Vxx.copy()
            for q in range(6):
                for r in range(6):
                    Vxx[r, q] = 0.5 * (Vxx[r, q] + Vxx[q, r])
            # `auto Qu`, `auto Quu` are lazy Eigen expressions (ilqr_optimizer.cc:349,352): at
            # :383-384 they are re-evaluated with the already-updated Vx / Vxx.
            Qu_l = self.Ju[i] + B.T @ Vx
            Quu_l = self.Hu[i] + B.T @ Vxx @ B
            dV[0] += float(k @ Qu_l)
            dV[1] += float(0.5 * k @ Quu_l @ k)
        self.dV = dV

    def forward(self, alpha, X, U):
        Xn, Un = X.copy(), U.copy()
        x = self.goals[0].copy()
        Xn[0] = x
        for i in range(self.N):
            Un[i] = U[i] + self.Ks[i] @ (x - X[i]) + alpha * self.ks[i]
            Un[i, 1] = normalize_angle(Un[i, 1])
            x = dynamics(x, Un[i])
            Xn[i + 1] = x
        return Xn, Un

    def iqr(self):
        N = self.N
        Q = np.diag([0.001, 0.001, 0.001, 0.001, 0.01, 0.005])
        R = np.diag([0.2, 0.05])  # off-diagonals uninitialised in the reference; 0 here
        Pm = Q.copy()
        Ks = [None] * N
        for i in range(N - 1, -1, -1):
            A, B = dynamics_jacobian(self.goals[i], np.zeros(2))
            S = R + B.T @ Pm @ B
            det = S[0, 0] * S[1, 1] - S[1, 0] * S[0, 1]
            inv = np.array([[S[1, 1], -S[0, 1]], [-S[1, 0], S[0, 0]]]) * (1.0 / det)
            Ks[i] = inv @ (B.T @ Pm @ A)
            Pm = Q + A.T @ Pm @ (A - B @ Ks[i])
        X, U = np.zeros((self.K, 6)), np.zeros((N, 2))
        x = self.goals[0].copy()
        X[0] = x
        for i in range(N):
            u = -Ks[i] @ (x - self.goals[i])
            u[0] = min(P.jerk_max, max(u[0], P.jerk_min))
            u[1] = min(P.delta_rate_max, max(u[1], P.delta_rate_min))
            U[i] = u
            x = dynamics(x, u)
            X[i + 1] = x
        return X, U

    def solve(self):
        X, U = self.iqr()
        init = (X.copy(), U.copy())
        c5 = self.total_cost(X, U)
        cost_old = c5[0]
        hist = [c5]
        lam, dlam = 1.0, 1.0
        updated = True
        status, alphas = 4, []
        it = 0
        while it < P.max_iter_num:
            if updated:
                self.linearize(X, U)
                updated = False
            self.backward(lam)
            gn = sum(max(abs(self.ks[i][0]) / (abs(U[i, 0]) + 1), abs(self.ks[i][1]) / (abs(U[i, 1]) + 1))
                     for i in range(self.N)) / self.N
            if gn < 1e-6 and lam < 1e-5:
                status = 2
                break
            done, ai_acc = False, 11
            for ai, alpha in enumerate(ALPHAS):
                Xn, Un = self.forward(alpha, X, U)
                c5n = self.total_cost(Xn, Un)
                dcost = cost_old - c5n[0]
                expected = -alpha * (self.dV[0] + alpha * self.dV[1])
                with np.errstate(divide="ignore", invalid="ignore"):
                    z = float(np.float64(dcost) / np.float64(expected))  # IEEE division like the C++
                if (1e-4 < z < 10.0) and dcost > 0.0:
                    done, ai_acc = True, ai
                    X, U = Xn, Un
                    break
            alphas.append(ai_acc)
            if done:
                dlam = min(dlam / 1.6, 1.0 / 1.6)
                lam = lam * dlam * (1.0 if lam > 1e-8 else 0.0)
                updated = True
                c5 = c5n
                hist.append(c5n)
                if dcost < P.abs_cost_tol or dcost / cost_old < P.rel_cost_tol:
                    status = 0 if dcost < P.abs_cost_tol else 1
                    cost_old = c5n[0]
                    break
                cost_old = c5n[0]
            else:
                dlam = max(dlam * 1.6, 1.6)
                lam = max(lam * dlam, 1e-8)
                if lam > 1e11:
                    status = 3
                    break
            it += 1
        return dict(states=X, controls=U, status=status, iters=it, cost=c5, cost_hist=np.array(hist),
                    alphas=alphas, init_states=init[0], init_controls=init[1])


def solve_scenario(batch, b):
    s = Solver(batch.start[b], batch.coarse[b], batch.corridor[b], batch.corridor_cnt[b], batch.lane_left[b],
               batch.lane_right[b])
    return s.solve()
