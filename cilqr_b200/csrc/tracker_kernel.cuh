// tracker_kernel.cuh -- batched Tracker initial guess for sm_100a.
//
// Replaces Tracker::Plan of mpt0816/Cilqr (algorithm/ilqr/tracker.cc:11-17, 169-215) and the copy
// IlqrOptimizer::InitGuess makes of its result (algorithm/ilqr/ilqr_optimizer.cc:107-139) for B scenarios: a
// closed-loop simulation of the vehicle along the coarse trajectory at 10 ms steps, steered by two discrete LQR
// controllers whose gains come from a fixed-point DARE iteration (math::SolveLQRProblem,
// algorithm/math/linear_quadratic_regulator.cc:30-70) at every step:
//   CalcaulateInitState  tracker.cc:20-60   (DiscretizedTrajectory::GetProjection / EvaluateTime,
//                                            algorithm/utils/discretized_trajectory.cpp:48-190)
//   LateralControl / LongitudinalControl    :62-88
//   VehicleDynamic (RK4)                    :90-141, vehicle_mode tracker.h:79-93
//
// Mapping: ONE THREAD = ONE SCENARIO.  The simulation is 800 strictly sequential steps, each a nearest-point scan
// over the K coarse points, ~30 iterations of 3x3 algebra and an RK4 step -- nothing inside a step is worth a
// warp, and there are B independent ones.  The x / y columns of the thread's coarse trajectory (the scan) live in
// shared memory laid out [knot][thread] (bank = thread: conflict free); everything else is read from global
// memory two records per step.  The longitudinal gain does not depend on the state (constant matrices,
// tracker.cc:79-88), so its DARE is solved once per scenario; the lateral one is re-solved whenever v changes, as
// the reference does.
//
// This translation unit is compiled with -fmad=false: every double expression keeps the reference's operation
// order and roundings (3x3 products as l(i,0) r(0,j) + l(i,1) r(1,j) + l(i,2) r(2,j), nested products inside out).
// cos / sin / tan / hypot / fmod are CUDA's (last-ulp differences from glibc).
#pragma once

#ifndef TRACKER_HOST_EMUL  // tools/tracker_host_emul.cc runs this code on the CPU (development aid)
#include <cuda_runtime.h>
#endif
#include <float.h>
#include <math.h>
#include <stdint.h>

namespace trk {

constexpr int TP = 13;  // time s x y theta kappa velocity a jerk delta delta_rate left_bound right_bound
enum { T_TIME, T_S, T_X, T_Y, T_THETA, T_KAPPA, T_V, T_A, T_JERK, T_DELTA, T_DRATE, T_LB, T_RB };
constexpr double kEps = 1e-10;  // math::kMathEpsilon
constexpr double kPi = 3.14159265358979323846;

struct Config {  // TrackerConfig (planner_config.h:18-43) + the VehicleParam fields the tracker reads
  double sumulation_dt, dt, tolerance;
  double lat_weight_l, lat_weight_theta, lat_weight_delta, lat_weight_delta_rate, lat_preview_time;
  double lon_weight_s, lon_weight_v, lon_weight_a, lon_weight_j;
  double wheel_base, delta_min, delta_max, min_acceleration, max_acceleration, delta_rate_min, delta_rate_max, jerk_min,
      jerk_max;
  int max_num_iteration;
};

struct Args {
  Config c;
  int B, K;
  const double* start;     // [B][4] x, y, theta, v
  const double* coarse;    // [B][K][13]
  double* traj;            // [B][K][13] or nullptr
  double* guess_states;    // [B][K][6] or nullptr     (InitGuess, ilqr_optimizer.cc:122-138)
  double* guess_controls;  // [B][K-1][2] or nullptr
  int* ok;                 // [B]
};

// math_utils.cpp:53-59
__device__ __forceinline__ double normalize_angle(double angle) {
  double a = fmod(angle + kPi, 2.0 * kPi);
  if (a < 0.0) a += (2.0 * kPi);
  return a - kPi;
}

// math_utils.h:208-225
__device__ __forceinline__ double slerp(double a0, double t0, double a1, double t1, double t) {
  if (fabs(t1 - t0) <= kEps) return normalize_angle(a0);
  const double a0_n = normalize_angle(a0);
  const double a1_n = normalize_angle(a1);
  double d = a1_n - a0_n;
  if (d > kPi) {
    d = d - 2 * kPi;
  } else if (d < -kPi) {
    d = d + 2 * kPi;
  }
  const double r = (t - t0) / (t1 - t0);
  const double a = a0_n + d * r;
  return normalize_angle(a);
}

// math::SolveLQRProblem (linear_quadratic_regulator.cc:30-70) for the tracker's systems: A = I + a01 e0 e1^T + a12 e1 e2^T
// is passed densely, B = (0, 0, b2)^T, Q diagonal in the reference's call sites but handled densely as well.
__device__ void solve_lqr(const double* A, const double* B, const double* Q, double R, double tolerance,
                          unsigned max_num_iteration, double* K) {
  double P[9], ATP[9], BTP[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) P[i] = Q[i];
  unsigned num_iteration = 0;
  double diff = DBL_MAX;
  while (num_iteration++ < max_num_iteration && diff > tolerance) {
    // P_next = AT*P*A - (AT*P*B + M) * (R + BT*P*B).inverse() * (BT*P*A + MT) + Q,   M = 0
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double r = A[0 * 3 + i] * P[0 * 3 + j];
        r += A[1 * 3 + i] * P[1 * 3 + j];
        r += A[2 * 3 + i] * P[2 * 3 + j];
        ATP[i * 3 + j] = r;
      }
    double ATPB[3], BTPA[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double r = ATP[i * 3 + 0] * B[0];
      r += ATP[i * 3 + 1] * B[1];
      r += ATP[i * 3 + 2] * B[2];
      ATPB[i] = r + 0.0;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double r = B[0] * P[0 * 3 + j];
      r += B[1] * P[1 * 3 + j];
      r += B[2] * P[2 * 3 + j];
      BTP[j] = r;
    }
    double btpb = BTP[0] * B[0];
    btpb += BTP[1] * B[1];
    btpb += BTP[2] * B[2];
    const double inv = 1.0 / (R + btpb);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double r = BTP[0] * A[0 * 3 + j];
      r += BTP[1] * A[1 * 3 + j];
      r += BTP[2] * A[2 * 3 + j];
      BTPA[j] = r + 0.0;
    }
    double maxc = -DBL_MAX;
    double Pn[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double xi = ATPB[i] * inv;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double a = ATP[i * 3 + 0] * A[0 * 3 + j];
        a += ATP[i * 3 + 1] * A[1 * 3 + j];
        a += ATP[i * 3 + 2] * A[2 * 3 + j];
        Pn[i * 3 + j] = a - xi * BTPA[j] + Q[i * 3 + j];
      }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const double d = Pn[i] - P[i];
      if (d > maxc) maxc = d;
      P[i] = Pn[i];
    }
    diff = fabs(maxc);
  }
  // *ptr_K = (R + BT*P*B).inverse() * (BT*P*A + MT)
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    double r = B[0] * P[0 * 3 + j];
    r += B[1] * P[1 * 3 + j];
    r += B[2] * P[2 * 3 + j];
    BTP[j] = r;
  }
  double btpb = BTP[0] * B[0];
  btpb += BTP[1] * B[1];
  btpb += BTP[2] * B[2];
  const double inv = 1.0 / (R + btpb);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    double r = BTP[0] * A[0 * 3 + j];
    r += BTP[1] * A[1 * 3 + j];
    r += BTP[2] * A[2 * 3 + j];
    K[j] = inv * (r + 0.0);
  }
}

struct VDot {
  double x, y, theta, v, delta, a;
};
// tracker.h:79-93
__device__ __forceinline__ VDot vehicle_mode(const Config& c, double theta, double v, double delta, double a, double j,
                                             double delta_rate) {
  VDot d;
  d.x = v * cos(theta);
  d.y = v * sin(theta);
  d.theta = v * tan(delta) / c.wheel_base;
  d.v = a;
  d.a = j;
  d.delta = delta_rate;
  return d;
}

struct State {  // the fields of cur_state the simulation carries
  double time, s, x, y, theta, kappa, v, a, delta;
};

// Tracker::VehicleDynamic, tracker.cc:90-141
__device__ State vehicle_dynamic(const Config& c, const State& cur, double delta_rate, double jerk) {
  const double dt = c.sumulation_dt;
  const double dt_2 = dt / 2.0;
  const VDot k1 = vehicle_mode(c, cur.theta, cur.v, cur.delta, cur.a, jerk, delta_rate);
  const VDot k2 = vehicle_mode(c, cur.theta + k1.theta * dt_2, cur.v + k1.v * dt_2, cur.delta + k1.delta * dt_2,
                               cur.a + k1.a * dt_2, jerk, delta_rate);
  const VDot k3 = vehicle_mode(c, cur.theta + k2.theta * dt_2, cur.v + k2.v * dt_2, cur.delta + k2.delta * dt_2,
                               cur.a + k2.a * dt_2, jerk, delta_rate);
  const VDot k4 = vehicle_mode(c, cur.theta + k3.theta * dt, cur.v + k3.v * dt, cur.delta + k3.delta * dt,
                               cur.a + k3.a * dt, jerk, delta_rate);
  State n;
  n.time = cur.time + dt;
  n.x = cur.x + (k1.x + k2.x * 2.0 + k3.x * 2.0 + k4.x) / 6.0 * dt;
  n.y = cur.y + (k1.y + k2.y * 2.0 + k3.y * 2.0 + k4.y) / 6.0 * dt;
  n.theta = normalize_angle(cur.theta + (k1.theta + k2.theta * 2.0 + k3.theta * 2.0 + k4.theta) / 6.0 * dt);
  n.v = fmax(0.0, cur.v + (k1.v + k2.v * 2.0 + k3.v * 2.0 + k4.v) / 6.0 * dt);
  n.delta = normalize_angle(
      fmin(c.delta_max, fmax(c.delta_min, cur.delta + (k1.delta + k2.delta * 2.0 + k3.delta * 2.0 + k4.delta) / 6.0 * dt)));
  n.a = fmin(c.max_acceleration, fmax(c.min_acceleration, cur.a + (k1.a + k2.a * 2.0 + k3.a * 2.0 + k4.a) / 6.0 * dt));
  n.kappa = tan(n.delta) / c.wheel_base;
  const double ds = hypot(n.x - cur.x, n.y - cur.y);
  n.s = cur.s + ds;
  return n;
}

__host__ __device__ inline size_t smem_bytes_per_thread(int K) { return sizeof(double) * 2 * (size_t)K; }

__device__ __forceinline__ void write_point(double* r, const State& s, double jerk, double drate) {
  r[T_TIME] = s.time;
  r[T_S] = s.s;
  r[T_X] = s.x;
  r[T_Y] = s.y;
  r[T_THETA] = s.theta;
  r[T_KAPPA] = s.kappa;
  r[T_V] = s.v;
  r[T_A] = s.a;
  r[T_JERK] = jerk;
  r[T_DELTA] = s.delta;
  r[T_DRATE] = drate;
  r[T_LB] = 0.0;
  r[T_RB] = 0.0;
}

__global__ void tracker_kernel(const Args a) {
#ifndef TRACKER_HOST_EMUL
  extern __shared__ __align__(16) double trk_smem[];  // [2][K][blockDim.x]: x then y columns, knot-major
#endif
  const Config& c = a.c;
  const int K = a.K, nt = blockDim.x, tid = threadIdx.x;
  double* sx = trk_smem + tid;
  double* sy = trk_smem + (size_t)K * nt + tid;
  for (int b = blockIdx.x * nt + tid; b - tid < a.B; b += gridDim.x * nt) {  // (whole CTAs iterate together)
    const bool live = b < a.B;
    const double* co = a.coarse + (size_t)(live ? b : 0) * K * TP;
    for (int k = 0; k < K; ++k) {
      sx[(size_t)k * nt] = co[(size_t)k * TP + T_X];
      sy[(size_t)k * nt] = co[(size_t)k * TP + T_Y];
    }
    if (!live) continue;
    // InitMatrix, tracker.cc:143-167
    double lat_A[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, lon_A[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double lat_B[3] = {0, 0, 1.0 * c.dt}, lon_B[3] = {0, 0, 1.0 * c.dt};
    double lat_Q[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, lon_Q[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    lat_Q[0] = c.lat_weight_l;
    lat_Q[4] = c.lat_weight_theta;
    lat_Q[8] = c.lat_weight_delta;
    lon_A[1] = c.dt;
    lon_A[5] = -c.dt;
    lon_Q[0] = c.lon_weight_s;
    lon_Q[4] = c.lon_weight_v;
    lon_Q[8] = c.lon_weight_a;
    double Kn[3];  // LongitudinalControl's gain: its matrices never change (tracker.cc:79-88)
    solve_lqr(lon_A, lon_B, lon_Q, c.lon_weight_j, c.tolerance, (unsigned)c.max_num_iteration, Kn);
    double Kl[3] = {0, 0, 0}, v_amend_cached = -1.0;

    State cur;
    const double* st = a.start + (size_t)b * 4;
    cur.time = 0.0;
    cur.s = 0.0;
    cur.x = st[0];
    cur.y = st[1];
    cur.theta = st[2];
    cur.v = st[3];
    cur.kappa = 0.0;
    cur.a = 0.0;
    cur.delta = 0.0;
    State pushed = cur;  // trajectory.back(): written out once its jerk / delta_rate are final
    int n_out = 1;
    const double start_time = co[T_TIME];
    const double end_time = co[(size_t)(K - 1) * TP + T_TIME];
    cur.time = start_time;
    cur.s = 0.0;
    double last_jerk = 0.0, last_drate = 0.0;
    bool failed = false;
    int i = 1;
    auto flush = [&](int k) {  // knot k of the result: the state pushed at that knot + the controls of the LAST step before the next push
      if (k >= K) return;
      if (a.traj) write_point(a.traj + ((size_t)b * K + k) * TP, pushed, last_jerk, last_drate);
      if (a.guess_states) {
        double* g = a.guess_states + ((size_t)b * K + k) * 6;
        g[0] = pushed.x; g[1] = pushed.y; g[2] = pushed.theta; g[3] = pushed.v; g[4] = pushed.a; g[5] = pushed.delta;
      }
      if (a.guess_controls && k < K - 1) {
        double* g = a.guess_controls + ((size_t)b * (K - 1) + k) * 2;
        g[0] = last_jerk;
        g[1] = last_drate;
      }
    };
    for (double t = start_time; t < end_time + kEps; t += c.sumulation_dt) {
      // ---- CalcaulateInitState, tracker.cc:20-60
      const double pvx = cur.x + cos(cur.theta) * cur.v * c.lat_preview_time;
      const double pvy = cur.y + sin(cur.theta) * cur.v * c.lat_preview_time;
      // GetProjection (discretized_trajectory.cpp:156-190): QueryNearestPoint, first minimum
      int idx = 0;
      double nearest = DBL_MAX;
      for (int k = 0; k < K; ++k) {
        const double dx = sx[(size_t)k * nt] - pvx, dy = sy[(size_t)k * nt] - pvy;
        const double distance = dx * dx + dy * dy;
        if (distance < nearest) {
          idx = k;
          nearest = distance;
        }
      }
      double pj_x = co[(size_t)idx * TP + T_X], pj_y = co[(size_t)idx * TP + T_Y], pj_th = co[(size_t)idx * TP + T_THETA],
             pj_s = co[(size_t)idx * TP + T_S];
      const int index_start = idx - 1 > 0 ? idx - 1 : 0;
      const int index_end = idx + 1 < K - 1 ? idx + 1 : K - 1;
      if (index_start < index_end) {
        const double* p0 = co + (size_t)index_start * TP;
        const double* p1 = co + (size_t)index_end * TP;
        const double v0x = pvx - p0[T_X], v0y = pvy - p0[T_Y];
        const double v1x = p1[T_X] - p0[T_X], v1y = p1[T_Y] - p0[T_Y];
        const double v1_norm = sqrt(v1x * v1x + v1y * v1y);
        const double dot = v0x * v1x + v0y * v1y;
        const double delta_s = dot / v1_norm;
        const double s = p0[T_S] + delta_s;
        // LinearInterpolateTrajectory, :62-84 (the fields the tracker reads: s, x, y, theta)
        const double s0 = p0[T_S], s1 = p1[T_S];
        if (fabs(s1 - s0) < kEps) {
          pj_x = p0[T_X]; pj_y = p0[T_Y]; pj_th = p0[T_THETA]; pj_s = p0[T_S];
        } else {
          const double weight = (s - s0) / (s1 - s0);
          pj_s = s;
          pj_x = (1 - weight) * p0[T_X] + weight * p1[T_X];
          pj_y = (1 - weight) * p0[T_Y] + weight * p1[T_Y];
          pj_th = slerp(p0[T_THETA], p0[T_S], p1[T_THETA], p1[T_S], s);
        }
      }
      const double dx = cur.x - pj_x;
      const double dy = cur.y - pj_y;
      const double l = sin(pj_th) * dx - cos(pj_th) * dy;
      const double theta_error = normalize_angle(pj_th - cur.theta);
      // EvaluateTime (:122-134, QueryLowerBoundTimePoint :48-60): the fields the tracker reads are s and velocity
      double m_s, m_v;
      {
        const double time = cur.time + 0.0;
        int it;
        if (time >= end_time) {
          it = K - 1;
        } else if (time < start_time) {
          it = 0;
        } else {
          int lo = 0, hi = K;
          while (lo < hi) {
            const int mid = lo + (hi - lo) / 2;
            if (co[(size_t)mid * TP + T_TIME] < time) lo = mid + 1; else hi = mid;
          }
          it = lo;
        }
        if (it == 0) it = 1;
        const double* p0 = co + (size_t)(it - 1) * TP;
        const double* p1 = co + (size_t)it * TP;
        const double time0 = p0[T_TIME], time1 = p1[T_TIME];
        if (fabs(time1 - time0) < kEps) {
          m_s = p0[T_S];
          m_v = p0[T_V];
        } else {
          const double weight = (time - time0) / (time1 - time0);
          m_s = (1 - weight) * p0[T_S] + weight * p1[T_S];
          m_v = (1 - weight) * p0[T_V] + weight * p1[T_V];
        }
      }
      const double v_error = m_v - cur.v;
      // ---- LateralControl, :62-77
      const double v_amend = fmax(2.0, cur.v);
      if (v_amend != v_amend_cached) {  // (same matrices -> same gain: skip the identical solve)
        const double dt = 0.1;
        lat_A[1] = v_amend * dt;
        lat_A[5] = -v_amend / c.wheel_base * dt;
        solve_lqr(lat_A, lat_B, lat_Q, c.lat_weight_delta_rate, c.tolerance, (unsigned)c.max_num_iteration, Kl);
        v_amend_cached = v_amend;
      }
      double ks = Kl[0] * l;
      ks += Kl[1] * theta_error;
      ks += Kl[2] * cur.delta;
      double delta_rate = -ks;
      // ---- LongitudinalControl, :79-88
      ks = Kn[0] * (m_s - pj_s);
      ks += Kn[1] * v_error;
      ks += Kn[2] * cur.a;
      double jerk = -ks;
      delta_rate = fmax(c.delta_rate_min, fmin(c.delta_rate_max, delta_rate));
      jerk = fmax(c.jerk_min, fmin(c.jerk_max, jerk));
      last_drate = delta_rate;  // trajectory.back().delta_rate / .jerk, :189-190
      last_jerk = jerk;
      cur = vehicle_dynamic(c, cur, delta_rate, jerk);
      cur.time = t;
      if (i >= K) {  // follow_trajectory_.trajectory().at(i) would throw, :198
        failed = true;
        break;
      }
      if (cur.time > co[(size_t)i * TP + T_TIME] - kEps) {
        flush(n_out - 1);  // the previous knot's controls are final now
        pushed = cur;
        last_jerk = 0.0;   // a freshly pushed point carries jerk = delta_rate = 0 until the next step sets them
        last_drate = 0.0;
        ++n_out;
        ++i;
      }
    }
    flush(n_out - 1);
    a.ok[b] = (!failed && n_out == K) ? 1 : 0;
  }
}

}  // namespace trk
