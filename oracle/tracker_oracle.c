/*
 * tracker_oracle.c -- CPU restatement of mpt0816/Cilqr's Tracker (the "better" initial guess of README.md:61).
 * TEST INFRASTRUCTURE ONLY.  Pinned bit for bit against the reference's own tracker.cc +
 * linear_quadratic_regulator.cc compiled against the Eigen stand-in (oracle/_ref/libcilqr_ref_tracker.so,
 * tests/test_reference_pins.py).  Paths are relative to the reference root.
 *
 *   Tracker::lqr                       algorithm/ilqr/tracker.cc:169-215
 *   Tracker::CalcaulateInitState       algorithm/ilqr/tracker.cc:20-60
 *   Tracker::LateralControl / LongitudinalControl   :62-88
 *   Tracker::VehicleDynamic (RK4)      :90-141,  vehicle_mode  tracker.h:79-93
 *   Tracker::InitMatrix                :143-167
 *   math::SolveLQRProblem              algorithm/math/linear_quadratic_regulator.cc:30-70
 *   DiscretizedTrajectory::GetProjection / EvaluateTime / QueryNearestPoint / LinearInterpolate*
 *                                      algorithm/utils/discretized_trajectory.cpp:48-190
 *   IlqrOptimizer::InitGuess           algorithm/ilqr/ilqr_optimizer.cc:107-139
 *
 * Matrices are 3x3 row-major; every product coefficient is l(i,0) r(0,j) + l(i,1) r(1,j) + l(i,2) r(2,j) in that
 * order and nested products are evaluated inside out (Eigen's coefficient-based products of small dynamic
 * matrices); M = 0 is added where the reference adds it (x + 0.0).
 */
#include "tracker_oracle.h"

#include <float.h>
#include <math.h>
#include <string.h>

#define TP 13 /* time s x y theta kappa velocity a jerk delta delta_rate left_bound right_bound */
enum { T_TIME, T_S, T_X, T_Y, T_THETA, T_KAPPA, T_V, T_A, T_JERK, T_DELTA, T_DRATE, T_LB, T_RB };

static const double kEps = 1e-10; /* math::kMathEpsilon */

void tracker_oracle_default_config(tracker_oracle_config* c) {
  c->sumulation_dt = 0.01; /* planner_config.h:36-43 */
  c->dt = 0.1;
  c->tolerance = 0.01;
  c->max_num_iteration = 150;
  c->lat_weight_l = 1e-1; /* :18-25 */
  c->lat_weight_theta = 1e-12;
  c->lat_weight_delta = 1e-12;
  c->lat_weight_delta_rate = 0.1;
  c->lat_preview_time = 0.2;
  c->lon_weight_s = 5.0 * 1e-1; /* :27-34 */
  c->lon_weight_v = 1e-12;
  c->lon_weight_a = 1e-12;
  c->lon_weight_j = 0.1;
  c->wheel_base = 1.0; /* vehicle_param.h:26-64 */
  c->delta_min = -40.0 / 180 * M_PI;
  c->delta_max = 40.0 / 180 * M_PI;
  c->min_acceleration = -5.0;
  c->max_acceleration = 5.0;
  c->delta_rate_min = c->delta_min / 3.0;
  c->delta_rate_max = c->delta_max / 3.0;
  c->jerk_min = -10.0;
  c->jerk_max = 10.0;
}

/* math_utils.cpp:53-59 */
static double normalize_angle(double angle) {
  double a = fmod(angle + M_PI, 2.0 * M_PI);
  if (a < 0.0) a += (2.0 * M_PI);
  return a - M_PI;
}

/* math_utils.h:208-225 */
static double slerp(double a0, double t0, double a1, double t1, double t) {
  if (fabs(t1 - t0) <= kEps) return normalize_angle(a0);
  const double a0_n = normalize_angle(a0);
  const double a1_n = normalize_angle(a1);
  double d = a1_n - a0_n;
  if (d > M_PI) {
    d = d - 2 * M_PI;
  } else if (d < -M_PI) {
    d = d + 2 * M_PI;
  }
  const double r = (t - t0) / (t1 - t0);
  const double a = a0_n + d * r;
  return normalize_angle(a);
}

/* LinearInterpolateTrajectory, discretized_trajectory.cpp:62-84 (fields it does not set keep their defaults, 0) */
static void interp_station(const double* p0, const double* p1, double s, double* pt) {
  const double s0 = p0[T_S], s1 = p1[T_S];
  if (fabs(s1 - s0) < kEps) {
    memcpy(pt, p0, sizeof(double) * TP);
    return;
  }
  memset(pt, 0, sizeof(double) * TP);
  const double weight = (s - s0) / (s1 - s0);
  pt[T_TIME] = (1 - weight) * p0[T_TIME] + weight * p1[T_TIME];
  pt[T_S] = s;
  pt[T_X] = (1 - weight) * p0[T_X] + weight * p1[T_X];
  pt[T_Y] = (1 - weight) * p0[T_Y] + weight * p1[T_Y];
  pt[T_THETA] = slerp(p0[T_THETA], p0[T_S], p1[T_THETA], p1[T_S], s);
  pt[T_KAPPA] = (1 - weight) * p0[T_KAPPA] + weight * p1[T_KAPPA];
  pt[T_V] = (1 - weight) * p0[T_V] + weight * p1[T_V];
  pt[T_LB] = (1 - weight) * p0[T_LB] + weight * p1[T_LB];
  pt[T_RB] = (1 - weight) * p0[T_RB] + weight * p1[T_RB];
}

/* LinearInterpolateTrajectoryWithTime, :86-108 */
static void interp_time(const double* p0, const double* p1, double time, double* pt) {
  const double time0 = p0[T_TIME], time1 = p1[T_TIME];
  if (fabs(time1 - time0) < kEps) {
    memcpy(pt, p0, sizeof(double) * TP);
    return;
  }
  memset(pt, 0, sizeof(double) * TP);
  const double weight = (time - time0) / (time1 - time0);
  pt[T_TIME] = time;
  pt[T_S] = (1 - weight) * p0[T_S] + weight * p1[T_S];
  pt[T_X] = (1 - weight) * p0[T_X] + weight * p1[T_X];
  pt[T_Y] = (1 - weight) * p0[T_Y] + weight * p1[T_Y];
  pt[T_THETA] = slerp(p0[T_THETA], p0[T_TIME], p1[T_THETA], p1[T_TIME], time);
  pt[T_KAPPA] = (1 - weight) * p0[T_KAPPA] + weight * p1[T_KAPPA];
  pt[T_V] = (1 - weight) * p0[T_V] + weight * p1[T_V];
  pt[T_LB] = (1 - weight) * p0[T_LB] + weight * p1[T_LB];
  pt[T_RB] = (1 - weight) * p0[T_RB] + weight * p1[T_RB];
}

/* EvaluateTime, :122-134 with QueryLowerBoundTimePoint :48-60 */
static void evaluate_time(const double* traj, int K, double time, double* pt) {
  int it;
  if (time >= traj[(size_t)(K - 1) * TP + T_TIME]) {
    it = K - 1;
  } else if (time < traj[T_TIME]) {
    it = 0;
  } else {
    int lo = 0, hi = K; /* std::lower_bound: first point whose time is not less than `time` */
    while (lo < hi) {
      const int mid = lo + (hi - lo) / 2;
      if (traj[(size_t)mid * TP + T_TIME] < time) lo = mid + 1; else hi = mid;
    }
    it = lo;
  }
  if (it == 0) it = 1;
  interp_time(traj + (size_t)(it - 1) * TP, traj + (size_t)it * TP, time, pt);
}

/* GetProjection, :156-190 with QueryNearestPoint :136-154 (first minimum) */
static void get_projection(const double* traj, int K, double x, double y, double* project) {
  int idx = 0;
  double nearest = DBL_MAX;
  for (int i = 0; i < K; ++i) {
    const double dx = traj[(size_t)i * TP + T_X] - x, dy = traj[(size_t)i * TP + T_Y] - y;
    const double distance = dx * dx + dy * dy;
    if (distance < nearest) {
      idx = i;
      nearest = distance;
    }
  }
  memcpy(project, traj + (size_t)idx * TP, sizeof(double) * TP);
  const int index_start = idx - 1 > 0 ? idx - 1 : 0;
  const int index_end = idx + 1 < K - 1 ? idx + 1 : K - 1;
  if (index_start < index_end) {
    const double* p0 = traj + (size_t)index_start * TP;
    const double* p1 = traj + (size_t)index_end * TP;
    const double v0x = x - p0[T_X], v0y = y - p0[T_Y];
    const double v1x = p1[T_X] - p0[T_X], v1y = p1[T_Y] - p0[T_Y];
    const double v1_norm = sqrt(v1x * v1x + v1y * v1y);
    const double dot = v0x * v1x + v0y * v1y;
    const double delta_s = dot / v1_norm;
    interp_station(p0, p1, p0[T_S] + delta_s, project);
  }
}

/* ---- small dense helpers: C = A(3x3) B(3x3); products in Eigen's coefficient order ---- */
static void mm33(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double r = A[i * 3 + 0] * B[0 * 3 + j];
      r += A[i * 3 + 1] * B[1 * 3 + j];
      r += A[i * 3 + 2] * B[2 * 3 + j];
      C[i * 3 + j] = r;
    }
}

/* math::SolveLQRProblem, linear_quadratic_regulator.cc:30-70, for a 3-state, 1-input system: K is 1x3 */
void tracker_oracle_solve_lqr(const double A[9], const double B[3], const double Q[9], double R, double tolerance,
                              unsigned max_num_iteration, double K[3], int* iterations) {
  double AT[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) AT[i * 3 + j] = A[j * 3 + i];
  const double M[3] = {0.0, 0.0, 0.0}; /* Matrix::Zero(Q.rows(), R.cols()), :66 */
  double P[9];
  memcpy(P, Q, sizeof(P));
  unsigned num_iteration = 0;
  double diff = DBL_MAX;
  double ATP[9], BTP[3];
  while (num_iteration++ < max_num_iteration && diff > tolerance) {
    /* P_next = AT*P*A - (AT*P*B + M) * (R + BT*P*B).inverse() * (BT*P*A + MT) + Q */
    mm33(AT, P, ATP);
    double ATPA[9];
    mm33(ATP, A, ATPA);
    double ATPB[3];
    for (int i = 0; i < 3; ++i) {
      double r = ATP[i * 3 + 0] * B[0];
      r += ATP[i * 3 + 1] * B[1];
      r += ATP[i * 3 + 2] * B[2];
      ATPB[i] = r + M[i];
    }
    for (int j = 0; j < 3; ++j) { /* BT*P: 1x3 */
      double r = B[0] * P[0 * 3 + j];
      r += B[1] * P[1 * 3 + j];
      r += B[2] * P[2 * 3 + j];
      BTP[j] = r;
    }
    double btpb = BTP[0] * B[0];
    btpb += BTP[1] * B[1];
    btpb += BTP[2] * B[2];
    const double inv = 1.0 / (R + btpb);
    double BTPA[3];
    for (int j = 0; j < 3; ++j) {
      double r = BTP[0] * A[0 * 3 + j];
      r += BTP[1] * A[1 * 3 + j];
      r += BTP[2] * A[2 * 3 + j];
      BTPA[j] = r + M[j];
    }
    double Pn[9];
    double maxc = -DBL_MAX;
    for (int i = 0; i < 3; ++i) {
      const double xi = ATPB[i] * inv; /* (3x1)(1x1) */
      for (int j = 0; j < 3; ++j) {
        Pn[i * 3 + j] = ATPA[i * 3 + j] - xi * BTPA[j] + Q[i * 3 + j];
      }
    }
    for (int i = 0; i < 9; ++i) {
      const double d = Pn[i] - P[i];
      if (d > maxc) maxc = d;
    }
    diff = fabs(maxc); /* fabs((P_next - P).maxCoeff()), :54 */
    memcpy(P, Pn, sizeof(P));
  }
  if (iterations) *iterations = (int)num_iteration;
  /* *ptr_K = (R + BT*P*B).inverse() * (BT*P*A + MT) */
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
  for (int j = 0; j < 3; ++j) {
    double r = BTP[0] * A[0 * 3 + j];
    r += BTP[1] * A[1 * 3 + j];
    r += BTP[2] * A[2 * 3 + j];
    K[j] = inv * (r + M[j]);
  }
}

typedef struct {
  double x, y, theta, v, delta, a;
} vdot;

/* tracker.h:79-93 */
static vdot vehicle_mode(const tracker_oracle_config* c, double theta, double v, double delta, double a, double j,
                         double delta_rate) {
  vdot d;
  d.x = v * cos(theta);
  d.y = v * sin(theta);
  d.theta = v * tan(delta) / c->wheel_base;
  d.v = a;
  d.a = j;
  d.delta = delta_rate;
  return d;
}

/* Tracker::VehicleDynamic, tracker.cc:90-141 */
static void vehicle_dynamic(const tracker_oracle_config* c, const double* cur, double delta_rate, double jerk,
                            double* next) {
  const double dt = c->sumulation_dt;
  const double dt_2 = dt / 2.0;
  const vdot k1 = vehicle_mode(c, cur[T_THETA], cur[T_V], cur[T_DELTA], cur[T_A], jerk, delta_rate);
  const vdot k2 = vehicle_mode(c, cur[T_THETA] + k1.theta * dt_2, cur[T_V] + k1.v * dt_2, cur[T_DELTA] + k1.delta * dt_2,
                               cur[T_A] + k1.a * dt_2, jerk, delta_rate);
  const vdot k3 = vehicle_mode(c, cur[T_THETA] + k2.theta * dt_2, cur[T_V] + k2.v * dt_2, cur[T_DELTA] + k2.delta * dt_2,
                               cur[T_A] + k2.a * dt_2, jerk, delta_rate);
  const vdot k4 = vehicle_mode(c, cur[T_THETA] + k3.theta * dt, cur[T_V] + k3.v * dt, cur[T_DELTA] + k3.delta * dt,
                               cur[T_A] + k3.a * dt, jerk, delta_rate);
  memset(next, 0, sizeof(double) * TP);
  next[T_TIME] = cur[T_TIME] + dt;
  next[T_X] = cur[T_X] + (k1.x + k2.x * 2.0 + k3.x * 2.0 + k4.x) / 6.0 * dt;
  next[T_Y] = cur[T_Y] + (k1.y + k2.y * 2.0 + k3.y * 2.0 + k4.y) / 6.0 * dt;
  next[T_THETA] = normalize_angle(cur[T_THETA] + (k1.theta + k2.theta * 2.0 + k3.theta * 2.0 + k4.theta) / 6.0 * dt);
  next[T_V] = fmax(0.0, cur[T_V] + (k1.v + k2.v * 2.0 + k3.v * 2.0 + k4.v) / 6.0 * dt);
  next[T_DELTA] = normalize_angle(
      fmin(c->delta_max, fmax(c->delta_min, cur[T_DELTA] + (k1.delta + k2.delta * 2.0 + k3.delta * 2.0 + k4.delta) / 6.0 * dt)));
  next[T_A] = fmin(c->max_acceleration,
                   fmax(c->min_acceleration, cur[T_A] + (k1.a + k2.a * 2.0 + k3.a * 2.0 + k4.a) / 6.0 * dt));
  next[T_KAPPA] = tan(next[T_DELTA]) / c->wheel_base;
  const double ds = hypot(next[T_X] - cur[T_X], next[T_Y] - cur[T_Y]);
  next[T_S] = cur[T_S] + ds;
}

/* Tracker::lqr, tracker.cc:169-215.  start: a TrajectoryPoint record (IlqrOptimizer::Plan hands start_state_ with x,
 * y, theta, velocity set, trajectory_planner.cpp:73-75); coarse / out: [K][13].  Returns 1 on success. */
int tracker_oracle_plan(const tracker_oracle_config* c, const double start[TP], const double* coarse, int K,
                        double* out, int* lqr_iterations_total) {
  /* InitMatrix, :143-167 */
  double lat_A[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, lon_A[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double lat_B[3] = {0, 0, 1.0 * c->dt}, lon_B[3] = {0, 0, 1.0 * c->dt};
  double lat_Q[9] = {0}, lon_Q[9] = {0};
  lat_Q[0] = c->lat_weight_l;
  lat_Q[4] = c->lat_weight_theta;
  lat_Q[8] = c->lat_weight_delta;
  lon_A[1] = c->dt;
  lon_A[5] = -c->dt;
  lon_Q[0] = c->lon_weight_s;
  lon_Q[4] = c->lon_weight_v;
  lon_Q[8] = c->lon_weight_a;

  double cur[TP];
  memcpy(cur, start, sizeof(cur));
  int n_out = 0;
  memcpy(out + (size_t)n_out++ * TP, cur, sizeof(cur)); /* trajectory.push_back(cur_state) BEFORE time / s are reset */
  const double start_time = coarse[T_TIME];
  const double end_time = coarse[(size_t)(K - 1) * TP + T_TIME];
  cur[T_TIME] = start_time;
  cur[T_S] = 0.0;
  int total_it = 0;
  int i = 1;
  for (double t = start_time; t < end_time + kEps; t += c->sumulation_dt) {
    /* CalcaulateInitState, :20-60 */
    const double preview_x = cur[T_X] + cos(cur[T_THETA]) * cur[T_V] * c->lat_preview_time;
    const double preview_y = cur[T_Y] + sin(cur[T_THETA]) * cur[T_V] * c->lat_preview_time;
    double project[TP], match[TP];
    get_projection(coarse, K, preview_x, preview_y, project);
    double dx = cur[T_X] - project[T_X];
    double dy = cur[T_Y] - project[T_Y];
    const double l = sin(project[T_THETA]) * dx - cos(project[T_THETA]) * dy;
    const double theta_error = normalize_angle(project[T_THETA] - cur[T_THETA]);
    const double lat_state[3] = {l, theta_error, cur[T_DELTA]};
    evaluate_time(coarse, K, cur[T_TIME] + 0.0, match);
    const double v_error = match[T_V] - cur[T_V];
    const double lon_state[3] = {match[T_S] - project[T_S], v_error, cur[T_A]};
    /* LateralControl, :62-77 */
    const double v_amend = fmax(2, cur[T_V]);
    const double dt = 0.1;
    lat_A[1] = v_amend * dt;
    lat_A[5] = -v_amend / c->wheel_base * dt;
    double Kl[3], Kn[3];
    int it = 0;
    tracker_oracle_solve_lqr(lat_A, lat_B, lat_Q, c->lat_weight_delta_rate, c->tolerance, (unsigned)c->max_num_iteration, Kl, &it);
    total_it += it;
    double ks = Kl[0] * lat_state[0];
    ks += Kl[1] * lat_state[1];
    ks += Kl[2] * lat_state[2];
    double delta_rate = -ks;
    /* LongitudinalControl, :79-88 */
    tracker_oracle_solve_lqr(lon_A, lon_B, lon_Q, c->lon_weight_j, c->tolerance, (unsigned)c->max_num_iteration, Kn, &it);
    total_it += it;
    ks = Kn[0] * lon_state[0];
    ks += Kn[1] * lon_state[1];
    ks += Kn[2] * lon_state[2];
    double jerk = -ks;
    delta_rate = fmax(c->delta_rate_min, fmin(c->delta_rate_max, delta_rate));
    jerk = fmax(c->jerk_min, fmin(c->jerk_max, jerk));
    out[(size_t)(n_out - 1) * TP + T_DRATE] = delta_rate; /* trajectory.back() */
    out[(size_t)(n_out - 1) * TP + T_JERK] = jerk;
    double next[TP];
    vehicle_dynamic(c, cur, delta_rate, jerk, next);
    memcpy(cur, next, sizeof(cur));
    cur[T_TIME] = t;
    if (i >= K) return 0; /* follow_trajectory_.trajectory().at(i) would throw */
    if (cur[T_TIME] > coarse[(size_t)i * TP + T_TIME] - kEps) {
      if (n_out < K) memcpy(out + (size_t)n_out * TP, cur, sizeof(cur));
      ++n_out;
      ++i;
    }
  }
  if (lqr_iterations_total) *lqr_iterations_total = total_it;
  return n_out == K ? 1 : 0;
}

/* IlqrOptimizer::InitGuess, ilqr_optimizer.cc:107-139: the tracker's trajectory as (states, controls) */
void tracker_oracle_init_guess(const double* traj, int K, double* states, double* controls) {
  for (int i = 0; i < K; ++i) {
    const double* p = traj + (size_t)i * TP;
    double* s = states + (size_t)i * 6;
    s[0] = p[T_X];
    s[1] = p[T_Y];
    s[2] = p[T_THETA];
    s[3] = p[T_V];
    s[4] = p[T_A];
    s[5] = p[T_DELTA];
    if (i < K - 1) {
      controls[(size_t)i * 2] = p[T_JERK];
      controls[(size_t)i * 2 + 1] = p[T_DRATE];
    }
  }
}
