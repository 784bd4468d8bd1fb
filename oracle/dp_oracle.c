/*
 * dp_oracle.c -- see dp_oracle.h.  TEST INFRASTRUCTURE ONLY.  -O2 -ffp-contract=off.
 */
#include "dp_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MATH_EPS 1e-10 /* math::kMathEpsilon, algorithm/math/math_utils.h */
#define DP_EPS 1e-3    /* the file-local kMathEpsilon of dp_planner.cpp:29 */
#define REF_STRIDE 7

void dp_default_config(dp_config* c) {
  c->tf = 8.0;
  c->delta_t = 0.1;
  c->dp_nominal_velocity = 10.0;
  c->dp_w_obstacle = 1000.0;
  c->dp_w_lateral = 0.1;
  c->dp_w_lateral_change = 0.5;
  c->dp_w_lateral_velocity_change = 1.0;
  c->dp_w_longitudinal_velocity_bias = 10.0;
  c->dp_w_longitudinal_velocity_change = 1.0;
  c->max_velocity = 20.0; /* vehicle_param.h:46 */
  c->width = 1.942;
  c->wheel_base = 1.0;
  c->front_hang_length = 0.96;
  c->rear_hang_length = 0.929;
}

/* ---- math_utils.cpp:53-59, math_utils.h:208-225 ------------------------------------------------ */
static double normalize_angle(double angle) {
  double a = fmod(angle + M_PI, 2.0 * M_PI);
  if (a < 0.0) a += 2.0 * M_PI;
  return a - M_PI;
}

static double slerp(double a0, double t0, double a1, double t1, double t) {
  if (fabs(t1 - t0) <= MATH_EPS) return normalize_angle(a0);
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

/* ---- DiscretizedTrajectory, utils/discretized_trajectory.cpp ----------------------------------- */
enum { F_S = 0, F_X, F_Y, F_TH, F_KAPPA, F_LB, F_RB };

/* QueryLowerBoundStationPoint, :34-46 */
static int lower_bound_station(int R, const double* ref, double station) {
  if (station >= ref[(size_t)(R - 1) * REF_STRIDE + F_S]) return R - 1;
  if (station < ref[F_S]) return 0;
  int lo = 0, hi = R; /* std::lower_bound: first index with s >= station */
  while (lo < hi) {
    const int mid = lo + (hi - lo) / 2;
    if (ref[(size_t)mid * REF_STRIDE + F_S] < station) lo = mid + 1; else hi = mid;
  }
  return lo;
}

/* LinearInterpolateTrajectory, :62-84 (time and velocity are not carried: always 0 in the reference line) */
static void interpolate(const double* p0, const double* p1, double s, double out[7]) {
  const double s0 = p0[F_S], s1 = p1[F_S];
  if (fabs(s1 - s0) < MATH_EPS) {
    memcpy(out, p0, sizeof(double) * 7);
    return;
  }
  const double weight = (s - s0) / (s1 - s0);
  out[F_S] = s;
  out[F_X] = (1 - weight) * p0[F_X] + weight * p1[F_X];
  out[F_Y] = (1 - weight) * p0[F_Y] + weight * p1[F_Y];
  out[F_TH] = slerp(p0[F_TH], p0[F_S], p1[F_TH], p1[F_S], s);
  out[F_KAPPA] = (1 - weight) * p0[F_KAPPA] + weight * p1[F_KAPPA];
  out[F_LB] = (1 - weight) * p0[F_LB] + weight * p1[F_LB];
  out[F_RB] = (1 - weight) * p0[F_RB] + weight * p1[F_RB];
}

/* EvaluateStation, :110-121 */
void dp_evaluate_station(int R, const double* ref, double station, double out[7]) {
  int it = lower_bound_station(R, ref, station);
  if (it == 0) it = 1;
  interpolate(ref + (size_t)(it - 1) * REF_STRIDE, ref + (size_t)it * REF_STRIDE, station, out);
}

/* GetCartesian, :192-196 */
static void get_cartesian(int R, const double* ref, double station, double lateral, double* x, double* y) {
  double r[7];
  dp_evaluate_station(R, ref, station, r);
  *x = r[F_X] - lateral * sin(r[F_TH]);
  *y = r[F_Y] + lateral * cos(r[F_TH]);
}

/* GetProjection, :156-190 (QueryNearestPoint :136-154: first minimum of the squared distance) */
void dp_get_projection(int R, const double* ref, double x, double y, double sl[2]) {
  long idx = 0;
  double nearest = DBL_MAX;
  for (int i = 0; i < R; ++i) {
    const double dx = ref[(size_t)i * REF_STRIDE + F_X] - x, dy = ref[(size_t)i * REF_STRIDE + F_Y] - y;
    const double d = dx * dx + dy * dy;
    if (d < nearest) {
      nearest = d;
      idx = i;
    }
  }
  double pp[7];
  memcpy(pp, ref + (size_t)idx * REF_STRIDE, sizeof(pp));
  const long index_start = idx - 1 > 0 ? idx - 1 : 0;
  const long index_end = idx + 1 < R - 1 ? idx + 1 : R - 1;
  if (index_start < index_end) {
    const double* a = ref + (size_t)index_start * REF_STRIDE;
    const double* b = ref + (size_t)index_end * REF_STRIDE;
    const double v0x = x - a[F_X], v0y = y - a[F_Y];
    const double v1x = b[F_X] - a[F_X], v1y = b[F_Y] - a[F_Y];
    const double v1_norm = sqrt(v1x * v1x + v1y * v1y);
    const double dot = v0x * v1x + v0y * v1y;
    const double delta_s = dot / v1_norm;
    interpolate(a, b, a[F_S] + delta_s, pp);
  }
  const double nr_x = x - pp[F_X], nr_y = y - pp[F_Y];
  sl[0] = pp[F_S];
  sl[1] = copysign(hypot(nr_x, nr_y), nr_y * cos(pp[F_TH]) - nr_x * sin(pp[F_TH]));
}

/* ---- Environment::set_reference, environment.cpp:24-49 ----------------------------------------- */
static int cmp_x(const void* a, const void* b) {
  const double xa = *(const double*)a, xb = *(const double*)b;
  return xa < xb ? -1 : (xa > xb ? 1 : 0);
}

int dp_build_barrier(int R, const double* ref, double* out, int cap) {
  const double start_s = ref[F_S], back_s = ref[(size_t)(R - 1) * REF_STRIDE + F_S];
  const int sample_points = (int)((back_s - start_s) / 0.1);
  int n = 0;
  for (int i = 0; i <= sample_points; ++i) {
    const double s = start_s + i * 0.1;
    double r[7];
    dp_evaluate_station(R, ref, s, r);
    if (n + 2 > cap) return -1;
    get_cartesian(R, ref, s, r[F_LB], &out[2 * n], &out[2 * n + 1]);
    ++n;
    get_cartesian(R, ref, s, -r[F_RB], &out[2 * n], &out[2 * n + 1]);
    ++n;
  }
  qsort(out, (size_t)n, 2 * sizeof(double), cmp_x); /* std::sort by x only; ties are unordered there too */
  return n;
}

/* ---- collision checks -------------------------------------------------------------------------- */
/* Polygon2d::IsPointIn, polygon2d.cpp:120-140 */
static int polygon_is_point_in(const double* p, int nv, double minx, double maxx, double miny, double maxy, double x,
                               double y) {
  if (x < minx || x > maxx || y < miny || y > maxy) return 0;
  int j = nv - 1, c = 0;
  for (int i = 0; i < nv; ++i) {
    const double xi = p[2 * i], yi = p[2 * i + 1], xj = p[2 * j], yj = p[2 * j + 1];
    if ((yi > y) != (yj > y)) {
      const double side = (xi - x) * (yj - y) - (yi - y) * (xj - x); /* CrossProd(point, p_i, p_j) */
      if (yi < yj ? side > 0.0 : side < 0.0) ++c;
    }
    j = i;
  }
  return c & 1;
}

/* Polygon2d::HasOverlap(const Box2d&), polygon2d.cpp:150-165, for the axis-aligned square
 * Box2d(AABox2d) of Environment::CheckOptimizationCollision (box2d.cpp:93-105, aabox2d.cpp:63-71) */
static int polygon_overlaps_box(const double* p, int nv, double cx, double cy, double half) {
  double minx = p[0], maxx = p[0], miny = p[1], maxy = p[1];
  for (int i = 1; i < nv; ++i) {
    minx = fmin(minx, p[2 * i]);
    maxx = fmax(maxx, p[2 * i]);
    miny = fmin(miny, p[2 * i + 1]);
    maxy = fmax(maxy, p[2 * i + 1]);
  }
  const double bminx = cx - half, bmaxx = cx + half, bminy = cy - half, bmaxy = cy + half;
  if (bmaxx < minx || bminx > maxx || bmaxy < miny || bminy > maxy) return 0;
  for (int i = 0; i < nv; ++i) { /* Box2d::IsPointIn, box2d.cpp:123-129, cos = 1, sin = 0 */
    const double x0 = p[2 * i] - cx, y0 = p[2 * i + 1] - cy;
    const double dx = fabs(x0 * 1.0 + y0 * 0.0), dy = fabs(-x0 * 0.0 + y0 * 1.0);
    if (dx <= half + MATH_EPS && dy <= half + MATH_EPS) return 1;
  }
  const double kx[4] = {cx + half, cx + half, cx - half, cx - half};
  const double ky[4] = {cy - half, cy + half, cy + half, cy - half};
  for (int q = 0; q < 4; ++q)
    if (polygon_is_point_in(p, nv, minx, maxx, miny, maxy, kx[q], ky[q])) return 1;
  return 0;
}

/* Environment::CheckStaticCollision, environment.cpp:51-87 */
static int check_static(const dp_env* e, double cx, double cy, double half) {
  for (int o = 0; o < e->n_static; ++o)
    if (polygon_overlaps_box(e->static_poly + (size_t)o * e->V * 2, e->static_nv[o], cx, cy, half)) return 1;
  if (e->NB == 0) return 0;
  const double minx = cx - half, maxx = cx + half;
  if (maxx < e->barrier[0] || minx > e->barrier[(size_t)(e->NB - 1) * 2]) return 0;
  /* std::upper_bound(val < a.x): first index with x > val */
  int lo = 0, hi = e->NB;
  while (lo < hi) {
    const int mid = lo + (hi - lo) / 2;
    if (minx < e->barrier[(size_t)mid * 2]) hi = mid; else lo = mid + 1;
  }
  int check_start = lo;
  lo = 0;
  hi = e->NB;
  while (lo < hi) {
    const int mid = lo + (hi - lo) / 2;
    if (maxx < e->barrier[(size_t)mid * 2]) hi = mid; else lo = mid + 1;
  }
  const int check_end = lo;
  if (check_start > 0) --check_start;
  for (int i = check_start; i < check_end; ++i) {
    const double x0 = e->barrier[(size_t)i * 2] - cx, y0 = e->barrier[(size_t)i * 2 + 1] - cy;
    const double dx = fabs(x0 * 1.0 + y0 * 0.0), dy = fabs(-x0 * 0.0 + y0 * 1.0);
    if (dx <= half + MATH_EPS && dy <= half + MATH_EPS) return 1;
  }
  return 0;
}

/* Environment::CheckDynamicCollision, environment.cpp:124-141.  When `time` equals the last sample's time the
 * reference dereferences end() (upper_bound finds nothing); the last sample is used here. */
static int check_dynamic(const dp_env* e, double time, double cx, double cy, double half) {
  for (int o = 0; o < e->n_dyn; ++o) {
    const int ns = e->dyn_samples[o];
    if (ns <= 0) continue;
    const double* tt = e->dyn_time + (size_t)o * e->T;
    if (tt[0] > time || tt[ns - 1] < time) continue;
    int lo = 0, hi = ns;
    while (lo < hi) {
      const int mid = lo + (hi - lo) / 2;
      if (time < tt[mid]) hi = mid; else lo = mid + 1;
    }
    if (lo >= ns) lo = ns - 1;
    if (polygon_overlaps_box(e->dyn_poly + ((size_t)o * e->T + lo) * e->V * 2, e->dyn_nv[o], cx, cy, half)) return 1;
  }
  return 0;
}

/* Environment::CheckOptimizationCollision, environment.cpp:99-122 (collision_buffer = 0);
 * VehicleParam::radius / f2x / r2x, vehicle_param.h:80-85; GetDiscPositions :88-95 */
int dp_check_optimization_collision(const dp_config* cfg, const dp_env* env, double time, double x, double y,
                                    double theta) {
  const double length = cfg->wheel_base + cfg->rear_hang_length + cfg->front_hang_length;
  const double radius = hypot(0.25 * length, 0.5 * cfg->width);
  const double r2x = 0.25 * length - cfg->rear_hang_length;
  const double f2x = 0.75 * length - cfg->rear_hang_length;
  const double half = (radius + 0.0 - (-radius - 0.0)) / 2.0; /* AABox2d(one_corner, opposite_corner): |dx| / 2 */
  const double xf = x + f2x * cos(theta), xr = x + r2x * cos(theta);
  const double yf = y + f2x * sin(theta), yr = y + r2x * sin(theta);
  /* initial_box centre = ((-r) + r) / 2 = 0, then Shift */
  const double c0 = (-radius - 0.0 + (radius + 0.0)) / 2.0;
  const double fx = c0 + xf, fy = c0 + yf, rx = c0 + xr, ry = c0 + yr;
  if (check_static(env, fx, fy, half) || check_static(env, rx, ry, half) || check_dynamic(env, time, fx, fy, half) ||
      check_dynamic(env, time, rx, ry, half))
    return 1;
  return 0;
}

/* ---- the planner ------------------------------------------------------------------------------- */
typedef struct {
  double cost, current_s;
  int parent_s_ind, parent_l_ind;
} cell_t;

typedef struct {
  const dp_config* cfg;
  const dp_env* env;
  double unit_time, time_[DP_NT], station_[DP_NS], lateral_[DP_NL - 1], safe_margin;
  int nseg[DP_NT];
  double start_s, start_l, start_theta;
  cell_t space[DP_NT][DP_NS][DP_NL];
} planner_t;

/* DpPlanner::DpPlanner, dp_planner.cpp:31-38; math::LinSpaced, math_utils.h:244-254 */
static void planner_init(planner_t* p, const dp_config* cfg, const dp_env* env) {
  p->cfg = cfg;
  p->env = env;
  p->unit_time = cfg->tf / DP_NT;
  {
    const double step = (cfg->tf - p->unit_time) / (DP_NT - 1);
    for (int i = 0; i < DP_NT; ++i) p->time_[i] = p->unit_time + step * i;
  }
  {
    const double step = (p->unit_time * cfg->max_velocity - 0) / (DP_NS - 1);
    for (int i = 0; i < DP_NS; ++i) p->station_[i] = 0 + step * i;
  }
  {
    const double step = (1.0 - 0) / (DP_NL - 1 - 1);
    for (int i = 0; i < DP_NL - 1; ++i) p->lateral_[i] = 0 + step * i;
  }
  p->safe_margin = cfg->width / 2 * 1.5;
  /* the segment counts of InterpolateLinearly (:287-298) depend only on the layer */
  for (int c = 0; c < DP_NT; ++c) {
    int nseg = 0;
    for (double t = 0.0; t < cfg->tf + cfg->delta_t - MATH_EPS; t += cfg->delta_t) {
      if (c == 0) {
        if (t > 0.0 - DP_EPS && t < p->unit_time + DP_EPS) ++nseg;
      } else {
        if (t > p->time_[c] - p->unit_time + MATH_EPS && t < p->time_[c] + MATH_EPS) ++nseg;
      }
    }
    p->nseg[c] = nseg;
  }
}

int dp_num_knots(const dp_config* cfg) {
  planner_t* p = (planner_t*)malloc(sizeof(planner_t));
  planner_init(p, cfg, NULL);
  int n = 0;
  for (int c = 0; c < DP_NT; ++c) n += p->nseg[c];
  free(p);
  return n;
}

/* GetLateralOffset, dp_planner.h:83-92 */
static double lateral_offset(const planner_t* p, double s, int l_ind) {
  if (l_ind == DP_NL - 1) return 0.0;
  double r[7];
  dp_evaluate_station(p->env->R, p->env->ref, s, r);
  const double lb = -r[F_RB] + p->safe_margin;
  const double ub = r[F_LB] - p->safe_margin;
  return lb + (ub - lb) * p->lateral_[l_ind];
}

/* InterpolateLinearly, dp_planner.cpp:283-320: path[i] = (s, l), i < nseg; returns nseg */
static int interpolate_linearly(const planner_t* p, double parent_s, int parent_l_ind, int cur_t_ind, int cur_s_ind,
                                int cur_l_ind, double* ps, double* pl) {
  const int nseg = p->nseg[cur_t_ind];
  double p_l = p->start_l, p_s = p->start_s;
  if (parent_l_ind >= 0) {
    p_s = parent_s;
    p_l = lateral_offset(p, p_s, parent_l_ind);
  }
  const double cur_s = p_s + p->station_[cur_s_ind];
  const double cur_l = lateral_offset(p, cur_s, cur_l_ind);
  const double s_step = p->station_[cur_s_ind] / nseg;
  const double l_step = (cur_l - p_l) / nseg;
  for (int i = 0; i < nseg; ++i) {
    ps[i] = p_s + i * s_step;
    pl[i] = p_l + i * l_step;
  }
  return nseg;
}

#define MAX_SEG 256

/* GetCollisionCost, dp_planner.cpp:40-85; parent t < 0 = the start state */
static double collision_cost(const planner_t* p, int pt, int psi, int pli, int ct, int csi, int cli) {
  double parent_s = p->start_s, grandparent_s = p->start_s;
  double last_l = p->start_l, last_s = p->start_s;
  double ps[MAX_SEG], pl[MAX_SEG];
  int parent_l_for_path = pli;
  if (pt >= 0) {
    const cell_t* cell = &p->space[pt][psi][pli];
    parent_s = cell->current_s;
    if (pt > 0) grandparent_s = p->space[pt - 1][cell->parent_s_ind][cell->parent_l_ind].current_s;
    const int n0 = interpolate_linearly(p, grandparent_s, cell->parent_l_ind, pt, psi, pli, ps, pl);
    last_l = pl[n0 - 1];
    last_s = ps[n0 - 1];
  }
  const int nseg = interpolate_linearly(p, parent_s, parent_l_for_path, ct, csi, cli, ps, pl);
  for (int i = 0; i < nseg; ++i) {
    const double dl = pl[i] - last_l;
    const double ds = fmax(ps[i] - last_s, DP_EPS);
    last_l = pl[i];
    last_s = ps[i];
    double cx, cy, r[7];
    get_cartesian(p->env->R, p->env->ref, ps[i], pl[i], &cx, &cy);
    dp_evaluate_station(p->env->R, p->env->ref, ps[i], r);
    const double lb = fmin(0.0, -r[F_RB] + p->safe_margin);
    const double ub = fmax(0.0, r[F_LB] - p->safe_margin);
    if (pl[i] < lb - DP_EPS || pl[i] > ub + DP_EPS) return p->cfg->dp_w_obstacle;
    const double heading = r[F_TH] + atan((dl / ds) / (1 - r[F_KAPPA] * pl[i]));
    const double parent_time = pt < 0 ? 0.0 : p->time_[pt];
    const double time = parent_time + i * (p->unit_time / nseg);
    if (dp_check_optimization_collision(p->cfg, p->env, time, cx, cy, heading)) return p->cfg->dp_w_obstacle;
  }
  return 0.0;
}

/* GetCost, dp_planner.cpp:87-133: returns delta cost, *cur_s_out = cur_s */
static double get_cost(const planner_t* p, int pt, int psi, int pli, int ct, int csi, int cli, double* cur_s_out) {
  const dp_config* cfg = p->cfg;
  double parent_s = p->start_s, grandparent_s = p->start_s;
  double parent_l = p->start_l, grandparent_l = p->start_l;
  if (pt >= 0) {
    const cell_t* cell = &p->space[pt][psi][pli];
    const int gs = cell->parent_s_ind, gl = cell->parent_l_ind;
    parent_s = cell->current_s;
    parent_l = lateral_offset(p, parent_s, pli);
    if (pt >= 1) {
      grandparent_s = p->space[pt - 1][gs][gl].current_s;
      grandparent_l = lateral_offset(p, grandparent_s, gl);
    }
  }
  const double cur_s = parent_s + p->station_[csi];
  const double cur_l = lateral_offset(p, cur_s, cli);
  const double ds1 = cur_s - parent_s;
  const double dl1 = cur_l - parent_l;
  const double ds0 = parent_s - grandparent_s;
  const double dl0 = parent_l - grandparent_l;
  *cur_s_out = cur_s;
  const double cost_obstacle = collision_cost(p, pt, psi, pli, ct, csi, cli);
  if (cost_obstacle >= cfg->dp_w_obstacle) return cfg->dp_w_obstacle;
  const double cost_lateral = fabs(cur_l);
  const double cost_lateral_change = fabs(parent_l - cur_l) / (p->station_[csi] + DP_EPS);
  const double cost_lateral_change_t = fabs(dl1 - dl0) / p->unit_time;
  const double cost_longitudinal_velocity = fabs(ds1 / p->unit_time - cfg->dp_nominal_velocity);
  const double cost_longitudinal_velocity_change = fabs((ds1 - ds0) / p->unit_time);
  return cfg->dp_w_lateral * cost_lateral + cfg->dp_w_lateral_change * cost_lateral_change +
         cfg->dp_w_lateral_velocity_change * cost_lateral_change_t +
         cfg->dp_w_longitudinal_velocity_bias * cost_longitudinal_velocity +
         cfg->dp_w_longitudinal_velocity_change * cost_longitudinal_velocity_change;
}

/* DiscretePointsMath::ComputePathProfile, discrete_points_math.cc:27-176 (speeds, accelerations, kappas) */
static void path_profile(double dt, int n, const double* x, const double* y, double* speeds, double* accel,
                         double* kappas) {
  double* acc_s = (double*)malloc(sizeof(double) * (size_t)n * 5);
  double *xds = acc_s + n, *yds = xds + n, *xdds = yds + n, *ydds = xdds + n;
  double distance = 0.0, fx = x[0], fy = y[0];
  acc_s[0] = distance;
  for (int i = 1; i < n; ++i) {
    const double nx = x[i], ny = y[i];
    const double end_segment_s = sqrt((fx - nx) * (fx - nx) + (fy - ny) * (fy - ny));
    acc_s[i] = end_segment_s + distance;
    distance += end_segment_s;
    fx = nx;
    fy = ny;
  }
  for (int i = 1; i < n; ++i) speeds[i - 1] = (acc_s[i] - acc_s[i - 1]) / dt;
  speeds[n - 1] = speeds[n - 2];
  for (int i = 1; i < n; ++i) accel[i - 1] = (speeds[i] - speeds[i - 1]) / dt;
  accel[n - 1] = accel[n - 2];
  for (int i = 0; i < n; ++i) {
    const int a = i == 0 ? 0 : i - 1, b = i == n - 1 ? n - 1 : i + 1;
    xds[i] = (x[b] - x[a]) / (acc_s[b] - acc_s[a]);
    yds[i] = (y[b] - y[a]) / (acc_s[b] - acc_s[a]);
  }
  for (int i = 0; i < n; ++i) {
    const int a = i == 0 ? 0 : i - 1, b = i == n - 1 ? n - 1 : i + 1;
    xdds[i] = (xds[b] - xds[a]) / (acc_s[b] - acc_s[a]);
    ydds[i] = (yds[b] - yds[a]) / (acc_s[b] - acc_s[a]);
  }
  for (int i = 0; i < n; ++i)
    kappas[i] = (xds[i] * ydds[i] - yds[i] * xdds[i]) /
                (sqrt(xds[i] * xds[i] + yds[i] * yds[i]) * (xds[i] * xds[i] + yds[i] * yds[i]) + 1e-6);
  free(acc_s);
}

/* DpPlanner::Plan, dp_planner.cpp:135-281 */
int dp_plan(const dp_config* cfg, const dp_env* env, double start_x, double start_y, double start_theta,
            double* trajectory, double* min_cost_out, double* waypoints) {
  planner_t* p = (planner_t*)malloc(sizeof(planner_t));
  planner_init(p, cfg, env);
  double sl[2];
  dp_get_projection(env->R, env->ref, start_x, start_y, sl);
  p->start_s = sl[0];
  p->start_l = sl[1];
  p->start_theta = start_theta;
  for (int i = 0; i < DP_NT; ++i)
    for (int j = 0; j < DP_NS; ++j)
      for (int k = 0; k < DP_NL; ++k) {
        cell_t* c = &p->space[i][j][k];
        c->cost = DBL_MAX;
        c->current_s = DBL_MIN; /* std::numeric_limits<double>::min(), dp_planner.h:43 */
        c->parent_s_ind = -1;
        c->parent_l_ind = -1;
      }
  /* first layer, :151-158 */
  for (int i = 0; i < DP_NS; ++i)
    for (int j = 0; j < DP_NL; ++j) {
      double cur_s;
      const double c = get_cost(p, -1, -1, -1, 0, i, j, &cur_s);
      p->space[0][i][j].current_s = cur_s;
      p->space[0][i][j].cost = c;
    }
  /* dynamic programming, :160-181 */
  for (int i = 0; i < DP_NT - 1; ++i)
    for (int j = 0; j < DP_NS; ++j)
      for (int k = 0; k < DP_NL; ++k)
        for (int m = 0; m < DP_NS; ++m)
          for (int n = 0; n < DP_NL; ++n) {
            double cur_s;
            const double delta_cost = get_cost(p, i, j, k, i + 1, m, n, &cur_s);
            const double cur_cost = p->space[i][j][k].cost + delta_cost;
            if (cur_cost < p->space[i + 1][m][n].cost) {
              cell_t* c = &p->space[i + 1][m][n];
              c->cost = cur_cost;
              c->current_s = cur_s;
              c->parent_s_ind = j;
              c->parent_l_ind = k;
            }
          }
  /* least cost in the final layer, :183-194 */
  double min_cost = DBL_MAX;
  int min_s_ind = 0, min_l_ind = 0;
  for (int i = 0; i < DP_NS; ++i)
    for (int j = 0; j < DP_NL; ++j) {
      const double cost = p->space[DP_NT - 1][i][j].cost;
      if (cost < min_cost) {
        min_s_ind = i;
        min_l_ind = j;
        min_cost = cost;
      }
    }
  /* trace back, :196-204 */
  int wp_s[DP_NT], wp_l[DP_NT];
  cell_t wp_cell[DP_NT];
  for (int i = DP_NT - 1; i >= 0; --i) {
    wp_cell[i] = p->space[i][min_s_ind][min_l_ind];
    wp_s[i] = min_s_ind;
    wp_l[i] = min_l_ind;
    min_s_ind = wp_cell[i].parent_s_ind;
    min_l_ind = wp_cell[i].parent_l_ind;
  }
  if (waypoints)
    for (int i = 0; i < DP_NT; ++i) {
      waypoints[3 * i] = wp_s[i];
      waypoints[3 * i + 1] = wp_l[i];
      waypoints[3 * i + 2] = wp_cell[i].current_s;
    }
  /* interpolation, :212-243 */
  int K = 0;
  for (int c = 0; c < DP_NT; ++c) K += p->nseg[c];
  double* xs = (double*)malloc(sizeof(double) * (size_t)K * 5);
  double *ys = xs + K, *speeds = ys + K, *accel = speeds + K, *kappas = accel + K;
  memset(trajectory, 0, sizeof(double) * (size_t)K * 13);
  double last_l = p->start_l, last_s = p->start_s;
  int n = 0;
  double ps[MAX_SEG], pl[MAX_SEG];
  for (int i = 0; i < DP_NT; ++i) {
    const double parent_s = i > 0 ? wp_cell[i - 1].current_s : p->start_s;
    const int ns = interpolate_linearly(p, parent_s, wp_cell[i].parent_l_ind, i, wp_s[i], wp_l[i], ps, pl);
    for (int j = 0; j < ns; ++j) {
      const double dl = pl[j] - last_l;
      const double ds = fmax(ps[j] - last_s, DP_EPS);
      last_l = pl[j];
      last_s = ps[j];
      double x, y, r[7];
      get_cartesian(env->R, env->ref, ps[j], pl[j], &x, &y);
      dp_evaluate_station(env->R, env->ref, ps[j], r);
      double* d = trajectory + (size_t)n * 13;
      d[0] = cfg->delta_t * n; /* time */
      d[1] = ps[j];            /* s */
      d[2] = x;
      d[3] = y;
      d[4] = r[F_TH] + atan((dl / ds) / (1 - r[F_KAPPA] * pl[j]));
      xs[n] = x;
      ys[n] = y;
      ++n;
    }
  }
  /* :244-273 */
  path_profile(cfg->delta_t, K, xs, ys, speeds, accel, kappas);
  for (int i = 0; i < K; ++i) {
    double* d = trajectory + (size_t)i * 13;
    d[5] = kappas[i];
    d[9] = atan(kappas[i] * cfg->wheel_base); /* delta */
    d[6] = speeds[i];
    d[7] = accel[i];
    d[8] = 0.0;  /* jerk */
    d[10] = 0.0; /* delta_rate */
  }
  if (min_cost_out) *min_cost_out = min_cost;
  const int ok = min_cost < cfg->dp_w_obstacle;
  free(xs);
  free(p);
  return ok;
}

/* ---- pieces exported for the pin tests against the compiled reference (oracle/_ref) ------------- */
int dp_polygon_overlaps_box(const double* p, int nv, double cx, double cy, double half) {
  return polygon_overlaps_box(p, nv, cx, cy, half);
}
int dp_polygon_is_point_in(const double* p, int nv, double x, double y) {
  double minx = p[0], maxx = p[0], miny = p[1], maxy = p[1];
  for (int i = 1; i < nv; ++i) {
    minx = fmin(minx, p[2 * i]);
    maxx = fmax(maxx, p[2 * i]);
    miny = fmin(miny, p[2 * i + 1]);
    maxy = fmax(maxy, p[2 * i + 1]);
  }
  return polygon_is_point_in(p, nv, minx, maxx, miny, maxy, x, y);
}
void dp_get_cartesian(int R, const double* ref, double station, double lateral, double xy[2]) {
  get_cartesian(R, ref, station, lateral, &xy[0], &xy[1]);
}
void dp_path_profile(double dt, int n, const double* x, const double* y, double* speeds, double* accel, double* kappas) {
  path_profile(dt, n, x, y, speeds, accel, kappas);
}
double dp_slerp(double a0, double t0, double a1, double t1, double t) { return slerp(a0, t0, a1, t1, t); }
