/*
 * cilqr_oracle.c -- CPU restatement of mpt0816/Cilqr's CILQR solve (see cilqr_oracle.h).
 * TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (the reference has no tests / golden vectors and is
 * not buildable here); every function cites the reference lines it follows.  Paths are relative
 * to the reference root.
 *
 * Conventions: matrices are row-major plain arrays (A[6][6] -> A[r*6+c]); all arithmetic is IEEE
 * double, compiled without FMA contraction and without -ffast-math so that the evaluation order
 * written here is the evaluation order executed.
 *
 * Eigen semantics that matter (Eigen 3.4, un-vendored dependency, README.md:12):
 *  - fixed-size 2x2 `.inverse()` is the closed form adj/det with invdet = 1/det;
 *  - `X = <expr containing a product of X>` evaluates the right-hand side into a temporary first;
 *  - `auto q = <expr>` keeps a LAZY expression holding references: it is re-evaluated, with the
 *    current operand values, at every use.  Backward() (ilqr_optimizer.cc:348-353) declares
 *    Qx,Qu,Qxx,Quu,Qux that way and reads Qu/Quu again at :383-384 AFTER Vx (:379) and Vxx
 *    (:380-381) have been overwritten, so delta_V_ is accumulated with the *updated* value
 *    function (quirk Q21, reproduced in backward() below).
 */
#include "cilqr_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#ifdef CILQR_PM_LIBM
/* libcilqr_oracle_pm.so: the same restatement with sin / cos / tan / log / hypot taken from the portable libm that
 * the STRICT build of the CUDA kernel uses too (cilqr_b200/csrc/pm_math.h: same source on host and device, hence
 * bit-identical results on both).  Strict GPU output is compared with THIS build bit for bit; the default build
 * (glibc libm, pinned against the compiled reference) differs from it only through the last bits of those five
 * functions -- tests/test_oracle_golden.py measures how far that moves whole solves.  fmod, sqrt, fabs, fmin, fmax are
 * exact / IEEE in glibc and in CUDA and stay as they are. */
#include "../cilqr_b200/csrc/pm_math.h"
#define sin pm_sin
#define cos pm_cos
#define tan pm_tan
#define log pm_log
#define hypot pm_hypot
#endif

#define NX 6
#define NU 2

static const double kMathEpsilon = 1e-10; /* algorithm/math/vec2d.h:33 */

/* ilqr_optimizer.cc:197 */
static const double kAlphaList[CILQR_ORACLE_NALPHA] = {1.0000, 0.5012, 0.2512, 0.1259, 0.0631, 0.0316,
                                                       0.0158, 0.0079, 0.0040, 0.0020, 0.0010};

void cilqr_oracle_default_params(cilqr_oracle_params* p) {
  /* vehicle_param.h:26-64 */
  p->front_hang_length = 0.96;
  p->wheel_base = 1.0;
  p->rear_hang_length = 0.929;
  p->width = 1.942;
  p->max_velocity = 20.0;
  p->min_acceleration = -5.0;
  p->max_acceleration = 5.0;
  p->jerk_min = -10.0;
  p->jerk_max = 10.0;
  p->delta_min = -40.0 / 180 * M_PI;
  p->delta_max = 40.0 / 180 * M_PI;
  p->delta_rate_min = p->delta_min / 3.0;
  p->delta_rate_max = p->delta_max / 3.0;
  /* planner_config.h:45-66 */
  p->safe_margin = 0.2;
  p->w_jerk = 1;
  p->w_delta_rate = 1;
  p->w_x_target = 0.5;
  p->w_y_target = 0.5;
  p->w_theta = 1e-3;
  p->w_v = 0.0;
  p->w_a = 0.0;
  p->w_delta = 0.0;
  p->abs_cost_tol = 1e-2;
  p->rel_cost_tol = 1e-2;
  /* barrier_function.h:143-146 (IlqrConfig::t is never read) */
  p->barrier_t = 5.0;
  p->barrier_eps = 0.01;
  p->delta_t = 0.1; /* planner_config.h:94 */
  p->num_of_disc = 5;
  p->max_iter_num = 200;
}

/* math_utils.cpp:53-59 */
double cilqr_oracle_normalize_angle(double angle) {
  double a = fmod(angle + M_PI, 2.0 * M_PI);
  if (a < 0.0) {
    a += (2.0 * M_PI);
  }
  return a - M_PI;
}

/* ilqr_optimizer.cc:97-104 */
double cilqr_oracle_disc_radius(const cilqr_oracle_params* p) {
  double length = p->front_hang_length + p->wheel_base + p->rear_hang_length;
  return hypot(p->width / 2.0, length / 2.0 / p->num_of_disc);
}

/* vehicle_model.cc:123-138 */
static void dynamics_continuous(const cilqr_oracle_params* p, const double* s, const double* u,
                                double* res) {
  double theta = cilqr_oracle_normalize_angle(s[2]);
  double v = s[3];
  double a = s[4];
  double delta = cilqr_oracle_normalize_angle(s[5]);
  res[0] = v * cos(theta);
  res[1] = v * sin(theta);
  res[2] = v * tan(delta) / p->wheel_base;
  res[3] = a;
  res[4] = u[0];
  res[5] = u[1];
}

/* vehicle_model.cc:88-121: midpoint RK2 with the same control at both stages, then wrap theta
 * and delta.  Safe for xn aliasing x (Forward calls it in place, ilqr_optimizer.cc:409). */
void cilqr_oracle_dynamics(const cilqr_oracle_params* p, const double x[6], const double u[2],
                           double xn[6]) {
  double k1[NX], mid[NX], k2[NX];
  const double dt = p->delta_t;
  dynamics_continuous(p, x, u, k1);
  for (int i = 0; i < NX; ++i) mid[i] = x[i] + 0.5 * dt * k1[i];
  dynamics_continuous(p, mid, u, k2);
  for (int i = 0; i < NX; ++i) xn[i] = x[i] + dt * k2[i];
  xn[2] = cilqr_oracle_normalize_angle(xn[2]);
  xn[5] = cilqr_oracle_normalize_angle(xn[5]);
}

/* vehicle_model.cc:21-86.  Hand-derived Jacobian; note v (not v_mid) in A(2,5)/B(2,1). */
void cilqr_oracle_dynamics_jacobian(const cilqr_oracle_params* p, const double x[6],
                                    const double u[2], double A[36], double B[12]) {
  const double L = p->wheel_base;
  const double delta_t_ = p->delta_t;
  double v = x[3];
  double theta = cilqr_oracle_normalize_angle(x[2]);
  double delta = cilqr_oracle_normalize_angle(x[5]);
  double a = x[4];
  double delta_rate = u[1];

  double theta_mid = theta + 0.5 * delta_t_ * v * tan(delta) / L;
  double tan_delta = tan(delta);
  double tan_delta_rate = tan(delta + 0.5 * delta_t_ * delta_rate);
  double cos_theta_mid = cos(theta_mid);
  double sin_theta_mid = sin(theta_mid);
  double tan_delta_square = tan_delta * tan_delta;
  double tan_delta_rate_square = tan_delta_rate * tan_delta_rate;
  double v_tan_delta_rate = v * (tan_delta_rate_square + 1);

  memset(A, 0, 36 * sizeof(double));
  memset(B, 0, 12 * sizeof(double));
  for (int i = 0; i < NX; ++i) A[i * 6 + i] = 1.0;

  A[0 * 6 + 2] = -delta_t_ * (0.5 * a * delta_t_ + v) * sin_theta_mid;
  A[0 * 6 + 3] = delta_t_ * cos_theta_mid -
                 0.5 * delta_t_ * delta_t_ * (0.5 * a * delta_t_ + v) * sin_theta_mid * tan_delta / L;
  A[0 * 6 + 4] = 0.5 * delta_t_ * delta_t_ * cos_theta_mid;
  A[0 * 6 + 5] = -0.5 * delta_t_ * delta_t_ * v * (0.5 * a * delta_t_ + v) * (tan_delta_square + 1) *
                 sin_theta_mid / L;

  A[1 * 6 + 2] = delta_t_ * (0.5 * a * delta_t_ + v) * cos_theta_mid;
  A[1 * 6 + 3] = delta_t_ * sin_theta_mid +
                 0.5 * delta_t_ * delta_t_ * (0.5 * a * delta_t_ + v) * cos_theta_mid * tan_delta / L;
  A[1 * 6 + 4] = 0.5 * delta_t_ * delta_t_ * sin_theta_mid;
  A[1 * 6 + 5] = 0.5 * delta_t_ * delta_t_ * v * (0.5 * a * delta_t_ + v) * (tan_delta_square + 1) *
                 cos_theta_mid / L;

  A[2 * 6 + 3] = delta_t_ * tan_delta_rate / L;
  A[2 * 6 + 4] = 0.5 * delta_t_ * delta_t_ * tan_delta_rate / L;
  A[2 * 6 + 5] = delta_t_ * v_tan_delta_rate / L;

  A[3 * 6 + 4] = delta_t_;

  B[2 * 2 + 1] = 0.5 * delta_t_ * delta_t_ * v * (tan_delta_rate_square + 1) / L;
  B[3 * 2 + 0] = 0.5 * delta_t_ * delta_t_;
  B[4 * 2 + 0] = delta_t_;
  B[5 * 2 + 1] = delta_t_;
}

/* barrier_function.h:104-113 (RelaxBarrierFunction::value), t and eps per :143-146 */
double cilqr_oracle_barrier_value(const cilqr_oracle_params* p, double x) {
  const double rt = 1.0 / p->barrier_t;
  const double eps = p->barrier_eps;
  if (x < -eps) {
    return -rt * log(-x);
  } else {
    double q = (-x - 2.0 * eps) / eps;
    return 0.5 * rt * (q * q - 1) - rt * log(eps);
  }
}

/* barrier_function.h:115-125: Jacbian(x, dx) = coef * dx; this returns coef */
double cilqr_oracle_barrier_dcoef(const cilqr_oracle_params* p, double x) {
  const double rt = 1.0 / p->barrier_t;
  const double eps = p->barrier_eps;
  if (x < -eps) {
    return -rt / x;
  } else {
    return rt * (x + 2.0 * eps) / eps / eps;
  }
}

/* barrier_function.h:127-140: Hessian(x, dx, ddx) = (c_outer*dx) dx^T - c_ddx * ddx.  The relaxed
 * branch reuses the gradient coefficient and drops ddx (quirk Q6): c_ddx = 0 there. */
void cilqr_oracle_barrier_hcoef(const cilqr_oracle_params* p, double x, double* c_outer,
                                double* c_ddx) {
  const double rt = 1.0 / p->barrier_t;
  const double eps = p->barrier_eps;
  if (x < -eps) {
    *c_outer = rt / x / x;
    *c_ddx = rt / x;
  } else {
    *c_outer = rt * (x + 2.0 * eps) / eps / eps;
    *c_ddx = 0.0;
  }
}

/* Derived fields of LineSegment2d (line_segment2d.cpp:40-49). */
typedef struct {
  double sx, sy, ex, ey, ux, uy, length;
} seg_t;

static void seg_init(seg_t* s, double x0, double y0, double x1, double y1) {
  s->sx = x0;
  s->sy = y0;
  s->ex = x1;
  s->ey = y1;
  const double dx = x1 - x0;
  const double dy = y1 - y0;
  s->length = hypot(dx, dy);
  if (s->length <= kMathEpsilon) {
    s->ux = 0;
    s->uy = 0;
  } else {
    s->ux = dx / s->length;
    s->uy = dy / s->length;
  }
}

/* line_segment2d.cpp:61-75 */
static double seg_distance(const seg_t* s, double px, double py) {
  if (s->length <= kMathEpsilon) {
    return hypot(px - s->sx, py - s->sy);
  }
  const double x0 = px - s->sx;
  const double y0 = py - s->sy;
  const double proj = x0 * s->ux + y0 * s->uy;
  if (proj <= 0.0) {
    return hypot(x0, y0);
  }
  if (proj >= s->length) {
    return hypot(px - s->ex, py - s->ey);
  }
  return fabs(x0 * s->uy - y0 * s->ux);
}

double cilqr_oracle_segment_distance(const double seg[7], double x, double y) {
  seg_t s;
  seg_init(&s, seg[3], seg[4], seg[5], seg[6]);
  return seg_distance(&s, x, y);
}

/* ------------------------------------------------------------------------------------------ */
/* Context = the members of IlqrOptimizer (ilqr_optimizer.h:170-214) for one Plan call.        */
struct cilqr_oracle_ctx {
  cilqr_oracle_params p;
  int N, K, M_max, S[2];
  double disc_radius;
  double* goals;     /* [K][6]                       goals_                 */
  double* corridor;  /* [K][M_max][3] shrunk+normalised  shrinked_corridor_ */
  int* cnt;          /* [K]                                                */
  double* lane_abc[2]; /* [S][3]  shrunk+normalised half-planes            */
  seg_t* lane_seg[2];  /* [S]                                              */
  double *As, *Bs, *Jx, *Ju, *Hx, *Hu; /* As Bs cost_Jx cost_Ju cost_Hx cost_Hu */
  double *Ks, *ks;   /* gains                                              */
  double dV[2];      /* delta_V_                                           */
};

/* ilqr_optimizer.cc:438-473 then :475-495 for one half-plane; `shrink` = r (+ margin) */
static void shrink_normalize(const double in[3], double shrink, double out[3]) {
  double e0 = in[0], e1 = in[1], e2 = in[2];
  e2 = e2 - shrink * (e0 * e0 + e1 * e1) / hypot(e0, e1);
  double norm = hypot(hypot(e0, e1), e2);
  out[0] = e0 / norm;
  out[1] = e1 / norm;
  out[2] = e2 / norm;
}

cilqr_oracle_ctx* cilqr_oracle_ctx_create(const cilqr_oracle_params* p,
                                          const cilqr_oracle_problem* pb) {
  cilqr_oracle_ctx* c = (cilqr_oracle_ctx*)calloc(1, sizeof(*c));
  c->p = *p;
  c->N = pb->N;
  c->K = pb->N + 1;
  c->M_max = pb->M_max;
  c->S[0] = pb->S_left;
  c->S[1] = pb->S_right;
  const int K = c->K, N = c->N;
  c->disc_radius = cilqr_oracle_disc_radius(p);

  /* TransformGoals, ilqr_optimizer.cc:141-152 */
  c->goals = (double*)malloc(sizeof(double) * K * NX);
  memcpy(c->goals, pb->coarse, sizeof(double) * K * NX);
  c->goals[0] = pb->start[0];
  c->goals[1] = pb->start[1];
  c->goals[2] = pb->start[2];
  c->goals[3] = pb->start[3];
  c->goals[4] = 0.0;
  c->goals[5] = 0.0;

  /* ShrinkConstraints + NormalizeHalfPlane, ilqr_optimizer.cc:163-164 */
  c->corridor = (double*)calloc((size_t)K * c->M_max * 3, sizeof(double));
  c->cnt = (int*)malloc(sizeof(int) * K);
  for (int k = 0; k < K; ++k) {
    c->cnt[k] = pb->corridor_cnt[k];
    for (int m = 0; m < c->cnt[k]; ++m) {
      size_t o = ((size_t)k * c->M_max + m) * 3;
      shrink_normalize(pb->corridor + o, c->disc_radius + p->safe_margin, c->corridor + o);
    }
  }
  for (int side = 0; side < 2; ++side) {
    const double* lane = side == 0 ? pb->lane_left : pb->lane_right;
    const int S = c->S[side];
    c->lane_abc[side] = (double*)malloc(sizeof(double) * 3 * (S > 0 ? S : 1));
    c->lane_seg[side] = (seg_t*)malloc(sizeof(seg_t) * (S > 0 ? S : 1));
    for (int s = 0; s < S; ++s) {
      shrink_normalize(lane + (size_t)s * 7, c->disc_radius, c->lane_abc[side] + (size_t)s * 3);
      seg_init(&c->lane_seg[side][s], lane[s * 7 + 3], lane[s * 7 + 4], lane[s * 7 + 5],
               lane[s * 7 + 6]);
    }
  }
  c->As = (double*)calloc((size_t)N * 36, sizeof(double));
  c->Bs = (double*)calloc((size_t)N * 12, sizeof(double));
  c->Jx = (double*)calloc((size_t)K * 6, sizeof(double));
  c->Ju = (double*)calloc((size_t)N * 2, sizeof(double));
  c->Hx = (double*)calloc((size_t)K * 36, sizeof(double));
  c->Hu = (double*)calloc((size_t)N * 4, sizeof(double));
  c->Ks = (double*)calloc((size_t)N * 12, sizeof(double));
  c->ks = (double*)calloc((size_t)N * 2, sizeof(double));
  return c;
}

void cilqr_oracle_ctx_destroy(cilqr_oracle_ctx* c) {
  if (!c) return;
  free(c->goals);
  free(c->corridor);
  free(c->cnt);
  for (int s = 0; s < 2; ++s) {
    free(c->lane_abc[s]);
    free(c->lane_seg[s]);
  }
  free(c->As);
  free(c->Bs);
  free(c->Jx);
  free(c->Ju);
  free(c->Hx);
  free(c->Hu);
  free(c->Ks);
  free(c->ks);
  free(c);
}

void cilqr_oracle_ctx_constraints(const cilqr_oracle_ctx* c, double* corridor, double* lane_left,
                                  double* lane_right) {
  if (corridor) memcpy(corridor, c->corridor, sizeof(double) * (size_t)c->K * c->M_max * 3);
  if (lane_left) memcpy(lane_left, c->lane_abc[0], sizeof(double) * 3 * c->S[0]);
  if (lane_right) memcpy(lane_right, c->lane_abc[1], sizeof(double) * 3 * c->S[1]);
}

/* ilqr_optimizer.cc:605-618: strict '<', first minimum wins. */
int cilqr_oracle_ctx_nearest(const cilqr_oracle_ctx* c, int side, double x, double y) {
  double min_dis = 1.7976931348623157e308; /* numeric_limits<double>::max() */
  int min_index = -1;
  const seg_t* segs = c->lane_seg[side];
  for (int i = 0; i < c->S[side]; ++i) {
    double dis = seg_distance(&segs[i], x, y);
    if (dis < min_dis) {
      min_dis = dis;
      min_index = i;
    }
  }
  return min_index;
}

/* disc offset along the heading, ilqr_optimizer.cc:556-557,564 (quirk Q7) */
static inline double disc_offset(const cilqr_oracle_params* p, int j) {
  double L = (p->rear_hang_length + p->wheel_base + p->front_hang_length) / p->num_of_disc;
  double rf = p->rear_hang_length;
  return (L * (j - 0.5) - rf);
}

/* ilqr_optimizer.cc:497-516 */
static double j_cost(const cilqr_oracle_ctx* c, const double* X, const double* U) {
  const cilqr_oracle_params* p = &c->p;
  double cost = 0.0;
  for (int i = 0; i < c->K; ++i) {
    double dx = X[i * 6 + 0] - c->goals[i * 6 + 0];
    double dy = X[i * 6 + 1] - c->goals[i * 6 + 1];
    double dth = X[i * 6 + 2] - c->goals[i * 6 + 2];
    cost += p->w_x_target * (dx * dx) + p->w_y_target * (dy * dy) + p->w_theta * (dth * dth);
  }
  for (int i = 0; i < c->N; ++i) {
    cost += p->w_jerk * (U[i * 2 + 0] * U[i * 2 + 0]) + p->w_delta_rate * (U[i * 2 + 1] * U[i * 2 + 1]);
  }
  return cost;
}

/* ilqr_optimizer.cc:518-551 */
static double dynamics_cost(const cilqr_oracle_ctx* c, const double* X, const double* U) {
  const cilqr_oracle_params* p = &c->p;
  double x_cost = 0.0;
  for (int i = 0; i < c->K; ++i) {
    const double* s = X + i * 6;
    x_cost += cilqr_oracle_barrier_value(p, -s[3]);
    x_cost += cilqr_oracle_barrier_value(p, s[3] - p->max_velocity);
    x_cost += cilqr_oracle_barrier_value(p, s[4] - p->max_acceleration);
    x_cost += cilqr_oracle_barrier_value(p, p->min_acceleration - s[4]);
    x_cost += cilqr_oracle_barrier_value(p, s[5] - p->delta_max);
    x_cost += cilqr_oracle_barrier_value(p, p->delta_min - s[5]);
  }
  double u_cost = 0.0;
  for (int i = 0; i < c->N; ++i) {
    const double* u = U + i * 2;
    u_cost += cilqr_oracle_barrier_value(p, u[0] - p->jerk_max);
    u_cost += cilqr_oracle_barrier_value(p, p->jerk_min - u[0]);
    u_cost += cilqr_oracle_barrier_value(p, u[1] - p->delta_rate_max);
    u_cost += cilqr_oracle_barrier_value(p, p->delta_rate_min - u[1]);
  }
  return x_cost + u_cost;
}

/* ilqr_optimizer.cc:553-581 */
static double corridor_cost(const cilqr_oracle_ctx* c, const double* X) {
  const cilqr_oracle_params* p = &c->p;
  double cost = 0.0;
  for (int i = 0; i < c->K; ++i) {
    const double* cons = c->corridor + (size_t)i * c->M_max * 3;
    for (int j = 0; j < p->num_of_disc; ++j) {
      double x = X[i * 6 + 0] + disc_offset(p, j) * cos(X[i * 6 + 2]);
      double y = X[i * 6 + 1] + disc_offset(p, j) * sin(X[i * 6 + 2]);
      for (int m = 0; m < c->cnt[i]; ++m) {
        const double* h = cons + m * 3;
        cost += cilqr_oracle_barrier_value(p, h[0] * x + h[1] * y - h[2]);
      }
    }
  }
  return cost;
}

/* ilqr_optimizer.cc:583-603 */
static double lane_boundary_cost(const cilqr_oracle_ctx* c, const double* X) {
  const cilqr_oracle_params* p = &c->p;
  double cost = 0.0;
  for (int i = 0; i < c->K; ++i) {
    for (int j = 0; j < p->num_of_disc; ++j) {
      double x = X[i * 6 + 0] + disc_offset(p, j) * cos(X[i * 6 + 2]);
      double y = X[i * 6 + 1] + disc_offset(p, j) * sin(X[i * 6 + 2]);
      for (int side = 0; side < 2; ++side) { /* left, then right */
        const double* h = c->lane_abc[side] + 3 * cilqr_oracle_ctx_nearest(c, side, x, y);
        cost += cilqr_oracle_barrier_value(p, h[0] * x + h[1] * y - h[2]);
      }
    }
  }
  return cost;
}

/* ilqr_optimizer.cc:417-436 */
double cilqr_oracle_ctx_total_cost(cilqr_oracle_ctx* c, const double* X, const double* U,
                                   double cost5[5]) {
  double j = j_cost(c, X, U);
  double d = dynamics_cost(c, X, U);
  double co = corridor_cost(c, X);
  double la = lane_boundary_cost(c, X);
  double total = j + d + co + la;
  if (cost5) {
    cost5[0] = total;
    cost5[1] = j;
    cost5[2] = d;
    cost5[3] = co;
    cost5[4] = la;
  }
  return total;
}

/* One half-plane barrier term acting on a disc: accumulates into Jx (if non-NULL) and the upper
 * left 3x3 block of Hx (if non-NULL).  dx = (a, b, -a*ls + b*lc, 0,0,0); ddx(2,2) = -a*lc - b*ls.
 * ilqr_optimizer.cc:703,723-724 and :741,744,762-767.  Entries outside the 3x3 block receive
 * exact zeros in the reference and are skipped here. */
static inline void plane_term(const cilqr_oracle_params* p, const double* h, double x, double y,
                              double lc, double ls, double* Jx, double* Hx) {
  const double g = h[0] * x + h[1] * y - h[2];
  const double d[3] = {h[0], h[1], -h[0] * ls + h[1] * lc};
  if (Jx) {
    const double cj = cilqr_oracle_barrier_dcoef(p, g);
    for (int r = 0; r < 3; ++r) Jx[r] += cj * d[r];
  }
  if (Hx) {
    double co, cd;
    cilqr_oracle_barrier_hcoef(p, g, &co, &cd);
    const double ddx22 = -h[0] * lc - h[1] * ls;
    for (int r = 0; r < 3; ++r) {
      const double cr = co * d[r];
      for (int q = 0; q < 3; ++q) {
        double v = cr * d[q];
        if (g < -p->barrier_eps) v = v - cd * ((r == 2 && q == 2) ? ddx22 : 0.0);
        Hx[r * 6 + q] += v;
      }
    }
  }
}

/* CostJacbian (ilqr_optimizer.cc:620-636) and CostHessian (:638-655) at one knot. */
static void cost_derivatives(const cilqr_oracle_ctx* c, int index, const double* s,
                             const double* u, double* Jx, double* Ju, double* Hx, double* Hu) {
  const cilqr_oracle_params* p = &c->p;
  const double* g = c->goals + index * 6;
  /* running cost */
  Jx[0] = 2.0 * p->w_x_target * (s[0] - g[0]);
  Jx[1] = 2.0 * p->w_y_target * (s[1] - g[1]);
  Jx[2] = 2.0 * p->w_theta * (s[2] - g[2]);
  Jx[3] = Jx[4] = Jx[5] = 0.0;
  Ju[0] = 2.0 * p->w_jerk * u[0];
  Ju[1] = 2.0 * p->w_delta_rate * u[1];
  memset(Hx, 0, 36 * sizeof(double));
  Hx[0] = 2.0 * p->w_x_target;
  Hx[7] = 2.0 * p->w_y_target;
  Hx[14] = 2.0 * p->w_theta;
  Hx[21] = 2.0 * p->w_v;
  Hx[28] = 2.0 * p->w_a;
  Hx[35] = 2.0 * p->w_delta;
  Hu[0] = 2.0 * p->w_jerk;
  Hu[1] = Hu[2] = 0.0;
  Hu[3] = 2.0 * p->w_delta_rate;

  /* DynamicsConsJacbian :657-671 / DynamicsConsHessian :673-688.  The reference sums the six
   * (four) vector terms first and then adds the sum; each pair below is that sum restricted to
   * the one component where it is non-zero. */
  {
    const double gx[6] = {0.0 - s[3],
                          s[3] - p->max_velocity,
                          p->min_acceleration - s[4],
                          s[4] - p->max_acceleration,
                          p->delta_min - s[5],
                          s[5] - p->delta_max};
    for (int q = 0; q < 3; ++q) {
      double c0 = cilqr_oracle_barrier_dcoef(p, gx[2 * q]);
      double c1 = cilqr_oracle_barrier_dcoef(p, gx[2 * q + 1]);
      Jx[3 + q] += c0 * -1.0 + c1 * 1.0;
      double h0, h1, unused;
      cilqr_oracle_barrier_hcoef(p, gx[2 * q], &h0, &unused);
      cilqr_oracle_barrier_hcoef(p, gx[2 * q + 1], &h1, &unused);
      Hx[(3 + q) * 6 + (3 + q)] += (h0 * -1.0) * -1.0 + (h1 * 1.0) * 1.0;
    }
    const double gu[4] = {p->jerk_min - u[0], u[0] - p->jerk_max, p->delta_rate_min - u[1],
                          u[1] - p->delta_rate_max};
    for (int q = 0; q < 2; ++q) {
      double c0 = cilqr_oracle_barrier_dcoef(p, gu[2 * q]);
      double c1 = cilqr_oracle_barrier_dcoef(p, gu[2 * q + 1]);
      Ju[q] += c0 * -1.0 + c1 * 1.0;
      double h0, h1, unused;
      cilqr_oracle_barrier_hcoef(p, gu[2 * q], &h0, &unused);
      cilqr_oracle_barrier_hcoef(p, gu[2 * q + 1], &h1, &unused);
      Hu[q * 2 + q] += (h0 * -1.0) * -1.0 + (h1 * 1.0) * 1.0;
    }
  }

  /* CorridorConsJacbian :690-706, LaneBoundaryConsJacbian :729-746 (all corridor terms of all
   * discs first, then all lane terms), same for the Hessians :708-727, :748-769.  Jacobian and
   * Hessian accumulate into different outputs, so interleaving the two sweeps is order-neutral. */
  const double* cons = c->corridor + (size_t)index * c->M_max * 3;
  for (int i = 0; i < p->num_of_disc; ++i) {
    double lc = disc_offset(p, i) * cos(s[2]);
    double ls = disc_offset(p, i) * sin(s[2]);
    double x = s[0] + lc;
    double y = s[1] + ls;
    for (int m = 0; m < c->cnt[index]; ++m) plane_term(p, cons + m * 3, x, y, lc, ls, Jx, Hx);
  }
  for (int i = 0; i < p->num_of_disc; ++i) {
    double lc = disc_offset(p, i) * cos(s[2]);
    double ls = disc_offset(p, i) * sin(s[2]);
    double x = s[0] + lc;
    double y = s[1] + ls;
    for (int side = 0; side < 2; ++side) {
      const double* h = c->lane_abc[side] + 3 * cilqr_oracle_ctx_nearest(c, side, x, y);
      plane_term(p, h, x, y, lc, ls, Jx, Hx);
    }
  }
}

/* ilqr_optimizer.cc:203-214 */
void cilqr_oracle_ctx_linearize(cilqr_oracle_ctx* c, const double* X, const double* U, double* As,
                                double* Bs, double* Jx, double* Ju, double* Hx, double* Hu) {
  const int N = c->N;
  for (int i = 0; i < N; ++i) {
    cilqr_oracle_dynamics_jacobian(&c->p, X + i * 6, U + i * 2, c->As + i * 36, c->Bs + i * 12);
    cost_derivatives(c, i, X + i * 6, U + i * 2, c->Jx + i * 6, c->Ju + i * 2, c->Hx + i * 36,
                     c->Hu + i * 4);
  }
  double zero_u[2] = {0.0, 0.0}, tmpJu[2], tmpHu[4];
  cost_derivatives(c, N, X + N * 6, zero_u, c->Jx + N * 6, tmpJu, c->Hx + N * 36, tmpHu);
  if (As) memcpy(As, c->As, sizeof(double) * N * 36);
  if (Bs) memcpy(Bs, c->Bs, sizeof(double) * N * 12);
  if (Jx) memcpy(Jx, c->Jx, sizeof(double) * (N + 1) * 6);
  if (Ju) memcpy(Ju, c->Ju, sizeof(double) * N * 2);
  if (Hx) memcpy(Hx, c->Hx, sizeof(double) * (N + 1) * 36);
  if (Hu) memcpy(Hu, c->Hu, sizeof(double) * N * 4);
}

/* small dense helpers, natural (row, inner) summation order */
static void mat_mul(const double* A, const double* B, double* C, int n, int m, int q) {
  /* C[n][q] = A[n][m] * B[m][q] */
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < q; ++j) {
      double s = A[i * m + 0] * B[0 * q + j];
      for (int k = 1; k < m; ++k) s += A[i * m + k] * B[k * q + j];
      C[i * q + j] = s;
    }
}
static void mat_T(const double* A, double* At, int n, int m) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) At[j * n + i] = A[i * m + j];
}

/* ilqr_optimizer.cc:334-390.  See the header comment for the lazy-expression quirk (Q21). */
void cilqr_oracle_ctx_backward(cilqr_oracle_ctx* c, double lambda, double* Ks_out, double* ks_out,
                               double dV_out[2]) {
  const int N = c->N;
  double Vx[6], Vxx[36];
  c->dV[0] = 0.0;
  c->dV[1] = 0.0;
  memcpy(Vx, c->Jx + N * 6, sizeof(Vx));
  memcpy(Vxx, c->Hx + N * 36, sizeof(Vxx));
  for (int i = N - 1; i >= 0; --i) {
    const double* A = c->As + i * 36;
    const double* B = c->Bs + i * 12;
    const double* Jx = c->Jx + i * 6;
    const double* Ju = c->Ju + i * 2;
    const double* Hx = c->Hx + i * 36;
    const double* Hu = c->Hu + i * 4;
    double At[36], Bt[12];
    mat_T(A, At, 6, 6);
    mat_T(B, Bt, 6, 2);

    double tmp6[6], tmp2[2], Qx[6], Qu[2];
    mat_mul(At, Vx, tmp6, 6, 6, 1);
    for (int r = 0; r < 6; ++r) Qx[r] = Jx[r] + tmp6[r];
    mat_mul(Bt, Vx, tmp2, 2, 6, 1);
    for (int r = 0; r < 2; ++r) Qu[r] = Ju[r] + tmp2[r];

    double AtV[36], BtV[12], t36[36], t4[4], Qxx[36], Quu[4], Qux[12];
    mat_mul(At, Vxx, AtV, 6, 6, 6);
    mat_mul(AtV, A, t36, 6, 6, 6);
    for (int r = 0; r < 36; ++r) Qxx[r] = Hx[r] + t36[r];
    mat_mul(Bt, Vxx, BtV, 2, 6, 6);
    mat_mul(BtV, B, t4, 2, 6, 2);
    for (int r = 0; r < 4; ++r) Quu[r] = Hu[r] + t4[r];
    mat_mul(BtV, A, Qux, 2, 6, 6);

    /* Quu_tem = Quu + lambda*I ; closed-form 2x2 inverse (Eigen compute_inverse<.,.,2>) */
    double T[4] = {Quu[0] + lambda * 1.0, Quu[1] + lambda * 0.0, Quu[2] + lambda * 0.0,
                   Quu[3] + lambda * 1.0};
    double det = T[0] * T[3] - T[2] * T[1];
    double invdet = 1.0 / det;
    double inv[4] = {T[3] * invdet, -T[1] * invdet, -T[2] * invdet, T[0] * invdet};
    double ninv[4] = {-inv[0], -inv[1], -inv[2], -inv[3]};
    double* Kg = c->Ks + i * 12;
    double* kg = c->ks + i * 2;
    mat_mul(ninv, Qux, Kg, 2, 2, 6);
    mat_mul(ninv, Qu, kg, 2, 2, 1);

    /* Vx = Qx + K^T Quu k + K^T Qu + Qux^T k          (:379, un-regularised Quu, quirk Q3) */
    double Kt[12], KtQuu[12], QuxT[12], a6[6], b6[6], c6[6], Vx_new[6];
    mat_T(Kg, Kt, 2, 6);
    mat_T(Qux, QuxT, 2, 6);
    mat_mul(Kt, Quu, KtQuu, 6, 2, 2);
    mat_mul(KtQuu, kg, a6, 6, 2, 1);
    mat_mul(Kt, Qu, b6, 6, 2, 1);
    mat_mul(QuxT, kg, c6, 6, 2, 1);
    for (int r = 0; r < 6; ++r) Vx_new[r] = Qx[r] + a6[r] + b6[r] + c6[r];
    /* Vxx = Qxx + K^T Quu K + K^T Qux + Qux^T K ; symmetrise                     (:380-381) */
    double a36[36], b36[36], c36[36], Vxx_new[36];
    mat_mul(KtQuu, Kg, a36, 6, 2, 6);
    mat_mul(Kt, Qux, b36, 6, 2, 6);
    mat_mul(QuxT, Kg, c36, 6, 2, 6);
    for (int r = 0; r < 36; ++r) Vxx_new[r] = Qxx[r] + a36[r] + b36[r] + c36[r];
    memcpy(Vx, Vx_new, sizeof(Vx));
    /* `Vxx = 0.5 * (Vxx + Vxx.transpose())` (:381) has no product in it, so Eigen assigns it coefficient by
     * coefficient, column-major, WITHOUT a temporary: the upper triangle reads lower-triangle entries that have
     * already been overwritten (quirk Q22).  The result differs from the exact mean only by the rounding-level
     * asymmetry of Vxx, but it is what the reference computes. */
    memcpy(Vxx, Vxx_new, sizeof(double) * 36);
    for (int q = 0; q < 6; ++q)
      for (int r = 0; r < 6; ++r) Vxx[r * 6 + q] = 0.5 * (Vxx[r * 6 + q] + Vxx[q * 6 + r]);

    /* :383-384 -- Qu and Quu are lazy expressions, re-evaluated here with the UPDATED Vx / Vxx */
    double Qu_new[2], Quu_new[4];
    mat_mul(Bt, Vx, tmp2, 2, 6, 1);
    for (int r = 0; r < 2; ++r) Qu_new[r] = Ju[r] + tmp2[r];
    mat_mul(Bt, Vxx, BtV, 2, 6, 6);
    mat_mul(BtV, B, t4, 2, 6, 2);
    for (int r = 0; r < 4; ++r) Quu_new[r] = Hu[r] + t4[r];
    c->dV[0] += kg[0] * Qu_new[0] + kg[1] * Qu_new[1];
    double hk[2] = {0.5 * kg[0], 0.5 * kg[1]};
    double hkQ[2] = {hk[0] * Quu_new[0] + hk[1] * Quu_new[2], hk[0] * Quu_new[1] + hk[1] * Quu_new[3]};
    c->dV[1] += hkQ[0] * kg[0] + hkQ[1] * kg[1];
  }
  if (Ks_out) memcpy(Ks_out, c->Ks, sizeof(double) * N * 12);
  if (ks_out) memcpy(ks_out, c->ks, sizeof(double) * N * 2);
  if (dV_out) {
    dV_out[0] = c->dV[0];
    dV_out[1] = c->dV[1];
  }
}

/* ilqr_optimizer.cc:392-415.  Restarts from goals_.front() (quirk Q15); wraps the delta-rate
 * control (quirk Q9). */
void cilqr_oracle_ctx_forward(cilqr_oracle_ctx* c, double alpha, const double* X, const double* U,
                              double* Xn, double* Un) {
  double x[6];
  memcpy(x, c->goals, sizeof(x));
  memcpy(Xn, x, sizeof(x));
  for (int i = 0; i < c->N; ++i) {
    const double* Kg = c->Ks + i * 12;
    const double* kg = c->ks + i * 2;
    double dx[6];
    for (int r = 0; r < 6; ++r) dx[r] = x[r] - X[i * 6 + r];
    for (int r = 0; r < 2; ++r) {
      double s = Kg[r * 6 + 0] * dx[0];
      for (int q = 1; q < 6; ++q) s += Kg[r * 6 + q] * dx[q];
      Un[i * 2 + r] = U[i * 2 + r] + s + alpha * kg[r];
    }
    Un[i * 2 + 1] = cilqr_oracle_normalize_angle(Un[i * 2 + 1]);
    cilqr_oracle_dynamics(&c->p, x, Un + i * 2, x);
    memcpy(Xn + (i + 1) * 6, x, sizeof(x));
  }
}

/* ilqr_optimizer.cc:793-842: time-varying LQR about the goals, clamped controls, RK2 rollout. */
void cilqr_oracle_ctx_iqr(cilqr_oracle_ctx* c, double* X, double* U) {
  const cilqr_oracle_params* p = &c->p;
  const int N = c->N;
  double* Ks = (double*)malloc(sizeof(double) * N * 12);
  double Q[36] = {0}, R[4] = {0}, P[36];
  Q[0] = 0.001;
  Q[7] = 0.001;
  Q[14] = 0.001;
  Q[21] = 0.001;
  Q[28] = 0.01;
  Q[35] = 0.005;
  R[0] = 0.2;
  R[3] = 0.05; /* off-diagonals: indeterminate in the reference, 0 here (documented deviation) */
  memcpy(P, Q, sizeof(P));
  const double zero_u[2] = {0.0, 0.0};
  for (int i = N - 1; i >= 0; --i) {
    double A[36], B[12], Bt[12], At[36];
    cilqr_oracle_dynamics_jacobian(p, c->goals + i * 6, zero_u, A, B);
    mat_T(B, Bt, 6, 2);
    mat_T(A, At, 6, 6);
    double BtP[12], S[4], G[12];
    mat_mul(Bt, P, BtP, 2, 6, 6);
    mat_mul(BtP, B, S, 2, 6, 2);
    for (int r = 0; r < 4; ++r) S[r] = R[r] + S[r];
    mat_mul(BtP, A, G, 2, 6, 6);
    double det = S[0] * S[3] - S[2] * S[1];
    double invdet = 1.0 / det;
    double inv[4] = {S[3] * invdet, -S[1] * invdet, -S[2] * invdet, S[0] * invdet};
    mat_mul(inv, G, Ks + i * 12, 2, 2, 6);
    /* P = Q + A^T P (A - B K) */
    double BK[36], AmBK[36], AtP[36], T[36];
    mat_mul(B, Ks + i * 12, BK, 6, 2, 6);
    for (int r = 0; r < 36; ++r) AmBK[r] = A[r] - BK[r];
    mat_mul(At, P, AtP, 6, 6, 6);
    mat_mul(AtP, AmBK, T, 6, 6, 6);
    for (int r = 0; r < 36; ++r) P[r] = Q[r] + T[r];
  }
  double x[6];
  memcpy(x, c->goals, sizeof(x));
  memcpy(X, x, sizeof(x));
  for (int i = 0; i < N; ++i) {
    const double* Kg = Ks + i * 12;
    double dx[6];
    for (int r = 0; r < 6; ++r) dx[r] = x[r] - c->goals[i * 6 + r];
    for (int r = 0; r < 2; ++r) {
      double s = -Kg[r * 6 + 0] * dx[0];
      for (int q = 1; q < 6; ++q) s += -Kg[r * 6 + q] * dx[q];
      U[i * 2 + r] = s;
    }
    U[i * 2 + 0] = fmin(p->jerk_max, fmax(U[i * 2 + 0], p->jerk_min));
    U[i * 2 + 1] = fmin(p->delta_rate_max, fmax(U[i * 2 + 1], p->delta_rate_min));
    cilqr_oracle_dynamics(p, x, U + i * 2, X + (i + 1) * 6);
    memcpy(x, X + (i + 1) * 6, sizeof(x));
  }
  free(Ks);
}

/* ilqr_optimizer.cc:322-332 */
static double gradient_norm(const cilqr_oracle_ctx* c, const double* U) {
  double acc = 0.0;
  for (int i = 0; i < c->N; ++i) {
    double v0 = fabs(c->ks[i * 2 + 0]) / (fabs(U[i * 2 + 0]) + 1);
    double v1 = fabs(c->ks[i * 2 + 1]) / (fabs(U[i * 2 + 1]) + 1);
    acc += (v0 > v1 ? v0 : v1); /* maxCoeff */
  }
  return acc / c->N;
}

static unsigned int fnv1a(unsigned int h, unsigned int byte) {
  return (h ^ (byte & 0xffu)) * 16777619u;
}

static void push_cost(cilqr_oracle_result* out, const double c5[5]) {
  if (out->cost_hist && out->cost_hist_len < out->cost_hist_cap) {
    memcpy(out->cost_hist + 5 * out->cost_hist_len, c5, 5 * sizeof(double));
  }
  out->cost_hist_len++;
}

/* Plan (ilqr_optimizer.cc:53-95) + Optimize (:154-320). */
int cilqr_oracle_solve(const cilqr_oracle_params* p, const cilqr_oracle_problem* pb,
                       cilqr_oracle_result* out) {
  /* guards :64-78 (knot-count mismatch cannot be expressed in this wire format) */
  if (!out || !out->states || !out->controls) return -1;
  if (pb->N < 1 || pb->S_left == 0 || pb->S_right == 0) return -1;

  cilqr_oracle_ctx* c = cilqr_oracle_ctx_create(p, pb);
  const int N = c->N, K = c->K;
  double* X = (double*)malloc(sizeof(double) * K * 6);
  double* U = (double*)malloc(sizeof(double) * N * 2);
  double* Xold = (double*)malloc(sizeof(double) * K * 6);
  double* Uold = (double*)malloc(sizeof(double) * N * 2);

  out->trace_len = 0;
  out->cost_hist_len = 0;
  out->accepted = 0;
  out->alpha_hash = 2166136261u;

  if (pb->init_mode == 2 && pb->init_states && pb->init_controls) { /* InitGuess, :168 (commented out there) */
    memcpy(X, pb->init_states, sizeof(double) * K * 6);
    memcpy(U, pb->init_controls, sizeof(double) * N * 2);
  } else if (pb->init_mode == 1 && pb->init_controls) { /* OpenLoopRollout, slover/ilqr.h:362-370 */
    memcpy(U, pb->init_controls, sizeof(double) * N * 2);
    memcpy(X, c->goals, sizeof(double) * 6);
    for (int i = 0; i < N; ++i) cilqr_oracle_dynamics(p, X + i * 6, U + i * 2, X + (i + 1) * 6);
  } else {
    cilqr_oracle_ctx_iqr(c, X, U); /* :169 */
  }
  if (out->init_states) memcpy(out->init_states, X, sizeof(double) * K * 6);
  if (out->init_controls) memcpy(out->init_controls, U, sizeof(double) * N * 2);

  double cost_data[5], cost_acc[5];
  double cost_old = cilqr_oracle_ctx_total_cost(c, X, U, cost_data); /* :172 */
  memcpy(cost_acc, cost_data, sizeof(cost_acc));
  memcpy(out->cost_init, cost_data, sizeof(cost_acc));
  push_cost(out, cost_data);

  int is_forward_pass_updated = 1;
  double dcost = 0.0, lambda = 1.0, dlambda = 1.0, z = 0.0, cost_new = 0.0;
  const double regularization_ratio = 1.6, regularization_min = 1e-8, regularization_max = 1e11;
  const double gradient_norm_min = 1e-6, beta_min = 1e-4, beta_max = 10.0;
  int status = CILQR_ORACLE_MAX_ITER;
  int iter = 0;
  for (; iter < p->max_iter_num; ++iter) {
    if (is_forward_pass_updated) { /* :203-214 */
      cilqr_oracle_ctx_linearize(c, X, U, 0, 0, 0, 0, 0, 0);
      is_forward_pass_updated = 0;
    }
    /* :216-233 -- Backward always reports "not diverged" (quirk Q2) */
    cilqr_oracle_ctx_backward(c, lambda, 0, 0, 0);

    double gnorm = gradient_norm(c, U); /* :235-241 */
    if (gnorm < gradient_norm_min && lambda < 1e-5) {
      status = CILQR_ORACLE_CONVERGED_GRAD;
      break;
    }

    int is_forward_pass_done = 0;
    int alpha_idx = CILQR_ORACLE_NALPHA;
    const double lambda_used = lambda;
    for (int i = 0; i < CILQR_ORACLE_NALPHA; ++i) { /* :246-265 */
      memcpy(Xold, X, sizeof(double) * K * 6);
      memcpy(Uold, U, sizeof(double) * N * 2);
      double alpha = kAlphaList[i];
      cilqr_oracle_ctx_forward(c, alpha, Xold, Uold, X, U);
      cost_new = cilqr_oracle_ctx_total_cost(c, X, U, cost_data);
      dcost = cost_old - cost_new;
      double expected = -alpha * (c->dV[0] + alpha * c->dV[1]);
      z = dcost / expected;
      if ((z > beta_min && z < beta_max) && dcost > 0.0) {
        is_forward_pass_done = 1;
        alpha_idx = i;
        break;
      }
      memcpy(X, Xold, sizeof(double) * K * 6);
      memcpy(U, Uold, sizeof(double) * N * 2);
    }
    out->alpha_hash = fnv1a(out->alpha_hash, (unsigned int)alpha_idx);
    if (out->trace && out->trace_len < out->trace_cap) {
      double* t = out->trace + 8 * out->trace_len;
      t[0] = iter;
      t[1] = alpha_idx;
      t[2] = cost_new;
      t[3] = dcost;
      t[4] = z;
      t[5] = lambda_used;
      t[6] = c->dV[0];
      t[7] = c->dV[1];
    }
    out->trace_len++;

    if (is_forward_pass_done) { /* :272-296 */
      dlambda = fmin(dlambda / regularization_ratio, 1.0 / regularization_ratio);
      lambda = lambda * dlambda * (lambda > regularization_min);
      is_forward_pass_updated = 1;
      out->accepted++;
      memcpy(cost_acc, cost_data, sizeof(cost_acc));
      if (dcost < p->abs_cost_tol || dcost / cost_old < p->rel_cost_tol) {
        push_cost(out, cost_data);
        status = dcost < p->abs_cost_tol ? CILQR_ORACLE_CONVERGED_ABS : CILQR_ORACLE_CONVERGED_REL;
        cost_old = cost_new;
        break;
      }
      cost_old = cost_new;
      push_cost(out, cost_data);
    } else { /* :297-308 */
      dlambda = fmax(dlambda * regularization_ratio, regularization_ratio);
      lambda = fmax(lambda * dlambda, regularization_min);
      if (lambda > regularization_max) {
        status = CILQR_ORACLE_LAMBDA_OVERFLOW;
        break;
      }
    }
  }
  out->status = status;
  out->iters = iter;
  out->lambda = lambda;
  memcpy(out->cost, cost_acc, sizeof(cost_acc));
  memcpy(out->states, X, sizeof(double) * K * 6);
  memcpy(out->controls, U, sizeof(double) * N * 2);
  free(X);
  free(U);
  free(Xold);
  free(Uold);
  cilqr_oracle_ctx_destroy(c);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
typedef struct {
  const cilqr_oracle_params* p;
  int B, N, M_max, S_left, S_right, b0, b1;
  const double *start, *coarse, *corridor, *lane_left, *lane_right;
  const int* corridor_cnt;
  double *states, *controls, *status_out;
  int converged;
  int init_mode;
  const double *init_states, *init_controls;
} batch_job;

static void* batch_worker(void* arg) {
  batch_job* j = (batch_job*)arg;
  const int K = j->N + 1;
  j->converged = 0;
  for (int b = j->b0; b < j->b1; ++b) {
    cilqr_oracle_problem pb;
    pb.N = j->N;
    pb.M_max = j->M_max;
    pb.S_left = j->S_left;
    pb.S_right = j->S_right;
    pb.start = j->start + (size_t)b * 4;
    pb.coarse = j->coarse + (size_t)b * K * 6;
    pb.corridor = j->corridor + (size_t)b * K * j->M_max * 3;
    pb.corridor_cnt = j->corridor_cnt + (size_t)b * K;
    pb.lane_left = j->lane_left + (size_t)b * j->S_left * 7;
    pb.lane_right = j->lane_right + (size_t)b * j->S_right * 7;
    pb.init_mode = j->init_mode;
    pb.init_states = j->init_states ? j->init_states + (size_t)b * K * 6 : 0;
    pb.init_controls = j->init_controls ? j->init_controls + (size_t)b * j->N * 2 : 0;
    cilqr_oracle_result r;
    memset(&r, 0, sizeof(r));
    r.states = j->states + (size_t)b * K * 6;
    r.controls = j->controls + (size_t)b * j->N * 2;
    cilqr_oracle_solve(j->p, &pb, &r);
    if (j->status_out) {
      double* s = j->status_out + (size_t)b * 8;
      s[0] = r.status;
      s[1] = r.iters;
      memcpy(s + 2, r.cost, 5 * sizeof(double));
      s[7] = (double)r.alpha_hash;
    }
    if (r.status <= CILQR_ORACLE_CONVERGED_GRAD) j->converged++;
  }
  return 0;
}

int cilqr_oracle_solve_batch(const cilqr_oracle_params* p, int B, int N, int M_max, int S_left,
                             int S_right, const double* start, const double* coarse,
                             const double* corridor, const int* corridor_cnt,
                             const double* lane_left, const double* lane_right, double* states,
                             double* controls, double* status_out, int nthreads) {
  return cilqr_oracle_solve_batch_init(p, B, N, M_max, S_left, S_right, start, coarse, corridor, corridor_cnt, lane_left,
                                       lane_right, 0, 0, 0, states, controls, status_out, nthreads);
}

int cilqr_oracle_solve_batch_init(const cilqr_oracle_params* p, int B, int N, int M_max, int S_left,
                                  int S_right, const double* start, const double* coarse,
                                  const double* corridor, const int* corridor_cnt,
                                  const double* lane_left, const double* lane_right, int init_mode,
                                  const double* init_states, const double* init_controls, double* states,
                                  double* controls, double* status_out, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > B) nthreads = B > 0 ? B : 1;
  batch_job* jobs = (batch_job*)calloc(nthreads, sizeof(batch_job));
  pthread_t* th = (pthread_t*)calloc(nthreads, sizeof(pthread_t));
  for (int t = 0; t < nthreads; ++t) {
    batch_job* j = &jobs[t];
    j->p = p;
    j->B = B;
    j->N = N;
    j->M_max = M_max;
    j->S_left = S_left;
    j->S_right = S_right;
    j->b0 = (int)((long long)B * t / nthreads);
    j->b1 = (int)((long long)B * (t + 1) / nthreads);
    j->start = start;
    j->coarse = coarse;
    j->corridor = corridor;
    j->corridor_cnt = corridor_cnt;
    j->lane_left = lane_left;
    j->lane_right = lane_right;
    j->states = states;
    j->controls = controls;
    j->status_out = status_out;
    j->init_mode = init_mode;
    j->init_states = init_states;
    j->init_controls = init_controls;
    if (nthreads == 1) {
      batch_worker(j);
    } else {
      pthread_create(&th[t], 0, batch_worker, j);
    }
  }
  int conv = 0;
  for (int t = 0; t < nthreads; ++t) {
    if (nthreads > 1) pthread_join(th[t], 0);
    conv += jobs[t].converged;
  }
  free(jobs);
  free(th);
  return conv;
}
