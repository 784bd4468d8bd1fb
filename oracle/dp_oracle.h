/*
 * dp_oracle.h -- CPU restatement (plain C99, IEEE double) of the coarse DP planner of mpt0816/Cilqr:
 * DpPlanner::Plan (algorithm/planner/dp_planner.cpp:135-281) and everything it executes: GetCost
 * (:87-133), GetCollisionCost (:40-85), InterpolateLinearly (:283-320), GetLateralOffset
 * (dp_planner.h:83-92), DiscretizedTrajectory::GetProjection / EvaluateStation / GetCartesian
 * (utils/discretized_trajectory.cpp:62-198, math_utils.h:208-225 slerp), the collision checks of
 * Environment (utils/environment.cpp:51-141: CheckOptimizationCollision, CheckStaticCollision,
 * CheckDynamicCollision; math/polygon2d.cpp:120-165 IsPointIn / HasOverlap(Box2d); math/box2d.cpp:93-129)
 * and DiscretePointsMath::ComputePathProfile (utils/discrete_points_math.cc:27-176).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rule as cilqr_oracle.h).
 *
 * PARITY PINNED AGAINST THE REFERENCE ITSELF: the reference's own dp_planner.cpp, environment.cpp, geometry,
 * reference-line and path-profile sources compile without ROS / Eigen / OpenCV once the visualization header is
 * replaced by a no-op stand-in (oracle/Makefile target `_ref`, wrappers oracle/ref_*_wrapper.cc).  tests/
 * test_reference_pins.py runs the reference's DpPlanner::Plan and Environment::CheckOptimizationCollision on the
 * same scenes: return value and every field of every trajectory point are bit-identical, and so is the committed
 * fixture tests/golden/dp_golden_v1.npz.  Additionally cross-checked against an independent Python restatement
 * (oracle/dp_python.py) and known-answer tests (tests/test_dp_oracle.py).
 */
#ifndef DP_ORACLE_H_
#define DP_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define DP_NT 5  /* dp_planner.h:27-29 */
#define DP_NS 7
#define DP_NL 10

/* PlannerConfig (params/planner_config.h:88-141) + the VehicleParam fields the planner reads
 * (params/vehicle_param.h:26-85). */
typedef struct dp_config {
  double tf, delta_t;                  /* 8, 0.1 */
  double dp_nominal_velocity;          /* 10 */
  double dp_w_obstacle;                /* 1000 */
  double dp_w_lateral;                 /* 0.1 */
  double dp_w_lateral_change;          /* 0.5 */
  double dp_w_lateral_velocity_change; /* 1 */
  double dp_w_longitudinal_velocity_bias;   /* 10 */
  double dp_w_longitudinal_velocity_change; /* 1 */
  double max_velocity, width, wheel_base, front_hang_length, rear_hang_length;
} dp_config;

void dp_default_config(dp_config* c);

/* The Environment as flat arrays (one scenario). */
typedef struct dp_env {
  int R;                 /* reference line points */
  const double* ref;     /* [R][7] s, x, y, theta, kappa, left_bound, right_bound (CenterLinePoint.msg) */
  int NB;                /* road barrier points, sorted by x (Environment::set_reference, environment.cpp:24-49) */
  const double* barrier; /* [NB][2] */
  int V;                 /* vertex pitch of the polygon arrays */
  int n_static;
  const double* static_poly; /* [n_static][V][2] */
  const int* static_nv;      /* [n_static] vertices used */
  int n_dyn, T;              /* dynamic obstacles, sample pitch */
  const double* dyn_time;    /* [n_dyn][T] sample times, ascending */
  const int* dyn_samples;    /* [n_dyn] samples used */
  const double* dyn_poly;    /* [n_dyn][T][V][2] */
  const int* dyn_nv;         /* [n_dyn] */
} dp_env;

/* Environment::set_reference's road barrier (environment.cpp:24-49): writes up to cap points [cap][2]
 * sorted by x, returns their number. */
int dp_build_barrier(int R, const double* ref, double* out, int cap);

/* pieces exposed for the known-answer tests */
void dp_evaluate_station(int R, const double* ref, double station, double out7[7]);
void dp_get_projection(int R, const double* ref, double x, double y, double sl[2]);
int dp_check_optimization_collision(const dp_config* cfg, const dp_env* env, double time, double x, double y,
                                    double theta);
int dp_num_knots(const dp_config* cfg); /* sum of the layers' segment counts (= tf/delta_t + 1 for the defaults) */

int dp_polygon_overlaps_box(const double* p, int nv, double cx, double cy, double half); /* polygon2d.cpp:150-165 */
int dp_polygon_is_point_in(const double* p, int nv, double x, double y);                  /* polygon2d.cpp:120-140 */
void dp_get_cartesian(int R, const double* ref, double station, double lateral, double xy[2]);
void dp_path_profile(double dt, int n, const double* x, const double* y, double* speeds, double* accel,
                     double* kappas);                                                     /* discrete_points_math.cc */
double dp_slerp(double a0, double t0, double a1, double t1, double t);                    /* math_utils.h:208-225 */

/* DpPlanner::Plan.  trajectory [K][13] in TrajectoryPoint field order (time, s, x, y, theta, kappa, velocity, a,
 * jerk, delta, delta_rate, left_bound, right_bound); waypoints [NT][3] = (s index, l index, current_s) of the
 * optimum (may be NULL).  Returns 1 when min_cost < dp_w_obstacle (the reference's return value), else 0. */
int dp_plan(const dp_config* cfg, const dp_env* env, double start_x, double start_y, double start_theta,
            double* trajectory, double* min_cost, double* waypoints);

#ifdef __cplusplus
}
#endif
#endif
