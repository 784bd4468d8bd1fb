/*
 * tracker_oracle.h -- CPU restatement of the reference's Tracker initial guess (algorithm/ilqr/tracker.cc,
 * algorithm/math/linear_quadratic_regulator.cc) and of IlqrOptimizer::InitGuess (ilqr_optimizer.cc:107-139).
 * TEST INFRASTRUCTURE ONLY -- see tracker_oracle.c.  Trajectory points are 13-double records in the field order
 * of TrajectoryPoint (discretized_trajectory.h:26-43): time, s, x, y, theta, kappa, velocity, a, jerk, delta,
 * delta_rate, left_bound, right_bound.
 */
#ifndef TRACKER_ORACLE_H_
#define TRACKER_ORACLE_H_
#ifdef __cplusplus
extern "C" {
#endif

typedef struct tracker_oracle_config { /* TrackerConfig (planner_config.h:18-43) + the VehicleParam fields the tracker reads */
  double sumulation_dt, dt, tolerance;
  int max_num_iteration;
  double lat_weight_l, lat_weight_theta, lat_weight_delta, lat_weight_delta_rate, lat_preview_time;
  double lon_weight_s, lon_weight_v, lon_weight_a, lon_weight_j;
  double wheel_base, delta_min, delta_max, min_acceleration, max_acceleration, delta_rate_min, delta_rate_max,
      jerk_min, jerk_max;
} tracker_oracle_config;

void tracker_oracle_default_config(tracker_oracle_config* c);
/* math::SolveLQRProblem for the tracker's 3-state / 1-input systems; K is 1x3 */
void tracker_oracle_solve_lqr(const double A[9], const double B[3], const double Q[9], double R, double tolerance,
                              unsigned max_num_iteration, double K[3], int* iterations);
/* Tracker::Plan: start [13], coarse [K][13] -> out [K][13]; returns 1 on success (trajectory.size() == K) */
int tracker_oracle_plan(const tracker_oracle_config* c, const double start[13], const double* coarse, int K,
                        double* out, int* lqr_iterations_total);
/* IlqrOptimizer::InitGuess: trajectory [K][13] -> states [K][6], controls [K-1][2] */
void tracker_oracle_init_guess(const double* traj, int K, double* states, double* controls);

#ifdef __cplusplus
}
#endif
#endif
