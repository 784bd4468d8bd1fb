// ref_tracker_wrapper.cc -- extern "C" access to the reference's OWN Tracker (algorithm/ilqr/tracker.cc) and DARE solver
// (algorithm/math/linear_quadratic_regulator.cc), compiled unmodified against the Eigen stand-in of ref_stubs/.
// TEST INFRASTRUCTURE ONLY (oracle/_ref/libcilqr_ref_tracker.so, built by `make -C oracle _ref`); pins
// oracle/tracker_oracle.c in tests/test_reference_pins.py.
#include <stdexcept>
#include <vector>

#include "algorithm/ilqr/tracker.h"

using namespace planning;

static TrajectoryPoint to_point(const double* r) {
  TrajectoryPoint p;
  p.time = r[0]; p.s = r[1]; p.x = r[2]; p.y = r[3]; p.theta = r[4]; p.kappa = r[5]; p.velocity = r[6]; p.a = r[7];
  p.jerk = r[8]; p.delta = r[9]; p.delta_rate = r[10]; p.left_bound = r[11]; p.right_bound = r[12];
  return p;
}

extern "C" {

// Tracker::Plan with the default TrackerConfig / VehicleParam.  Returns 1 on success, 0 on "tracker failed", -1 when the
// reference throws (trajectory().at(i) past the end, tracker.cc:198).
int ref_tracker_plan(const double* start13, const double* coarse, int K, double* out) {
  TrackerConfig cfg;
  VehicleParam veh;
  Tracker tracker(cfg, veh);
  std::vector<TrajectoryPoint> pts(K);
  for (int k = 0; k < K; ++k) pts[k] = to_point(coarse + (size_t)k * 13);
  DiscretizedTrajectory follow(pts), res;
  bool ok = false;
  try {
    ok = tracker.Plan(to_point(start13), follow, &res);
  } catch (const std::out_of_range&) {
    return -1;
  }
  if (!ok) return 0;
  for (int k = 0; k < K && k < (int)res.trajectory().size(); ++k) {
    const TrajectoryPoint& p = res.trajectory()[k];
    double* r = out + (size_t)k * 13;
    r[0] = p.time; r[1] = p.s; r[2] = p.x; r[3] = p.y; r[4] = p.theta; r[5] = p.kappa; r[6] = p.velocity; r[7] = p.a;
    r[8] = p.jerk; r[9] = p.delta; r[10] = p.delta_rate; r[11] = p.left_bound; r[12] = p.right_bound;
  }
  return 1;
}

}  // extern "C"
