// ref_dp_wrapper.cc -- extern "C" access to the REFERENCE'S OWN DpPlanner and Environment, compiled unmodified from
// /root/reference (algorithm/planner/dp_planner.cpp, algorithm/utils/environment.cpp and the geometry sources;
// only the visualization header is replaced by the no-op stand-in in oracle/ref_stubs).  With it the DP planner
// restatement (oracle/dp_oracle.c) is pinned against the reference itself on whole plans.
// TEST INFRASTRUCTURE ONLY.  This file contains no reference code: it only calls it.
#include <memory>
#include <vector>

#include "algorithm/planner/dp_planner.h"
#include "algorithm/visualization/plot.h"
#include "algorithm/utils/environment.h"

using namespace planning;

namespace planning {
namespace visualization {
Color Color::Grey, Color::Magenta, Color::White, Color::Cyan;
}
}  // namespace planning

static Env make_env(const PlannerConfig& config, int R, const double* ref, int V, int n_static,
                    const double* static_poly, const int* static_nv, int n_dyn, int T, const double* dyn_time,
                    const int* dyn_samples, const double* dyn_poly, const int* dyn_nv) {
  Env env = std::make_shared<Environment>(config);
  std::vector<TrajectoryPoint> pts(R);
  for (int i = 0; i < R; ++i) {  // PlanningNode::CenterLineCallback, planning_node.cc:33-48
    const double* r = ref + (size_t)i * 7;
    pts[i].s = r[0]; pts[i].x = r[1]; pts[i].y = r[2]; pts[i].theta = r[3]; pts[i].kappa = r[4];
    pts[i].left_bound = r[5]; pts[i].right_bound = r[6];
  }
  env->set_reference(DiscretizedTrajectory(pts));
  for (int o = 0; o < n_static; ++o) {  // ObstaclesCallback, :51-61
    std::vector<math::Vec2d> p;
    for (int v = 0; v < static_nv[o]; ++v)
      p.emplace_back(static_poly[((size_t)o * V + v) * 2], static_poly[((size_t)o * V + v) * 2 + 1]);
    env->obstacles().emplace_back(p);
  }
  for (int o = 0; o < n_dyn; ++o) {  // DynamicObstaclesCallback, :63-80
    Environment::DynamicObstacle ob;
    for (int t = 0; t < dyn_samples[o]; ++t) {
      std::vector<math::Vec2d> p;
      for (int v = 0; v < dyn_nv[o]; ++v) {
        const double* q = dyn_poly + (((size_t)o * T + t) * V + v) * 2;
        p.emplace_back(q[0], q[1]);
      }
      ob.emplace_back(dyn_time[(size_t)o * T + t], math::Polygon2d(p));
    }
    env->dynamic_obstacles().push_back(ob);
  }
  return env;
}

extern "C" {

// DpPlanner::Plan with the default PlannerConfig.  trajectory [K][11]: time, s, x, y, theta, kappa, velocity, a,
// jerk, delta, delta_rate.  Returns the number of knots; *ok = Plan's return value.
int ref_dp_plan(int R, const double* ref, int V, int n_static, const double* static_poly, const int* static_nv,
                int n_dyn, int T, const double* dyn_time, const int* dyn_samples, const double* dyn_poly,
                const int* dyn_nv, double sx, double sy, double stheta, double* trajectory, int cap, int* ok) {
  PlannerConfig config;
  Env env = make_env(config, R, ref, V, n_static, static_poly, static_nv, n_dyn, T, dyn_time, dyn_samples, dyn_poly,
                     dyn_nv);
  DpPlanner dp(config, env);
  DiscretizedTrajectory result;
  *ok = dp.Plan(sx, sy, stheta, result) ? 1 : 0;
  const auto& pts = result.trajectory();
  const int K = (int)pts.size();
  for (int k = 0; k < K && k < cap; ++k) {
    const TrajectoryPoint& p = pts[k];
    const double row[11] = {p.time, p.s, p.x, p.y, p.theta, p.kappa, p.velocity, p.a, p.jerk, p.delta, p.delta_rate};
    for (int j = 0; j < 11; ++j) trajectory[(size_t)k * 11 + j] = row[j];
  }
  return K;
}

// Environment::CheckOptimizationCollision (collision_buffer = 0)
int ref_check_optimization_collision(int R, const double* ref, int V, int n_static, const double* static_poly,
                                     const int* static_nv, int n_dyn, int T, const double* dyn_time,
                                     const int* dyn_samples, const double* dyn_poly, const int* dyn_nv, int n,
                                     const double* queries /* [n][4] time, x, y, theta */, int* out) {
  PlannerConfig config;
  Env env = make_env(config, R, ref, V, n_static, static_poly, static_nv, n_dyn, T, dyn_time, dyn_samples, dyn_poly,
                     dyn_nv);
  for (int i = 0; i < n; ++i) {
    const double* q = queries + (size_t)i * 4;
    out[i] = env->CheckOptimizationCollision(q[0], math::Pose(q[1], q[2], q[3])) ? 1 : 0;
  }
  return 0;
}

// Environment::set_reference's road barrier is private; its left / right halves are public (same points)
int ref_road_barrier(int R, const double* ref, double* out, int cap) {
  PlannerConfig config;
  Env env = make_env(config, R, ref, 1, 0, nullptr, nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr);
  int n = 0;
  for (const auto& p : env->left_road_barrier()) {
    if (n >= cap) return -1;
    out[2 * n] = p.x(); out[2 * n + 1] = p.y(); ++n;
  }
  for (const auto& p : env->right_road_barrier()) {
    if (n >= cap) return -1;
    out[2 * n] = p.x(); out[2 * n + 1] = p.y(); ++n;
  }
  return n;
}

}  // extern "C"
