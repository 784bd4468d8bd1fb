// dp_planner_b200.h -- header-compatible replacement of planning::DpPlanner
// (reference: algorithm/planner/dp_planner.h:31-97, algorithm/planner/dp_planner.cpp) that runs the lattice
// search on a B200 through the C ABI of include/cilqr_b200.h.
//
// Drop-in use inside the reference tree: include this header instead of
// "algorithm/planner/dp_planner.h" in algorithm/planner/trajectory_planner.h and drop
// algorithm/planner/dp_planner.cpp from CMakeLists.txt; trajectory_planner.cpp:24 (member initialiser
// `dp_(config, env)`) and :32 (`dp_.Plan(state.x, state.y, state.theta, coarse_trajectory)`) compile unchanged.
//
// The adapter only flattens the Environment (centre line, road barrier, obstacle polygons, dynamic obstacle
// trajectories -- utils/environment.h:24-88) into the arrays of the ABI (one scenario), calls
// cilqr_dp_plan_batch and rebuilds the DiscretizedTrajectory.  It contains no planner arithmetic; on any error
// (no device, CUDA failure) Plan() returns false and leaves `result` untouched.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "algorithm/math/polygon2d.h"
#include "algorithm/params/planner_config.h"
#include "algorithm/utils/discretized_trajectory.h"
#include "algorithm/utils/environment.h"
#include "cilqr_b200.h"

namespace planning {

class DpPlanner {
 public:
  DpPlanner(const PlannerConfig& config, const Env& env) : env_(env), config_(config) {}
  DpPlanner(const DpPlanner& o) : env_(o.env_), config_(o.config_), device_(o.device_) {}
  DpPlanner& operator=(const DpPlanner& o) {
    if (this != &o) {
      Release();
      env_ = o.env_;
      config_ = o.config_;
      device_ = o.device_;
    }
    return *this;
  }
  ~DpPlanner() { Release(); }

  void set_device(int device) {
    Release();
    device_ = device;
  }

  // dp_planner.cpp:135-281
  bool Plan(const double start_x, const double start_y, const double start_theta, DiscretizedTrajectory& result) {
    if (!EnsureHandle()) return false;
    CilqrDpConfig cfg;
    cilqr_dp_default_config(&cfg);
    cfg.tf = config_.tf;
    cfg.delta_t = config_.delta_t;
    cfg.dp_nominal_velocity = config_.dp_nominal_velocity;
    cfg.dp_w_obstacle = config_.dp_w_obstacle;
    cfg.dp_w_lateral = config_.dp_w_lateral;
    cfg.dp_w_lateral_change = config_.dp_w_lateral_change;
    cfg.dp_w_lateral_velocity_change = config_.dp_w_lateral_velocity_change;
    cfg.dp_w_longitudinal_velocity_bias = config_.dp_w_longitudinal_velocity_bias;
    cfg.dp_w_longitudinal_velocity_change = config_.dp_w_longitudinal_velocity_change;
    cfg.max_velocity = config_.vehicle.max_velocity;
    cfg.width = config_.vehicle.width;
    cfg.wheel_base = config_.vehicle.wheel_base;
    cfg.front_hang_length = config_.vehicle.front_hang_length;
    cfg.rear_hang_length = config_.vehicle.rear_hang_length;
    const int K = cilqr_dp_num_knots(&cfg);
    if (K < 2) return false;

    // ---- centre line (Environment::reference()) and road barrier (road_barrier_ is private in the reference:
    // it is the left and right barrier points sorted by x, environment.cpp:24-49)
    const auto& line = env_->reference().trajectory();
    if (line.size() < 2) return false;
    std::vector<double> ref(line.size() * 7);
    for (size_t i = 0; i < line.size(); ++i) {
      double* r = &ref[i * 7];
      r[0] = line[i].s; r[1] = line[i].x; r[2] = line[i].y; r[3] = line[i].theta; r[4] = line[i].kappa;
      r[5] = line[i].left_bound; r[6] = line[i].right_bound;
    }
    std::vector<std::pair<double, double>> bar;
    for (const auto& p : env_->left_road_barrier()) bar.emplace_back(p.x(), p.y());
    for (const auto& p : env_->right_road_barrier()) bar.emplace_back(p.x(), p.y());
    std::stable_sort(bar.begin(), bar.end(),
                     [](const std::pair<double, double>& a, const std::pair<double, double>& b) { return a.first < b.first; });
    std::vector<double> barrier(bar.size() * 2 + 2);
    for (size_t i = 0; i < bar.size(); ++i) {
      barrier[2 * i] = bar[i].first;
      barrier[2 * i + 1] = bar[i].second;
    }
    // ---- obstacles
    const auto& statics = env_->obstacles();
    const auto& dynamics = env_->dynamic_obstacles();
    size_t V = 1, T = 1;
    for (const auto& o : statics) V = std::max(V, o.points().size());
    for (const auto& o : dynamics) {
      T = std::max(T, o.size());
      for (const auto& s : o) V = std::max(V, s.second.points().size());
    }
    std::vector<double> spoly(std::max<size_t>(statics.size(), 1) * V * 2, 0.0);
    std::vector<int32_t> snv(std::max<size_t>(statics.size(), 1), 0);
    for (size_t o = 0; o < statics.size(); ++o) {
      snv[o] = static_cast<int32_t>(statics[o].points().size());
      for (size_t v = 0; v < statics[o].points().size(); ++v) {
        spoly[(o * V + v) * 2] = statics[o].points()[v].x();
        spoly[(o * V + v) * 2 + 1] = statics[o].points()[v].y();
      }
    }
    const size_t nd = dynamics.size();
    std::vector<double> dtime(std::max<size_t>(nd, 1) * T, 0.0), dpoly(std::max<size_t>(nd, 1) * T * V * 2, 0.0);
    std::vector<int32_t> dsamples(std::max<size_t>(nd, 1), 0), dnv(std::max<size_t>(nd, 1), 0);
    for (size_t o = 0; o < nd; ++o) {
      dsamples[o] = static_cast<int32_t>(dynamics[o].size());
      for (size_t t = 0; t < dynamics[o].size(); ++t) {
        dtime[o * T + t] = dynamics[o][t].first;
        const auto& pts = dynamics[o][t].second.points();
        dnv[o] = static_cast<int32_t>(pts.size());  // one polygon shape per obstacle (planning_node.cc:63-78)
        for (size_t v = 0; v < pts.size(); ++v) {
          dpoly[((o * T + t) * V + v) * 2] = pts[v].x();
          dpoly[((o * T + t) * V + v) * 2 + 1] = pts[v].y();
        }
      }
    }
    const double start[3] = {start_x, start_y, start_theta};
    CilqrDpIn in;
    in.B = 1; in.R = static_cast<int32_t>(line.size()); in.NB = static_cast<int32_t>(bar.size());
    in.V = static_cast<int32_t>(V); in.n_static = static_cast<int32_t>(statics.size());
    in.n_dyn = static_cast<int32_t>(nd); in.T = static_cast<int32_t>(T);
    in.ref = ref.data(); in.barrier = barrier.data(); in.start = start; in.static_poly = spoly.data();
    in.static_nv = snv.data(); in.dyn_time = dtime.data(); in.dyn_samples = dsamples.data();
    in.dyn_poly = dpoly.data(); in.dyn_nv = dnv.data();
    std::vector<double> traj(static_cast<size_t>(K) * CILQR_TRAJPOINT_DOUBLES);
    int32_t ok = 0;
    double min_cost = 0.0;
    CilqrDpOut out;
    out.trajectory = traj.data(); out.coarse = nullptr; out.xytheta = nullptr; out.ok = &ok; out.cost = &min_cost;
    out.waypoints = nullptr;
    const int rc = cilqr_dp_plan_batch(handle_, &cfg, &in, &out);
    if (rc != CILQR_OK) {
      std::fprintf(stderr, "cilqr_b200: %s (%s)\n", cilqr_strerror(rc), cilqr_last_cuda_error(handle_));
      return false;
    }
    std::vector<TrajectoryPoint> data(K);
    for (int k = 0; k < K; ++k) {
      const double* r = &traj[static_cast<size_t>(k) * CILQR_TRAJPOINT_DOUBLES];
      TrajectoryPoint& p = data[k];
      p.time = r[0]; p.s = r[1]; p.x = r[2]; p.y = r[3]; p.theta = r[4]; p.kappa = r[5]; p.velocity = r[6];
      p.a = r[7]; p.jerk = r[8]; p.delta = r[9]; p.delta_rate = r[10];
    }
    result = DiscretizedTrajectory(data);
    min_cost_ = min_cost;
    return ok != 0;  // min_cost < config_.dp_w_obstacle, dp_planner.cpp:280
  }

  double min_cost() const { return min_cost_; }

 private:
  bool EnsureHandle() {
    if (handle_ != nullptr) return true;
    CilqrParams p;
    cilqr_default_params(&p);
    const int rc = cilqr_create(&p, device_, 1, 1, 1, 1, &handle_);
    if (rc != CILQR_OK) {
      std::fprintf(stderr, "cilqr_create failed: %s\n", cilqr_strerror(rc));
      handle_ = nullptr;
      return false;
    }
    return true;
  }
  void Release() {
    if (handle_ != nullptr) cilqr_destroy(handle_);
    handle_ = nullptr;
  }

  Env env_;
  PlannerConfig config_;
  cilqr_handle* handle_ = nullptr;
  int device_ = 0;
  double min_cost_ = 0.0;
};

}  // namespace planning
