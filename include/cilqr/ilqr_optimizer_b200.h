// ilqr_optimizer_b200.h -- header-compatible replacement of planning::IlqrOptimizer
// (reference: algorithm/ilqr/ilqr_optimizer.h:31-52) that runs the solve on a B200 through the
// C ABI of include/cilqr_b200.h.
//
// Drop-in use inside the reference tree:
//   * algorithm/planner/trajectory_planner.h:  #include "cilqr/ilqr_optimizer_b200.h"  instead of
//     "algorithm/ilqr/ilqr_optimizer.h"; remove algorithm/ilqr/ilqr_optimizer.cc from CMakeLists.txt;
//     link libcilqr_b200.so.  trajectory_planner.cpp:26,80-86,97 compile unchanged.
//   * Same constructor / Init / Plan / cost() signatures, same guards and the same observable
//     outputs: opt_trajectory (K TrajectoryPoints, ilqr_optimizer.cc:771-791), iter_trajs
//     (initial guess first, then every accepted non-final iterate, :170,294) and cost()
//     (:173,283,296).
//
// Host side stays C++/Eigen: this header only packs the caller's containers into the flat POD
// arrays of the ABI (batch of one), calls cilqr_plan_batch, and unpacks.  It contains no solver
// arithmetic; if the CUDA library reports an error Plan() returns false and leaves
// *opt_trajectory empty, which is the caller's failure signal (trajectory_planner.cpp:91-94).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <utility>
#include <vector>

#include <Eigen/Core>

#include "algorithm/math/line_segment2d.h"
#include "algorithm/params/planner_config.h"
#include "algorithm/utils/discretized_trajectory.h"
#include "cilqr_b200.h"

namespace planning {

// Same aliases as algorithm/ilqr/corridor.h:18-25 (re-declaring an alias to the same type is legal,
// so this header coexists with corridor.h).  Half-plane convention: a*x + b*y < c, stored (a,b,c).
using Constraints = std::vector<Eigen::Vector3d>;
using CorridorConstraints = std::vector<Constraints>;
using LaneConstraints = std::vector<std::pair<Eigen::Vector3d, math::LineSegment2d>>;

#ifndef CILQR_B200_HAVE_REFERENCE_COST
// algorithm/ilqr/ilqr_optimizer.h:14-27
struct Cost {
  double total_cost = 0.0;
  double target_cost = 0.0;
  double dynamic_cost = 0.0;
  double corridor_cost = 0.0;
  double lane_boundary_cost = 0.0;
  Cost() = default;
  Cost(const double c0, const double c1, const double c2, const double c3, const double c4)
      : total_cost(c0), target_cost(c1), dynamic_cost(c2), corridor_cost(c3), lane_boundary_cost(c4) {}
};
#endif

class IlqrOptimizer {
 public:
  IlqrOptimizer() = default;

  IlqrOptimizer(const IlqrConfig& config, const VehicleParam& param, const double horizon, const double dt) {
    Init(config, param, horizon, dt);
  }

  IlqrOptimizer(const IlqrOptimizer& o) { *this = o; }
  // TrajectoryPlanner assigns a temporary (trajectory_planner.cpp:26): configuration is copied,
  // the device handle is not shared -- each object lazily creates its own.
  IlqrOptimizer& operator=(const IlqrOptimizer& o) {
    if (this != &o) {
      Release();
      config_ = o.config_;
      vehicle_param_ = o.vehicle_param_;
      horizon_ = o.horizon_;
      delta_t_ = o.delta_t_;
      num_of_knots_ = o.num_of_knots_;
      device_ = o.device_;
      cost_ = o.cost_;
      last_status_ = o.last_status_;
      last_iterations_ = o.last_iterations_;
    }
    return *this;
  }

  ~IlqrOptimizer() { Release(); }

  // ilqr_optimizer.cc:13-51
  void Init(const IlqrConfig& config, const VehicleParam& param, const double horizon, const double dt) {
    Release();
    config_ = config;
    vehicle_param_ = param;
    horizon_ = horizon;
    delta_t_ = dt;
    num_of_knots_ = static_cast<int>(std::floor(horizon_ / delta_t_ + 1));
    cost_.clear();
  }

  void set_device(int device) {
    Release();
    device_ = device;
  }

  // ilqr_optimizer.cc:53-95.  Returns false on the reference's three guards (and on a CUDA
  // failure); true otherwise (the reference falls off the end there).
  bool Plan(const TrajectoryPoint& start_state, const DiscretizedTrajectory& coarse_traj,
            const CorridorConstraints& corridor, const LaneConstraints& left_lane_cons,
            const LaneConstraints& right_lane_cons, DiscretizedTrajectory* const opt_trajectory,
            std::vector<DiscretizedTrajectory>* const iter_trajs) {
    cost_.clear();
    if (opt_trajectory == nullptr || iter_trajs == nullptr) return false;
    if (corridor.size() == 0 || left_lane_cons.size() == 0 || right_lane_cons.size() == 0) {
      std::fprintf(stderr, "ilqr input constraints error\n");
      return false;
    }
    if (static_cast<size_t>(num_of_knots_) != coarse_traj.trajectory().size()) {
      std::fprintf(stderr, "ilqr input coarse_traj error\n");
      return false;
    }
    const int K = num_of_knots_, N = K - 1;
    // The reference indexes corridor[i] for every knot (ilqr_optimizer.cc:561); a shorter
    // corridor is undefined behaviour there and an input error here.
    if (N < 1 || corridor.size() < static_cast<size_t>(K)) {
      std::fprintf(stderr, "ilqr input corridor error\n");
      return false;
    }
    int M_max = 1;
    for (int k = 0; k < K; ++k) M_max = std::max<int>(M_max, static_cast<int>(corridor[k].size()));
    const int S_left = static_cast<int>(left_lane_cons.size()), S_right = static_cast<int>(right_lane_cons.size());

    // ---- pack (wire format of cilqr_b200.h)
    const double start[4] = {start_state.x, start_state.y, start_state.theta, start_state.velocity};
    std::vector<double> coarse(static_cast<size_t>(K) * 6);
    for (int k = 0; k < K; ++k) {
      const TrajectoryPoint& pt = coarse_traj.trajectory()[k];  // fields of TransformGoals, :148
      double* g = &coarse[static_cast<size_t>(k) * 6];
      g[0] = pt.x; g[1] = pt.y; g[2] = pt.theta; g[3] = pt.velocity; g[4] = pt.a; g[5] = pt.delta;
    }
    std::vector<double> planes(static_cast<size_t>(K) * M_max * 3, 0.0);
    std::vector<int32_t> cnt(K);
    for (int k = 0; k < K; ++k) {
      cnt[k] = static_cast<int32_t>(corridor[k].size());
      for (int m = 0; m < cnt[k]; ++m) {
        double* p = &planes[(static_cast<size_t>(k) * M_max + m) * 3];
        p[0] = corridor[k][m][0]; p[1] = corridor[k][m][1]; p[2] = corridor[k][m][2];
      }
    }
    auto pack_lane = [](const LaneConstraints& lane) {
      std::vector<double> out(lane.size() * 7);
      for (size_t s = 0; s < lane.size(); ++s) {
        double* p = &out[s * 7];
        p[0] = lane[s].first[0]; p[1] = lane[s].first[1]; p[2] = lane[s].first[2];
        p[3] = lane[s].second.start().x(); p[4] = lane[s].second.start().y();
        p[5] = lane[s].second.end().x();   p[6] = lane[s].second.end().y();
      }
      return out;
    };
    const std::vector<double> ll = pack_lane(left_lane_cons), lr = pack_lane(right_lane_cons);

    if (!EnsureHandle(N, M_max, std::max(S_left, S_right))) return false;

    const int H = config_.max_iter_num + 2;  // initial + at most one entry per iteration
    std::vector<double> states(static_cast<size_t>(K) * 6), controls(static_cast<size_t>(N) * 2), status(CILQR_STATUS_DOUBLES);
    std::vector<double> traj(static_cast<size_t>(K) * CILQR_TRAJPOINT_DOUBLES);
    std::vector<double> cost_hist(static_cast<size_t>(H) * 5), it_x(static_cast<size_t>(H) * K * 6), it_u(static_cast<size_t>(H) * N * 2);
    int32_t hist_len[2] = {0, 0};
    CilqrBatchIn in;
    in.B = 1; in.N = N; in.M_max = M_max; in.S_left = S_left; in.S_right = S_right;
    in.start = start; in.coarse = coarse.data(); in.corridor = planes.data(); in.corridor_cnt = cnt.data();
    in.lane_left = ll.data(); in.lane_right = lr.data();
    in.init_mode = CILQR_INIT_IQR; in.init_states = nullptr; in.init_controls = nullptr;  // the live line :169
    CilqrBatchOut out;
    out.states = states.data(); out.controls = controls.data(); out.status = status.data();
    out.trajectory = traj.data(); out.init_states = nullptr; out.init_controls = nullptr;
    out.cost_hist = cost_hist.data(); out.iter_states = it_x.data(); out.iter_controls = it_u.data();
    out.hist_len = hist_len; out.hist_cap = H; out.result = nullptr;
    const int rc = cilqr_plan_batch(handle_, &in, &out);
    if (rc != CILQR_OK) {
      std::fprintf(stderr, "cilqr_b200: %s (%s)\n", cilqr_strerror(rc), cilqr_last_cuda_error(handle_));
      return false;
    }
    last_status_ = static_cast<int>(status[CILQR_ST_FLAG]);
    last_iterations_ = static_cast<int>(status[CILQR_ST_ITERS]);

    // ---- unpack: cost_ (:173,283,296), iter_trajs (:170,294), opt_trajectory (every exit path)
    for (int i = 0; i < hist_len[0] && i < H; ++i) {
      const double* c = &cost_hist[static_cast<size_t>(i) * 5];
      cost_.emplace_back(c[0], c[1], c[2], c[3], c[4]);
    }
    for (int i = 0; i < hist_len[1] && i < H; ++i) {
      iter_trajs->emplace_back(ToTrajectory(&it_x[static_cast<size_t>(i) * K * 6], &it_u[static_cast<size_t>(i) * N * 2]));
    }
    std::vector<TrajectoryPoint> pts(K);
    for (int k = 0; k < K; ++k) {
      const double* r = &traj[static_cast<size_t>(k) * CILQR_TRAJPOINT_DOUBLES];
      TrajectoryPoint& p = pts[k];
      p.time = r[0]; p.s = r[1]; p.x = r[2]; p.y = r[3]; p.theta = r[4]; p.kappa = r[5]; p.velocity = r[6];
      p.a = r[7]; p.jerk = r[8]; p.delta = r[9]; p.delta_rate = r[10]; p.left_bound = r[11]; p.right_bound = r[12];
    }
    *opt_trajectory = DiscretizedTrajectory(pts);
    return true;
  }

  std::vector<Cost> cost() { return cost_; }

  // extras (not in the reference): which exit Optimize() took (CILQR_CONVERGED_* ...) and `iter`
  int last_status() const { return last_status_; }
  int last_iterations() const { return last_iterations_; }

 private:
  // TransformToTrajectory, ilqr_optimizer.cc:771-791 (record layout only; no solver arithmetic)
  DiscretizedTrajectory ToTrajectory(const double* X, const double* U) const {
    const int K = num_of_knots_;
    std::vector<TrajectoryPoint> traj(K);
    for (int i = 0; i < K; ++i) {
      const double* x = X + static_cast<size_t>(i) * 6;
      traj[i].time = i * delta_t_;
      traj[i].x = x[0]; traj[i].y = x[1]; traj[i].theta = x[2]; traj[i].velocity = x[3];
      traj[i].a = x[4]; traj[i].delta = x[5];
      traj[i].kappa = std::tan(x[5]) / vehicle_param_.wheel_base;
      if (i < K - 1) {
        traj[i].jerk = U[static_cast<size_t>(i) * 2];
        traj[i].delta_rate = U[static_cast<size_t>(i) * 2 + 1];
      }
    }
    return DiscretizedTrajectory(traj);
  }

  bool EnsureHandle(int N, int M, int S) {
    if (handle_ && N <= cap_N_ && M <= cap_M_ && S <= cap_S_) return true;
    Release();
    CilqrParams p;
    cilqr_default_params(&p);
    p.front_hang_length = vehicle_param_.front_hang_length;
    p.wheel_base = vehicle_param_.wheel_base;
    p.rear_hang_length = vehicle_param_.rear_hang_length;
    p.width = vehicle_param_.width;
    p.max_velocity = vehicle_param_.max_velocity;
    p.min_acceleration = vehicle_param_.min_acceleration;
    p.max_acceleration = vehicle_param_.max_acceleration;
    p.jerk_min = vehicle_param_.jerk_min;
    p.jerk_max = vehicle_param_.jerk_max;
    p.delta_min = vehicle_param_.delta_min;
    p.delta_max = vehicle_param_.delta_max;
    p.delta_rate_min = vehicle_param_.delta_rate_min;
    p.delta_rate_max = vehicle_param_.delta_rate_max;
    p.safe_margin = config_.safe_margin;
    p.w_jerk = config_.weights.jerk;
    p.w_delta_rate = config_.weights.delta_rate;
    p.w_x_target = config_.weights.x_target;
    p.w_y_target = config_.weights.y_target;
    p.w_theta = config_.weights.theta;
    p.w_v = config_.weights.v;
    p.w_a = config_.weights.a;
    p.w_delta = config_.weights.delta;
    p.abs_cost_tol = config_.abs_cost_tol;
    p.rel_cost_tol = config_.rel_cost_tol;
    // barrier_t / barrier_eps keep the library defaults: the reference never reads IlqrConfig::t
    // (barrier_function.h:143-146 hard-wires t = 5, eps = 0.01).
    p.delta_t = delta_t_;
    p.num_of_disc = config_.num_of_disc;
    p.max_iter_num = config_.max_iter_num;
    cap_N_ = N;
    cap_M_ = std::max(M, 32);
    cap_S_ = std::max(S, 64);
    const int rc = cilqr_create(&p, device_, cap_N_, cap_M_, cap_S_, /*B_max=*/1, &handle_);
    if (rc != CILQR_OK) {
      std::fprintf(stderr, "cilqr_b200: cilqr_create failed: %s\n", cilqr_strerror(rc));
      handle_ = nullptr;
      return false;
    }
    return true;
  }

  void Release() {
    if (handle_) cilqr_destroy(handle_);
    handle_ = nullptr;
    cap_N_ = cap_M_ = cap_S_ = 0;
  }

  IlqrConfig config_;
  VehicleParam vehicle_param_;
  double horizon_ = 0.0;
  double delta_t_ = 0.0;
  int num_of_knots_ = 0;
  int device_ = 0;
  std::vector<Cost> cost_;
  int last_status_ = -1;
  int last_iterations_ = 0;
  cilqr_handle* handle_ = nullptr;
  int cap_N_ = 0, cap_M_ = 0, cap_S_ = 0;
};

}  // namespace planning
