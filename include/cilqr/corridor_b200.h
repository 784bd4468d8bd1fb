// corridor_b200.h -- header-compatible replacement of planning::Corridor
// (reference: algorithm/ilqr/corridor.h:27-91, algorithm/ilqr/corridor.cc) that builds the safe
// corridor on a B200 through the C ABI of include/cilqr_b200.h.
//
// Drop-in use inside the reference tree: include this header instead of
// "algorithm/ilqr/corridor.h" in algorithm/planner/trajectory_planner.h, drop algorithm/ilqr/corridor.cc
// (and with it the OpenCV dependency of the planner library) from CMakeLists.txt.  The call sites
// trajectory_planner.cpp:25 (member initialiser), :49-57 (Plan) and planning_node.cc:87-103
// (convex polygons / points_for_corridors for plotting) compile
// unchanged.
//
// The environment queries stay on the host exactly as in the reference (BuildCorridorConstraints,
// corridor.cc:56-87: QueryStaticObstaclesPoints once, QueryDynamicObstaclesPoints per knot at pt.time);
// this header packs the per-knot point clouds into the flat arrays of the ABI (batch of one trajectory),
// calls cilqr_corridor_batch / cilqr_lane_constraints, and unpacks.  It contains no hull arithmetic.
// Failure (any knot's code != 0, a lane boundary with fewer than two sampled points, a CUDA error, no
// device) makes Plan() return false, like the reference.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <utility>
#include <vector>

#include <Eigen/Core>

#include "algorithm/math/line_segment2d.h"
#include "algorithm/params/planner_config.h"
#include "algorithm/utils/discretized_trajectory.h"
#include "algorithm/utils/environment.h"
#include "cilqr_b200.h"

namespace planning {

// Same aliases as algorithm/ilqr/corridor.h:18-25.
using ConvexPolygon = std::vector<Eigen::Vector2d>;
using ConvexPolygons = std::vector<ConvexPolygon>;
using Constraints = std::vector<Eigen::Vector3d>;
using CorridorConstraints = std::vector<Constraints>;
using LaneConstraints = std::vector<std::pair<Eigen::Vector3d, math::LineSegment2d>>;

class Corridor {
 public:
  Corridor() = default;
  Corridor(const CorridorConfig& config, const Env& env) : config_(config), env_(env) {}
  Corridor(const Corridor& o) { *this = o; }
  Corridor& operator=(const Corridor& o) {  // configuration is copied, the device handle is not shared
    if (this != &o) {
      Release();
      config_ = o.config_;
      env_ = o.env_;
      device_ = o.device_;
      points_for_corridors_ = o.points_for_corridors_;
    }
    return *this;
  }
  ~Corridor() { Release(); }

  void Init(const CorridorConfig& config, const Env& env) {
    config_ = config;
    env_ = env;
  }
  void set_device(int device) {
    Release();
    device_ = device;
  }

  // corridor.cc:17-54
  bool Plan(const DiscretizedTrajectory& trajectory, CorridorConstraints* const corridor_constraints,
            ConvexPolygons* const convex_polygons, LaneConstraints* const left_lane_constraints,
            LaneConstraints* const right_lane_constraints) {
    if (trajectory.empty()) {
      std::fprintf(stderr, "Corridor failed: Trajectory is empty!\n");
      return false;
    }
    if (corridor_constraints == nullptr || convex_polygons == nullptr || left_lane_constraints == nullptr ||
        right_lane_constraints == nullptr) {
      std::fprintf(stderr, "Corridor failed: Input ptr is nullptr!\n");
      return false;
    }
    if (config_.is_multiple_sample) {
      std::fprintf(stderr, "Corridor failed: is_multiple_sample is not supported by the B200 builder\n");
      return false;
    }
    if (!EnsureHandle()) return false;
    if (!BuildCorridorConstraints(trajectory, corridor_constraints, convex_polygons)) {
      std::fprintf(stderr, "Corridor failed: Safe Corridors Build Failed!\n");
      return false;
    }
    if (!LaneSide(env_->left_road_barrier(), true, left_lane_constraints)) {
      std::fprintf(stderr, "Corridor failed: Left Lane Boundary Constraints Failed!\n");
      return false;
    }
    if (!LaneSide(env_->right_road_barrier(), false, right_lane_constraints)) {
      std::fprintf(stderr, "Corridor failed: Right Lane Boundary Constraints Failed!\n");
      return false;
    }
    return true;
  }

  std::vector<std::vector<math::Vec2d>> points_for_corridors() { return points_for_corridors_; }

 private:
  CilqrCorridorConfig AbiConfig() const {
    CilqrCorridorConfig c;
    cilqr_corridor_default_config(&c);
    c.max_diff_x = config_.max_diff_x;
    c.max_diff_y = config_.max_diff_y;
    c.radius = config_.radius;
    c.max_axis_x = config_.max_axis_x;
    c.max_axis_y = config_.max_axis_y;
    c.lane_segment_length = config_.lane_segment_length;
    return c;
  }

  // corridor.cc:56-87
  bool BuildCorridorConstraints(const DiscretizedTrajectory& trajectory, CorridorConstraints* const cc,
                                ConvexPolygons* const polys) {
    points_for_corridors_.clear();
    cc->clear();
    polys->clear();
    const int K = static_cast<int>(trajectory.trajectory().size());
    std::vector<math::Vec2d> static_points;
    env_->QueryStaticObstaclesPoints(&static_points, false);
    std::vector<std::vector<math::Vec2d>> per_knot(K);
    size_t P_max = 0;
    for (int k = 0; k < K; ++k) {
      per_knot[k] = static_points;
      env_->QueryDynamicObstaclesPoints(trajectory.trajectory()[k].time, &per_knot[k], false);
      P_max = std::max(P_max, per_knot[k].size());
    }
    std::vector<double> traj(static_cast<size_t>(K) * 3), pts(static_cast<size_t>(K) * std::max<size_t>(P_max, 1) * 2, 0.0);
    std::vector<int32_t> cnt(K);
    for (int k = 0; k < K; ++k) {
      const TrajectoryPoint& pt = trajectory.trajectory()[k];
      traj[3 * k] = pt.x;
      traj[3 * k + 1] = pt.y;
      traj[3 * k + 2] = pt.theta;
      cnt[k] = static_cast<int32_t>(per_knot[k].size());
      for (size_t i = 0; i < per_knot[k].size(); ++i) {
        pts[(static_cast<size_t>(k) * P_max + i) * 2] = per_knot[k][i].x();
        pts[(static_cast<size_t>(k) * P_max + i) * 2 + 1] = per_knot[k][i].y();
      }
    }
    // a convex polygon around a knot has at most as many edges as points in its window
    const int M_max = static_cast<int>(std::min<size_t>(P_max + 8, 128));
    std::vector<double> planes(static_cast<size_t>(K) * M_max * 3), poly(static_cast<size_t>(K) * M_max * 2);
    std::vector<int32_t> pcnt(K), code(K);
    CilqrCorridorConfig cfg = AbiConfig();
    cfg.point_cap = static_cast<int32_t>(std::min<size_t>(std::max<size_t>(P_max + 16, 64), 250));
    CilqrCorridorIn in;
    in.B = 1; in.K = K; in.P_max = static_cast<int32_t>(P_max); in.M_max = M_max;
    in.traj = traj.data(); in.obs_points = pts.data(); in.obs_cnt = cnt.data();
    CilqrCorridorOut out;
    out.corridor = planes.data(); out.corridor_cnt = pcnt.data(); out.polygon = poly.data(); out.code = code.data();
    const int rc = cilqr_corridor_batch(handle_, &cfg, &in, &out);
    if (rc != CILQR_OK) {
      std::fprintf(stderr, "cilqr_b200: %s (%s)\n", cilqr_strerror(rc), cilqr_last_cuda_error(handle_));
      return false;
    }
    for (int k = 0; k < K; ++k) {
      // points_for_corridors_ (visualisation, planning_node.cc:88,103): the knot's obstacle points
      // followed by the eight box points of AddCorridorPoints (corridor.cc:89-120)
      AppendBoxPoints(trajectory.trajectory()[k], &per_knot[k]);
      points_for_corridors_.push_back(per_knot[k]);
      if (code[k] != CILQR_CORR_OK) {
        std::fprintf(stderr, "Corridor failed: BuildCorridor Failed! (knot %d, code %d)\n", k, code[k]);
        return false;
      }
      Constraints cons;
      ConvexPolygon pg;
      for (int m = 0; m < pcnt[k]; ++m) {
        const double* p = &planes[(static_cast<size_t>(k) * M_max + m) * 3];
        cons.push_back(Eigen::Vector3d(p[0], p[1], p[2]));
        const double* q = &poly[(static_cast<size_t>(k) * M_max + m) * 2];
        pg.push_back(Eigen::Vector2d(q[0], q[1]));
      }
      cc->push_back(cons);
      polys->push_back(pg);
    }
    return true;
  }

  void AppendBoxPoints(const TrajectoryPoint& pt, std::vector<math::Vec2d>* const points) const {
    const double c = std::cos(pt.theta), s = std::sin(pt.theta);
    const double dx1 = c * config_.max_axis_x, dy1 = s * config_.max_axis_x;
    const double dx2 = s * config_.max_axis_y, dy2 = -c * config_.max_axis_y;
    const double cx[4] = {pt.x + dx1 + dx2, pt.x + dx1 - dx2, pt.x - dx1 - dx2, pt.x - dx1 + dx2};
    const double cy[4] = {pt.y + dy1 + dy2, pt.y + dy1 - dy2, pt.y - dy1 - dy2, pt.y - dy1 + dy2};
    for (int i = 0; i < 4; ++i) {
      points->emplace_back(cx[i], cy[i]);
      points->emplace_back(cx[(i + 1) % 4], cy[(i + 1) % 4]);
    }
  }

  // corridor.cc:265-307
  bool LaneSide(const std::vector<math::Vec2d>& boundary, bool is_left, LaneConstraints* const lane) {
    lane->clear();
    if (boundary.empty()) return false;
    const int n = static_cast<int>(boundary.size());
    std::vector<double> b(static_cast<size_t>(n) * 2);
    for (int i = 0; i < n; ++i) {
      b[2 * i] = boundary[i].x();
      b[2 * i + 1] = boundary[i].y();
    }
    const int S_max = n;  // a sampled polyline has fewer segments than the boundary has points
    std::vector<double> seg(static_cast<size_t>(S_max) * 7);
    int32_t count = 0;
    const CilqrCorridorConfig cfg = AbiConfig();
    const int rc = cilqr_lane_constraints(handle_, &cfg, 1, n, S_max, is_left ? 1 : 0, b.data(), seg.data(), &count);
    if (rc != CILQR_OK) {
      std::fprintf(stderr, "cilqr_b200: %s (%s)\n", cilqr_strerror(rc), cilqr_last_cuda_error(handle_));
      return false;
    }
    if (count < 1) return false;
    for (int i = 0; i < count; ++i) {
      const double* p = &seg[static_cast<size_t>(i) * 7];
      lane->push_back(std::make_pair(Eigen::Vector3d(p[0], p[1], p[2]),
                                     math::LineSegment2d(math::Vec2d(p[3], p[4]), math::Vec2d(p[5], p[6]))));
    }
    return true;
  }

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

  CorridorConfig config_;
  Env env_;
  std::vector<std::vector<math::Vec2d>> points_for_corridors_;
  cilqr_handle* handle_ = nullptr;
  int device_ = 0;
};

}  // namespace planning
