// Stand-in for algorithm/utils/environment.h:24-88: only the four members planning::Corridor uses
// (QueryStaticObstaclesPoints, QueryDynamicObstaclesPoints, left/right_road_barrier), same names and
// meaning (utils/environment.cpp:163-194).  Obstacles are stored as their corner points; dynamic ones as
// (time, points) samples, the query takes the first sample whose time exceeds the argument
// (std::upper_bound with kMathEpsilon, environment.cpp:147-151).
#pragma once
#include <memory>
#include <utility>
#include <vector>
#include "algorithm/math/line_segment2d.h"
#include "algorithm/math/polygon2d.h"
#include "algorithm/utils/discretized_trajectory.h"
namespace planning {
class Environment {
 public:
  using DynamicObstaclePoints = std::vector<std::pair<double, std::vector<math::Vec2d>>>;
  bool QueryDynamicObstaclesPoints(const double time, std::vector<math::Vec2d>* const points,
                                   const bool /*is_multiple_sample*/ = false) {
    if (points == nullptr) return false;
    for (const auto& ob : dynamic_) {
      if (ob.front().first > time + 1e-10 || ob.back().first < time - 1e-10) continue;
      size_t i = 0;
      while (i + 1 < ob.size() && !(time < ob[i].first + 1e-10)) ++i;
      points->insert(points->end(), ob[i].second.begin(), ob[i].second.end());
    }
    return true;
  }
  bool QueryStaticObstaclesPoints(std::vector<math::Vec2d>* const points, const bool /*is_multiple_sample*/ = false) {
    if (points == nullptr) return false;
    points->insert(points->end(), static_.begin(), static_.end());
    return true;
  }
  const std::vector<math::Vec2d>& left_road_barrier() { return left_; }
  const std::vector<math::Vec2d>& right_road_barrier() { return right_; }
  // the accessors planning::DpPlanner reads (utils/environment.h:30-33,45-59)
  using DynamicObstacle = std::vector<std::pair<double, math::Polygon2d>>;
  std::vector<math::Polygon2d>& obstacles() { return obstacles_; }
  std::vector<DynamicObstacle>& dynamic_obstacles() { return dynamic_obstacles_; }
  const DiscretizedTrajectory& reference() const { return reference_; }
  std::vector<math::Polygon2d> obstacles_;
  std::vector<DynamicObstacle> dynamic_obstacles_;
  DiscretizedTrajectory reference_;
  std::vector<math::Vec2d> static_, left_, right_;
  std::vector<DynamicObstaclePoints> dynamic_;
};
using Env = std::shared_ptr<Environment>;
}  // namespace planning
