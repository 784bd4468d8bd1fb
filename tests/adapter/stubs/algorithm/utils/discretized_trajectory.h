// Stand-in for algorithm/utils/discretized_trajectory.h:26-93 (TrajectoryPoint + container).
#pragma once
#include <utility>
#include <vector>
namespace planning {
struct TrajectoryPoint {
  double time = 0.0;
  double s = 0.0;
  double x = 0.0;
  double y = 0.0;
  double theta = 0.0;
  double kappa = 0.0;
  double velocity = 0.0;
  double a = 0.0;
  double jerk = 0.0;
  double delta = 0.0;
  double delta_rate = 0.0;
  double left_bound = 0.0;
  double right_bound = 0.0;
};
class DiscretizedTrajectory {
 public:
  typedef std::vector<TrajectoryPoint> DataType;
  DiscretizedTrajectory() = default;
  explicit DiscretizedTrajectory(const std::vector<TrajectoryPoint>& points) : trajectory_(points) {}
  inline const DataType& trajectory() const { return trajectory_; }
  bool empty() const { return trajectory_.empty(); }
 protected:
  std::vector<TrajectoryPoint> trajectory_;
};
}  // namespace planning
