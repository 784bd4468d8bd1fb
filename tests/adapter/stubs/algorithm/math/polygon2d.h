// Stand-in for algorithm/math/polygon2d.h:30-66 (constructor from points + points()).
#pragma once
#include <vector>
#include "algorithm/math/line_segment2d.h"
namespace planning { namespace math {
class Polygon2d {
 public:
  Polygon2d() = default;
  explicit Polygon2d(std::vector<Vec2d> points) : points_(std::move(points)) {}
  const std::vector<Vec2d>& points() const { return points_; }
 private:
  std::vector<Vec2d> points_;
};
}  // namespace math
}  // namespace planning
