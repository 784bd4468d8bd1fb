// Stand-in for algorithm/math/line_segment2d.h:52-64 and vec2d.h (start()/end() accessors only).
#pragma once
namespace planning { namespace math {
class Vec2d {
 public:
  Vec2d() = default;
  Vec2d(double x, double y) : x_(x), y_(y) {}
  double x() const { return x_; }
  double y() const { return y_; }
 private:
  double x_ = 0.0, y_ = 0.0;
};
class LineSegment2d {
 public:
  LineSegment2d() = default;
  LineSegment2d(const Vec2d& start, const Vec2d& end) : start_(start), end_(end) {}
  const Vec2d& start() const { return start_; }
  const Vec2d& end() const { return end_; }
 private:
  Vec2d start_, end_;
};
}  // namespace math
}  // namespace planning
