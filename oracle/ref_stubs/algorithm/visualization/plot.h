// No-op stand-in for the reference's algorithm/visualization/plot.h (which needs ROS, visualization_msgs and Eigen),
// placed FIRST on the include path when oracle/Makefile compiles the reference's own utils/environment.cpp and
// planner/dp_planner.cpp into oracle/_ref.  Only Environment::Visualize() and Corridor::CheckLaneConstraints() -- neither
// called on the planning path -- use it.  Everything else in those translation units is the reference's unmodified source.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <string>
#include <vector>

#include <ros/ros.h>  // the real plot.h brings the ROS logging macros into its includers

#include "algorithm/math/line_segment2d.h"
#include "algorithm/math/polygon2d.h"
#include "algorithm/math/vec2d.h"

namespace planning {
namespace visualization {

class Color {
 public:
  Color() = default;
  Color(double, double, double) {}
  static Color Grey, Magenta, White, Cyan;
  static Color fromHSV(int, double, double) { return Color(); }
  void set_alpha(double) {}
};

using Vector = std::vector<double>;
inline void Plot(const Vector&, const Vector&, double = 0.1, Color = Color(), int = -1, const std::string& = "") {}
inline void PlotPolygon(const math::Polygon2d&, double = 0.1, Color = Color(), int = -1, const std::string& = "") {}
inline void PlotPoint(const math::Vec2d&, double = 0.1, Color = Color(), int = -1, const std::string& = "") {}
inline void PlotLineSegment(const math::LineSegment2d&, double = 0.1, Color = Color(), int = -1, const std::string& = "") {}
inline void Trigger() {}

}  // namespace visualization
}  // namespace planning
