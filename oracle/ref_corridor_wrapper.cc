// ref_corridor_wrapper.cc -- extern "C" access to the REFERENCE'S OWN Corridor (algorithm/ilqr/corridor.cc),
// compiled unmodified against the stand-ins of oracle/ref_stubs: Eigen-lite, no-op ROS / visualization, and an
// OpenCV header whose cv::convexHull is defined HERE by the hull restatement of oracle/corridor_oracle.c -- the one
// piece that is pinned separately, index for index, against the real OpenCV through python cv2.  With it the
// arithmetic around the hulls (AddCorridorPoints, the sphere flip, visible-vertex planes, dual points, polygon and
// half-planes, lane sampling) is the reference's own.  TEST INFRASTRUCTURE ONLY; contains no reference code.
#include <memory>
#include <vector>

#define private public
#include "algorithm/ilqr/corridor.h"
#undef private
#include "algorithm/visualization/plot.h"
#include "corridor_oracle.h"

using namespace planning;

namespace cv {
void convexHull(const std::vector<Point2f>& points, std::vector<int>& hull, bool clockwise, bool) {
  std::vector<float> p(points.size() * 2 + 2);
  for (size_t i = 0; i < points.size(); ++i) {
    p[2 * i] = points[i].x;
    p[2 * i + 1] = points[i].y;
  }
  std::vector<int> idx(points.size() + 1);
  const int n = corr_convex_hull_f32(p.data(), (int)points.size(), clockwise ? 1 : 0, idx.data());
  hull.assign(idx.begin(), idx.begin() + n);
}
void convexHull(const std::vector<Point2f>& points, std::vector<Point2f>& hull, bool clockwise, bool) {
  std::vector<int> idx;
  convexHull(points, idx, clockwise, false);
  hull.clear();
  for (int i : idx) hull.push_back(points[i]);
}
}  // namespace cv

extern "C" {

// Corridor::AddCorridorPoints + Corridor::BuildCorridor for one knot: obstacle points [n][2] in, constraints
// [cap][3] and polygon [cap][2] out; returns the number of planes, or -1 when BuildCorridor returns false, -2 on
// capacity overflow.
int ref_build_corridor(double x, double y, double theta, int n, const double* pts, double* constraints,
                       double* polygon, int cap) {
  CorridorConfig config;
  Corridor c(config, Env());
  TrajectoryPoint tp;
  tp.x = x; tp.y = y; tp.theta = theta;
  std::vector<math::Vec2d> points;
  for (int i = 0; i < n; ++i) points.emplace_back(pts[2 * i], pts[2 * i + 1]);
  c.AddCorridorPoints(tp, &points);
  Constraints cons;
  ConvexPolygon poly;
  if (!c.BuildCorridor(x, y, points, &cons, &poly)) return -1;
  if ((int)cons.size() > cap) return -2;
  for (size_t i = 0; i < cons.size(); ++i) {
    constraints[3 * i] = cons[i][0]; constraints[3 * i + 1] = cons[i][1]; constraints[3 * i + 2] = cons[i][2];
    polygon[2 * i] = poly[i][0]; polygon[2 * i + 1] = poly[i][1];
  }
  return (int)cons.size();
}

// Corridor::LaneBoundarySample + HalfPlaneConstraint as CalLeft/RightLaneConstraints combine them
// (corridor.cc:265-331); boundary [n][2] -> out [cap][7]; returns the number of segments, -1 / -2 as above
int ref_lane_constraints(int n, const double* boundary, int is_left, double* out, int cap) {
  CorridorConfig config;
  Corridor c(config, Env());
  std::vector<math::Vec2d> b;
  for (int i = 0; i < n; ++i) b.emplace_back(boundary[2 * i], boundary[2 * i + 1]);
  const std::vector<math::Vec2d> s = c.LaneBoundarySample(b);
  if (s.size() < 2) return -1;
  if ((int)s.size() - 1 > cap) return -2;
  for (size_t i = 1; i < s.size(); ++i) {
    const math::Vec2d& st = is_left ? s[i] : s[i - 1];
    const math::Vec2d& en = is_left ? s[i - 1] : s[i];
    const math::LineSegment2d seg(st, en);
    const Eigen::Vector3d h = c.HalfPlaneConstraint(st, en);
    double* o = out + (i - 1) * 7;
    o[0] = h[0]; o[1] = h[1]; o[2] = h[2];
    o[3] = seg.start().x(); o[4] = seg.start().y(); o[5] = seg.end().x(); o[6] = seg.end().y();
  }
  return (int)s.size() - 1;
}

}  // extern "C"

namespace planning {
namespace visualization {
Color Color::Grey, Color::Magenta, Color::White, Color::Cyan;
}
}  // namespace planning
