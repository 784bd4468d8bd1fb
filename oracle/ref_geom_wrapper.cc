// ref_geom_wrapper.cc -- extern "C" access to the REFERENCE'S OWN geometry / reference-line classes, compiled from
// the sources where they lie under /root/reference (oracle/Makefile target `_ref`), so that the restatements in
// oracle/ can be pinned against the reference itself for the pieces that build without ROS / Eigen / OpenCV:
//   algorithm/math/{vec2d,aabox2d,box2d,line_segment2d,polygon2d,math_utils}.cpp
//   algorithm/utils/discretized_trajectory.cpp, algorithm/utils/discrete_points_math.cc
// TEST INFRASTRUCTURE ONLY.  This file contains no reference code: it only calls it.
#include <utility>
#include <vector>

#include "algorithm/math/box2d.h"
#include "algorithm/math/line_segment2d.h"
#include "algorithm/math/math_utils.h"
#include "algorithm/math/polygon2d.h"
#include "algorithm/utils/discrete_points_math.h"
#include "algorithm/utils/discretized_trajectory.h"

using planning::DiscretizedTrajectory;
using planning::TrajectoryPoint;
namespace math = planning::math;

static DiscretizedTrajectory make_line(int R, const double* ref) {
  std::vector<TrajectoryPoint> pts(R);
  for (int i = 0; i < R; ++i) {
    const double* r = ref + (size_t)i * 7;
    pts[i].s = r[0]; pts[i].x = r[1]; pts[i].y = r[2]; pts[i].theta = r[3]; pts[i].kappa = r[4];
    pts[i].left_bound = r[5]; pts[i].right_bound = r[6];
  }
  return DiscretizedTrajectory(pts);
}

extern "C" {

double ref_normalize_angle(double a) { return math::NormalizeAngle(a); }
double ref_slerp(double a0, double t0, double a1, double t1, double t) { return math::slerp(a0, t0, a1, t1, t); }

// LineSegment2d::DistanceTo (line_segment2d.cpp:61-75), the distance FindNeastLaneSegment minimises
double ref_segment_distance(double x0, double y0, double x1, double y1, double px, double py) {
  return math::LineSegment2d(math::Vec2d(x0, y0), math::Vec2d(x1, y1)).DistanceTo(math::Vec2d(px, py));
}

// DiscretizedTrajectory::EvaluateStation for n stations -> out [n][7] (s, x, y, theta, kappa, left, right)
void ref_evaluate_stations(int R, const double* ref, int n, const double* stations, double* out) {
  const DiscretizedTrajectory line = make_line(R, ref);
  for (int i = 0; i < n; ++i) {
    const TrajectoryPoint p = line.EvaluateStation(stations[i]);
    double* o = out + (size_t)i * 7;
    o[0] = p.s; o[1] = p.x; o[2] = p.y; o[3] = p.theta; o[4] = p.kappa; o[5] = p.left_bound; o[6] = p.right_bound;
  }
}

// GetProjection for n points [n][2] -> out [n][2] (s, l); GetCartesian for n (s, l) pairs -> out [n][2]
void ref_get_projections(int R, const double* ref, int n, const double* xy, double* out) {
  const DiscretizedTrajectory line = make_line(R, ref);
  for (int i = 0; i < n; ++i) {
    const math::Vec2d sl = line.GetProjection(math::Vec2d(xy[2 * i], xy[2 * i + 1]));
    out[2 * i] = sl.x();
    out[2 * i + 1] = sl.y();
  }
}
void ref_get_cartesians(int R, const double* ref, int n, const double* sl, double* out) {
  const DiscretizedTrajectory line = make_line(R, ref);
  for (int i = 0; i < n; ++i) {
    const math::Vec2d p = line.GetCartesian(sl[2 * i], sl[2 * i + 1]);
    out[2 * i] = p.x();
    out[2 * i + 1] = p.y();
  }
}

// Polygon2d::HasOverlap(Box2d) with the box built exactly as Environment::CheckOptimizationCollision builds it
// (environment.cpp:101-112): AABox2d({-r, -r}, {r, r}), Shift(centre), Box2d(aabox)
int ref_polygon_overlaps_disc_box(int nv, const double* pts, double cx, double cy, double radius) {
  std::vector<math::Vec2d> p;
  for (int i = 0; i < nv; ++i) p.emplace_back(pts[2 * i], pts[2 * i + 1]);
  const math::Polygon2d poly(p);
  math::AABox2d box({-radius, -radius}, {radius, radius});
  box.Shift({cx, cy});
  return poly.HasOverlap(math::Box2d(box)) ? 1 : 0;
}
int ref_polygon_is_point_in(int nv, const double* pts, double x, double y) {
  std::vector<math::Vec2d> p;
  for (int i = 0; i < nv; ++i) p.emplace_back(pts[2 * i], pts[2 * i + 1]);
  return math::Polygon2d(p).IsPointIn(math::Vec2d(x, y)) ? 1 : 0;
}
int ref_disc_box_is_point_in(double cx, double cy, double radius, double x, double y) {
  math::AABox2d box({-radius, -radius}, {radius, radius});
  box.Shift({cx, cy});
  return math::Box2d(box).IsPointIn(math::Vec2d(x, y)) ? 1 : 0;
}

// DiscretePointsMath::ComputePathProfile -> speeds, accelerations, kappas (each [n]); returns 1 on success
int ref_compute_path_profile(double dt, int n, const double* xy, double* speeds, double* accel, double* kappas) {
  std::vector<std::pair<double, double>> pts;
  for (int i = 0; i < n; ++i) pts.emplace_back(xy[2 * i], xy[2 * i + 1]);
  std::vector<double> h, s, v, a, k;
  if (!planning::DiscretePointsMath::ComputePathProfile(dt, pts, &h, &s, &v, &a, &k)) return 0;
  for (int i = 0; i < n; ++i) {
    speeds[i] = v[i];
    accel[i] = a[i];
    kappas[i] = k[i];
  }
  return 1;
}

}  // extern "C"
