// Stand-in for <opencv2/opencv.hpp>: exactly what algorithm/ilqr/corridor.cc uses -- cv::Point2f and the two
// cv::convexHull overloads -- declared here and DEFINED in oracle/ref_corridor_wrapper.cc by calling the hull
// restatement of oracle/corridor_oracle.c, which is itself pinned index for index against the real OpenCV
// (python cv2, tests/test_corridor_oracle.py).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <vector>

namespace cv {
struct Point2f {
  float x, y;
  Point2f() : x(0), y(0) {}
  Point2f(float x_, float y_) : x(x_), y(y_) {}
  Point2f(double x_, double y_) : x(static_cast<float>(x_)), y(static_cast<float>(y_)) {}
  Point2f(int x_, int y_) : x(static_cast<float>(x_)), y(static_cast<float>(y_)) {}
};
inline Point2f operator-(const Point2f& a, const Point2f& b) { return Point2f(a.x - b.x, a.y - b.y); }
// convexHull(points, hull, clockwise, returnPoints): an int hull receives indices; a Point2f hull receives points
// (OpenCV forces returnPoints for a fixed-type non-int output, imgproc/convhull.cpp)
void convexHull(const std::vector<Point2f>& points, std::vector<int>& hull, bool clockwise, bool returnPoints);
void convexHull(const std::vector<Point2f>& points, std::vector<Point2f>& hull, bool clockwise, bool returnPoints);
}  // namespace cv
