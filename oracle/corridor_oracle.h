/*
 * corridor_oracle.h -- CPU restatement (plain C99) of the safe-corridor builder of mpt0816/Cilqr:
 * Corridor::Plan (algorithm/ilqr/corridor.cc:17-54) and everything it executes: the per-knot
 * BuildCorridor (corridor.cc:122-263), AddCorridorPoints (:89-120), the lane sampling and
 * half-planes (:265-331), and -- because the reference calls it three times per knot --
 * cv::convexHull on CV_32F points (OpenCV imgproc, convhull.cpp; un-vendored dependency, the
 * reference pins no version: `find_package(OpenCV REQUIRED)`, CMakeLists.txt).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rule as cilqr_oracle.h).
 *
 * PINNING, in two steps that together cover the whole path:
 *  (1) the convex hull restatement (corr_convex_hull_f32) is pinned index-for-index against the real OpenCV
 *      (python cv2 4.13.0 in this image) by tests/test_corridor_oracle.py on random, gridded, duplicated and
 *      collinear inputs, and through committed cv2 outputs (tests/golden/corridor_hull_v1.npz);
 *  (2) everything around the hulls is pinned against the REFERENCE'S OWN corridor.cc, compiled unmodified into
 *      oracle/_ref (Makefile target `_ref`, wrapper oracle/ref_corridor_wrapper.cc) against an Eigen stand-in,
 *      no-op ROS / visualization headers and an OpenCV header whose cv::convexHull is (1):
 *      tests/test_reference_pins.py compares AddCorridorPoints + BuildCorridor on > 1400 knots, the committed
 *      fixture tests/golden/corridor_golden_v1.npz and the lane constraints -- all bit-identical.
 * Additionally cross-checked against an independent NumPy float32/float64 restatement that calls cv2.convexHull
 * itself (oracle/corridor_numpy.py).
 */
#ifndef CORRIDOR_ORACLE_H_
#define CORRIDOR_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* CorridorConfig, algorithm/params/planner_config.h:75-86 (is_multiple_sample = false only). */
typedef struct corr_config {
  double max_diff_x, max_diff_y; /* 25, 25 */
  double radius;                 /* 150 */
  double max_axis_x, max_axis_y; /* 10, 10 */
  double lane_segment_length;    /* 5 */
} corr_config;

void corr_default_config(corr_config* c);

/* per-knot result codes */
enum {
  CORR_OK = 0,
  CORR_E_NO_POINTS = 1,     /* corridor.cc:127-130 */
  CORR_E_FEW_POINTS = 2,    /* fewer than 4 flipped points, corridor.cc:179-182 */
  CORR_E_ORIGIN_UB = 3,     /* the origin is a hull vertex and the reference indexes filterd_points out of
                               range (corridor.cc:193-209 with more than one zero slot in flipData) */
  CORR_E_CAPACITY = 4       /* more constraints than the caller's row pitch */
};

/* cv::convexHull(points, hull, clockwise, returnPoints=false) for n CV_32F points [n][2].
 * Writes the hull's point indices to hull_idx (capacity n) and returns their number. */
int corr_convex_hull_f32(const float* pts, int n, int clockwise, int* hull_idx);

/* Corridor::AddCorridorPoints (corridor.cc:89-120), is_multiple_sample = false: appends the 8 box
 * points (each corner twice) to points[*n ...]. */
void corr_add_corridor_points(const corr_config* cfg, double x, double y, double theta, double* points, int* n);

/* Corridor::BuildCorridor (corridor.cc:122-263).  points [n][2]; constraints [cap][3] (a,b,c:
 * a x + b y < c, corridor.h:20), polygon [cap][2].  Returns a CORR_* code. */
int corr_build_corridor(const corr_config* cfg, double origin_x, double origin_y, const double* points, int n,
                        double* constraints, double* polygon, int cap, int* count);

/* BuildCorridorConstraints (corridor.cc:56-87) for one trajectory, the environment queries already
 * done by the caller: traj [K][3] (x, y, theta); obs_points [K][P_max][2] with obs_cnt[K] valid
 * (static obstacle points first, then the dynamic ones at pt.time, as QueryStatic/DynamicObstaclesPoints
 * return them, environment.cpp:163-194).  Outputs constraints [K][M_max][3], cnt [K], polygon
 * [K][M_max][2] (may be NULL), code [K].  Returns the first non-zero code (the reference returns false
 * at that knot) or 0. */
int corr_plan(const corr_config* cfg, int K, const double* traj, const double* obs_points, const int* obs_cnt,
              int P_max, int M_max, double* constraints, int* cnt, double* polygon, int* code);

/* CalLeft/RightLaneConstraints (corridor.cc:265-307) + LaneBoundarySample (:309-322) +
 * HalfPlaneConstraint (:324-331).  boundary [n][2]; out [cap][7] = a,b,c,x0,y0,x1,y1 (segment start,
 * end as constructed).  Returns the number of segments, or -1 when fewer than 2 sampled points
 * (the reference returns false), or -2 on capacity overflow. */
int corr_lane_constraints(const corr_config* cfg, const double* boundary, int n, int is_left, double* out, int cap);

#ifdef __cplusplus
}
#endif
#endif
