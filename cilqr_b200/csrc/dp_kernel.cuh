// dp_kernel.cuh -- batched coarse DP planner for sm_100a.
//
// Replaces DpPlanner::Plan of mpt0816/Cilqr (algorithm/planner/dp_planner.cpp:135-281) for B scenarios:
// the 5 x 7 x 10 lattice search with GetCost (:87-133), GetCollisionCost (:40-85), InterpolateLinearly
// (:283-320), GetLateralOffset (dp_planner.h:83-92), the reference-line queries of DiscretizedTrajectory
// (utils/discretized_trajectory.cpp: EvaluateStation :110-121, GetCartesian :192-196, GetProjection :156-190,
// slerp math_utils.h:208-225), the collision checks of Environment (utils/environment.cpp:51-141 with
// Polygon2d::HasOverlap(Box2d) / IsPointIn, math/polygon2d.cpp:120-165) and ComputePathProfile
// (utils/discrete_points_math.cc:27-176).
//
// Mapping: ONE CTA = ONE SCENARIO.  A layer of the lattice is 70 x 70 independent (parent, child)
// transitions, each a walk of 16-17 path points with an early exit at the first collision; the threads of the
// CTA take transitions round-robin and write their cost into a shared 70 x 70 table, then 70 threads pick
// every child's best parent in the reference's (s, l) parent order with its strict '<' (first minimum wins).
// The lattice (350 cells) and the table live in shared memory; the environment (centre line, road barrier:
// shared by the batch, L2 resident; the scenario's obstacle polygons) is read from global memory.
//
// This translation unit is compiled with -fmad=false: every double expression is evaluated with the
// reference's operation order and roundings.  cos/sin/atan/fmod/hypot are CUDA's (last-ulp differences from
// glibc on rare arguments).
#pragma once

#ifndef DP_HOST_EMUL  // tools/dp_host_emul.cc runs this code on the CPU (development aid)
#include <cuda_runtime.h>
#endif
#include <float.h>
#include <math.h>
#include <stdint.h>

#include <vector>

namespace dp {

constexpr int NT = 5, NS = 7, NL = 10, NP = NS * NL;  // dp_planner.h:27-29
constexpr double kMathEps = 1e-10;                    // math::kMathEpsilon
constexpr double kDpEps = 1e-3;                       // the file-local kMathEpsilon of dp_planner.cpp:29
constexpr double kPi = 3.14159265358979323846;

struct Lattice {  // DpPlanner::DpPlanner, dp_planner.cpp:31-38, computed on the host with the same expressions
  double unit_time, time_[NT], station_[NS], lateral_[NL - 1], safe_margin;
  double radius, f2x, r2x;  // VehicleParam(), vehicle_param.h:80-85
  int nseg[NT];             // InterpolateLinearly's segment count per layer, :287-298
  int K;
};

// DpPlanner::DpPlanner (dp_planner.cpp:31-38), math::LinSpaced (math_utils.h:244-254), VehicleParam()
// (vehicle_param.h:80-85) and the segment counts of InterpolateLinearly (dp_planner.cpp:287-298); host side
inline void make_lattice(double tf, double delta_t, double max_velocity, double width, double wheel_base,
                         double front_hang_length, double rear_hang_length, Lattice* L) {
  L->unit_time = tf / NT;
  {
    const double step = (tf - L->unit_time) / (NT - 1);
    for (int i = 0; i < NT; ++i) L->time_[i] = L->unit_time + step * i;
  }
  {
    const double step = (L->unit_time * max_velocity - 0) / (NS - 1);
    for (int i = 0; i < NS; ++i) L->station_[i] = 0 + step * i;
  }
  {
    const double step = (1.0 - 0) / (NL - 1 - 1);
    for (int i = 0; i < NL - 1; ++i) L->lateral_[i] = 0 + step * i;
  }
  L->safe_margin = width / 2 * 1.5;
  const double length = wheel_base + rear_hang_length + front_hang_length;
  L->radius = hypot(0.25 * length, 0.5 * width);
  L->r2x = 0.25 * length - rear_hang_length;
  L->f2x = 0.75 * length - rear_hang_length;
  L->K = 0;
  for (int k = 0; k < NT; ++k) {
    int nseg = 0;
    for (double t = 0.0; t < tf + delta_t - kMathEps; t += delta_t) {
      if (k == 0) {
        if (t > 0.0 - kDpEps && t < L->unit_time + kDpEps) ++nseg;
      } else {
        if (t > L->time_[k] - L->unit_time + kMathEps && t < L->time_[k] + kMathEps) ++nseg;
      }
    }
    L->nseg[k] = nseg;
    L->K += nseg;
  }
}

struct Args {
  int B, R, NB, V, n_static, n_dyn, T;
  double tf, delta_t, nominal_velocity, w_obstacle, w_lateral, w_lateral_change, w_lateral_velocity_change,
      w_longitudinal_velocity_bias, w_longitudinal_velocity_change, wheel_base;
  double ref_s0, ref_inv_ds;  // first station and 1 / mean spacing: a guess for the station search
  Lattice lat;
  const double* ref;      // [R][7]
  const double* barrier;  // [NB][2]
  // uniform grid over the barrier points (built by the library from `barrier`): cell (ix, iy) holds the sorted-
  // barrier indices grid_idx[grid_start[iy * gnx + ix] .. grid_start[iy * gnx + ix + 1])
  double gx0, gy0, ginv;
  int gnx, gny;
  const int* grid_start;
  const int* grid_idx;
  const double* grid_xy;  // [NB][2]: the points in grid order (grid_xy[k] = barrier[grid_idx[k]]), one load level less per point
  const double* start;    // [B][3]
  const double* static_poly;
  const int* static_nv;
  const double* dyn_time;
  const int* dyn_samples;
  const double* dyn_poly;
  const int* dyn_nv;
  double* trajectory;  // [B][K][13] or nullptr
  double* coarse;      // [B][K][6] or nullptr
  double* xytheta;     // [B][K][3] or nullptr
  int* ok;
  double* cost;
  double* waypoints;
  int use_sample_bounds;  // the bounds of every sample of every dynamic obstacle fit the CTA's shared memory
};

// Host side: buckets the barrier points into square cells of a quarter of the collision box's side (a box reaches
// at most 5 x 5 cells; a row of cells is one contiguous range of the CSR list).  cell(p) = floor((p - origin) * ginv) is monotone in p, and the kernel computes the
// cells of a query with the same expression, so every point the box test can accept lies in a scanned cell.
inline void build_grid(const double* barrier, int NB, double half, Args* a, std::vector<int>* start,
                       std::vector<int>* idx, std::vector<double>* xy) {
  const double cs = (2.0 * half + 1e-6) / 4.0;
  double minx = 0, maxx = 0, miny = 0, maxy = 0;
  for (int i = 0; i < NB; ++i) {
    const double x = barrier[2 * i], y = barrier[2 * i + 1];
    if (i == 0 || x < minx) minx = x;
    if (i == 0 || x > maxx) maxx = x;
    if (i == 0 || y < miny) miny = y;
    if (i == 0 || y > maxy) maxy = y;
  }
  a->gx0 = minx - cs;
  a->gy0 = miny - cs;
  a->ginv = 1.0 / cs;
  a->gnx = NB ? (int)floor((maxx - a->gx0) * a->ginv) + 2 : 1;
  a->gny = NB ? (int)floor((maxy - a->gy0) * a->ginv) + 2 : 1;
  const size_t cells = (size_t)a->gnx * a->gny;
  start->assign(cells + 1, 0);
  idx->assign(NB > 0 ? NB : 1, 0);
  std::vector<int> cell(NB > 0 ? NB : 1);
  for (int i = 0; i < NB; ++i) {
    const int ix = (int)floor((barrier[2 * i] - a->gx0) * a->ginv), iy = (int)floor((barrier[2 * i + 1] - a->gy0) * a->ginv);
    cell[i] = iy * a->gnx + ix;
    ++(*start)[cell[i] + 1];
  }
  for (size_t c = 0; c < cells; ++c) (*start)[c + 1] += (*start)[c];
  std::vector<int> fill(start->begin(), start->end() - 1);
  for (int i = 0; i < NB; ++i) (*idx)[fill[cell[i]]++] = i;
  xy->assign(2 * (size_t)(NB > 0 ? NB : 1), 0.0);
  for (int k = 0; k < NB; ++k) {
    (*xy)[2 * (size_t)k] = barrier[2 * (size_t)(*idx)[k]];
    (*xy)[2 * (size_t)k + 1] = barrier[2 * (size_t)(*idx)[k] + 1];
  }
}

struct Cell {
  double cost, current_s;
  int ps, pl;
};

struct RefPoint {
  double s, x, y, theta, kappa, lb, rb;
};

// fmod(x, y) for y > 0, bit-exact: fmod is an exact operation, and for |x| < 2y its result is x itself or
// x -/+ y, a difference that is exactly representable (Sterbenz) -- the general (slow, iterative) routine is only
// needed for larger arguments, which headings never are
// (out of line: libdevice's fmod is ~100 instructions, and normalize_angle is inlined at a dozen sites -- the kernel is
// bound by instruction fetch, profiles/r02_dp_ncu_summary.txt: stall_no_instruction 7 of 15 cycles per issue)
__device__ __noinline__ double fmod_general(double x, double y) { return fmod(x, y); }
__device__ __forceinline__ double fmod_small(double x, double y) {
  const double ax = fabs(x);
  if (ax < y) return x;
  if (ax < 2.0 * y) return copysign(ax - y, x);
  return fmod_general(x, y);
}

// math_utils.cpp:53-59
__device__ __forceinline__ double normalize_angle(double angle) {
  double a = fmod_small(angle + kPi, 2.0 * kPi);
  if (a < 0.0) a += 2.0 * kPi;
  return a - kPi;
}

// math_utils.h:208-225
__device__ __forceinline__ double slerp(double a0, double t0, double a1, double t1, double t) {
  if (fabs(t1 - t0) <= kMathEps) return normalize_angle(a0);
  const double a0_n = normalize_angle(a0);
  const double a1_n = normalize_angle(a1);
  double d = a1_n - a0_n;
  if (d > kPi) {
    d = d - 2 * kPi;
  } else if (d < -kPi) {
    d = d + 2 * kPi;
  }
  const double r = (t - t0) / (t1 - t0);
  const double a = a0_n + d * r;
  return normalize_angle(a);
}

// LinearInterpolateTrajectory, discretized_trajectory.cpp:62-84
__device__ __forceinline__ RefPoint interpolate(const double* p0, const double* p1, double s) {
  RefPoint o;
  const double s0 = p0[0], s1 = p1[0];
  if (fabs(s1 - s0) < kMathEps) {
    o.s = p0[0]; o.x = p0[1]; o.y = p0[2]; o.theta = p0[3]; o.kappa = p0[4]; o.lb = p0[5]; o.rb = p0[6];
    return o;
  }
  const double weight = (s - s0) / (s1 - s0);
  o.s = s;
  o.x = (1 - weight) * p0[1] + weight * p1[1];
  o.y = (1 - weight) * p0[2] + weight * p1[2];
  o.theta = slerp(p0[3], p0[0], p1[3], p1[0], s);
  o.kappa = (1 - weight) * p0[4] + weight * p1[4];
  o.lb = (1 - weight) * p0[5] + weight * p1[5];
  o.rb = (1 - weight) * p0[6] + weight * p1[6];
  return o;
}

// EvaluateStation, :110-121 with QueryLowerBoundStationPoint :34-46.  std::lower_bound's answer (the first
// index whose s is not less than the station) is unique for the sorted line, so it is found from a guess
// (uniform spacing) corrected by stepping -- the same index in 2-3 loads instead of 12.
__device__ __forceinline__ RefPoint evaluate_station(const Args& a, double station) {
  const double* ref = a.ref;
  const int R = a.R;
  int it;
  if (station >= ref[(size_t)(R - 1) * 7]) {
    it = R - 1;
  } else if (station < ref[0]) {
    it = 0;
  } else {
    double g = (station - a.ref_s0) * a.ref_inv_ds;
    it = g < 0.0 ? 0 : (g > (double)(R - 1) ? R - 1 : (int)g);
    while (it > 0 && !(ref[(size_t)(it - 1) * 7] < station)) --it;
    while (it < R && ref[(size_t)it * 7] < station) ++it;
  }
  if (it == 0) it = 1;
  return interpolate(ref + (size_t)(it - 1) * 7, ref + (size_t)it * 7, station);
}

// Polygon2d::IsPointIn, polygon2d.cpp:120-140
__device__ __forceinline__ bool polygon_is_point_in(const double* p, int nv, double minx, double maxx, double miny,
                                                    double maxy, double x, double y) {
  if (x < minx || x > maxx || y < miny || y > maxy) return false;
  int j = nv - 1, c = 0;
#pragma unroll 1
  for (int i = 0; i < nv; ++i) {
    const double xi = p[2 * i], yi = p[2 * i + 1], xj = p[2 * j], yj = p[2 * j + 1];
    if ((yi > y) != (yj > y)) {
      const double side = (xi - x) * (yj - y) - (yi - y) * (xj - x);
      if (yi < yj ? side > 0.0 : side < 0.0) ++c;
    }
    j = i;
  }
  return c & 1;
}

// Box2d::IsPointIn for Box2d(AABox2d): cos = 1, sin = 0 (box2d.cpp:93-105,123-129)
__device__ __forceinline__ bool box_is_point_in(double px, double py, double cx, double cy, double half) {
  const double x0 = px - cx, y0 = py - cy;
  const double dx = fabs(x0 * 1.0 + y0 * 0.0), dy = fabs(-x0 * 0.0 + y0 * 1.0);
  return dx <= half + kMathEps && dy <= half + kMathEps;
}

// a polygon's axis-aligned bounds as Polygon2d::BuildFromPoints computes them (polygon2d.cpp:246-256)
__device__ __noinline__ void polygon_aabb(const double* p, int nv, double* o) {
  double minx = p[0], maxx = p[0], miny = p[1], maxy = p[1];
#pragma unroll 1
  for (int i = 1; i < nv; ++i) {
    minx = fmin(minx, p[2 * i]);
    maxx = fmax(maxx, p[2 * i]);
    miny = fmin(miny, p[2 * i + 1]);
    maxy = fmax(maxy, p[2 * i + 1]);
  }
  o[0] = minx; o[1] = maxx; o[2] = miny; o[3] = maxy;
}

__device__ __forceinline__ bool aabb_disjoint(const double* o, double cx, double cy, double half) {
  const double bminx = cx - half, bmaxx = cx + half, bminy = cy - half, bmaxy = cy + half;
  return bmaxx < o[0] || bminx > o[1] || bmaxy < o[2] || bminy > o[3];
}

// Polygon2d::HasOverlap(const Box2d&), polygon2d.cpp:150-165, after its bounding-box rejection
// (out of line: reached only when the bounding boxes overlap)
__device__ __noinline__ bool polygon_overlaps_box(const double* p, int nv, const double* bb, double cx, double cy, double half) {
  if (aabb_disjoint(bb, cx, cy, half)) return false;
#pragma unroll 1
  for (int i = 0; i < nv; ++i)
    if (box_is_point_in(p[2 * i], p[2 * i + 1], cx, cy, half)) return true;
  if (polygon_is_point_in(p, nv, bb[0], bb[1], bb[2], bb[3], cx + half, cy - half)) return true;  // aabox2d.cpp:63-71
  if (polygon_is_point_in(p, nv, bb[0], bb[1], bb[2], bb[3], cx + half, cy + half)) return true;
  if (polygon_is_point_in(p, nv, bb[0], bb[1], bb[2], bb[3], cx - half, cy + half)) return true;
  if (polygon_is_point_in(p, nv, bb[0], bb[1], bb[2], bb[3], cx - half, cy - half)) return true;
  return false;
}

// std::upper_bound over the barrier's x: first index with x > val
__device__ __forceinline__ int barrier_upper_bound(const double* bar, int NB, double val) {
  int lo = 0, hi = NB;
  while (lo < hi) {
    const int mid = lo + (hi - lo) / 2;
    if (val < bar[(size_t)mid * 2]) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// Environment::CheckStaticCollision, environment.cpp:51-87
// obb: per scenario, in shared memory: the bounds of every static polygon, then for every dynamic obstacle the
// bounds of ALL its samples (a box that misses those misses every sample's own bounding box, which is the
// reference's first test)
__device__ __forceinline__ int lowest_bit(unsigned m) {
#ifdef __CUDA_ARCH__
  return __ffs(m) - 1;
#else
  return __builtin_ctz(m);
#endif
}
// near: bit o set = static obstacle o's bounds reach the union of the two disc boxes (an obstacle whose bit is clear is
// disjoint from either disc box, i.e. the loop below would skip it); all = true: every obstacle is visited (> 32 obstacles)
__device__ bool check_static(const Args& a, int b, const double* obb, unsigned near, bool all, double cx, double cy, double half) {
  const double* polys = a.static_poly + (size_t)b * a.n_static * a.V * 2;
  const int* nv = a.static_nv + (size_t)b * a.n_static;
  if (all) {
    for (int o = 0; o < a.n_static; ++o) {
      if (aabb_disjoint(obb + 4 * o, cx, cy, half)) continue;
      if (polygon_overlaps_box(polys + (size_t)o * a.V * 2, nv[o], obb + 4 * o, cx, cy, half)) return true;
    }
  } else {
    for (unsigned m = near; m != 0; m &= m - 1) {
      const int o = lowest_bit(m);
      if (aabb_disjoint(obb + 4 * o, cx, cy, half)) continue;
      if (polygon_overlaps_box(polys + (size_t)o * a.V * 2, nv[o], obb + 4 * o, cx, cy, half)) return true;
    }
  }
  if (a.NB == 0) return false;
  const double minx = cx - half, maxx = cx + half;
  if (maxx < a.barrier[0] || minx > a.barrier[(size_t)(a.NB - 1) * 2]) return false;
  // The reference tests the sorted points with index in [upper_bound(minx) - 1, upper_bound(maxx)), i.e. those with
  // minx < x <= maxx plus the last one with x <= minx.  The same points are found through the grid: only the (at
  // most 5 x 5) cells the box can reach are scanned, a hit counts if x <= maxx and (x > minx or it is that one
  // extra point -- the binary search is only run when such a candidate shows up).
  const double m = half + 1e-9;  // the box test accepts |d| <= half + 1e-10
  int ix0 = (int)floor((cx - m - a.gx0) * a.ginv), ix1 = (int)floor((cx + m - a.gx0) * a.ginv);
  int iy0 = (int)floor((cy - m - a.gy0) * a.ginv), iy1 = (int)floor((cy + m - a.gy0) * a.ginv);
  if (ix1 < 0 || iy1 < 0 || ix0 >= a.gnx || iy0 >= a.gny) return false;
  ix0 = ix0 < 0 ? 0 : ix0;
  iy0 = iy0 < 0 ? 0 : iy0;
  ix1 = ix1 >= a.gnx ? a.gnx - 1 : ix1;
  iy1 = iy1 >= a.gny ? a.gny - 1 : iy1;
  for (int iy = iy0; iy <= iy1; ++iy) {
    const int k1 = a.grid_start[iy * a.gnx + ix1 + 1];
    for (int k = a.grid_start[iy * a.gnx + ix0]; k < k1; ++k) {
      const double px = a.grid_xy[(size_t)k * 2], py = a.grid_xy[(size_t)k * 2 + 1];
      if (!box_is_point_in(px, py, cx, cy, half)) continue;
      if (!(px <= maxx)) continue;  // index >= upper_bound(maxx)
      if (px > minx) return true;
      if (a.grid_idx[k] == barrier_upper_bound(a.barrier, a.NB, minx) - 1) return true;
    }
  }
  return false;
}

// Environment::CheckDynamicCollision, environment.cpp:124-141 (query time == last sample time: the reference
// dereferences end(); the last sample is used).  The sample a query time selects (upper_bound over the obstacle's
// sample times, :131-137) depends only on (obstacle, time); all transitions of a lattice layer query the same
// nseg times, so the selection is tabulated once per layer (sidx[o * kMaxSeg + point], -1: outside the obstacle's
// time range), and the bounds of every sample once per scenario (sbb, when they fit shared memory).
constexpr int kMaxSeg = 32;  // points of one lattice segment (InterpolateLinearly): 16-17 for the shipped configuration
__device__ __forceinline__ int dynamic_sample(const Args& a, size_t ob, double time) {
  const int ns = a.dyn_samples[ob];
  if (ns <= 0) return -1;
  const double* tt = a.dyn_time + ob * a.T;
  if (tt[0] > time || tt[ns - 1] < time) return -1;
  int lo = 0, hi = ns;
  while (lo < hi) {
    const int mid = lo + (hi - lo) / 2;
    if (time < tt[mid]) hi = mid; else lo = mid + 1;
  }
  return lo >= ns ? ns - 1 : lo;
}
// near / all: as in check_static, bit o = dynamic obstacle o
__device__ bool check_dynamic(const Args& a, int b, const double* obb, const double* sbb, const int* sidx, int point,
                              double time, unsigned near, bool all, double cx, double cy, double half) {
  unsigned m = near;
  for (int oi = 0; all ? oi < a.n_dyn : m != 0; ++oi, m &= m - 1) {
    const int o = all ? oi : lowest_bit(m);
    if (aabb_disjoint(obb + 4 * (a.n_static + o), cx, cy, half)) continue;
    const size_t ob = (size_t)b * a.n_dyn + o;
    const int lo = sidx ? sidx[o * kMaxSeg + point] : dynamic_sample(a, ob, time);
    if (lo < 0) continue;
    double bb_local[4];
    const double* bb = bb_local;
    if (sbb) bb = sbb + ((size_t)o * a.T + lo) * 4;
    else polygon_aabb(a.dyn_poly + (ob * a.T + lo) * a.V * 2, a.dyn_nv[ob], bb_local);
    if (polygon_overlaps_box(a.dyn_poly + (ob * a.T + lo) * a.V * 2, a.dyn_nv[ob], bb, cx, cy, half)) return true;
  }
  return false;
}

// Environment::CheckOptimizationCollision, environment.cpp:99-122; GetDiscPositions, vehicle_param.h:88-95
__device__ bool check_optimization_collision(const Args& a, int b, const double* obb, const double* sbb, const int* sidx,
                                             int point, double time, double x, double y, double theta) {
  const double radius = a.lat.radius;
  const double half = (radius + 0.0 - (-radius - 0.0)) / 2.0;
  const double ct = cos(theta), st = sin(theta);
  const double xf = x + a.lat.f2x * ct, xr = x + a.lat.r2x * ct;
  const double yf = y + a.lat.f2x * st, yr = y + a.lat.r2x * st;
  const double c0 = (-radius - 0.0 + (radius + 0.0)) / 2.0;
  const double fx = c0 + xf, fy = c0 + yf, rx = c0 + xr, ry = c0 + yr;
  // obstacles whose bounds miss the union of the two disc boxes cannot overlap either of them
  const double ucx = 0.5 * (fx + rx), ucy = 0.5 * (fy + ry);
  const double uhx = 0.5 * fabs(fx - rx) + half + 1e-9, uhy = 0.5 * fabs(fy - ry) + half + 1e-9;
  unsigned near = 0;  // bit o: obstacle o is near (up to 32 obstacles are pre-screened, the rest always checked)
  const int nobs = a.n_static + a.n_dyn;
  for (int o = 0; o < nobs && o < 32; ++o) {
    const double* q = obb + 4 * o;
    if (!(ucx + uhx < q[0] || ucx - uhx > q[1] || ucy + uhy < q[2] || ucy - uhy > q[3])) near |= 1u << o;
  }
  if (nobs > 32) near = 0xffffffffu;
  const unsigned near_static = a.n_static >= 32 ? near : (near & ((1u << a.n_static) - 1u));
  const unsigned near_dyn = nobs > 32 ? 1u : (a.n_static >= 32 ? 0u : near >> a.n_static);
  // front disc, then rear disc, static before dynamic (the reference's order); one copy of each check in the code
  // (only the obstacles that passed the pre-screen are visited: the others are disjoint from either disc box)
  const bool all = nobs > 32;
#pragma unroll 1
  for (int w = 0; w < 2; ++w)
    if (check_static(a, b, obb, near_static, all, w == 0 ? fx : rx, w == 0 ? fy : ry, half)) return true;
  if (near_dyn == 0) return false;
#pragma unroll 1
  for (int w = 0; w < 2; ++w)
    if (check_dynamic(a, b, obb, sbb, sidx, point, time, near_dyn, all, w == 0 ? fx : rx, w == 0 ? fy : ry, half)) return true;
  return false;
}

struct Start {
  double s, l;
};

// GetLateralOffset, dp_planner.h:83-92
__device__ __forceinline__ double lateral_offset(const Args& a, double s, int l_ind) {
  if (l_ind == NL - 1) return 0.0;
  const RefPoint r = evaluate_station(a, s);
  const double lb = -r.rb + a.lat.safe_margin;
  const double ub = r.lb - a.lat.safe_margin;
  return lb + (ub - lb) * a.lat.lateral_[l_ind];
}

// InterpolateLinearly, dp_planner.cpp:283-320, as the segment's generator: point i = (s0 + i ds, l0 + i dl)
struct Segment {
  double p_s, p_l, s_step, l_step;
  int nseg;
};
__device__ __forceinline__ Segment make_segment(const Args& a, const Start& st, double parent_s, int parent_l_ind,
                                                int cur_t_ind, int cur_s_ind, int cur_l_ind) {
  Segment g;
  g.nseg = a.lat.nseg[cur_t_ind];
  g.p_l = st.l;
  g.p_s = st.s;
  if (parent_l_ind >= 0) {
    g.p_s = parent_s;
    g.p_l = lateral_offset(a, g.p_s, parent_l_ind);
  }
  const double cur_s = g.p_s + a.lat.station_[cur_s_ind];
  const double cur_l = lateral_offset(a, cur_s, cur_l_ind);
  g.s_step = a.lat.station_[cur_s_ind] / g.nseg;
  g.l_step = (cur_l - g.p_l) / g.nseg;
  return g;
}

// the same segment from end points the caller already holds (GetCost evaluates the lateral offsets of the grandparent,
// the parent and the child itself -- the same function of the same stations, so the values are the same bits)
__device__ __forceinline__ Segment segment_of(double p_s, double p_l, double cur_l, double station, int nseg) {
  Segment g;
  g.nseg = nseg;
  g.p_s = p_s;
  g.p_l = p_l;
  g.s_step = station / nseg;
  g.l_step = (cur_l - p_l) / nseg;
  return g;
}

// GetCollisionCost, dp_planner.cpp:40-85; pt < 0: the parent is the start state.  (grandparent_s, grandparent_l),
// (parent_s, parent_l), cur_l: the end points GetCost computed (start state where there is no such ancestor) -- seven
// reference-line evaluations per transition become three, and four inlined copies of evaluate_station go away.
__device__ double collision_cost(const Args& a, int b, const double* obb, const double* sbb, const int* sidx,
                                 const Start& st, int pt, int psi, int ct, int csi, double grandparent_s,
                                 double grandparent_l, double parent_s, double parent_l, double cur_l) {
  double last_l = st.l, last_s = st.s;
  if (pt >= 0) {
    const Segment prev = segment_of(grandparent_s, grandparent_l, parent_l, a.lat.station_[psi], a.lat.nseg[pt]);
    last_l = prev.p_l + (prev.nseg - 1) * prev.l_step;
    last_s = prev.p_s + (prev.nseg - 1) * prev.s_step;
  }
  const Segment g = segment_of(parent_s, parent_l, cur_l, a.lat.station_[csi], a.lat.nseg[ct]);
  const double parent_time = pt < 0 ? 0.0 : a.lat.time_[pt];
  for (int i = 0; i < g.nseg; ++i) {
    const double ps = g.p_s + i * g.s_step, pl = g.p_l + i * g.l_step;
    const double dl = pl - last_l;
    const double ds = fmax(ps - last_s, kDpEps);
    last_l = pl;
    last_s = ps;
    const RefPoint r = evaluate_station(a, ps);  // GetCartesian (:192-196) evaluates the same station
    const double cx = r.x - pl * sin(r.theta);
    const double cy = r.y + pl * cos(r.theta);
    const double lb = fmin(0.0, -r.rb + a.lat.safe_margin);
    const double ub = fmax(0.0, r.lb - a.lat.safe_margin);
    if (pl < lb - kDpEps || pl > ub + kDpEps) return a.w_obstacle;
    const double heading = r.theta + atan((dl / ds) / (1 - r.kappa * pl));
    const double time = parent_time + i * (a.lat.unit_time / g.nseg);
    if (check_optimization_collision(a, b, obb, sbb, sidx, i, time, cx, cy, heading)) return a.w_obstacle;
  }
  return 0.0;
}

// GetCost, dp_planner.cpp:87-133
__device__ double get_cost(const Args& a, int b, const double* obb, const double* sbb, const int* sidx, const Start& st,
                           const Cell* cells, int pt, int psi, int pli, int ct, int csi, int cli, double* cur_s_out) {
  double parent_s = st.s, grandparent_s = st.s;
  double parent_l = st.l, grandparent_l = st.l;
  if (pt >= 0) {
    const Cell cell = cells[pt * NP + psi * NL + pli];
    parent_s = cell.current_s;
    parent_l = lateral_offset(a, parent_s, pli);
    if (pt >= 1) {
      grandparent_s = cells[(pt - 1) * NP + cell.ps * NL + cell.pl].current_s;
      grandparent_l = lateral_offset(a, grandparent_s, cell.pl);
    }
  }
  const double cur_s = parent_s + a.lat.station_[csi];
  const double cur_l = lateral_offset(a, cur_s, cli);
  const double ds1 = cur_s - parent_s;
  const double dl1 = cur_l - parent_l;
  const double ds0 = parent_s - grandparent_s;
  const double dl0 = parent_l - grandparent_l;
  *cur_s_out = cur_s;
  const double cost_obstacle = collision_cost(a, b, obb, sbb, sidx, st, pt, psi, ct, csi, grandparent_s, grandparent_l,
                                              parent_s, parent_l, cur_l);
  if (cost_obstacle >= a.w_obstacle) return a.w_obstacle;
  const double cost_lateral = fabs(cur_l);
  const double cost_lateral_change = fabs(parent_l - cur_l) / (a.lat.station_[csi] + kDpEps);
  const double cost_lateral_change_t = fabs(dl1 - dl0) / a.lat.unit_time;
  const double cost_longitudinal_velocity = fabs(ds1 / a.lat.unit_time - a.nominal_velocity);
  const double cost_longitudinal_velocity_change = fabs((ds1 - ds0) / a.lat.unit_time);
  return a.w_lateral * cost_lateral + a.w_lateral_change * cost_lateral_change +
         a.w_lateral_velocity_change * cost_lateral_change_t +
         a.w_longitudinal_velocity_bias * cost_longitudinal_velocity +
         a.w_longitudinal_velocity_change * cost_longitudinal_velocity_change;
}

// 512 threads x 2 CTAs per SM = 32 warps per SM at 64 registers per thread: the walk is latency bound (dependent FP64 chains, L1/L2
// loads, half the lanes of a warp active), and doubling the resident warps pays more than the registers cost
// (profiles/r02_dp_ab_5_occupancy.log: 256 x 2 at 128 registers 32.7k plans/s, 384 x 2 38.0k, 512 x 2 38.6k)
constexpr int kMaxThreads = 512;
constexpr int kMaxKnots = 512;

// shared-memory layout (bytes); K <= kMaxKnots
__host__ __device__ inline size_t smem_bytes(int K, int n_obstacles, int n_dyn, int T, bool sample_bounds) {
  return sizeof(double) * 4 * (size_t)n_obstacles + sizeof(Cell) * NT * NP + sizeof(double) * NP * NP + sizeof(double) * kMaxThreads + sizeof(int) * kMaxThreads +
         sizeof(double) * 8 + sizeof(int) * 3 * NT + sizeof(double) * 10 * (size_t)K + 64 +
         sizeof(int) * kMaxSeg * (size_t)n_dyn + 16 + (sample_bounds ? sizeof(double) * 4 * (size_t)n_dyn * T : 0);
}

__global__ void __launch_bounds__(kMaxThreads, 2) dp_plan_kernel(const __grid_constant__ Args a) {
  extern __shared__ __align__(16) unsigned char dp_smem[];
  Cell* cells = reinterpret_cast<Cell*>(dp_smem);                        // [NT][NP]
  double* delta = reinterpret_cast<double*>(cells + NT * NP);             // [NP parents][NP children]
  double* red_d = delta + NP * NP;                                        // [threads]
  double* misc = red_d + kMaxThreads;                                     // start_s, start_l, min_cost
  int* red_i = reinterpret_cast<int*>(misc + 8);                          // [threads]
  int* wp = red_i + kMaxThreads;                                          // [NT][3]: s index, l index, parent l index
  int* queue = wp + 3 * NT;                                               // next unclaimed transition of the layer
  double* kn = reinterpret_cast<double*>(wp + 3 * NT + 1);                // 10 arrays of K doubles
  kn = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(kn) + 15) & ~(uintptr_t)15);
  double* obb = kn + 10 * (size_t)a.lat.K;                                // [n_static + n_dyn][4] obstacle bounds
  double* sbb = a.use_sample_bounds ? obb + 4 * (size_t)(a.n_static + a.n_dyn) : nullptr;  // [n_dyn][T][4]
  int* sidx = reinterpret_cast<int*>(obb + 4 * (size_t)(a.n_static + a.n_dyn) + (a.use_sample_bounds ? 4 * (size_t)a.n_dyn * a.T : 0));  // [n_dyn][kMaxSeg]
  bool use_sidx = true;
  for (int k = 0; k < NT; ++k) use_sidx = use_sidx && a.lat.nseg[k] <= kMaxSeg;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int K = a.lat.K;
  double *xs = kn, *ys = kn + K, *acc_s = kn + 2 * K, *speeds = kn + 3 * K, *accel = kn + 4 * K, *xds = kn + 5 * K,
         *yds = kn + 6 * K, *xdds = kn + 7 * K, *ydds = kn + 8 * K, *th = kn + 9 * K;

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    __syncthreads();
    const double sx = a.start[(size_t)b * 3], sy = a.start[(size_t)b * 3 + 1];
    // ---- obstacle bounds of this scenario
    for (int o = tid; o < a.n_static + a.n_dyn; o += nt) {
      if (o < a.n_static) {
        polygon_aabb(a.static_poly + ((size_t)b * a.n_static + o) * a.V * 2, a.static_nv[(size_t)b * a.n_static + o],
                     obb + 4 * o);
      } else {
        const size_t ob = (size_t)b * a.n_dyn + (o - a.n_static);
        const int ns = a.dyn_samples[ob], nv = a.dyn_nv[ob];
        double acc[4] = {DBL_MAX, -DBL_MAX, DBL_MAX, -DBL_MAX};  // no sample: never reached by a box
        for (int t = 0; t < ns; ++t) {
          double bb[4];
          polygon_aabb(a.dyn_poly + (ob * a.T + t) * a.V * 2, nv, bb);
          acc[0] = fmin(acc[0], bb[0]);
          acc[1] = fmax(acc[1], bb[1]);
          acc[2] = fmin(acc[2], bb[2]);
          acc[3] = fmax(acc[3], bb[3]);
        }
        for (int q = 0; q < 4; ++q) obb[4 * o + q] = acc[q];
      }
    }
    // ---- bounds of every sample of every dynamic obstacle (Polygon2d::BuildFromPoints, polygon2d.cpp:246-256)
    if (sbb) {
      for (int i = tid; i < a.n_dyn * a.T; i += nt) {
        const int o = i / a.T, t = i - o * a.T;
        const size_t ob = (size_t)b * a.n_dyn + o;
        if (t < a.dyn_samples[ob]) polygon_aabb(a.dyn_poly + (ob * a.T + t) * a.V * 2, a.dyn_nv[ob], sbb + (size_t)i * 4);
      }
    }
    // ---- GetProjection, discretized_trajectory.cpp:156-190; QueryNearestPoint (:136-154): first minimum
    {
      double best = DBL_MAX;
      int bi = 0x7fffffff;
      for (int i = tid; i < a.R; i += nt) {
        const double dx = a.ref[(size_t)i * 7 + 1] - sx, dy = a.ref[(size_t)i * 7 + 2] - sy;
        const double d = dx * dx + dy * dy;
        if (d < best) {
          best = d;
          bi = i;
        }
      }
      red_d[tid] = best;
      red_i[tid] = bi;
    }
    __syncthreads();
    if (tid == 0) {
      double best = DBL_MAX;
      int idx = 0;
      bool any = false;
      for (int t = 0; t < nt; ++t) {
        if (red_i[t] == 0x7fffffff) continue;
        if (!any || red_d[t] < best || (red_d[t] == best && red_i[t] < idx)) {
          best = red_d[t];
          idx = red_i[t];
          any = true;
        }
      }
      const double* pr = a.ref + (size_t)idx * 7;
      RefPoint pp;
      pp.s = pr[0]; pp.x = pr[1]; pp.y = pr[2]; pp.theta = pr[3]; pp.kappa = pr[4]; pp.lb = pr[5]; pp.rb = pr[6];
      const int index_start = idx - 1 > 0 ? idx - 1 : 0;
      const int index_end = idx + 1 < a.R - 1 ? idx + 1 : a.R - 1;
      if (index_start < index_end) {
        const double* p0 = a.ref + (size_t)index_start * 7;
        const double* p1 = a.ref + (size_t)index_end * 7;
        const double v0x = sx - p0[1], v0y = sy - p0[2];
        const double v1x = p1[1] - p0[1], v1y = p1[2] - p0[2];
        const double v1_norm = sqrt(v1x * v1x + v1y * v1y);
        const double dot = v0x * v1x + v0y * v1y;
        const double delta_s = dot / v1_norm;
        pp = interpolate(p0, p1, p0[0] + delta_s);
      }
      const double nr_x = sx - pp.x, nr_y = sy - pp.y;
      misc[0] = pp.s;
      misc[1] = copysign(hypot(nr_x, nr_y), nr_y * cos(pp.theta) - nr_x * sin(pp.theta));
    }
    __syncthreads();
    Start st;
    st.s = misc[0];
    st.l = misc[1];

    // the sample every dynamic obstacle shows at each of a layer's query times (GetCollisionCost :76: time =
    // parent_time + i * unit_time / nseg)
    auto tabulate_samples = [&](int layer) {
      if (!use_sidx) return;
      const int ns = a.lat.nseg[layer];
      const double parent_time = layer == 0 ? 0.0 : a.lat.time_[layer - 1];
      for (int q = tid; q < a.n_dyn * ns; q += nt) {
        const int o = q / ns, i = q - o * ns;
        const double time = parent_time + i * (a.lat.unit_time / ns);
        sidx[o * kMaxSeg + i] = dynamic_sample(a, (size_t)b * a.n_dyn + o, time);
      }
    };
    const int* sidx_c = use_sidx ? sidx : nullptr;
    tabulate_samples(0);
    __syncthreads();
    // ---- first layer, dp_planner.cpp:151-158
    for (int p = tid; p < NP; p += nt) {
      double cur_s;
      const double c = get_cost(a, b, obb, sbb, sidx_c, st, cells, -1, -1, -1, 0, p / NL, p % NL, &cur_s);
      Cell cell;
      cell.cost = c;
      cell.current_s = cur_s;
      cell.ps = -1;
      cell.pl = -1;
      cells[p] = cell;
    }
    __syncthreads();
    // ---- dynamic programming, :160-181
    for (int i = 0; i < NT - 1; ++i) {
      // transitions are claimed from a shared counter: one that collides at its first point costs a fraction of
      // one that walks all its points, and a static split would leave most threads waiting for the unlucky ones
      if (tid == 0) *queue = 0;
      tabulate_samples(i + 1);
      __syncthreads();
      for (int q = atomicAdd(queue, 1); q < NP * NP; q = atomicAdd(queue, 1)) {
        const int parent = q / NP, child = q - parent * NP;
        double cur_s;
        delta[q] = get_cost(a, b, obb, sbb, sidx_c, st, cells, i, parent / NL, parent % NL, i + 1, child / NL, child % NL, &cur_s);
      }
      __syncthreads();
      for (int child = tid; child < NP; child += nt) {
        Cell best;
        best.cost = DBL_MAX;
        best.current_s = DBL_MIN;
        best.ps = -1;
        best.pl = -1;
        for (int parent = 0; parent < NP; ++parent) {  // (j, k) order; strict '<': the first minimum wins
          const double cur_cost = cells[i * NP + parent].cost + delta[parent * NP + child];
          if (cur_cost < best.cost) {
            best.cost = cur_cost;
            best.current_s = cells[i * NP + parent].current_s + a.lat.station_[child / NL];
            best.ps = parent / NL;
            best.pl = parent % NL;
          }
        }
        cells[(i + 1) * NP + child] = best;
      }
      __syncthreads();
    }
    // ---- least cost in the final layer (:183-194) and trace back (:196-204)
    if (tid == 0) {
      double min_cost = DBL_MAX;
      int ms = 0, ml = 0;
      for (int p = 0; p < NP; ++p) {
        const double c = cells[(NT - 1) * NP + p].cost;
        if (c < min_cost) {
          ms = p / NL;
          ml = p % NL;
          min_cost = c;
        }
      }
      misc[2] = min_cost;
      for (int i = NT - 1; i >= 0; --i) {
        const Cell c = cells[i * NP + ms * NL + ml];
        wp[3 * i] = ms;
        wp[3 * i + 1] = ml;
        wp[3 * i + 2] = c.pl;
        ms = c.ps;
        ml = c.pl;
      }
      a.ok[b] = min_cost < a.w_obstacle ? 1 : 0;
      if (a.cost) a.cost[b] = min_cost;
      if (a.waypoints)
        for (int i = 0; i < NT; ++i) {
          a.waypoints[((size_t)b * NT + i) * 3] = wp[3 * i];
          a.waypoints[((size_t)b * NT + i) * 3 + 1] = wp[3 * i + 1];
          a.waypoints[((size_t)b * NT + i) * 3 + 2] = cells[i * NP + wp[3 * i] * NL + wp[3 * i + 1]].current_s;
        }
    }
    __syncthreads();
    // ---- interpolation of the optimum, :212-243: knot n = point j of layer i; (dl, ds) against the previous point
    for (int n = tid; n < K; n += nt) {
      int i = 0, j = n;
      while (j >= a.lat.nseg[i]) {
        j -= a.lat.nseg[i];
        ++i;
      }
      const double parent_s = i > 0 ? cells[(i - 1) * NP + wp[3 * (i - 1)] * NL + wp[3 * (i - 1) + 1]].current_s : st.s;
      const Segment g = make_segment(a, st, parent_s, wp[3 * i + 2], i, wp[3 * i], wp[3 * i + 1]);
      const double ps = g.p_s + j * g.s_step, pl = g.p_l + j * g.l_step;
      double last_s, last_l;
      if (j > 0) {
        last_s = g.p_s + (j - 1) * g.s_step;
        last_l = g.p_l + (j - 1) * g.l_step;
      } else if (i > 0) {
        const double gp_s =
            i > 1 ? cells[(i - 2) * NP + wp[3 * (i - 2)] * NL + wp[3 * (i - 2) + 1]].current_s : st.s;
        const Segment pg = make_segment(a, st, gp_s, wp[3 * (i - 1) + 2], i - 1, wp[3 * (i - 1)], wp[3 * (i - 1) + 1]);
        last_s = pg.p_s + (pg.nseg - 1) * pg.s_step;
        last_l = pg.p_l + (pg.nseg - 1) * pg.l_step;
      } else {
        last_s = st.s;
        last_l = st.l;
      }
      const double dl = pl - last_l;
      const double ds = fmax(ps - last_s, kDpEps);
      const RefPoint r = evaluate_station(a, ps);
      xs[n] = r.x - pl * sin(r.theta);
      ys[n] = r.y + pl * cos(r.theta);
      th[n] = r.theta + atan((dl / ds) / (1 - r.kappa * pl));
      if (a.trajectory) a.trajectory[((size_t)b * K + n) * 13 + 1] = ps;
    }
    __syncthreads();
    // ---- ComputePathProfile, discrete_points_math.cc:27-176 (the running sum is serial in the reference)
    if (tid == 0) {
      double distance = 0.0, fx = xs[0], fy = ys[0];
      acc_s[0] = distance;
      for (int i = 1; i < K; ++i) {
        const double nx = xs[i], ny = ys[i];
        const double end_segment_s = sqrt((fx - nx) * (fx - nx) + (fy - ny) * (fy - ny));
        acc_s[i] = end_segment_s + distance;
        distance += end_segment_s;
        fx = nx;
        fy = ny;
      }
    }
    __syncthreads();
    for (int i = tid + 1; i < K; i += nt) speeds[i - 1] = (acc_s[i] - acc_s[i - 1]) / a.delta_t;
    __syncthreads();
    if (tid == 0) speeds[K - 1] = speeds[K - 2];
    __syncthreads();
    for (int i = tid + 1; i < K; i += nt) accel[i - 1] = (speeds[i] - speeds[i - 1]) / a.delta_t;
    for (int i = tid; i < K; i += nt) {
      const int lo = i == 0 ? 0 : i - 1, hi = i == K - 1 ? K - 1 : i + 1;
      xds[i] = (xs[hi] - xs[lo]) / (acc_s[hi] - acc_s[lo]);
      yds[i] = (ys[hi] - ys[lo]) / (acc_s[hi] - acc_s[lo]);
    }
    __syncthreads();
    if (tid == 0) accel[K - 1] = accel[K - 2];
    for (int i = tid; i < K; i += nt) {
      const int lo = i == 0 ? 0 : i - 1, hi = i == K - 1 ? K - 1 : i + 1;
      xdds[i] = (xds[hi] - xds[lo]) / (acc_s[hi] - acc_s[lo]);
      ydds[i] = (yds[hi] - yds[lo]) / (acc_s[hi] - acc_s[lo]);
    }
    __syncthreads();
    for (int i = tid; i < K; i += nt) {
      const double kappa = (xds[i] * ydds[i] - yds[i] * xdds[i]) /
                           (sqrt(xds[i] * xds[i] + yds[i] * yds[i]) * (xds[i] * xds[i] + yds[i] * yds[i]) + 1e-6);
      const double delta_w = atan(kappa * a.wheel_base);  // :270
      if (a.trajectory) {
        double* d = a.trajectory + ((size_t)b * K + i) * 13;
        d[0] = a.delta_t * i;
        d[2] = xs[i];
        d[3] = ys[i];
        d[4] = th[i];
        d[5] = kappa;
        d[6] = speeds[i];
        d[7] = accel[i];
        d[8] = 0.0;
        d[9] = delta_w;
        d[10] = 0.0;
        d[11] = 0.0;
        d[12] = 0.0;
      }
      if (a.coarse) {
        double* d = a.coarse + ((size_t)b * K + i) * 6;
        d[0] = xs[i];
        d[1] = ys[i];
        d[2] = th[i];
        d[3] = speeds[i];
        d[4] = accel[i];
        d[5] = delta_w;
      }
      if (a.xytheta) {
        double* d = a.xytheta + ((size_t)b * K + i) * 3;
        d[0] = xs[i];
        d[1] = ys[i];
        d[2] = th[i];
      }
    }
  }
}

}  // namespace dp
