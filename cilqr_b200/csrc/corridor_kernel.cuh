// corridor_kernel.cuh -- batched safe-corridor builder for sm_100a.
//
// Replaces, for B trajectories x K knots at once, the per-knot body of Corridor::Plan of
// mpt0816/Cilqr: AddCorridorPoints (algorithm/ilqr/corridor.cc:89-120) + BuildCorridor (:122-263),
// including the three cv::convexHull calls on CV_32F points (OpenCV imgproc convhull.cpp: sort by
// (x, y, address), four Sklansky scans, cyclic shift of the index sequence), and the lane-boundary
// sampling / half-planes of CalLeft/RightLaneConstraints (:265-331).
//
// Mapping: ONE THREAD = ONE KNOT.  B*K = 6.6 M independent builds at the roofline-capture config,
// each a sort of <= ~60 points and three strictly serial stack scans: there is no parallelism inside
// a build worth a warp, and plenty across builds.  Every thread keeps its working set (one float
// pair per slot -- flipped points, later the visible vertices followed by the dual points -- and
// four byte-wide index arrays packed into one 32-bit word per slot: 12 bytes per point) in SHARED
// memory laid out [slot][thread]: thread t only ever touches bank t % 32, so the
// data-dependent indexing of sort and scans is bank-conflict free, and nothing spills to local memory.
//
// Arithmetic: float and double expressions are evaluated exactly as the reference's x86-64 build
// evaluates them (cv::Point2f / Eigen::Vector2f,3f are float, everything else double; no FMA
// contraction -> explicit round-to-nearest intrinsics below), so hull decisions are bit-identical to
// the CPU restatement.  The only non-IEEE inputs are cos/sin of the knot heading (CUDA's differ from
// glibc's in the last ulp on rare arguments).
#pragma once

#ifndef CORRIDOR_HOST_EMUL  // tools/corridor_host_emul.cc runs the build logic on the CPU (development aid)
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace corridor {

constexpr double kMathEpsilon = 1e-10;  // algorithm/math/math_utils.h

// per-knot result codes (CILQR_CORR_* in include/cilqr_b200.h)
enum { OK = 0, E_NO_POINTS = 1, E_FEW_POINTS = 2, E_ORIGIN = 3, E_CAPACITY = 4, E_POINT_CAPACITY = 5 };

struct Args {
  int B, K, P_max, M_max;
  int cap;  // slots per thread for points (filtered points + the origin); hull stack uses cap + 2
  double max_diff_x, max_diff_y, radius, max_axis_x, max_axis_y;
  const double* traj;        // [B][K][3] x, y, theta
  const double* obs_points;  // [B][K][P_max][2]
  const int* obs_cnt;        // [B][K]
  double* corridor;          // [B][K][M_max][3]
  int* corridor_cnt;         // [B][K]
  double* polygon;           // [B][K][M_max][2] or nullptr
  int* code;                 // [B][K]
};

__host__ __device__ inline size_t smem_bytes_per_thread(int cap) { return (size_t)(cap + 2) * 12; }

// ---- exact-rounding helpers (never contracted into FMA) ----------------------------------------
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }

__device__ __forceinline__ int sgnf(float v) { return (v > 0.0f) - (v < 0.0f); }
__device__ __forceinline__ int sgnd(double v) { return (v > 0.0) - (v < 0.0); }

// ---- the thread's slice of shared memory --------------------------------------------------------
enum { F_ORD = 0, F_STK = 1, F_HUL = 2, F_SRC = 3 };

struct Slice {
  float* ax;    // flipped points, then visible vertices (vertexData) followed by the dual points
  float* ay;
  uint32_t* w;  // packed bytes: sorted order | scan stack | hull indices | filtered -> source index
  int T;        // threads per CTA = stride between slots
  __device__ __forceinline__ int get(int field, int slot) const { return (w[slot * T] >> (8 * field)) & 0xff; }
  __device__ __forceinline__ void set(int field, int slot, int v) const {
    const uint32_t sh = 8 * field;
    uint32_t x = w[slot * T];
    x = (x & ~(0xffu << sh)) | ((uint32_t)(v & 0xff) << sh);
    w[slot * T] = x;
  }
};

// CHullCmpPoints<float>: x, then y, then address
__device__ __forceinline__ bool pt_less(float xi, float yi, int i, float xj, float yj, int j) {
  if (xi != xj) return xi < xj;
  if (yi != yj) return yi < yj;
  return i < j;
}

// Sklansky_<float, double> on sorted positions; the stack lives in field F_STK at slots sbase...,
// entries stored +1 (the look-ahead entry can be -1)
__device__ int sklansky(const Slice& s, const float* X, const float* Y, int start, int end, int sbase, int nsign,
                        int sign2) {
  const int T = s.T;
#define PX(i) X[s.get(F_ORD, (i)) * T]
#define PY(i) Y[s.get(F_ORD, (i)) * T]
#define STK(i) (s.get(F_STK, sbase + (i)) - 1)
#define SETSTK(i, v) s.set(F_STK, sbase + (i), (v) + 1)
  const int incr = end > start ? 1 : -1;
  int pprev = start, pcur = pprev + incr, pnext = pcur + incr;
  int stacksize = 3;
  if (start == end || (PX(start) == PX(end) && PY(start) == PY(end))) {
    SETSTK(0, start);
    return 1;
  }
  SETSTK(0, pprev);
  SETSTK(1, pcur);
  SETSTK(2, pnext);
  end += incr;
  float prevx = PX(pprev), prevy = PY(pprev), curx = PX(pcur), cury = PY(pcur);
  // the point after pnext is fetched one step ahead: every branch but the pop advances pnext by incr, so
  // the two dependent shared-memory loads (order byte, then coordinates) overlap the arithmetic of this step
  float nextx = 0.0f, nexty = 0.0f, aheadx = 0.0f, aheady = 0.0f;
  if (pnext != end) { nextx = PX(pnext); nexty = PY(pnext); }
  bool have_ahead = false;
  while (pnext != end) {
    if (!have_ahead) {
      const int pa = pnext + incr;
      if (pa != end) { aheadx = PX(pa); aheady = PY(pa); }
      have_ahead = true;
    }
    const float by = fsub(nexty, cury);
    bool advance = true;
    if (sgnf(by) != nsign) {
      const float ax = fsub(curx, prevx);
      const float bx = fsub(nextx, curx);
      const float ay = fsub(cury, prevy);
      const double convexity = dsub(dmul((double)ay, (double)bx), dmul((double)ax, (double)by));
      if (sgnd(convexity) == sign2 && (ax != 0.0f || ay != 0.0f)) {
        pprev = pcur; prevx = curx; prevy = cury;
        pcur = pnext; curx = nextx; cury = nexty;
        pnext += incr;
        SETSTK(stacksize, pnext);
        stacksize++;
      } else {
        if (pprev == start) {
          pcur = pnext; curx = nextx; cury = nexty;
          SETSTK(1, pcur);
          pnext += incr;
          SETSTK(2, pnext);
        } else {
          SETSTK(stacksize - 2, pnext);
          pcur = pprev; curx = prevx; cury = prevy;
          pprev = STK(stacksize - 4);
          prevx = PX(pprev); prevy = PY(pprev);
          stacksize--;
          advance = false;  // pnext stays: the point fetched ahead stays ahead
        }
      }
    } else {
      pnext += incr;
      SETSTK(stacksize - 1, pnext);
    }
    if (advance) {
      nextx = aheadx; nexty = aheady;
      have_ahead = false;
    }
  }
  return --stacksize;
#undef PX
#undef PY
#undef STK
#undef SETSTK
}

// cv::convexHull(points, hull, clockwise, returnPoints = false) on the thread's n points (X, Y);
// the hull's point indices land in field F_HUL, their number is returned.
__device__ int convex_hull(const Slice& s, const float* X, const float* Y, int total, bool clockwise) {
  const int T = s.T;
  if (total <= 0) return 0;
  // sort of the index permutation by rank counting: the comparator is a strict total order, so the rank
  // of a point is its sorted position.  O(n^2) comparisons, but every one of them is independent (the
  // insertion sort this replaces was a chain of dependent shared-memory read-modify-writes, and with only a
  // few warps per SM -- the per-thread slice is ~1 KB -- latency, not issue rate, is what costs).  Four
  // points are ranked per sweep so that each loaded (x, y) is used four times.
  // The index tie-break of the comparator is folded into the sweep: against the points before the block a
  // coincident point ranks first ("<="), against the points after it last ("<"), so both sweeps are three
  // predicate-combining float compares per pair; only the block's own 4 x 4 pairs use the general rule.
  for (int i = 0; i < total; i += 4) {
    float xi[4], yi[4];
    int r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ii = i + q < total ? i + q : total - 1;
      xi[q] = X[ii * T];
      yi[q] = Y[ii * T];
      r[q] = 0;
    }
#pragma unroll 4
    for (int j = 0; j < i; ++j) {
      const float xj = X[j * T], yj = Y[j * T];
#pragma unroll
      for (int q = 0; q < 4; ++q) r[q] += (xj < xi[q]) | ((xj == xi[q]) & (yj <= yi[q]));
    }
    const int jend = i + 4 < total ? i + 4 : total;
    for (int j = i; j < jend; ++j) {
      const float xj = X[j * T], yj = Y[j * T];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        r[q] += (xj < xi[q]) | ((xj == xi[q]) & ((yj < yi[q]) | ((yj == yi[q]) & (j < i + q))));
    }
#pragma unroll 4
    for (int j = i + 4; j < total; ++j) {
      const float xj = X[j * T], yj = Y[j * T];
#pragma unroll
      for (int q = 0; q < 4; ++q) r[q] += (xj < xi[q]) | ((xj == xi[q]) & (yj < yi[q]));
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (i + q < total) s.set(F_ORD, r[q], i + q);
  }
  int miny_ind = 0, maxy_ind = 0;
  {
    float miny = Y[s.get(F_ORD, 0) * T], maxy = miny;
    for (int i = 1; i < total; ++i) {
      const float y = Y[s.get(F_ORD, i) * T];
      if (miny > y) { miny = y; miny_ind = i; }
      if (maxy < y) { maxy = y; maxy_ind = i; }
    }
  }
  int nout = 0;
  const int o0 = s.get(F_ORD, 0), oL = s.get(F_ORD, total - 1);
  if (X[o0 * T] == X[oL * T] && Y[o0 * T] == Y[oL * T]) {
    s.set(F_HUL, nout++, 0);
    return nout;
  }
#define ORD(i) s.get(F_ORD, (i))
#define STK(i) (s.get(F_STK, (i)) - 1)
  // upper half
  int tl_base = 0;
  int tl_count = sklansky(s, X, Y, 0, maxy_ind, 0, -1, 1);
  int tr_base = tl_count;
  int tr_count = sklansky(s, X, Y, total - 1, maxy_ind, tr_base, -1, -1);
  if (!clockwise) {
    int t = tl_base; tl_base = tr_base; tr_base = t;
    t = tl_count; tl_count = tr_count; tr_count = t;
  }
  for (int i = 0; i < tl_count - 1; ++i) s.set(F_HUL, nout++, ORD(STK(tl_base + i)));
  for (int i = tr_count - 1; i > 0; --i) s.set(F_HUL, nout++, ORD(STK(tr_base + i)));
  const int stop_idx = tr_count > 2 ? STK(tr_base + 1) : tl_count > 2 ? STK(tl_base + tl_count - 2) : -1;

  // lower half
  int bl_base = 0;
  int bl_count = sklansky(s, X, Y, 0, miny_ind, 0, 1, -1);
  int br_base = bl_count;
  int br_count = sklansky(s, X, Y, total - 1, miny_ind, br_base, 1, 1);
  if (clockwise) {
    int t = bl_base; bl_base = br_base; br_base = t;
    t = bl_count; bl_count = br_count; br_count = t;
  }
  if (stop_idx >= 0) {
    const int check_idx =
        bl_count > 2 ? STK(bl_base + 1) : bl_count + br_count > 2 ? STK(br_base + 2 - bl_count) : -1;
    bool same = check_idx == stop_idx;
    if (!same && check_idx >= 0) {
      const int oc = ORD(check_idx), os = ORD(stop_idx);
      same = X[oc * T] == X[os * T] && Y[oc * T] == Y[os * T];
    }
    if (same) {
      bl_count = bl_count < 2 ? bl_count : 2;
      br_count = br_count < 2 ? br_count : 2;
    }
  }
  for (int i = 0; i < bl_count - 1; ++i) s.set(F_HUL, nout++, ORD(STK(bl_base + i)));
  for (int i = br_count - 1; i > 0; --i) s.set(F_HUL, nout++, ORD(STK(br_base + i)));
#undef ORD
#undef STK

  // cyclic shift towards an ascending / descending index sequence
  if (nout >= 3) {
    int min_idx = 0, max_idx = 0, lt = 0, i;
    int hmin = s.get(F_HUL, 0), hmax = hmin, prev = hmin;
    for (i = 1; i < nout; ++i) {
      const int idx = s.get(F_HUL, i);
      lt += prev < idx;
      prev = idx;
      if (lt > 1 && lt <= i - 2) break;
      if (idx < hmin) { hmin = idx; min_idx = i; }
      if (idx > hmax) { hmax = idx; max_idx = i; }
    }
    const int mmdist = max_idx > min_idx ? max_idx - min_idx : min_idx - max_idx;
    if ((mmdist == 1 || mmdist == nout - 1) && (lt <= 1 || lt >= nout - 2)) {
      const int ascending = (max_idx + 1) % nout == min_idx;
      const int i0 = ascending ? min_idx : max_idx;
      int j = i0;
      if (i0 > 0) {
        for (i = 0; i < nout; ++i) {
          const int curr_idx = s.get(F_HUL, j);
          s.set(F_STK, i, curr_idx);
          const int next_j = j + 1 < nout ? j + 1 : 0;
          const int next_idx = s.get(F_HUL, next_j);
          if (i < nout - 1 && (ascending != (curr_idx < next_idx))) break;
          j = next_j;
        }
        if (i == nout)
          for (i = 0; i < nout; ++i) s.set(F_HUL, i, s.get(F_STK, i));
      }
    }
  }
  return nout;
}

__global__ void __launch_bounds__(256) corridor_build_kernel(const Args a) {
  extern __shared__ __align__(16) unsigned char corr_smem[];
  const int T = blockDim.x;
  const int slots = a.cap + 2;
  Slice s;
  {
    float* f = reinterpret_cast<float*>(corr_smem);
    s.ax = f + threadIdx.x;
    s.ay = f + (size_t)slots * T + threadIdx.x;
    s.w = reinterpret_cast<uint32_t*>(f + (size_t)2 * slots * T) + threadIdx.x;
    s.T = T;
  }
  const long long total_items = (long long)a.B * a.K;
  for (long long item = (long long)blockIdx.x * T + threadIdx.x; item < total_items;
       item += (long long)gridDim.x * T) {
    const double ox = a.traj[item * 3], oy = a.traj[item * 3 + 1], theta = a.traj[item * 3 + 2];
    const int cnt = a.obs_cnt[item];
    const double* pts = a.obs_points + (size_t)item * a.P_max * 2;
    double* cons = a.corridor + (size_t)item * a.M_max * 3;
    double* poly = a.polygon ? a.polygon + (size_t)item * a.M_max * 2 : nullptr;
    int rc = OK, nd = 0;

    // AddCorridorPoints, corridor.cc:89-120: the box corners, each emitted twice (ratio 0 and 1)
    double sn, cs;
    sincos(theta, &sn, &cs);
    const double dx1 = dmul(cs, a.max_axis_x), dy1 = dmul(sn, a.max_axis_x);
    const double dx2 = dmul(sn, a.max_axis_y), dy2 = dmul(-cs, a.max_axis_y);
    const double cxs[4] = {dadd(dadd(ox, dx1), dx2), dsub(dadd(ox, dx1), dx2), dsub(dsub(ox, dx1), dx2),
                           dadd(dsub(ox, dx1), dx2)};
    const double cys[4] = {dadd(dadd(oy, dy1), dy2), dsub(dadd(oy, dy1), dy2), dsub(dsub(oy, dy1), dy2),
                           dadd(dsub(oy, dy1), dy2)};
    const int n = cnt + 8;
    auto point = [&](int i, double* x, double* y) {
      if (i < cnt) {
        const double2 p = *reinterpret_cast<const double2*>(pts + 2 * i);
        *x = p.x;
        *y = p.y;
      } else {
        const int c = ((i - cnt + 1) >> 1) & 3;  // c0 c1 c1 c2 c2 c3 c3 c0
        *x = c == 0 ? cxs[0] : c == 1 ? cxs[1] : c == 2 ? cxs[2] : cxs[3];
        *y = c == 0 ? cys[0] : c == 1 ? cys[1] : c == 2 ? cys[2] : cys[3];
      }
    };

    // :135-177 filter + sphere flip about the knot.  flipData's unused slots are (0,0) = the knot
    // itself; they are interior to the hull of the flipped box points, so one of them is kept.
    int nf = 0;
    double qx, qy;  // the next point is loaded one step ahead of the double-precision work on this one
    point(0, &qx, &qy);
    for (int i = 0; i < n; ++i) {
      const double px = qx, py = qy;
      if (i + 1 < n) point(i + 1, &qx, &qy);
      const double dx = dsub(px, ox), dy = dsub(py, oy);
      if (fabs(dx) > a.max_diff_x || fabs(dy) > a.max_diff_y) continue;
      const double norm2 = __dsqrt_rn(dadd(dmul(dx, dx), dmul(dy, dy)));
      if (fabs(norm2) < kMathEpsilon) continue;
      if (nf >= a.cap - 1) { rc = E_POINT_CAPACITY; break; }
      const double k2 = dmul(2.0, dsub(a.radius, norm2));
      s.ax[nf * T] = __double2float_rn(dadd(dx, __ddiv_rn(dmul(k2, dx), norm2)));
      s.ay[nf * T] = __double2float_rn(dadd(dy, __ddiv_rn(dmul(k2, dy), norm2)));
      s.set(F_SRC, nf, i);
      ++nf;
    }
    if (rc == OK && nf < 4) rc = E_FEW_POINTS;  // :179-182
    if (rc == OK) {
      s.ax[nf * T] = 0.0f;
      s.ay[nf * T] = 0.0f;
      // :184-199 visible vertices
      const int nv = convex_hull(s, s.ax, s.ay, nf + 1, false);
      for (int i = 0; i < nv && rc == OK; ++i)
        if (s.get(F_HUL, i) >= nf) rc = E_ORIGIN;  // cannot happen: the box surrounds the knot
      if (rc == OK) {
        for (int i = 0; i < nv; ++i) {
          double px, py;
          point(s.get(F_SRC, s.get(F_HUL, i)), &px, &py);
          s.ax[i * T] = __double2float_rn(px);  // cv::Point2f(filterd_points[v].x(), ...)
          s.ay[i * T] = __double2float_rn(py);
        }
        const double ix = ox, iy = oy;  // :213-216 interior point = the knot
        // :218-234 hull of the visible vertices; every vertex between two consecutive hull vertices gets
        // the plane parallel to that hull edge through itself; :236-240 its dual point
        const int nh2 = convex_hull(s, s.ax, s.ay, nv, false);
        int ntc = 0;
        float* const bx = s.ax + nv * T;  // the dual points go behind the visible vertices
        float* const by = s.ay + nv * T;
        for (int j = 0; j < nh2 && rc == OK; ++j) {
          const int va = s.get(F_HUL, j), vb = s.get(F_HUL, j + 1 < nh2 ? j + 1 : 0);
          const float rx = fsub(s.ax[vb * T], s.ax[va * T]);
          const float ry = fsub(s.ay[vb * T], s.ay[va * T]);
          float n0 = ry, n1 = -rx;
          const float z = fadd(fmul(n0, n0), fmul(n1, n1));  // Eigen normalize(): if (z > 0) v /= sqrt(z)
          if (z > 0.0f) {
            const float sq = __fsqrt_rn(z);
            n0 = __fdiv_rn(n0, sq);
            n1 = __fdiv_rn(n1, sq);
          }
          int idx = va;
          while (idx != vb) {
            if (nv + ntc >= a.cap) { rc = E_POINT_CAPACITY; break; }
            const double c = dadd(dmul(dsub((double)s.ax[idx * T], ix), (double)n0),
                                  dmul(dsub((double)s.ay[idx * T], iy), (double)n1));
            const float cf = __double2float_rn(c);
            bx[ntc * T] = __fdiv_rn(n0, cf);
            by[ntc * T] = __fdiv_rn(n1, cf);
            ++ntc;
            idx = idx + 1 < nv ? idx + 1 : 0;
          }
        }
        if (rc == OK) {
          // :241-243 hull of the dual points (clockwise); :245-261 polygon vertices and their edges' planes
          nd = convex_hull(s, bx, by, ntc, true);
          if (nd > a.M_max) {
            rc = E_CAPACITY;
            nd = 0;
          }
          double p0x = 0, p0y = 0, ppx = 0, ppy = 0;
          for (int i = 0; i <= nd && nd > 0; ++i) {
            double qx, qy;
            if (i < nd) {
              const int h0 = s.get(F_HUL, i), h1 = s.get(F_HUL, i + 1 < nd ? i + 1 : 0);
              const float dvx = bx[h0 * T], dvy = by[h0 * T];
              const float rx = fsub(bx[h1 * T], dvx), ry = fsub(by[h1 * T], dvy);
              const double c = (double)fsub(fmul(ry, dvx), fmul(rx, dvy));  // float expression widened (:248)
              qx = dadd(ix, __ddiv_rn((double)ry, c));
              qy = dsub(iy, __ddiv_rn((double)rx, c));
              if (poly) {
                poly[2 * i] = qx;
                poly[2 * i + 1] = qy;
              }
              if (i == 0) { p0x = qx; p0y = qy; }
            } else {
              qx = p0x;
              qy = p0y;
            }
            if (i > 0) {
              const double r0 = dsub(qx, ppx), r1 = dsub(qy, ppy);
              double* o = cons + 3 * (i - 1);
              o[0] = -r1;
              o[1] = r0;
              o[2] = dadd(dmul(-r1, ppx), dmul(r0, ppy));
            }
            ppx = qx;
            ppy = qy;
          }
        }
      }
    }
    a.corridor_cnt[item] = rc == OK ? nd : 0;
    a.code[item] = rc;
  }
}

#ifndef CORRIDOR_HOST_EMUL
// ---- lane constraints: LaneBoundarySample (:309-322) + HalfPlaneConstraint (:324-331) --------------
// One warp per boundary polyline: the greedy "next point >= 5 m from the last kept one" scan is
// serial in the kept points but each search is a 32-wide ballot over the following points.
struct LaneArgs {
  int B, n, S_max, is_left;
  double seg_len;
  const double* boundary;  // [B][n][2]
  double* out;             // [B][S_max][7]
  int* count;              // [B]: segments, -1: fewer than 2 sampled points, -2: more than S_max
};

__global__ void lane_constraints_kernel(const LaneArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.B) return;
  const double* p = a.boundary + (size_t)warp * a.n * 2;
  double* out = a.out + (size_t)warp * a.S_max * 7;
  double lx = p[0], ly = p[1];
  int ns = 1, i = 0, rc = 0;
  const double thr = a.seg_len - kMathEpsilon;
  while (i < a.n) {
    const int j = i + lane;
    bool hit = false;
    double qx = 0, qy = 0;
    if (j < a.n) {
      qx = p[2 * j];
      qy = p[2 * j + 1];
      hit = hypot(dsub(qx, lx), dsub(qy, ly)) >= thr;
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (!m) {
      i += 32;
      continue;
    }
    const int first = __ffs(m) - 1;
    const double nx_ = __shfl_sync(0xffffffffu, qx, first), ny_ = __shfl_sync(0xffffffffu, qy, first);
    if (ns - 1 >= a.S_max) {
      rc = -2;
      break;
    }
    if (lane == 0) {
      // left: segment(pt[i], pt[i-1]); right: segment(pt[i-1], pt[i])  (corridor.cc:279,300)
      const double sx = a.is_left ? nx_ : lx, sy = a.is_left ? ny_ : ly;
      const double ex = a.is_left ? lx : nx_, ey = a.is_left ? ly : ny_;
      const double dx = dsub(ex, sx), dy = dsub(ey, sy);
      const double A = dy, Bc = -dx;
      double* o = out + (size_t)(ns - 1) * 7;
      o[0] = A;
      o[1] = Bc;
      o[2] = dadd(dmul(A, sx), dmul(Bc, sy));
      o[3] = sx;
      o[4] = sy;
      o[5] = ex;
      o[6] = ey;
    }
    lx = nx_;
    ly = ny_;
    ++ns;
    i += first + 1;
  }
  if (lane == 0) a.count[warp] = rc ? rc : (ns < 2 ? -1 : ns - 1);
}

#endif  // CORRIDOR_HOST_EMUL

}  // namespace corridor
