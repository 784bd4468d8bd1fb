/*
 * corridor_oracle.c -- see corridor_oracle.h.  TEST INFRASTRUCTURE ONLY.
 *
 * Build: -O2 -ffp-contract=off (no FMA contraction: float and double expressions are evaluated as
 * the reference's x86-64 -O2 build evaluates them, FLT_EVAL_METHOD = 0).
 */
#include "corridor_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define K_MATH_EPSILON 1e-10 /* algorithm/math/math_utils.h */

void corr_default_config(corr_config* c) {
  c->max_diff_x = 25.0;
  c->max_diff_y = 25.0;
  c->radius = 150.0;
  c->max_axis_x = 10.0;
  c->max_axis_y = 10.0;
  c->lane_segment_length = 5.0;
}

/* ---------------------------------------------------------------------------------------------
 * cv::convexHull for CV_32F input (OpenCV imgproc/src/convhull.cpp): pointers sorted by
 * (x, y, address); Sklansky's scan run four times (upper-left, upper-right, lower-left,
 * lower-right chains towards the max-y / min-y points); the chains are concatenated and the result
 * cyclically shifted so that the indices ascend or descend when that is possible.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* p; /* all points */
  int* ord;       /* sorted order: ord[i] = index of the i-th point */
} hull_ctx;

/* CHullCmpPoints<float>: by x, then y, then pointer address (= index) */
static int pt_less(const float* p, int i, int j) {
  const float xi = p[2 * i], xj = p[2 * j];
  if (xi != xj) return xi < xj;
  const float yi = p[2 * i + 1], yj = p[2 * j + 1];
  if (yi != yj) return yi < yj;
  return i < j;
}

/* the comparator is a strict total order, so any sorting algorithm yields std::sort's result */
static void sort_points(const float* p, int* ord, int n, int* tmp) {
  for (int w = 1; w < n; w *= 2) {
    for (int lo = 0; lo < n; lo += 2 * w) {
      const int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
      int a = lo, b = mid, o = lo;
      while (a < mid && b < hi) tmp[o++] = pt_less(p, ord[b], ord[a]) ? ord[b++] : ord[a++];
      while (a < mid) tmp[o++] = ord[a++];
      while (b < hi) tmp[o++] = ord[b++];
    }
    memcpy(ord, tmp, sizeof(int) * (size_t)n);
  }
}

static int sgn_f(float v) { return (v > 0) - (v < 0); }
static int sgn_d(double v) { return (v > 0) - (v < 0); }

/* Sklansky_<float, double>: array[] = sorted points, scan from start towards end. */
static int sklansky(const hull_ctx* h, int start, int end, int* stack, int nsign, int sign2) {
#define PX(i) (h->p[2 * h->ord[(i)]])
#define PY(i) (h->p[2 * h->ord[(i)] + 1])
  const int incr = end > start ? 1 : -1;
  int pprev = start, pcur = pprev + incr, pnext = pcur + incr;
  int stacksize = 3;

  if (start == end || (PX(start) == PX(end) && PY(start) == PY(end))) {
    stack[0] = start;
    return 1;
  }
  stack[0] = pprev;
  stack[1] = pcur;
  stack[2] = pnext;
  end += incr; /* one past the end */

  while (pnext != end) {
    const float cury = PY(pcur);
    const float nexty = PY(pnext);
    const float by = nexty - cury;
    if (sgn_f(by) != nsign) {
      const float ax = PX(pcur) - PX(pprev);
      const float bx = PX(pnext) - PX(pcur);
      const float ay = cury - PY(pprev);
      const double convexity = (double)ay * bx - (double)ax * by;
      if (sgn_d(convexity) == sign2 && (ax != 0 || ay != 0)) {
        pprev = pcur;
        pcur = pnext;
        pnext += incr;
        stack[stacksize] = pnext;
        stacksize++;
      } else {
        if (pprev == start) {
          pcur = pnext;
          stack[1] = pcur;
          pnext += incr;
          stack[2] = pnext;
        } else {
          stack[stacksize - 2] = pnext;
          pcur = pprev;
          pprev = stack[stacksize - 4];
          stacksize--;
        }
      }
    } else {
      pnext += incr;
      stack[stacksize - 1] = pnext;
    }
  }
  return --stacksize;
#undef PX
#undef PY
}

int corr_convex_hull_f32(const float* pts, int total, int clockwise, int* hullbuf) {
  if (total <= 0) return 0;
  int* ord = (int*)malloc(sizeof(int) * (size_t)total);
  int* stack = (int*)malloc(sizeof(int) * (size_t)(total + 2));
  for (int i = 0; i < total; ++i) ord[i] = i;
  sort_points(pts, ord, total, stack);
  hull_ctx h = {pts, ord};
  int nout = 0, miny_ind = 0, maxy_ind = 0;
  for (int i = 1; i < total; ++i) {
    const float y = pts[2 * ord[i] + 1];
    if (pts[2 * ord[miny_ind] + 1] > y) miny_ind = i;
    if (pts[2 * ord[maxy_ind] + 1] < y) maxy_ind = i;
  }
  if (pts[2 * ord[0]] == pts[2 * ord[total - 1]] && pts[2 * ord[0] + 1] == pts[2 * ord[total - 1] + 1]) {
    hullbuf[nout++] = 0;
  } else {
    /* upper half */
    int* tl_stack = stack;
    int tl_count = sklansky(&h, 0, maxy_ind, tl_stack, -1, 1);
    int* tr_stack = stack + tl_count;
    int tr_count = sklansky(&h, total - 1, maxy_ind, tr_stack, -1, -1);
    if (!clockwise) {
      int* t = tl_stack; tl_stack = tr_stack; tr_stack = t;
      int c = tl_count; tl_count = tr_count; tr_count = c;
    }
    for (int i = 0; i < tl_count - 1; ++i) hullbuf[nout++] = ord[tl_stack[i]];
    for (int i = tr_count - 1; i > 0; --i) hullbuf[nout++] = ord[tr_stack[i]];
    const int stop_idx = tr_count > 2 ? tr_stack[1] : tl_count > 2 ? tl_stack[tl_count - 2] : -1;

    /* lower half */
    int* bl_stack = stack;
    int bl_count = sklansky(&h, 0, miny_ind, bl_stack, 1, -1);
    int* br_stack = stack + bl_count;
    int br_count = sklansky(&h, total - 1, miny_ind, br_stack, 1, 1);
    if (clockwise) {
      int* t = bl_stack; bl_stack = br_stack; br_stack = t;
      int c = bl_count; bl_count = br_count; br_count = c;
    }
    if (stop_idx >= 0) {
      const int check_idx = bl_count > 2 ? bl_stack[1] : bl_count + br_count > 2 ? br_stack[2 - bl_count] : -1;
      if (check_idx == stop_idx ||
          (check_idx >= 0 && pts[2 * ord[check_idx]] == pts[2 * ord[stop_idx]] &&
           pts[2 * ord[check_idx] + 1] == pts[2 * ord[stop_idx] + 1])) {
        /* all points on one line: the bottom part mirrors the top part */
        bl_count = bl_count < 2 ? bl_count : 2;
        br_count = br_count < 2 ? br_count : 2;
      }
    }
    for (int i = 0; i < bl_count - 1; ++i) hullbuf[nout++] = ord[bl_stack[i]];
    for (int i = br_count - 1; i > 0; --i) hullbuf[nout++] = ord[br_stack[i]];

    /* cyclic shift towards an ascending / descending index sequence */
    if (nout >= 3) {
      int min_idx = 0, max_idx = 0, lt = 0, i;
      for (i = 1; i < nout; ++i) {
        const int idx = hullbuf[i];
        lt += hullbuf[i - 1] < idx;
        if (lt > 1 && lt <= i - 2) break;
        if (idx < hullbuf[min_idx]) min_idx = i;
        if (idx > hullbuf[max_idx]) max_idx = i;
      }
      const int mmdist = abs(max_idx - min_idx);
      if ((mmdist == 1 || mmdist == nout - 1) && (lt <= 1 || lt >= nout - 2)) {
        const int ascending = (max_idx + 1) % nout == min_idx;
        const int i0 = ascending ? min_idx : max_idx;
        int j = i0;
        if (i0 > 0) {
          for (i = 0; i < nout; ++i) {
            const int curr_idx = stack[i] = hullbuf[j];
            const int next_j = j + 1 < nout ? j + 1 : 0;
            const int next_idx = hullbuf[next_j];
            if (i < nout - 1 && (ascending != (curr_idx < next_idx))) break;
            j = next_j;
          }
          if (i == nout) memcpy(hullbuf, stack, sizeof(int) * (size_t)nout);
        }
      }
    }
  }
  free(ord);
  free(stack);
  return nout;
}

/* ---------------------------------------------------------------------------------------------
 * Corridor::AddCorridorPoints, corridor.cc:89-120 (is_multiple_sample = false: kSampleMultiple = 1,
 * ratio takes the values 0 and 1, so every corner is emitted twice).
 * ------------------------------------------------------------------------------------------- */
void corr_add_corridor_points(const corr_config* cfg, double x, double y, double theta, double* points, int* n) {
  const double cos_heading = cos(theta);
  const double sin_heading = sin(theta);
  const double dx1 = cos_heading * cfg->max_axis_x;
  const double dy1 = sin_heading * cfg->max_axis_x;
  const double dx2 = sin_heading * cfg->max_axis_y;
  const double dy2 = -cos_heading * cfg->max_axis_y;
  const double cx[4] = {x + dx1 + dx2, x + dx1 - dx2, x - dx1 - dx2, x - dx1 + dx2};
  const double cy[4] = {y + dy1 + dy2, y + dy1 - dy2, y - dy1 - dy2, y - dy1 + dy2};
  const double ratio_step = 1.0 / 1.0;
  for (int i = 0; i < 4; ++i) {
    const int nx = (i + 1) % 4;
    for (double ratio = 0.0; ratio < 1.0 + K_MATH_EPSILON; ratio += ratio_step) {
      points[2 * *n] = cx[i] * (1 - ratio) + cx[nx] * ratio;
      points[2 * *n + 1] = cy[i] * (1 - ratio) + cy[nx] * ratio;
      ++*n;
    }
  }
}

/* ---------------------------------------------------------------------------------------------
 * Corridor::BuildCorridor, corridor.cc:122-263.
 * ------------------------------------------------------------------------------------------- */
int corr_build_corridor(const corr_config* cfg, double origin_x, double origin_y, const double* points, int n,
                        double* constraints, double* polygon, int cap, int* count) {
  *count = 0;
  if (n == 0) return CORR_E_NO_POINTS;

  /* :135-148 filterd_points */
  int* fsrc = (int*)malloc(sizeof(int) * (size_t)(n + 1));
  int nf = 0;
  for (int i = 0; i < n; ++i) {
    const double dx = points[2 * i] - origin_x;
    const double dy = points[2 * i + 1] - origin_y;
    if (fabs(dx) > cfg->max_diff_x || fabs(dy) > cfg->max_diff_y) continue;
    const double norm2 = sqrt(dx * dx + dy * dy);
    if (fabs(norm2) < K_MATH_EPSILON) continue;
    fsrc[nf++] = i;
  }
#define FX(i) points[2 * fsrc[(i)]]
#define FY(i) points[2 * fsrc[(i)] + 1]

  /* :153-177 sphere flipping about the origin; flipData has points.size()+1 slots, the unused ones
   * stay (0,0) = the origin itself */
  double safe_radius = cfg->radius;
  const int nflip = n + 1;
  float* flip = (float*)calloc((size_t)nflip * 2, sizeof(float));
  int sum = 0;
  for (int i = 0; i < nf; ++i) {
    const double dx = FX(i) - origin_x;
    const double dy = FY(i) - origin_y;
    const double norm2 = sqrt(dx * dx + dy * dy);
    if (norm2 < cfg->radius) safe_radius = norm2; /* the last such point wins, not the nearest (:168-170) */
    flip[2 * i] = (float)(dx + 2 * (cfg->radius - norm2) * dx / norm2);
    flip[2 * i + 1] = (float)(dy + 2 * (cfg->radius - norm2) * dy / norm2);
    ++sum;
  }
  int rc = CORR_OK;
  int *vidx = NULL, *vidx2 = NULL;
  float *vdata = NULL, *tcons = NULL, *dual = NULL;
  int* dhull = NULL;
  if (sum < 4) { rc = CORR_E_FEW_POINTS; goto done; }

  /* :184-199 visible vertices */
  vidx = (int*)malloc(sizeof(int) * (size_t)nflip);
  const int nv = corr_convex_hull_f32(flip, nflip, 0, vidx);
  vdata = (float*)malloc(sizeof(float) * 2 * (size_t)nv);
  int origin_vertex = 0, origin_index = -1;
  for (int i = 0; i < nv; ++i) {
    const int v = vidx[i];
    if (v == nf) {
      origin_vertex = 1;
      origin_index = i;
      vdata[2 * i] = (float)origin_x;
      vdata[2 * i + 1] = (float)origin_y;
    } else if (v > nf) {
      rc = CORR_E_ORIGIN_UB; /* filterd_points[v] out of range in the reference */
      goto done;
    } else {
      vdata[2 * i] = (float)FX(v);
      vdata[2 * i + 1] = (float)FY(v);
    }
  }

  /* :201-216 interior point */
  double interior_x, interior_y;
  if (origin_vertex) {
    /* (OriginIndex - 1) % vertexIndice.size(): int converted to size_t before the modulo */
    const int last_index = (int)((size_t)(long long)(origin_index - 1) % (size_t)nv);
    const int next_index = (int)((size_t)(origin_index + 1) % (size_t)nv);
    if (vidx[last_index] >= nf || vidx[next_index] >= nf) { rc = CORR_E_ORIGIN_UB; goto done; }
    const double dx = (FX(vidx[last_index]) + origin_x + FX(vidx[next_index])) / 3 - origin_x;
    const double dy = (FY(vidx[last_index]) + origin_y + FY(vidx[next_index])) / 3 - origin_y;
    const double d = sqrt(dx * dx + dy * dy);
    interior_x = 0.99 * safe_radius * dx / d + origin_x;
    interior_y = 0.99 * safe_radius * dy / d + origin_y;
  } else {
    interior_x = origin_x;
    interior_y = origin_y;
  }

  /* :218-234 hull of the visible vertices; every visible vertex between two consecutive hull
   * vertices gets the plane parallel to that hull edge through itself */
  vidx2 = (int*)malloc(sizeof(int) * (size_t)nv);
  const int nh2 = corr_convex_hull_f32(vdata, nv, 0, vidx2);
  int ntc = 0;
  const int tc_cap = nv * (nh2 > 0 ? nh2 : 1) + 1;
  tcons = (float*)malloc(sizeof(float) * 3 * (size_t)tc_cap);
  for (int j = 0; j < nh2; ++j) {
    const int jplus1 = (j + 1) % nh2;
    const float rx = vdata[2 * vidx2[jplus1]] - vdata[2 * vidx2[j]];
    const float ry = vdata[2 * vidx2[jplus1] + 1] - vdata[2 * vidx2[j] + 1];
    float n0 = ry, n1 = -rx; /* points outside */
    const float z = n0 * n0 + n1 * n1; /* Eigen MatrixBase::normalize(): if (z > 0) v /= sqrt(z) */
    if (z > 0.0f) {
      const float s = sqrtf(z);
      n0 = n0 / s;
      n1 = n1 / s;
    }
    int indexJ = vidx2[j];
    while (indexJ != vidx2[jplus1]) {
      const double c = (vdata[2 * indexJ] - interior_x) * n0 + (vdata[2 * indexJ + 1] - interior_y) * n1;
      tcons[3 * ntc] = n0;
      tcons[3 * ntc + 1] = n1;
      tcons[3 * ntc + 2] = (float)c;
      ++ntc;
      indexJ = (indexJ + 1) % nv;
    }
  }

  /* :236-243 dual points, their hull (clockwise, points returned) */
  dual = (float*)malloc(sizeof(float) * 2 * (size_t)(ntc + 1));
  for (int i = 0; i < ntc; ++i) {
    dual[2 * i] = tcons[3 * i] / tcons[3 * i + 2];
    dual[2 * i + 1] = tcons[3 * i + 1] / tcons[3 * i + 2];
  }
  dhull = (int*)malloc(sizeof(int) * (size_t)(ntc + 1));
  const int nd = corr_convex_hull_f32(dual, ntc, 1, dhull);
  if (nd > cap) { rc = CORR_E_CAPACITY; goto done; }

  /* :245-250 polygon vertices = duals of the dual hull's edges */
  for (int i = 0; i < nd; ++i) {
    const int iplus1 = (i + 1) % nd;
    const float dvx = dual[2 * dhull[i]], dvy = dual[2 * dhull[i] + 1];
    const float rx = dual[2 * dhull[iplus1]] - dvx;
    const float ry = dual[2 * dhull[iplus1] + 1] - dvy;
    const float cf = ry * dvx - rx * dvy; /* a float expression assigned to a double (:248) */
    const double c = cf;
    polygon[2 * i] = interior_x + ry / c;
    polygon[2 * i + 1] = interior_y - rx / c;
  }
  /* :252-261 half-planes of the polygon's edges */
  for (int i = 0; i < nd; ++i) {
    const int iplus1 = (i + 1) % nd;
    const double r0 = polygon[2 * iplus1] - polygon[2 * i];
    const double r1 = polygon[2 * iplus1 + 1] - polygon[2 * i + 1];
    const double c = -r1 * polygon[2 * i] + r0 * polygon[2 * i + 1];
    constraints[3 * i] = -r1;
    constraints[3 * i + 1] = r0;
    constraints[3 * i + 2] = c;
  }
  *count = nd;
done:
  free(fsrc); free(flip); free(vidx); free(vdata); free(vidx2); free(tcons); free(dual); free(dhull);
  return rc;
#undef FX
#undef FY
}

/* BuildCorridorConstraints, corridor.cc:56-87 */
int corr_plan(const corr_config* cfg, int K, const double* traj, const double* obs_points, const int* obs_cnt,
              int P_max, int M_max, double* constraints, int* cnt, double* polygon, int* code) {
  int first = 0;
  double* pts = (double*)malloc(sizeof(double) * 2 * (size_t)(P_max + 8));
  double* poly = (double*)malloc(sizeof(double) * 2 * (size_t)(M_max > 0 ? M_max : 1));
  for (int k = 0; k < K; ++k) {
    int n = obs_cnt[k];
    memcpy(pts, obs_points + (size_t)k * P_max * 2, sizeof(double) * 2 * (size_t)n);
    corr_add_corridor_points(cfg, traj[3 * k], traj[3 * k + 1], traj[3 * k + 2], pts, &n);
    int m = 0;
    const int rc = corr_build_corridor(cfg, traj[3 * k], traj[3 * k + 1], pts, n,
                                       constraints + (size_t)k * M_max * 3,
                                       polygon ? polygon + (size_t)k * M_max * 2 : poly, M_max, &m);
    cnt[k] = m;
    if (code) code[k] = rc;
    if (rc && !first) first = rc;
  }
  free(pts);
  free(poly);
  return first;
}

/* LaneBoundarySample (:309-322), HalfPlaneConstraint (:324-331), CalLeft/RightLaneConstraints (:265-307);
 * LineSegment2d keeps start/end as given (line_segment2d.cpp:40-49). */
int corr_lane_constraints(const corr_config* cfg, const double* boundary, int n, int is_left, double* out, int cap) {
  if (n < 1) return -1;
  double* s = (double*)malloc(sizeof(double) * 2 * (size_t)(n + 1));
  int ns = 0;
  double lx = boundary[0], ly = boundary[1];
  s[0] = lx; s[1] = ly; ns = 1;
  for (int i = 0; i < n; ++i) {
    const double px = boundary[2 * i], py = boundary[2 * i + 1];
    if (hypot(px - lx, py - ly) >= cfg->lane_segment_length - K_MATH_EPSILON) {
      s[2 * ns] = px; s[2 * ns + 1] = py; ++ns;
      lx = px; ly = py;
    }
  }
  int rc;
  if (ns < 2) { rc = -1; goto done; }
  if (ns - 1 > cap) { rc = -2; goto done; }
  for (int i = 1; i < ns; ++i) {
    /* left: segment(pt[i], pt[i-1]); right: segment(pt[i-1], pt[i]) */
    const double sx = is_left ? s[2 * i] : s[2 * (i - 1)], sy = is_left ? s[2 * i + 1] : s[2 * (i - 1) + 1];
    const double ex = is_left ? s[2 * (i - 1)] : s[2 * i], ey = is_left ? s[2 * (i - 1) + 1] : s[2 * i + 1];
    const double nx = ex - sx, ny = ey - sy;
    const double a = ny, b = -nx;
    const double c = a * sx + b * sy;
    double* o = out + (size_t)(i - 1) * 7;
    o[0] = a; o[1] = b; o[2] = c; o[3] = sx; o[4] = sy; o[5] = ex; o[6] = ey;
  }
  rc = ns - 1;
done:
  free(s);
  return rc;
}
