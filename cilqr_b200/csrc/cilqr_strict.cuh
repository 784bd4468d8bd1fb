// cilqr_strict.cuh -- the reference-ordered arithmetic of the STRICT build (CILQR_STRICT = 1; see the top of
// cilqr_kernel.cuh).  Included from cilqr_kernel.cuh inside namespace cilqr, after Ctx.
//
// Every function here evaluates exactly the expressions of the reference, in the reference's order, as the CPU
// restatement oracle/cilqr_oracle.c does (which is pinned bit for bit against the reference's own sources,
// tests/test_reference_pins.py); speed is not a goal of this build.  Reference lines:
//   RelaxBarrierFunction value / Jacbian / Hessian   algorithm/ilqr/barrier_function.h:104-140
//   LineSegment2d::DistanceTo                        algorithm/math/line_segment2d.cpp:61-75
//   FindNeastLaneSegment                             algorithm/ilqr/ilqr_optimizer.cc:605-618
//   TotalCost = JCost + DynamicsCost + CorridorCost + LaneBoundaryCost   :417-436, 497-603
//   CostJacbian / CostHessian (+ the three *Cons* pairs)                 :620-769
//   Backward                                         :334-390  (lazy-expression quirk Q21, in-place Vxx Q22)
//   CalGradientNorm                                  :322-332
//   iqr                                              :793-824
#pragma once

// barrier_function.h:104-113
__device__ __forceinline__ double s_bar_value(double g, const DevParams& P) {
  if (g < -P.eps) return -P.rt * nt_log(-g);
  const double q = (-g - 2.0 * P.eps) / P.eps;
  return 0.5 * P.rt * (q * q - 1) - P.rt_log_eps;
}
// barrier_function.h:115-125: coefficient of dx
__device__ __forceinline__ double s_bar_dcoef(double g, const DevParams& P) {
  if (g < -P.eps) return -P.rt / g;
  return P.rt * (g + 2.0 * P.eps) / P.eps / P.eps;
}
// barrier_function.h:127-140: coefficient of dx dx^T and of ddx (0 on the relaxed branch, quirk Q6)
__device__ __forceinline__ void s_bar_hcoef(double g, const DevParams& P, double& co, double& cd) {
  if (g < -P.eps) {
    co = P.rt / g / g;
    cd = P.rt / g;
  } else {
    co = P.rt * (g + 2.0 * P.eps) / P.eps / P.eps;
    cd = 0.0;
  }
}

// line_segment2d.cpp:61-75 on a staged segment record (sx sy ex ey ux uy len a b c)
__device__ __forceinline__ double s_seg_distance(const double* sg, double px, double py) {
  if (sg[6] <= 1e-10) return nt_hypot(px - sg[0], py - sg[1]);
  const double x0 = px - sg[0], y0 = py - sg[1];
  const double proj = x0 * sg[4] + y0 * sg[5];
  if (proj <= 0.0) return nt_hypot(x0, y0);
  if (proj >= sg[6]) return nt_hypot(px - sg[2], py - sg[3]);
  return fabs(x0 * sg[5] - y0 * sg[4]);
}
// ilqr_optimizer.cc:605-618: brute force, strict '<', first minimum wins
__device__ __noinline__ int s_nearest(const double* sg0, int S, double x, double y) {
  double min_dis = 1.7976931348623157e308;
  int bi = -1;
#pragma unroll 1
  for (int i = 0; i < S; ++i) {
    const double dis = s_seg_distance(sg0 + i * kSegStride, x, y);
    if (dis < min_dis) {
      min_dis = dis;
      bi = i;
    }
  }
  return bi < 0 ? 0 : bi;  // (every distance NaN: the reference indexes [-1]; a blown-up candidate is rejected anyway)
}

// ------------------------------------------------------------------------------------------
// TotalCost (:417-436).  Terms are computed by all lanes into shared memory and ADDED by one lane per
// accumulator in the reference's order: knot-major, disc-major, plane-minor, left lane before right.
//   buf  [32][bw]   bw = M_max + 2: per item (knot, disc) its corridor terms, then the two lane terms
__device__ __noinline__ void eval_cost(const Ctx& c, const double* Xs, const unsigned char* /*guess*/,
                                       unsigned char* nidx, double cost5[5]) {
  const KernelArgs& a = c.a;
  const DevParams& P = a.P;
  const int K = a.N + 1, N = a.N, lane = c.lane;
  const double* seg = c.smp() + a.sm.seg;
  double* trig = c.smp() + a.sm.trig;  // [K][2] sin, cos
  double* buf = c.smp() + a.sm.pl_e;
  const int bw = strict_row_width(a.M_max);
  // ---- JCost (:497-516) state terms and the state half of DynamicsCost (:518-538); sin / cos of the heading
  double acc_j = 0.0, acc_x = 0.0;
#pragma unroll 1
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int k = k0 + lane;
    if (k < K) {
      const double px = Xs[k], py = Xs[a.Kc + k], th = Xs[2 * a.Kc + k], v = Xs[3 * a.Kc + k], ac = Xs[4 * a.Kc + k],
                   de = Xs[5 * a.Kc + k];
      const double dx = px - c.goal(k, 0), dy = py - c.goal(k, 1), dth = th - c.goal(k, 2);
      double* row = buf + lane * bw;
      row[0] = P.wx * (dx * dx) + P.wy * (dy * dy) + P.wth * (dth * dth);
      row[1] = s_bar_value(-v, P);
      row[2] = s_bar_value(v - P.vmax, P);
      row[3] = s_bar_value(ac - P.amax, P);
      row[4] = s_bar_value(P.amin - ac, P);
      row[5] = s_bar_value(de - P.dmax, P);
      row[6] = s_bar_value(P.dmin - de, P);
      const double2 sc = nt_sincos(th);
      trig[k * 2] = sc.x;
      trig[k * 2 + 1] = sc.y;
    }
    __syncwarp();
    const int n = K - k0 < 32 ? K - k0 : 32;
    if (lane == 0) {
      for (int i = 0; i < n; ++i) acc_j += buf[i * bw];
    } else if (lane == 1) {
      for (int i = 0; i < n; ++i)
        for (int q = 1; q <= 6; ++q) acc_x += buf[i * bw + q];
    }
    __syncwarp();
  }
  // ---- control terms: JCost continues its accumulator (:509-512), DynamicsCost starts u_cost (:539-549)
  double acc_u = 0.0;
#pragma unroll 1
  for (int k0 = 0; k0 < N; k0 += 32) {
    const int k = k0 + lane;
    if (k < N) {
      const double u0 = Xs[6 * a.Kc + k], u1 = Xs[7 * a.Kc + k];
      double* row = buf + lane * bw;
      row[0] = P.wj * (u0 * u0) + P.wdr * (u1 * u1);
      row[1] = s_bar_value(u0 - P.jmax, P);
      row[2] = s_bar_value(P.jmin - u0, P);
      row[3] = s_bar_value(u1 - P.drmax, P);
      row[4] = s_bar_value(P.drmin - u1, P);
    }
    __syncwarp();
    const int n = N - k0 < 32 ? N - k0 : 32;
    if (lane == 0) {
      for (int i = 0; i < n; ++i) acc_j += buf[i * bw];
    } else if (lane == 1) {
      for (int i = 0; i < n; ++i)
        for (int q = 1; q <= 4; ++q) acc_u += buf[i * bw + q];
    }
    __syncwarp();
  }
  // ---- CorridorCost (:553-581) and LaneBoundaryCost (:583-603): item = (knot, disc)
  const int items = K * kDisc;
  double acc_c = 0.0, acc_l = 0.0;
  const double* planes = c.planes();
#pragma unroll 1
  for (int j0 = 0; j0 < items; j0 += 32) {
    const int j = j0 + lane;
    if (j < items) {
      const int k = j / kDisc, d = j - k * kDisc;
      const double o = P.off[d];
      const double xd = Xs[k] + o * trig[k * 2 + 1];
      const double yd = Xs[a.Kc + k] + o * trig[k * 2];
      double* row = buf + lane * bw;
      const int M = c.cnt[k];
#pragma unroll 1
      for (int m = 0; m < M; ++m) {
        const double pa = planes[(m * 3 + 0) * a.Kc + k], pb = planes[(m * 3 + 1) * a.Kc + k],
                     pc = planes[(m * 3 + 2) * a.Kc + k];
        row[m] = s_bar_value(pa * xd + pb * yd - pc, P);
      }
#pragma unroll 1
      for (int side = 0; side < 2; ++side) {
        const int S = side == 0 ? a.S_left : a.S_right;
        const double* sg0 = seg + (side == 0 ? 0 : a.S_left) * kSegStride;
        const int bi = s_nearest(sg0, S, xd, yd);
        const double* sg = sg0 + bi * kSegStride;
        row[a.M_max + side] = s_bar_value(sg[7] * xd + sg[8] * yd - sg[9], P);
        nidx[j * 2 + side] = (unsigned char)bi;
      }
    }
    __syncwarp();
    const int n = items - j0 < 32 ? items - j0 : 32;
    if (lane == 0) {
      for (int i = 0; i < n; ++i) {
        const int M = c.cnt[(j0 + i) / kDisc];
        for (int m = 0; m < M; ++m) acc_c += buf[i * bw + m];
      }
    } else if (lane == 1) {
      for (int i = 0; i < n; ++i) {
        acc_l += buf[i * bw + a.M_max];
        acc_l += buf[i * bw + a.M_max + 1];
      }
    }
    __syncwarp();
  }
  const double jc = __shfl_sync(kFull, acc_j, 0);
  const double dc = __shfl_sync(kFull, acc_x, 1) + __shfl_sync(kFull, acc_u, 1);
  const double co = __shfl_sync(kFull, acc_c, 0), la = __shfl_sync(kFull, acc_l, 1);
  cost5[0] = jc + dc + co + la;
  cost5[1] = jc;
  cost5[2] = dc;
  cost5[3] = co;
  cost5[4] = la;
}

// ------------------------------------------------------------------------------------------
// DynamicsJacbian + CostJacbian + CostHessian of ONE knot by ONE lane, everything accumulated in the reference's
// order (:620-769): running cost, bound barriers, then every corridor plane of every disc, then the nearest lane
// segments of every disc (left, right).  The record carries the full (not bit-symmetric) 3x3 block of Hx.
__device__ __forceinline__ void s_plane_term(const DevParams& P, double h0, double h1, double h2, double x, double y,
                                             double lc, double ls, double* J, double* H) {
  const double g = h0 * x + h1 * y - h2;
  const double d[3] = {h0, h1, -h0 * ls + h1 * lc};
  const double cj = s_bar_dcoef(g, P);
#pragma unroll
  for (int r = 0; r < 3; ++r) J[r] += cj * d[r];
  double co, cd;
  s_bar_hcoef(g, P, co, cd);
  const double ddx22 = -h0 * lc - h1 * ls;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double cr = co * d[r];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      double v = cr * d[q];
      if (g < -P.eps) v = v - cd * ((r == 2 && q == 2) ? ddx22 : 0.0);
      H[r * 3 + q] += v;
    }
  }
}

__device__ __noinline__ void s_linearize_knot(const Ctx& c, int k, const double* Xs, const unsigned char* nidx, double* rec) {
  const KernelArgs& a = c.a;
  const DevParams& P = a.P;
  const int N = a.N;
  double x[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) x[i] = Xs[i * a.Kc + k];
  const double u0 = k < N ? Xs[6 * a.Kc + k] : 0.0, u1 = k < N ? Xs[7 * a.Kc + k] : 0.0;
  if (k < N) {
    double A11[11], b21;
    dynamics_jacobian(P, x, u1, A11, &b21);
#pragma unroll
    for (int i = 0; i < 11; ++i) rec[LA + i] = A11[i];
    rec[LB21] = b21;
  }
  // DynamicsConsJacbian / Hessian (:657-688): per component one (lower, upper) pair of bounds
  auto bound_pair = [&](double lo_g, double hi_g, double& dj, double& dh) {
    const double c0 = s_bar_dcoef(lo_g, P), c1 = s_bar_dcoef(hi_g, P);
    double h0, h1, unused;
    s_bar_hcoef(lo_g, P, h0, unused);
    s_bar_hcoef(hi_g, P, h1, unused);
    dj = c0 * -1.0 + c1 * 1.0;
    dh = (h0 * -1.0) * -1.0 + (h1 * 1.0) * 1.0;
  };
  double dj, dh;
  bound_pair(0.0 - x[3], x[3] - P.vmax, dj, dh);
  rec[LJX + 3] = 0.0 + dj;
  rec[LHX + 6] = 2.0 * P.wv + dh;
  bound_pair(P.amin - x[4], x[4] - P.amax, dj, dh);
  rec[LJX + 4] = 0.0 + dj;
  rec[LHX + 7] = 2.0 * P.wa + dh;
  bound_pair(P.dmin - x[5], x[5] - P.dmax, dj, dh);
  rec[LJX + 5] = 0.0 + dj;
  rec[LHX + 8] = 2.0 * P.wd + dh;
  bound_pair(P.jmin - u0, u0 - P.jmax, dj, dh);
  rec[LJU + 0] = 2.0 * P.wj * u0 + dj;
  rec[LHU + 0] = 2.0 * P.wj + dh;
  bound_pair(P.drmin - u1, u1 - P.drmax, dj, dh);
  rec[LJU + 1] = 2.0 * P.wdr * u1 + dj;
  rec[LHU + 1] = 2.0 * P.wdr + dh;
  const double2 sc = nt_sincos(x[2]);
  rec[LSN] = sc.x;
  rec[LCS] = sc.y;
  rec[LZ] = 0.0;
  rec[LO] = 1.0;
  rec[LDT] = P.dt;
  rec[LB30] = 0.5 * P.dt * P.dt;
  double J[3] = {2.0 * P.wx * (x[0] - c.goal(k, 0)), 2.0 * P.wy * (x[1] - c.goal(k, 1)), 2.0 * P.wth * (x[2] - c.goal(k, 2))};
  double H[9] = {2.0 * P.wx, 0.0, 0.0, 0.0, 2.0 * P.wy, 0.0, 0.0, 0.0, 2.0 * P.wth};
  const double* planes = c.planes();
  const double* seg = c.gseg();
  const int M = c.cnt[k];
#pragma unroll 1
  for (int d = 0; d < kDisc; ++d) {  // CorridorConsJacbian / Hessian (:690-727)
    const double lc = P.off[d] * sc.y, ls = P.off[d] * sc.x;
    const double xd = x[0] + lc, yd = x[1] + ls;
#pragma unroll 1
    for (int m = 0; m < M; ++m)
      s_plane_term(P, planes[(m * 3 + 0) * a.Kc + k], planes[(m * 3 + 1) * a.Kc + k], planes[(m * 3 + 2) * a.Kc + k], xd, yd,
                   lc, ls, J, H);
  }
#pragma unroll 1
  for (int d = 0; d < kDisc; ++d) {  // LaneBoundaryConsJacbian / Hessian (:729-769)
    const double lc = P.off[d] * sc.y, ls = P.off[d] * sc.x;
    const double xd = x[0] + lc, yd = x[1] + ls;
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
      const double* sg = seg + ((side == 0 ? 0 : a.S_left) + nidx[(k * kDisc + d) * 2 + side]) * kSegStride;
      s_plane_term(P, sg[7], sg[8], sg[9], xd, yd, lc, ls, J, H);
    }
  }
  rec[LJX + 0] = J[0];
  rec[LJX + 1] = J[1];
  rec[LJX + 2] = J[2];
  rec[LHX + 0] = H[0];
  rec[LHX + 1] = H[1];
  rec[LHX + 2] = H[2];
  rec[LHX + 3] = H[4];
  rec[LHX + 4] = H[5];
  rec[LHX + 5] = H[8];
  rec[LH10 + 0] = H[3];
  rec[LH10 + 1] = H[6];
  rec[LH10 + 2] = H[7];
}

__device__ __noinline__ void linearize_window(const Ctx& c, int k0, const double* Xs, const unsigned char* nidx,
                                              const DebugPtrs* dbg, int b) {
  const KernelArgs& a = c.a;
  const int N = a.N, K = N + 1;
  const int lane = c.lane;
  double* lin = c.smp() + a.sm.lin;
  double* R = c.linrec();
  const int k = k0 + lane;
  const int nk = K - k0 < kWin ? K - k0 : kWin;
  const bool lin_lane = lane < kWin && k < K;
  __syncwarp();
  if (lin_lane) s_linearize_knot(c, k, Xs, nidx, lin + lane * kLinStride);
  __syncwarp();
  if (dbg && lin_lane) {
    const double* rec = lin + lane * kLinStride;
    if (k < N) {
      if (dbg->A11) for (int i = 0; i < 12; ++i) dbg->A11[((size_t)b * N + k) * 12 + i] = rec[LA + i];
      if (dbg->Ju) for (int i = 0; i < 2; ++i) dbg->Ju[((size_t)b * N + k) * 2 + i] = rec[LJU + i];
      if (dbg->Hu) for (int i = 0; i < 2; ++i) dbg->Hu[((size_t)b * N + k) * 2 + i] = rec[LHU + i];
    }
    if (dbg->Jx) for (int i = 0; i < 6; ++i) dbg->Jx[((size_t)b * K + k) * 6 + i] = rec[LJX + i];
    if (dbg->Hx) for (int i = 0; i < 9; ++i) dbg->Hx[((size_t)b * K + k) * 9 + i] = rec[LHX + i];
  }
  static_assert(kLinStride == kRecStride, "strict records are flushed one to one");
#pragma unroll 1
  for (int idx = lane; idx < nk * kRecStride; idx += 32) R[(size_t)k0 * kRecStride + idx] = lin[idx];
  __syncwarp();
}

__device__ __forceinline__ void linearize_all(const Ctx& c, const double* Xs, const unsigned char* nidx,
                                              const DebugPtrs* dbg, int b) {
  for (int k0 = 0; k0 <= c.a.N; k0 += kWin) linearize_window(c, k0, Xs, nidx, dbg, b);
}

// ------------------------------------------------------------------------------------------
// C[n][q] = A[n][m] * B[m][q] with element (i, k) of A at A[i*ar + k*ac] (so a transpose is a stride swap) and
// element (k, j) of B at B[k*br + j*bc]; each coefficient is l(i,0) r(0,j) + l(i,1) r(1,j) + ... in that order
// (Eigen's lazy product of fixed-size matrices; oracle mat_mul).  One output element per lane.
__device__ __forceinline__ void s_mm(const double* A, int ar, int ac, const double* B, int br, int bc, double* C, int n,
                                     int m, int q, int lane) {
  for (int e = lane; e < n * q; e += 32) {
    const int i = e / q, j = e - i * q;
    double s = A[i * ar] * B[j * bc];
    for (int k = 1; k < m; ++k) s += A[i * ar + k * ac] * B[k * br + j * bc];
    C[e] = s;
  }
  __syncwarp();
}

// record -> dense A (6x6), B (6x2), Jx, Ju, Hx (6x6), Hu (2x2) in shared memory
__device__ __forceinline__ void s_expand(const double* rec, double* A, double* B, double* Jx, double* Ju, double* Hx,
                                         double* Hu, int lane) {
  for (int e = lane; e < 36; e += 32) {
    const int r = e / 6, cc = e - r * 6;
    A[e] = rec[f_off(r, cc)];
    double h = 0.0;
    if (r == cc) h = r < 3 ? rec[LHX + (r == 0 ? 0 : r == 1 ? 3 : 5)] : rec[LHX + 6 + r - 3];
    else if (r < 3 && cc < 3) h = r < cc ? rec[LHX + (r == 0 ? cc : 4)] : rec[LH10 + (r == 1 ? 0 : cc + 1)];
    Hx[e] = h;
  }
  if (lane < 12) B[lane] = rec[f_off(lane / 2, 6 + (lane & 1))];
  if (lane < 6) Jx[lane] = rec[LJX + lane];
  if (lane < 2) Ju[lane] = rec[LJU + lane];
  if (lane < 4) Hu[lane] = (lane == 0 || lane == 3) ? rec[LHU + (lane == 3)] : 0.0;
  __syncwarp();
}

// Backward (:334-390), a transcription of oracle cilqr_oracle_ctx_backward.
__device__ __noinline__ void backward_pass(const Ctx& c, double lambda, double dV[2]) {
  const KernelArgs& a = c.a;
  const int N = a.N, lane = c.lane;
  double* w = c.smp() + a.sm.scr;
  const double* R = c.linrec();
  double* gains = c.gains();
  double *A = w, *B = A + 36, *Jx = B + 12, *Ju = Jx + 6, *Hx = Ju + 2, *Hu = Hx + 36;             // 96
  double *Vx = Hu + 4, *Vxx = Vx + 6, *tmp6 = Vxx + 36, *tmp2 = tmp6 + 6, *Qx = tmp2 + 2, *Qu = Qx + 6;  // +58
  double *AtV = Qu + 2, *t36 = AtV + 36, *Qxx = t36 + 36, *BtV = Qxx + 36, *t4 = BtV + 12, *Quu = t4 + 4, *Qux = Quu + 4;
  double *Kg = Qux + 12, *kg = Kg + 12, *ninv = kg + 2, *KtQuu = ninv + 4, *a6 = KtQuu + 12, *b6 = a6 + 6, *c6 = b6 + 6;
  double *a36 = c6 + 6, *b36 = a36 + 36, *c36 = b36 + 36, *Vxn = c36 + 36, *Vxxn = Vxn + 6;
  static_assert(96 + 58 + 36 * 3 + 12 + 4 + 4 + 12 + 12 + 2 + 4 + 12 + 18 + 36 * 3 + 6 + 36 <= kScratch, "strict scratch");
  double dV0 = 0.0, dV1 = 0.0;
  {  // Vx = cost_Jx.back(), Vxx = cost_Hx.back()   (:343-344)
    const double* rec = R + (size_t)N * kRecStride;
    s_expand(rec, A, B, Jx, Ju, Hx, Hu, lane);
    if (lane < 6) Vx[lane] = Jx[lane];
    for (int e = lane; e < 36; e += 32) Vxx[e] = Hx[e];
    __syncwarp();
  }
#pragma unroll 1
  for (int i = N - 1; i >= 0; --i) {
    s_expand(R + (size_t)i * kRecStride, A, B, Jx, Ju, Hx, Hu, lane);
    s_mm(A, 1, 6, Vx, 1, 1, tmp6, 6, 6, 1, lane);  // A^T Vx
    s_mm(B, 1, 2, Vx, 1, 1, tmp2, 2, 6, 1, lane);  // B^T Vx
    if (lane < 6) Qx[lane] = Jx[lane] + tmp6[lane];
    if (lane < 2) Qu[lane] = Ju[lane] + tmp2[lane];
    s_mm(A, 1, 6, Vxx, 6, 1, AtV, 6, 6, 6, lane);  // A^T Vxx
    s_mm(AtV, 6, 1, A, 6, 1, t36, 6, 6, 6, lane);
    for (int e = lane; e < 36; e += 32) Qxx[e] = Hx[e] + t36[e];
    s_mm(B, 1, 2, Vxx, 6, 1, BtV, 2, 6, 6, lane);  // B^T Vxx
    s_mm(BtV, 6, 1, B, 2, 1, t4, 2, 6, 2, lane);
    if (lane < 4) Quu[lane] = Hu[lane] + t4[lane];
    s_mm(BtV, 6, 1, A, 6, 1, Qux, 2, 6, 6, lane);
    {  // (Quu + lambda I)^-1, closed form (:361-366)
      const double T0 = Quu[0] + lambda * 1.0, T1 = Quu[1] + lambda * 0.0, T2 = Quu[2] + lambda * 0.0,
                   T3 = Quu[3] + lambda * 1.0;
      const double det = T0 * T3 - T2 * T1;
      const double invdet = 1.0 / det;
      if (lane == 0) {
        ninv[0] = -(T3 * invdet);
        ninv[1] = -(-T1 * invdet);
        ninv[2] = -(-T2 * invdet);
        ninv[3] = -(T0 * invdet);
      }
      __syncwarp();
    }
    s_mm(ninv, 2, 1, Qux, 6, 1, Kg, 2, 2, 6, lane);
    s_mm(ninv, 2, 1, Qu, 1, 1, kg, 2, 2, 1, lane);
    if (lane < 12) gains[i * kGainStride + lane] = Kg[lane];
    if (lane < 2) gains[i * kGainStride + 12 + lane] = kg[lane];
    // Vx = Qx + K^T Quu k + K^T Qu + Qux^T k ; Vxx = Qxx + K^T Quu K + K^T Qux + Qux^T K   (:379-380)
    s_mm(Kg, 1, 6, Quu, 2, 1, KtQuu, 6, 2, 2, lane);
    s_mm(KtQuu, 2, 1, kg, 1, 1, a6, 6, 2, 1, lane);
    s_mm(Kg, 1, 6, Qu, 1, 1, b6, 6, 2, 1, lane);
    s_mm(Qux, 1, 6, kg, 1, 1, c6, 6, 2, 1, lane);
    if (lane < 6) Vxn[lane] = Qx[lane] + a6[lane] + b6[lane] + c6[lane];
    s_mm(KtQuu, 2, 1, Kg, 6, 1, a36, 6, 2, 6, lane);
    s_mm(Kg, 1, 6, Qux, 6, 1, b36, 6, 2, 6, lane);
    s_mm(Qux, 1, 6, Kg, 6, 1, c36, 6, 2, 6, lane);
    for (int e = lane; e < 36; e += 32) Vxxn[e] = Qxx[e] + a36[e] + b36[e] + c36[e];
    __syncwarp();
    if (lane < 6) Vx[lane] = Vxn[lane];
    // Vxx = 0.5 * (Vxx + Vxx^T), assigned in place, column-major, no temporary (quirk Q22, :381)
    if (lane == 0) {
      for (int e = 0; e < 36; ++e) Vxx[e] = Vxxn[e];
      for (int q = 0; q < 6; ++q)
        for (int r = 0; r < 6; ++r) Vxx[r * 6 + q] = 0.5 * (Vxx[r * 6 + q] + Vxx[q * 6 + r]);
    }
    __syncwarp();
    // delta_V (:383-384): Qu, Quu are lazy expressions, re-evaluated with the UPDATED Vx / Vxx (quirk Q21)
    s_mm(B, 1, 2, Vx, 1, 1, tmp2, 2, 6, 1, lane);
    s_mm(B, 1, 2, Vxx, 6, 1, BtV, 2, 6, 6, lane);
    s_mm(BtV, 6, 1, B, 2, 1, t4, 2, 6, 2, lane);
    {
      const double qn0 = Ju[0] + tmp2[0], qn1 = Ju[1] + tmp2[1];
      const double q0 = Hu[0] + t4[0], q1 = Hu[1] + t4[1], q2 = Hu[2] + t4[2], q3 = Hu[3] + t4[3];
      dV0 += kg[0] * qn0 + kg[1] * qn1;
      const double hk0 = 0.5 * kg[0], hk1 = 0.5 * kg[1];
      const double hq0 = hk0 * q0 + hk1 * q2, hq1 = hk0 * q1 + hk1 * q3;
      dV1 += hq0 * kg[0] + hq1 * kg[1];
    }
    __syncwarp();
  }
  dV[0] = dV0;
  dV[1] = dV1;
}

// iqr (:793-824), a transcription of oracle cilqr_oracle_ctx_iqr's sweep: K_k = (R + B'PB)^-1 B'PA,
// P = Q + A'P(A - BK).  The rollout adds -K (x - goal): the gains are stored negated, k = 0.  A_k, B_k about
// (goal_k, u = 0) are the records INIT wrote (iqr_records).
__device__ __noinline__ void s_iqr_sweep(const Ctx& c) {
  const KernelArgs& a = c.a;
  const int N = a.N, lane = c.lane;
  double* w = c.smp() + a.sm.scr;
  const double* R = c.linrec();
  double* gains = c.gains();
  double *A = w, *B = A + 36, *Jx = B + 12, *Ju = Jx + 6, *Hx = Ju + 2, *Hu = Hx + 36;
  double *Pm = Hu + 4, *BtP = Pm + 36, *S = BtP + 12, *G = S + 4, *inv = G + 12, *Kk = inv + 4, *BK = Kk + 12;
  double *AmBK = BK + 36, *AtP = AmBK + 36, *T = AtP + 36;
  static_assert(96 + 36 + 12 + 4 + 12 + 4 + 12 + 36 * 4 <= kScratch, "strict scratch");
  const double Qd[6] = {0.001, 0.001, 0.001, 0.001, 0.01, 0.005};
  for (int e = lane; e < 36; e += 32) Pm[e] = (e / 6 == e % 6) ? Qd[e / 6] : 0.0;
  __syncwarp();
#pragma unroll 1
  for (int i = N - 1; i >= 0; --i) {
    s_expand(R + (size_t)i * kRecStride, A, B, Jx, Ju, Hx, Hu, lane);
    s_mm(B, 1, 2, Pm, 6, 1, BtP, 2, 6, 6, lane);
    s_mm(BtP, 6, 1, B, 2, 1, S, 2, 6, 2, lane);
    s_mm(BtP, 6, 1, A, 6, 1, G, 2, 6, 6, lane);
    {
      const double S0 = 0.2 + S[0], S1 = 0.0 + S[1], S2 = 0.0 + S[2], S3 = 0.05 + S[3];
      const double det = S0 * S3 - S2 * S1;
      const double invdet = 1.0 / det;
      __syncwarp();
      if (lane == 0) {
        inv[0] = S3 * invdet;
        inv[1] = -S1 * invdet;
        inv[2] = -S2 * invdet;
        inv[3] = S0 * invdet;
      }
      __syncwarp();
    }
    s_mm(inv, 2, 1, G, 6, 1, Kk, 2, 2, 6, lane);
    if (lane < 12) gains[i * kGainStride + lane] = -Kk[lane];
    if (lane < 2) gains[i * kGainStride + 12 + lane] = 0.0;
    s_mm(B, 2, 1, Kk, 6, 1, BK, 6, 2, 6, lane);
    for (int e = lane; e < 36; e += 32) AmBK[e] = A[e] - BK[e];
    __syncwarp();
    s_mm(A, 1, 6, Pm, 6, 1, AtP, 6, 6, 6, lane);
    s_mm(AtP, 6, 1, AmBK, 6, 1, T, 6, 6, 6, lane);
    for (int e = lane; e < 36; e += 32) Pm[e] = ((e / 6 == e % 6) ? Qd[e / 6] : 0.0) + T[e];
    __syncwarp();
  }
}

// CalGradientNorm (:322-332): sequential mean of max_i |k_i| / (|u_i| + 1)
__device__ __noinline__ double s_gradient_norm(const Ctx& c, const double* Xs) {
  const KernelArgs& a = c.a;
  const double* gains = c.gains();
  double acc = 0.0;
  if (c.lane == 0) {
    for (int k = 0; k < a.N; ++k) {
      const double v0 = fabs(gains[k * kGainStride + 12]) / (fabs(Xs[6 * a.Kc + k]) + 1);
      const double v1 = fabs(gains[k * kGainStride + 13]) / (fabs(Xs[7 * a.Kc + k]) + 1);
      acc += (v0 > v1 ? v0 : v1);
    }
    acc = acc / a.N;
  }
  return __shfl_sync(kFull, acc, 0);
}
