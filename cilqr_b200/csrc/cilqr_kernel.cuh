// cilqr_kernel.cuh -- device code of the batched CILQR solver (sm_100a).
//
// One warp runs one trajectory PHASE at a time.  A persistent grid of one CTA per SM (16 warps) owns
// 128 scenario CONTEXTS that live in a global workspace (L2 / HBM): shrunk corridor planes, five
// trajectory slots (the iterate and four line-search candidates), feedback gains, linearisation
// records, lane segments, nearest-segment indices and a small header with the solver scalars.  The
// solve of a scenario is cut into phases -- INIT (load, shrink, LQ records of the initial guess), LIN
// (linearise + quadratise), BACK (Riccati sweep), ROLL (speculative rollout of four step sizes, eight
// contexts per warp), EVAL (cost of one candidate + accept/reject logic) -- and at any moment the CTA
// runs ONE phase type: warps claim waiting contexts of that type from a shared table and only move on
// to the type with the most waiting work when none is left (no barriers; see cilqr_solve_kernel).
// Reason (measured): a B200 SM feeds unaligned instruction streams at full rate
// only while their combined hot code fits ~32 KB; the whole solver is ~75 KB of fp64 code, so
// free-running warps at different phases are instruction-fetch bound at 1 warp per scheduler (issue
// 20 %, identical from 4 to 12 warps/SM), while warps that run the same phase together scale (2.7x at
// 12 warps/SM).  With contexts in global memory the shared-memory stage per warp is only what the
// running phase needs (lane segments + heading table + a plane tile for EVAL, a record window + plane
// tile for LIN, a record ring + Riccati scratch for BACK, cp.async rings of gains / nominal trajectory
// for ROLL), so 16 warps fit at any horizon.
//
// All arithmetic is IEEE double like the reference (Eigen Matrix<double,...>); no tensor cores.
//
// Reference functions re-created here (file:line relative to the reference root):
//   ShrinkConstraints / NormalizeHalfPlane   algorithm/ilqr/ilqr_optimizer.cc:438-495
//   iqr (LQR initial guess)                  algorithm/ilqr/ilqr_optimizer.cc:793-842
//   Dynamics / DynamicsJacbian               algorithm/ilqr/vehicle_model.cc:88-121, 21-86
//   RelaxBarrierFunction                     algorithm/ilqr/barrier_function.h:104-140
//   TotalCost (J, Dynamics, Corridor, Lane)  algorithm/ilqr/ilqr_optimizer.cc:417-436, 497-603
//   FindNeastLaneSegment / DistanceTo        algorithm/ilqr/ilqr_optimizer.cc:605-618,
//                                            algorithm/math/line_segment2d.cpp:61-75
//   CostJacbian / CostHessian                algorithm/ilqr/ilqr_optimizer.cc:620-769
//   Backward / Forward / CalGradientNorm     algorithm/ilqr/ilqr_optimizer.cc:334-415, 322-332
//   Optimize (line search, lambda schedule)  algorithm/ilqr/ilqr_optimizer.cc:154-320
//   TransformToTrajectory                    algorithm/ilqr/ilqr_optimizer.cc:771-791
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

// CILQR_STRICT = 1 builds the PARITY INSTRUMENT (libcilqr_b200_strict.so, compiled with -fmad=false): the same
// scheduler, contexts, line search and data flow as the production kernel, but every floating-point expression is
// evaluated in the reference's own order (sequential sums over knots / discs / planes, one log per barrier term,
// IEEE division, the dense Eigen-ordered Riccati step with the in-place symmetrisation Q22, the reference's iqr
// update) and the five libm functions come from the portable pm_math.h that the oracle build
// libcilqr_oracle_pm.so shares -- so its output can be compared with that oracle BIT FOR BIT
// (tests/test_gpu_strict.py).  It proves the kernel's logic; the production build (0) re-associates sums, fuses
// multiply-adds and uses CUDA's libm, which moves results by rounding only.
#ifndef CILQR_STRICT
#define CILQR_STRICT 0
#endif
#if CILQR_STRICT
#include "pm_math.h"
#define CILQR_FMA(a, b, c) ((a) * (b) + (c))
#define CILQR_DIVL(P, x) ((x) / (P).L)
#define CILQR_LQ(P) ((P).L)          // the wheel-base operand as a value ...
#define CILQR_DIVQ(q, x) ((x) / (q))  // ... and the division by it
#else
#define CILQR_FMA(a, b, c) fma((a), (b), (c))
#define CILQR_DIVL(P, x) ((x) * (P).inv_L)
#define CILQR_LQ(P) ((P).inv_L)
#define CILQR_DIVQ(q, x) ((x) * (q))
#endif

namespace cilqr {

// The CTA's dynamic shared memory: one stage per warp.  Declared at namespace scope and addressed as
// smem_cta + offset inside every phase function, so that the compiler KNOWS the accesses are to shared memory and
// emits LDS / STS with 32-bit addresses; a stage pointer handed through structs and __noinline__ calls is a generic
// pointer and costs a 64-bit address computation plus a generic LD / ST per access (profiles/r02_sass_summary.txt).
extern __shared__ __align__(16) double smem_cta[];

constexpr int kNX = 6;
constexpr int kNU = 2;
constexpr int kDisc = 5;
constexpr int kNAlpha = 11;
constexpr int kSpec = 4;          // step sizes rolled out speculatively per ROLL phase
constexpr int kTrajSlots = kSpec + 1;  // iterate + candidates
#if CILQR_STRICT
constexpr int kLinStride = 40;    // strict: the lower triangle of Hx's 3x3 block is carried too (it is not bit-symmetric)
#else
constexpr int kLinStride = 37;    // doubles per knot in the linearisation window
#endif
constexpr int kSegStride = 10;    // sx sy ex ey ux uy len a b c
static_assert(kSegStride % 2 == 0, "seg_dist2 reads a record with 16-byte loads");
constexpr int kGainStride = 16;   // K (2x6), k (2), pad (2): one 128-byte record per knot
constexpr int kRollChunk = 4;     // knots per cp.async stage of the rollout ring
constexpr int kRecStride = 40;    // doubles per knot of the linearisation records in the context (16-byte pieces)
constexpr int kBackChunk = 4;     // knots per cp.async stage of the Riccati ring
constexpr int kRingDoubles = 2 * (8 * kRollChunk + kGainStride * kRollChunk);
#ifndef CILQR_GROUP
#define CILQR_GROUP 8
#endif
constexpr int kGroup = CILQR_GROUP;  // lane segments per bounding-circle group of the pruned nearest search
#if CILQR_STRICT
constexpr int kScratch = 768;     // strict: dense 6x6 temporaries of the Eigen-ordered Riccati step
#else
constexpr int kScratch = 192;     // doubles of per-warp Riccati scratch
#endif
constexpr int kPlaneKnots = 8;    // knots per staged tile of corridor planes
constexpr int kPlaneTile = 3 * kPlaneKnots;  // doubles per plane (a, b, c rows) in a tile; x M_max per buffer
#ifndef CILQR_TILE_BUFS
#define CILQR_TILE_BUFS 1
#endif
// 2: the next tile is copied while the whole current chunk is processed; 1: (less shared memory, for 16
// warps per SM) the next tile is copied into the same buffer behind the lane-boundary part of the chunk
constexpr int kTileBufs = CILQR_TILE_BUFS;
constexpr int kHdrDoubles = 24;   // sizeof(CtxHdr) / 8
#ifndef CILQR_MAX_CTX
#define CILQR_MAX_CTX 128
#endif
constexpr int kMaxCtx = CILQR_MAX_CTX;  // contexts per CTA (power of two, multiple of 32)
constexpr int kCtxWords = kMaxCtx / 32;
static_assert(kMaxCtx % 32 == 0 && (kMaxCtx & (kMaxCtx - 1)) == 0 && kMaxCtx <= 256, "context table size");
constexpr unsigned kFull = 0xffffffffu;
#ifndef CILQR_LIN_WINDOW
#define CILQR_LIN_WINDOW 16
#endif
constexpr int kWin = CILQR_LIN_WINDOW;  // knots per linearisation window (<= 32: lane == knot inside a window)
static_assert(kWin >= 1 && kWin <= 32, "linearisation window is at most one knot per lane");
#ifndef CILQR_CTA_WARPS
#define CILQR_CTA_WARPS 16
#endif
constexpr int kCtaWarps = CILQR_CTA_WARPS;  // warps per CTA (one CTA per SM)

enum Phase : int { PH_INIT = 0, PH_BACK = 1, PH_ROLL = 2, PH_EVAL = 3, PH_LIN = 4, PH_DONE = 5 };
constexpr int kNumTypes = 5;

// linearisation record offsets
constexpr int LA = 0;    // A02 A03 A04 A05 A12 A13 A14 A15 A23 A24 A25
constexpr int LB21 = 11;
constexpr int LJX = 12;  // 6
constexpr int LJU = 18;  // 2
constexpr int LHX = 20;  // H00 H01 H02 H11 H12 H22 H33 H44 H55
constexpr int LHU = 29;  // 2
constexpr int LZ = 31;   // constants 0, 1, dt, dt^2/2 so that A, B, H can be gathered by offset
constexpr int LO = 32;
constexpr int LDT = 33;
constexpr int LB30 = 34;
constexpr int LSN = 35;  // sin, cos of the heading (consumed by linearize_discs)
constexpr int LCS = 36;
constexpr int LH10 = 37;  // strict build only: H10 H20 H21
// strict build: doubles per lane of the term buffer of eval_cost (corridor terms + two lane terms, at least the
// seven per-knot terms of the first pass)
__host__ __device__ constexpr int strict_row_width(int M_max) { return M_max + 2 < 8 ? 8 : M_max + 2; }

struct DevParams {
  double dt, L, inv_L, rt, eps, inv_eps, inv_eps2, relax_c;  // relax_c = -0.5*rt - rt*log(eps)
  double rt_log_eps;  // rt * log(eps): the constant of the relaxed barrier branch as the reference writes it (strict build)
  double vmax, amin, amax, dmin, dmax, jmin, jmax, drmin, drmax;
  double wx, wy, wth, wv, wa, wd, wj, wdr;
  double abs_tol, rel_tol;
  double off[kDisc];  // L_disc*(j-0.5) - rear_hang   (ilqr_optimizer.cc:556-565)
  double shrink_corr, shrink_lane;
  int max_iter;
};

// Per-warp shared-memory stage, offsets in doubles.  Phases alias each other's regions:
//   EVAL: seg, grp, trig, pl_e   BACK: lin, scr, pl_b   ROLL: ring   INIT: seg, grp (built here), scr
// pl_*: two buffers of M_max * kPlaneTile doubles, the cp.async double buffer of corridor-plane tiles
struct SmemLayout {
  int seg, grp, trig, lin, scr, ring, pl_e, pl_b, bring, red;
  int total_bytes;
};

// One scenario context in the global workspace, offsets in doubles from the context base.
struct CtxLayout {
  int planes;  // [M_max][3][Kc]    shrunk + normalised half-planes, knot-minor
  int slots;   // [5][8][Kc]        trajectories x0..x5,u0,u1 component-major (lane == knot coalesces)
  int gains;   // [Npad][16]        K (2x6 row-major), k (2), pad
  int lin;     // [Kpad][40]        linearisation records of the iterate (LIN -> BACK)
  int seg;     // [S_left+S_right][10]
  int grp;     // [groups][3]       bounding circles cx, cy, r
  int nidx;    // one byte array [K][5][2] of nearest-segment indices per trajectory slot
  int hdr;     // CtxHdr
  int nidx_bytes;
  int stride;  // doubles per context
};

// Solver scalars of one context (the locals of IlqrOptimizer::Optimize, ilqr_optimizer.cc:182-199).
struct CtxHdr {
  double lambda, dlambda, cost_old, dV0, dV1;
  double cost_acc[5];
  double g0[6];              // goals_[0] = (x0, y0, theta0, v0, 0, 0)     :151
  unsigned int b, ahash;     // scenario id, FNV-1a over the line-search outcome per iteration
  int iter, status, cur, nflip, ai, gb, rmode, emode, n_cost, n_iter_traj;
  unsigned int retired, deferred;  // line-search lanes whose rollout blew up / needs the general wrap
  int imode;  // 1 while the context builds its initial guess (the BACK phase then runs the iqr sweep)
  int wait;   // the phase this context waits for, written when it is handed to a later launch (drain relay)
};
static_assert(sizeof(CtxHdr) == kHdrDoubles * 8, "CtxHdr size");

struct DebugPtrs {
  double *corridor, *lanes, *X0, *U0, *cost0, *A11, *Jx, *Ju, *Hx, *Hu, *Kg, *kg, *dV, *Xn, *Un, *costn;
  int32_t* nearest;
  double* gnorm;
};

struct KernelArgs {
  DevParams P;
  SmemLayout sm;
  CtxLayout cl;
  int B, N, M_max, S_left, S_right, Kc;  // Kc: knot pitch of the plane / trajectory arrays
  int ctx_per_cta;
  int hot_live;  // when at most this many contexts of the CTA are alive, every context is treated as hot
  int hot_iter;  // a scenario that has run this many iterations keeps its warp through all phases until it exits
  const double* start;
  const double* coarse;
  const double* corridor;
  const int32_t* corridor_cnt;
  const double* lane_left;
  const double* lane_right;
  int init_mode;                 // 0 iqr, 1 open-loop rollout of guess_controls, 2 (guess_states, guess_controls) as given
  const double* guess_states;    // [B][K][6]
  const double* guess_controls;  // [B][N][2]
  double* states;
  double* controls;
  double* status;
  double* trajectory;
  double* result;
  double* init_states;
  double* init_controls;
  double* cost_hist;
  double* iter_states;
  double* iter_controls;
  int32_t* hist_len;
  int hist_cap;
  double* ws;            // [gridDim.x][ctx_per_cta][cl.stride]
  unsigned int* ticket;  // [0] scenario counter, [1] error bits (kErr*), [2] scenarios finished
  const unsigned int* ready;  // host path: scenarios below *ready have arrived on the device (NULL: all)
  unsigned long long watchdog_ns;  // how long INIT waits for the watermark before it flags kErrStarved
  // Drain relay (see cilqr_solve_kernel): once the batch's ticket is exhausted and a CTA has at most donate_thr
  // unfinished contexts left, it appends them to `donate` ([0] = count, then global context indices) and exits;
  // the next launch of the relay adopts `resume` (same layout), resume_per_cta contexts per CTA.
  int donate_thr;
  unsigned int* donate;
  const unsigned int* resume;
  int resume_per_cta;
  unsigned long long* stats;  // optional [8 + 2 + 256]: completion-time histogram (2 ms buckets) after the counters; [8]: scheduler passes, idle polls, failed claims, phases run by type (4), type switches
  DebugPtrs dbg;
  int debug;             // 1: stop after the first line-search evaluation and dump stages
};

// error bits a launch can raise in ticket[1]; the host turns any of them into CILQR_E_TIMEOUT
constexpr unsigned kErrStarved = 1u;   // INIT gave up waiting for the host path's watermark
constexpr unsigned kErrIdle = 2u;      // a warp's idle poll ran into its watchdog
constexpr unsigned kErrHelp = 4u;      // a help-board wait ran into its watchdog
__device__ __forceinline__ void raise_error(const KernelArgs& a, unsigned bits) { atomicOr(a.ticket + 1, bits); }
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__constant__ double kAlphaList[kNAlpha] = {1.0000, 0.5012, 0.2512, 0.1259, 0.0631, 0.0316,
                                           0.0158, 0.0079, 0.0040, 0.0020, 0.0010};

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return __shfl_sync(kFull, v, 0);
}

// ---- out-of-line math: ONE copy of each double-precision libdevice routine in the kernel.
// Inlined at every call site they made the kernel ~300 KB of SASS, which thrashed the instruction
// cache (ncu: stall_no_instruction 5.4 cycles per issued instruction); see DESIGN.md.
#if CILQR_STRICT
__device__ __noinline__ double nt_tan(double x) { return pm_tan(x); }
__device__ __noinline__ double2 nt_tan2(double x, double y) { return make_double2(pm_tan(x), pm_tan(y)); }
__device__ __noinline__ double nt_log(double x) { return pm_log(x); }
__device__ __noinline__ double nt_hypot(double x, double y) { return pm_hypot(x, y); }
__device__ __noinline__ double2 nt_sincos(double x) {
  double s, c;
  pm_sincos(x, &s, &c);
  return make_double2(s, c);
}
#else
__device__ __noinline__ double nt_tan(double x) { return tan(x); }
// two independent tangents in one call: the two polynomial chains interleave (the rollout is a serial
// dependency chain, so instruction-level parallelism inside a step is all there is)
__device__ __noinline__ double2 nt_tan2(double x, double y) { return make_double2(tan(x), tan(y)); }
__device__ __noinline__ double nt_log(double x) { return log(x); }
__device__ __noinline__ double nt_hypot(double x, double y) { return hypot(x, y); }
__device__ __noinline__ double2 nt_sincos(double x) {
  double s, c;
  sincos(x, &s, &c);
  return make_double2(s, c);
}
#endif
// general branch of NormalizeAngle (math_utils.cpp:53-59) on a = angle + pi
__device__ __noinline__ double nt_wrap_general(double a) {
  const double kTwoPi = 2.0 * 3.14159265358979323846;
  a = fmod(a, kTwoPi);
  if (a < 0.0) a += kTwoPi;
  return a;
}

// math_utils.cpp:53-59.  fmod(a + pi, 2pi) == a + pi exactly whenever 0 <= a + pi < 2pi, which is
// the only case a sane rollout produces.  allow_general = false: never enter the fmod branch; `slow`
// reports that it was needed (the caller then discards the rollout and repeats it faithfully).
__device__ __forceinline__ double wrap_angle(double angle, bool allow_general, bool& slow) {
  const double kPi = 3.14159265358979323846;
  const double kTwoPi = 2.0 * 3.14159265358979323846;
  double a = angle + kPi;
  if (!(a >= 0.0 && a < kTwoPi)) {
    if (allow_general) a = nt_wrap_general(a);
    else slow = true;
  }
  return a - kPi;
}
__device__ __forceinline__ double normalize_angle(double angle) {
  bool unused = false;
  return wrap_angle(angle, true, unused);
}

// vehicle_model.cc:88-138: midpoint RK2, same control at both stages, wrap theta and delta.
// Division by the wheel base is a multiplication by its reciprocal (bit-identical for the
// reference's L_w = 1.0, vehicle_param.h:30; <= 1 ulp otherwise).
// (theta enters k1 only through k1x, k1y, which the midpoint step never uses.)
// dt and lq = CILQR_LQ(P) are values the caller loaded once (see bar_add).
__device__ __forceinline__ void rollout_step(double dt, double lq, double* x, double u0, double u1, bool allow_general,
                                             bool& slow) {
  const double h = 0.5 * dt;
  const double m5 = x[5] + h * u1;
  const double de = wrap_angle(x[5], allow_general, slow);
  const double dem = wrap_angle(m5, allow_general, slow);
  const double2 tt = nt_tan2(de, dem);
  const double k1t = CILQR_DIVQ(lq, x[3] * tt.x);
  const double m2 = x[2] + h * k1t, m3 = x[3] + h * x[4], m4 = x[4] + h * u0;
  const double thm = wrap_angle(m2, allow_general, slow);
  const double2 sc = nt_sincos(thm);
  const double k2x = m3 * sc.y, k2y = m3 * sc.x, k2t = CILQR_DIVQ(lq, m3 * tt.y);
  x[0] = x[0] + dt * k2x;
  x[1] = x[1] + dt * k2y;
  x[2] = wrap_angle(x[2] + dt * k2t, allow_general, slow);
  x[3] = x[3] + dt * m4;
  x[4] = x[4] + dt * u0;
  x[5] = wrap_angle(x[5] + dt * u1, allow_general, slow);
}

// vehicle_model.cc:21-86.  Writes the 11 state-dependent entries of A and B(2,1).
__device__ __noinline__ void dynamics_jacobian(const DevParams& P, const double* x, double u1,
                                               double* A11, double* b21) {
  const double dt = P.dt;
  const double v = x[3];
  const double theta = normalize_angle(x[2]);
  const double delta = normalize_angle(x[5]);
  const double a = x[4];
  const double2 tt = nt_tan2(delta, delta + 0.5 * dt * u1);
  const double tan_delta = tt.x, tan_dr = tt.y;
  const double theta_mid = theta + CILQR_DIVL(P, 0.5 * dt * v * tan_delta);
  const double2 scm = nt_sincos(theta_mid);
  const double sm = scm.x, cm = scm.y;
  const double td2 = tan_delta * tan_delta;
  const double tr2 = tan_dr * tan_dr;
  const double vm = 0.5 * a * dt + v;
  A11[0] = -dt * vm * sm;
  A11[1] = dt * cm - CILQR_DIVL(P, 0.5 * dt * dt * vm * sm * tan_delta);
  A11[2] = 0.5 * dt * dt * cm;
  A11[3] = CILQR_DIVL(P, -0.5 * dt * dt * v * vm * (td2 + 1) * sm);
  A11[4] = dt * vm * cm;
  A11[5] = dt * sm + CILQR_DIVL(P, 0.5 * dt * dt * vm * cm * tan_delta);
  A11[6] = 0.5 * dt * dt * sm;
  A11[7] = CILQR_DIVL(P, 0.5 * dt * dt * v * vm * (td2 + 1) * cm);
  A11[8] = CILQR_DIVL(P, dt * tan_dr);
  A11[9] = CILQR_DIVL(P, 0.5 * dt * dt * tan_dr);
  A11[10] = CILQR_DIVL(P, dt * (v * (tr2 + 1)));
  *b21 = CILQR_DIVL(P, 0.5 * dt * dt * v * (tr2 + 1));
}

// Barrier value accumulator: sum of -rt*log(-g) over the log branch is -rt*log(prod(-g)).
struct BarAcc {
  double prod, quad;
};
// quadratic extension of the barrier for g >= -eps (barrier_function.h:108-112): rare, kept out of line
__device__ __noinline__ double bar_quad(double g, const DevParams& P) {
  const double q = (-g - 2.0 * P.eps) * P.inv_eps;
  return fma(0.5 * P.rt * q, q, P.relax_c);
}
// eps / rt are passed as VALUES the caller loaded once: a member of `const DevParams&` is re-read from the parameter block
// (a generic load, two R2UR and a long-scoreboard wait) after every store or call the compiler cannot disambiguate
// from it -- once per half-plane in the loops below (profiles/r02_q_ncu_summary.txt, line 375 of that revision).
__device__ __forceinline__ void bar_add(BarAcc& a, double g, double eps, const DevParams& P) {
  if (g < -eps) a.prod *= -g;
  else a.quad += bar_quad(g, P);
}
__device__ __forceinline__ double bar_value(const BarAcc& a, double rt) {
  return a.quad - rt * nt_log(a.prod);
}
// 1 / g for a normal, non-zero g (here g < -eps): hardware seed (relative error < 1e-6) + two Newton steps.  Measured on
// B200 over 6.2e8 log-uniform operands of both signs: 0 ulp from the IEEE division, with two steps as with three
// (tools/microbench/rcp_steps.cu, profiles/r02_s_rcp_steps.json) -- a fifth of the division's instructions, no slow-path call.
__device__ __forceinline__ double fast_rcp(double g) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(g));
  double e = fma(-g, r, 1.0);
  r = fma(r, e, r);
  e = fma(-g, r, 1.0);
  r = fma(r, e, r);
  return r;
}
// barrier_function.h:115-140: coefficient of dx in the Jacobian (cj), of dx dx^T (co) and of ddx (cd)
__device__ __forceinline__ void bar_coef(double g, double eps, double rt, double inv_eps2, double& cj, double& co,
                                         double& cd) {
  if (g < -eps) {
    const double inv = fast_rcp(g);
    const double q = rt * inv;
    cj = -q;
    co = q * inv;
    cd = q;
  } else {
    cj = rt * (g + 2.0 * eps) * inv_eps2;
    co = cj;
    cd = 0.0;
  }
}

// the same coefficients with the Jacobian's negated (mj = -cj): in the log branch mj = cd = rt / g needs no sign flip (an
// FP64 negation is a DADD), and the accumulation uses the free operand negation of the multiply-add
__device__ __forceinline__ void bar_coef_neg(double g, double eps, double rt, double inv_eps2, double& mj, double& co,
                                             double& cd) {
  if (g < -eps) {
    const double inv = fast_rcp(g);
    const double q = rt * inv;
    mj = q;
    co = q * inv;
    cd = q;
  } else {
    co = rt * (g + 2.0 * eps) * inv_eps2;
    mj = -co;
    cd = 0.0;
  }
}

// squared distance point -> segment with the case split of line_segment2d.cpp:61-75
// sg: a segment record in the shared-memory stage (80-byte records from a 16-byte aligned base: three 16-byte loads)
__device__ __forceinline__ double seg_dist2(const double* sg, double px, double py) {
  const double2 s = *reinterpret_cast<const double2*>(sg), e = *reinterpret_cast<const double2*>(sg + 2);
  const double2 u = *reinterpret_cast<const double2*>(sg + 4);
  const double x0 = px - s.x, y0 = py - s.y;
  const double x1 = px - e.x, y1 = py - e.y;
  const double proj = x0 * u.x + y0 * u.y;
  const double cr = x0 * u.y - y0 * u.x;
  const double d0 = fma(x0, x0, y0 * y0);
  const double d1 = fma(x1, x1, y1 * y1);
  double d = cr * cr;
  d = (proj >= sg[6]) ? d1 : d;
  d = (proj <= 0.0) ? d0 : d;
  return d;
}

// ---- cp.async (LDGSTS) helpers: 16-byte global -> shared copies that bypass registers and L1
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// What one warp needs to run one phase of one context.
struct Ctx {
  const KernelArgs& a;
  int sm_off;   // this warp's shared-memory stage: smem_cta + sm_off (doubles)
  double* cx;   // this context in the global workspace
  CtxHdr* h;
  int lane;
  bool seg_staged;      // this warp's stage already holds the context's lane segments (chained EVALs)
  const double* goals;  // global [K][6] (row 0 is replaced by h->g0)
  const int32_t* cnt;   // global [K]
  __device__ __forceinline__ double* smp() const { return smem_cta + sm_off; }
  __device__ Ctx(const KernelArgs& a_, int s, double* c, int l)
      : a(a_), sm_off(s), cx(c), h(reinterpret_cast<CtxHdr*>(c + a_.cl.hdr)), lane(l), seg_staged(false), goals(nullptr), cnt(nullptr) {}
  __device__ __forceinline__ void bind(unsigned b) {
    goals = a.coarse + (size_t)b * (a.N + 1) * 6;
    cnt = a.corridor_cnt + (size_t)b * (a.N + 1);
  }
  __device__ __forceinline__ double goal(int k, int c) const { return k == 0 ? h->g0[c] : goals[k * 6 + c]; }
  __device__ __forceinline__ double* planes() const { return cx + a.cl.planes; }
  __device__ __forceinline__ double* slot(int i) const { return cx + a.cl.slots + i * 8 * a.Kc; }
  __device__ __forceinline__ double* gains() const { return cx + a.cl.gains; }
  __device__ __forceinline__ double* linrec() const { return cx + a.cl.lin; }
  __device__ __forceinline__ double* gseg() const { return cx + a.cl.seg; }
  __device__ __forceinline__ double* ggrp() const { return cx + a.cl.grp; }
  __device__ __forceinline__ unsigned char* nidx(int i) const {
    return reinterpret_cast<unsigned char*>(cx + a.cl.nidx) + i * a.cl.nidx_bytes;
  }
};
// physical trajectory slot of line-search candidate `ai` while slot `cur` holds the iterate
__device__ __forceinline__ int cand_slot(int cur, int ai) {
  const int s = cur + 1 + (ai & (kSpec - 1));
  return s >= kTrajSlots ? s - kTrajSlots : s;
}

// Corridor planes of knots [k_lo, k_lo + nk), planes m < Mrows: context (global, [m][a|b|c][knot]) ->
// shared tile buf[(m * 3 + comp) * kPlaneKnots + knot - k_lo] by cp.async, one commit group.  The
// consumer overlaps the copy of the next tile with the arithmetic of the current one, so the inner
// loops never wait on L2 / HBM.
// (planes = c.planes() and Kc as values: the callers load them once, not once per tile)
__device__ __forceinline__ void stage_planes(const double* planes, int Kc, int lane, double* buf, int k_lo, int nk, int Mrows) {
  static_assert(32 % kPlaneKnots == 0, "a lane keeps its knot column");
  constexpr int kRowsPerIter = 32 / kPlaneKnots;
  const int kk = lane % kPlaneKnots, r0 = lane / kPlaneKnots;
  if (kk < nk) {
    const double* src = planes + k_lo + kk + (size_t)r0 * Kc;
    unsigned dst = (unsigned)__cvta_generic_to_shared(buf + lane);
    const size_t sstep = (size_t)kRowsPerIter * Kc;
#pragma unroll 1  // (asynchronous copies: nothing to overlap by unrolling, and this body is inlined four times)
    for (int row = r0; row < Mrows * 3; row += kRowsPerIter) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
      src += sstep;
      dst += 32 * 8;
    }
  }
  cp_async_commit();
}

// ------------------------------------------------------------------------------------------
// Nearest lane segment of one disc centre (FindNeastLaneSegment, ilqr_optimizer.cc:605-618) without
// scanning all S segments: segments are bundled in groups of kGroup with a bounding circle (centre,
// radius) built at scenario load.  The exact distance to the segment that was nearest in the current
// iterate (`guess`) is an upper bound ub on the minimum; a group whose circle is farther from the disc
// centre than ub + radius cannot contain the minimiser -- nor tie with it -- and is skipped.  Surviving
// groups are scanned in index order with the reference's strict '<', so the arg-min (first minimum)
// is the brute-force one.
// Fast path ("certificate"): cert[s] (built at scenario load, see phase_init) is the square of half a LOWER bound on
// the distance between segment s and every segment more than kNearWin indices away from it.  If the disc centre p
// is closer than that to the guess segment g, then for any point r of such a far segment
//     |p - r| >= dist(seg g, seg j) - dist(p, seg g) > dist(p, seg g) >= min over all segments,
// strictly -- a far segment can neither be the minimiser nor tie with it -- so the first minimum of the brute-force
// scan lies in the window [g - kNearWin, g + kNearWin], which is scanned in index order with the reference's
// strict '<'.  The bounds are deflated by 1e-9 relative + 1e-9 absolute, nine orders above the rounding of the
// distances that are compared.
constexpr int kNearWin = 3;
__device__ __forceinline__ int nearest_segment(const double* sg0, const double* gp, const double* cert, int S, int guess,
                                               double xd, double yd) {
  const int gi = guess < S ? guess : S - 1;
  const double ddg = seg_dist2(sg0 + gi * kSegStride, xd, yd);
  if (ddg < cert[gi]) {  // (NaN compares false: the general search below handles it)
    const int lo = gi - kNearWin > 0 ? gi - kNearWin : 0;
    const int hi = gi + kNearWin < S - 1 ? gi + kNearWin : S - 1;
    double best = 1.7976931348623157e308;
    int bi = gi;
#pragma unroll 1  // (unroll 2: -4.5 %, profiles/r02_t_ab_rcp_lds128_unroll.log -- eval_cost is at the edge of the instruction cache)
    for (int s = lo; s <= hi; ++s) {
      const double dd = s == gi ? ddg : seg_dist2(sg0 + s * kSegStride, xd, yd);
      if (dd < best) {
        best = dd;
        bi = s;
      }
    }
    return bi;
  }
  const double ub = sqrt(ddg);
  const int ng = (S + kGroup - 1) / kGroup;
  double best = 1.7976931348623157e308;
  int bi = 0;
#pragma unroll 1
  for (int g = 0; g < ng; ++g) {
    const double dcx = xd - gp[g * 3], dcy = yd - gp[g * 3 + 1];
    const double thr = ub + gp[g * 3 + 2];
    if (fma(dcx, dcx, dcy * dcy) > thr * thr) continue;  // (NaN compares false: never pruned)
    const int s_hi = (g + 1) * kGroup < S ? (g + 1) * kGroup : S;
#pragma unroll 1  // (code size before ILP: profiles/r02_u_ab_unroll_reduction.log)
    for (int s = g * kGroup; s < s_hi; ++s) {
      const double dd = seg_dist2(sg0 + s * kSegStride, xd, yd);
      if (dd < best) {
        best = dd;
        bi = s;
      }
    }
  }
  return bi;
}

#if !CILQR_STRICT
// ------------------------------------------------------------------------------------------
// TotalCost of the trajectory in slot Xs (ilqr_optimizer.cc:417-436), [8][Kc]: x0..x5, u0, u1, in two
// passes:
//   pass 1, lane == knot:          JCost + DynamicsCost (:497-551), sin/cos of the heading -> trig
//   pass 2, lane == (knot, disc):  CorridorCost + LaneBoundaryCost (:553-603) of ONE disc per lane
// (one disc per lane keeps the code five times smaller than a disc-unrolled knot-per-lane body and
// fills 505 of 512 lane slots at K = 101 instead of 101 of 128).  Also records the nearest lane
// segment of every (knot, disc, side) in nidx for the linearisation that follows an accepted step.
// The lane segments / group circles must have been staged in shared memory (stage_segments).
__device__ __noinline__ void eval_cost(const Ctx& c_ref, const double* Xs, const unsigned char* guess,
                                       unsigned char* nidx, double cost5[5]) {
  const Ctx c = c_ref;  // a private copy (registers): the caller's Ctx lives in local memory and every char store below could alias it
  const KernelArgs& a = c.a;
  const DevParams& P = a.P;
  const int K = a.N + 1, N = a.N, Kc = a.Kc, S_left = a.S_left, S_right = a.S_right;
  const double eps = P.eps, rt = P.rt;  // loaded once (see bar_add)
  const double* seg = c.smp() + a.sm.seg;
  double* trig = c.smp() + a.sm.trig;  // [K][2] sin, cos
  double sum_j = 0.0, sum_d = 0.0, sum_c = 0.0, sum_l = 0.0;
#pragma unroll 1
  for (int k = c.lane; k < K; k += 32) {
    const double* xc = Xs + k;
    const double px = xc[0], py = xc[Kc], th = xc[2 * Kc], v = xc[3 * Kc], ac = xc[4 * Kc], de = xc[5 * Kc];
    const double dx = px - c.goal(k, 0), dy = py - c.goal(k, 1), dth = th - c.goal(k, 2);
    double tj = P.wx * (dx * dx) + P.wy * (dy * dy) + P.wth * (dth * dth);
    BarAcc bd = {1.0, 0.0};
    bar_add(bd, -v, eps, P);
    bar_add(bd, v - P.vmax, eps, P);
    bar_add(bd, ac - P.amax, eps, P);
    bar_add(bd, P.amin - ac, eps, P);
    bar_add(bd, de - P.dmax, eps, P);
    bar_add(bd, P.dmin - de, eps, P);
    if (k < N) {
      const double u0 = Xs[6 * Kc + k], u1 = Xs[7 * Kc + k];
      tj += P.wj * (u0 * u0) + P.wdr * (u1 * u1);
      bar_add(bd, u0 - P.jmax, eps, P);
      bar_add(bd, P.jmin - u0, eps, P);
      bar_add(bd, u1 - P.drmax, eps, P);
      bar_add(bd, P.drmin - u1, eps, P);
    }
    sum_j += tj;
    sum_d += bar_value(bd, rt);
    const double2 scth = nt_sincos(th);
    trig[k * 2] = scth.x;
    trig[k * 2 + 1] = scth.y;
  }
  __syncwarp();
  const int items = K * kDisc;
  const int ngl = (S_left + kGroup - 1) / kGroup, ngr = (S_right + kGroup - 1) / kGroup;
  const double* grp = c.smp() + a.sm.grp;
  const double* cert = grp + (ngl + ngr) * 3;  // [S_left + S_right] certificate radii, behind the group circles
  double* pbuf = c.smp() + a.sm.pl_e;
  const int pstride = a.M_max * kPlaneTile;
  const double* planes_g = c.planes();
  const int32_t* cnt = c.cnt;
  // tile of chunk j0: knots j0/5 .. (j0+31)/5 (at most kPlaneKnots), planes below the chunk's largest count
  auto chunk_M = [&](int j0) {
    const int j = j0 + c.lane;
    return j < items ? cnt[j / kDisc] : 0;
  };
  auto chunk_stage = [&](int j0, int Mw, int s) {
    const int k_lo = j0 / kDisc;
    int k_hi = (j0 + 31) / kDisc;
    k_hi = k_hi < K ? k_hi : K - 1;
    stage_planes(planes_g, Kc, c.lane, pbuf + s * pstride, k_lo, k_hi - k_lo + 1, Mw);
  };
  int M = chunk_M(0);
  int Mw = __reduce_max_sync(kFull, M);
  chunk_stage(0, Mw, 0);
  int stage = 0;
  // x, y of this lane's knot in the NEXT chunk travel from the context while the current chunk is processed
  auto item_knot = [&](int j0) {
    const int j = j0 + c.lane;
    return (j < items ? j : items - 1) / kDisc;
  };
  double px_n = Xs[item_knot(0)], py_n = Xs[Kc + item_knot(0)];
#pragma unroll 1
  for (int j0 = 0; j0 < items; j0 += 32, stage ^= 1) {
    int M_next = 0, Mw_next = 0;
    if (kTileBufs == 2 && j0 + 32 < items) {
      M_next = chunk_M(j0 + 32);
      Mw_next = __reduce_max_sync(kFull, M_next);
      chunk_stage(j0 + 32, Mw_next, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    const int j = j0 + c.lane;
    const bool act = j < items;
    const int jj = act ? j : items - 1;
    const int k = jj / kDisc, d = jj - k * kDisc;
    // issued now, consumed after the corridor loop: the next chunk's plane count and this item's guesses
    const int M_pre = (kTileBufs == 1 && j0 + 32 < items) ? chunk_M(j0 + 32) : 0;
    const unsigned short g2 = *reinterpret_cast<const unsigned short*>(guess + jj * 2);
    const double o = P.off[d];
    const double xd = fma(o, trig[k * 2 + 1], px_n);
    const double yd = fma(o, trig[k * 2], py_n);
    if (j0 + 32 < items) {
      const int kn = item_knot(j0 + 32);
      px_n = Xs[kn];
      py_n = Xs[Kc + kn];
    }
    // corridor half-planes of this knot
    BarAcc bc = {1.0, 0.0}, bc2 = {1.0, 0.0};  // two running products: two independent multiply chains
    const double* w = pbuf + (kTileBufs == 2 ? stage * pstride : 0) + (k - j0 / kDisc);
#pragma unroll 1  // (code size before ILP: profiles/r02_u_ab_unroll_reduction.log)
    for (int m = 0; m < Mw; m += 2) {
      if (m < M) {
        const double pa = w[m * kPlaneTile], pb = w[m * kPlaneTile + kPlaneKnots], pc = w[m * kPlaneTile + 2 * kPlaneKnots];
        bar_add(bc, fma(pb, yd, pa * xd) - pc, eps, P);
      }
      if (m + 1 < M) {
        const double* w1 = w + kPlaneTile;
        const double pa = w1[m * kPlaneTile], pb = w1[m * kPlaneTile + kPlaneKnots], pc = w1[m * kPlaneTile + 2 * kPlaneKnots];
        bar_add(bc2, fma(pb, yd, pa * xd) - pc, eps, P);
      }
    }
    bc.prod *= bc2.prod;
    bc.quad += bc2.quad;
    if (kTileBufs == 1 && j0 + 32 < items) {
      __syncwarp();  // every lane is done with the tile
      M_next = M_pre;
      Mw_next = __reduce_max_sync(kFull, M_next);
      chunk_stage(j0 + 32, Mw_next, 0);
    }
    // nearest lane segment per side (strict '<': first minimum wins)
    BarAcc bl = {1.0, 0.0};
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
      const int S = side == 0 ? S_left : S_right;
      const double* sg0 = seg + (side == 0 ? 0 : S_left) * kSegStride;
      const int bi = nearest_segment(sg0, grp + (side == 0 ? 0 : ngl) * 3, cert + (side == 0 ? 0 : S_left), S,
                                     side == 0 ? (g2 & 0xff) : (g2 >> 8), xd, yd);
      const double* sg = sg0 + bi * kSegStride;
      bar_add(bl, fma(sg[8], yd, sg[7] * xd) - sg[9], eps, P);
      if (act) nidx[jj * 2 + side] = (unsigned char)bi;
    }
    const double tc = bar_value(bc, rt), tl = bar_value(bl, rt);
    if (act) {
      sum_c += tc;
      sum_l += tl;
    }
    __syncwarp();  // the tile is dead: the next iteration's prefetch may overwrite it
    M = M_next;
    Mw = Mw_next;
  }
  const double j = warp_sum(sum_j), d = warp_sum(sum_d), co = warp_sum(sum_c), la = warp_sum(sum_l);
  cost5[0] = j + d + co + la;
  cost5[1] = j;
  cost5[2] = d;
  cost5[3] = co;
  cost5[4] = la;
}

// ------------------------------------------------------------------------------------------
// CostJacbian + CostHessian + DynamicsJacbian (ilqr_optimizer.cc:620-769, vehicle_model.cc:21-86) of one
// window of knots -> 37-double records, in two passes:
//   linearize_knot, lane == knot:          A, B, the running-cost and bound-barrier terms, sin/cos heading
//   linearize_discs, lane == (disc, knot): corridor and lane-boundary barrier terms of ONE disc per lane,
//                                          reduced over the five disc lanes of a knot by shuffles
// Xs is the iterate's slot ([8][Kc], component-major, global).
__device__ void linearize_knot(const Ctx& c, int k, const double* Xs, double* rec) {
  const KernelArgs& a = c.a;
  const DevParams& P = a.P;
  const int N = a.N, Kc = a.Kc;
  double x[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) x[i] = Xs[i * Kc + k];
  const double u0 = k < N ? Xs[6 * Kc + k] : 0.0, u1 = k < N ? Xs[7 * Kc + k] : 0.0;
  if (k < N) {
    double A11[11], b21;
    dynamics_jacobian(P, x, u1, A11, &b21);
#pragma unroll
    for (int i = 0; i < 11; ++i) rec[LA + i] = A11[i];
    rec[LB21] = b21;
  }
  // every parameter and goal is loaded HERE, in one batch (one load latency): below, each would be a generic load
  // serialised behind the record stores that precede it
  const double eps = P.eps, rt = P.rt, inv_eps2 = P.inv_eps2, dt = P.dt;
  const double vmax = P.vmax, amin = P.amin, amax = P.amax, dmin = P.dmin, dmax = P.dmax;
  const double jmin = P.jmin, jmax = P.jmax, drmin = P.drmin, drmax = P.drmax;
  const double wx2 = 2.0 * P.wx, wy2 = 2.0 * P.wy, wth2 = 2.0 * P.wth, wv2 = 2.0 * P.wv, wa2 = 2.0 * P.wa, wd2 = 2.0 * P.wd;
  const double wj2 = 2.0 * P.wj, wdr2 = 2.0 * P.wdr;
  const double gx = c.goal(k, 0), gy = c.goal(k, 1), gth = c.goal(k, 2);
  // DynamicsConsJacbian / Hessian: each state (control) component carries a pair of bounds
  auto bound_pair = [&](double lo_g, double hi_g, double& dj, double& dh) {
    double cj0, cj1, co0, co1, cd;
    bar_coef(lo_g, eps, rt, inv_eps2, cj0, co0, cd);
    bar_coef(hi_g, eps, rt, inv_eps2, cj1, co1, cd);
    dj = cj1 - cj0;
    dh = co0 + co1;
  };
  double dj, dh;
  bound_pair(0.0 - x[3], x[3] - vmax, dj, dh);
  rec[LJX + 3] = dj;
  rec[LHX + 6] = wv2 + dh;
  bound_pair(amin - x[4], x[4] - amax, dj, dh);
  rec[LJX + 4] = dj;
  rec[LHX + 7] = wa2 + dh;
  bound_pair(dmin - x[5], x[5] - dmax, dj, dh);
  rec[LJX + 5] = dj;
  rec[LHX + 8] = wd2 + dh;
  bound_pair(jmin - u0, u0 - jmax, dj, dh);
  rec[LJU + 0] = wj2 * u0 + dj;
  rec[LHU + 0] = wj2 + dh;
  bound_pair(drmin - u1, u1 - drmax, dj, dh);
  rec[LJU + 1] = wdr2 * u1 + dj;
  rec[LHU + 1] = wdr2 + dh;
  const double2 scth = nt_sincos(x[2]);
  rec[LSN] = scth.x;
  rec[LCS] = scth.y;
  rec[LZ] = 0.0;
  rec[LO] = 1.0;
  rec[LDT] = dt;
  rec[LB30] = 0.5 * dt * dt;
  rec[LJX + 0] = wx2 * (x[0] - gx);
  rec[LJX + 1] = wy2 * (x[1] - gy);
  rec[LJX + 2] = wth2 * (x[2] - gth);
  rec[LHX + 0] = wx2;
  rec[LHX + 1] = 0.0;
  rec[LHX + 2] = 0.0;
  rec[LHX + 3] = wy2;
  rec[LHX + 4] = 0.0;
  rec[LHX + 5] = wth2;
}

constexpr int kKnotsPerPass = 6;  // 6 knots x 5 discs = 30 lanes per pass of linearize_discs

// Barrier terms of the half-planes acting on disc d of knot k (CorridorConsJacbian/Hessian :690-727,
// LaneBoundaryConsJacbian/Hessian :729-769).  dx = (a, b, off*t) with t = -a sin + b cos;
// ddx(2,2) = -off*(a cos + b sin).  lane = d * 6 + (knot inside the group of six).
__device__ void linearize_discs(const Ctx& c, int k0, int nk, const double* Xs, const unsigned char* nidx,
                                double* lin) {
  const KernelArgs& a = c.a;
  const DevParams& P = a.P;
  const double* seg = c.gseg();  // only (a, b, c) of ten segments per knot: read from the context
  double* pbuf = c.smp() + a.sm.pl_b;
  double* rb = c.smp() + a.sm.red;  // [9][32]
  const int pstride = a.M_max * kPlaneTile;
  const int d = c.lane / kKnotsPerPass, kl = c.lane - d * kKnotsPerPass;
  // launch constants as values, loaded once (see bar_add)
  const int Kc = a.Kc, S_left = a.S_left;
  const double eps = P.eps, rt = P.rt, inv_eps2 = P.inv_eps2;
  const double o = P.off[d < kDisc ? d : 0];
  const double* planes_g = c.planes();
  const int32_t* cnt = c.cnt;
  auto pass_M = [&](int g0) { return (d < kDisc && g0 + kl < nk) ? cnt[k0 + g0 + kl] : 0; };
  auto pass_stage = [&](int g0, int Mw, int s) {
    const int n = nk - g0 < kKnotsPerPass ? nk - g0 : kKnotsPerPass;
    stage_planes(planes_g, Kc, c.lane, pbuf + s * pstride, k0 + g0, n, Mw);
  };
  int M = pass_M(0);
  int Mw = __reduce_max_sync(kFull, M);
  pass_stage(0, Mw, 0);
  int stage = 0;
#pragma unroll 1
  for (int g0 = 0; g0 < nk; g0 += kKnotsPerPass, stage ^= 1) {
    int M_next = 0, Mw_next = 0;
    if (kTileBufs == 2 && g0 + kKnotsPerPass < nk) {
      M_next = pass_M(g0 + kKnotsPerPass);
      Mw_next = __reduce_max_sync(kFull, M_next);
      pass_stage(g0 + kKnotsPerPass, Mw_next, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    const int M_pre = (kTileBufs == 1 && g0 + kKnotsPerPass < nk) ? pass_M(g0 + kKnotsPerPass) : 0;  // used after the plane loop
    const bool act = d < kDisc && g0 + kl < nk;
    const int ko = act ? g0 + kl : 0;  // knot inside the window
    const int k = k0 + ko;
    double* rec = lin + ko * kLinStride;
    const double sn = rec[LSN], cs = rec[LCS];
    const double xd = fma(o, cs, Xs[k]), yd = fma(o, sn, Xs[Kc + k]);
    // Per half-plane only the sums over its normal (a, b) are accumulated; the heading components follow once per
    // disc, after the loops.  With t = off (b cos - a sin) and w = off (a cos + b sin), tc = off cos, ts = off sin:
    //   J2  = sum cj t   = tc J1 - ts J0
    //   H02 = sum co a t = tc H01 - ts H00        H12 = sum co b t = tc H11 - ts H01
    //   H22 = sum co t^2 + sum cd w = (tc H12 - ts H02) + (tc D0 + ts D1),   D = sum cd (a, b)
    // -- the sums of CorridorConsJacbian / Hessian (:690-727) re-associated: 9 multiply-adds per half-plane instead of 21
    // (the LIN epoch of a CTA is bound by the FP64 pipe: all sixteen warps run this loop at the same time).
    double J0 = 0.0, J1 = 0.0, H00 = 0.0, H01 = 0.0, H11 = 0.0, D0 = 0.0, D1 = 0.0;
    auto plane = [&](double pa, double pb, double pc) {
      const double g = fma(pb, yd, pa * xd) - pc;
      double mj, co, cdd;
      bar_coef_neg(g, eps, rt, inv_eps2, mj, co, cdd);
      const double ca = pa * co, cb = pb * co;
      J0 = fma(-pa, mj, J0);
      J1 = fma(-pb, mj, J1);
      H00 = fma(pa, ca, H00);
      H01 = fma(pa, cb, H01);
      H11 = fma(pb, cb, H11);
      D0 = fma(pa, cdd, D0);
      D1 = fma(pb, cdd, D1);
    };
    const double* w = pbuf + (kTileBufs == 2 ? stage * pstride : 0) + (act ? kl : 0);
#pragma unroll 1  // (code size before ILP: profiles/r02_u_ab_unroll_reduction.log)
    for (int m = 0; m < Mw; ++m) {
      // (loaded ahead of the validity branch: rows below Mw are staged for every knot of the tile)
      const double pa = w[m * kPlaneTile], pb = w[m * kPlaneTile + kPlaneKnots], pc = w[m * kPlaneTile + 2 * kPlaneKnots];
      if (m < M) plane(pa, pb, pc);
    }
    if (kTileBufs == 1 && g0 + kKnotsPerPass < nk) {
      __syncwarp();  // every lane is done with the tile
      M_next = M_pre;
      Mw_next = __reduce_max_sync(kFull, M_next);
      pass_stage(g0 + kKnotsPerPass, Mw_next, 0);
    }
    if (act) {
#pragma unroll 1
      for (int side = 0; side < 2; ++side) {
        const double* sg = seg + ((side == 0 ? 0 : S_left) + nidx[(k * kDisc + d) * 2 + side]) * kSegStride;
        plane(sg[7], sg[8], sg[9]);
      }
    }
    const double tc = o * cs, ts = o * sn;
    const double J2 = tc * J1 - ts * J0;
    const double H02 = tc * H01 - ts * H00, H12 = tc * H11 - ts * H01;
    const double H22 = (tc * H12 - ts * H02) + (tc * D0 + ts * D1);
    // sum over the five disc lanes of each knot through shared memory, in the fixed order
    // ((d0 + d3) + (d1 + d4)) + d2
    rb[0 * 32 + c.lane] = J0;
    rb[1 * 32 + c.lane] = J1;
    rb[2 * 32 + c.lane] = J2;
    rb[3 * 32 + c.lane] = H00;
    rb[4 * 32 + c.lane] = H01;
    rb[5 * 32 + c.lane] = H02;
    rb[6 * 32 + c.lane] = H11;
    rb[7 * 32 + c.lane] = H12;
    rb[8 * 32 + c.lane] = H22;
    __syncwarp();
    if (d == 0 && act) {
#pragma unroll 1
      for (int v = 0; v < 9; ++v) {
        const double* r = rb + v * 32 + kl;
        const double sum = ((r[0] + r[3 * kKnotsPerPass]) + (r[kKnotsPerPass] + r[4 * kKnotsPerPass])) + r[2 * kKnotsPerPass];
        rec[v < 3 ? LJX + v : LHX + v - 3] += sum;
      }
    }
    __syncwarp();  // records updated; the tile is dead
    M = M_next;
    Mw = Mw_next;
  }
}

#endif  // !CILQR_STRICT

// ---- Backward (ilqr_optimizer.cc:334-390) in augmented, lane-uniform form --------------------
// With z = (x, 1) and F = [A | B] (6 x 8) one knot of the recursion is
//   M   = [Vxx ; Vx^T]                       7 x 6   value function (row 6 = gradient)
//   G   = M F                                7 x 8
//   Qh  = Hh + F^T G[0:6]  (8 x 8, sym)      = [[Qxx, Qux^T], [Qux, Quu]]
//   ql  = Jh + G[6]        (8)               = (Qx, Qu)
//   Kh  = -(Quu + lambda I)^-1 [Qux | Qu]    2 x 7   = (K, k)            (:361-366)
//   M' (i,j) = Qz(i,j) + sum_a Kh(a,i) T(a,j) + sum_a Qu_(a,i) Kh(a,j),  T = Quu Kh + [Qux | Qu]
//                                            (:379-381, un-regularised Quu, quirk Q3)
// so that every lane evaluates the same expression on operands gathered by per-lane offsets that
// are fixed before the knot loop.  delta_V (:383-384) is accumulated per lane from the UPDATED
// value function (lazy-expression quirk Q21): k.Qu' = k.Ju + y.Vx', k.Quu'.k = k.Hu.k + y.Vxx'.y,
// y = B k, and reduced once at the end of the pass.
__device__ __forceinline__ int f_off(int r, int cc) {  // offset of F[r][cc] inside a linearisation record
  if (cc < 6) {
    if (r == cc) return LO;
    if (r == 0 && cc >= 2) return LA + cc - 2;
    if (r == 1 && cc >= 2) return LA + 4 + cc - 2;
    if (r == 2 && cc >= 3) return LA + 8 + cc - 3;
    if (r == 3 && cc == 4) return LDT;
    return LZ;
  }
  if (cc == 6) return r == 3 ? LB30 : r == 4 ? LDT : LZ;
  return r == 2 ? LB21 : r == 5 ? LDT : LZ;
}
__device__ __forceinline__ int h_off(int p, int q) {  // offset of Hh[p][q], p <= q
  if (p == q) {
    if (p < 3) return LHX + (p == 0 ? 0 : p == 1 ? 3 : 5);
    if (p < 6) return LHX + 6 + p - 3;
    return LHU + p - 6;
  }
  if (q < 3) return LHX + (p == 0 ? q : 4);
  return LZ;
}

#if CILQR_STRICT
#include "cilqr_strict.cuh"
#else
constexpr int SM_ = 0;    // 42  M   [7][6]
constexpr int SG = 42;    // 56  G   [7][8]
constexpr int SQH = 98;   // 64  Qh  [8][8]
constexpr int SQL = 162;  // 8   ql
constexpr int SKH = 170;  // 14  Kh  [2][7]
static_assert(SKH + 14 <= kScratch, "scratch overflow");

// LIN phase body: linearise + quadratise every knot of the iterate in windows of kWin knots (shared
// memory), and flush the records to the context, where the Riccati sweep of the BACK phase streams
// them from.  (One phase for both was 57 KB of code -- more than the ~32 KB an SM's instruction cache
// feeds to unaligned warps -- and spilled at 128 registers.)
__device__ __noinline__ void linearize_window(const Ctx& c_ref, int k0, const double* Xs, const unsigned char* nidx,
                                              const DebugPtrs* dbg, int b) {
  const Ctx c = c_ref;  // a private copy (registers): the caller's Ctx lives in local memory and every char store below could alias it
  const KernelArgs& a = c.a;
  const int N = a.N, K = N + 1;
  const int lane = c.lane;
  double* lin = c.smp() + a.sm.lin;
  double* R = c.linrec();
  const int k = k0 + lane;
  const int nk = K - k0 < kWin ? K - k0 : kWin;
  const bool lin_lane = lane < kWin && k < K;
  __syncwarp();
  if (lin_lane) linearize_knot(c, k, Xs, lin + lane * kLinStride);
  __syncwarp();
  linearize_discs(c, k0, nk, Xs, nidx, lin);
  if (dbg) {
    if (lin_lane) {
      const double* rec = lin + lane * kLinStride;
      if (k < N) {
        if (dbg->A11) for (int i = 0; i < 12; ++i) dbg->A11[((size_t)b * N + k) * 12 + i] = rec[LA + i];
        if (dbg->Ju) for (int i = 0; i < 2; ++i) dbg->Ju[((size_t)b * N + k) * 2 + i] = rec[LJU + i];
        if (dbg->Hu) for (int i = 0; i < 2; ++i) dbg->Hu[((size_t)b * N + k) * 2 + i] = rec[LHU + i];
      }
      if (dbg->Jx) for (int i = 0; i < 6; ++i) dbg->Jx[((size_t)b * K + k) * 6 + i] = rec[LJX + i];
      if (dbg->Hx) for (int i = 0; i < 9; ++i) dbg->Hx[((size_t)b * K + k) * 9 + i] = rec[LHX + i];
    }
  }
  // window -> context, coalesced (record stride 37 in shared memory, 40 in the context)
#pragma unroll 1
  for (int idx = lane; idx < nk * kRecStride; idx += 32) {
    const int kk = idx / kRecStride, i = idx - kk * kRecStride;
    R[(size_t)(k0 + kk) * kRecStride + i] = i < kLinStride ? lin[kk * kLinStride + i] : 0.0;
  }
  __syncwarp();
}

__device__ __forceinline__ void linearize_all(const Ctx& c, const double* Xs, const unsigned char* nidx,
                                              const DebugPtrs* dbg, int b) {
  for (int k0 = 0; k0 <= c.a.N; k0 += kWin) linearize_window(c, k0, Xs, nidx, dbg, b);
}

__device__ __noinline__ void backward_pass(const Ctx& c_ref, double lambda, double dV[2]) {
  const Ctx c = c_ref;  // a private copy (registers): the caller's Ctx lives in local memory and every char store below could alias it
  const KernelArgs& a = c.a;
  const int N = a.N;
  const int lane = c.lane;
  double* scr = c.smp() + a.sm.scr;
  double* ring = c.smp() + a.sm.bring;  // two stages of kBackChunk records
  const double* R = c.linrec();
  double* gains = c.gains();  // global: K, k of every knot are consumed by the next ROLL phase
  double* M = scr + SM_;
  double* G = scr + SG;
  double* Qh = scr + SQH;
  double* ql = scr + SQL;
  double* Kh = scr + SKH;

  // ---- per-lane roles (fixed for the whole pass)
  // S1: G[i][cc], cc = lane & 7, i = lane >> 3 and 4 + (lane >> 3)
  const int s1c = lane & 7, s1i = lane >> 3;
  int fo1[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) fo1[r] = f_off(r, s1c);
  // S2 round A: pair e = lane of the 36 (p <= q); round B: lanes 0..3 pairs 32..35, lanes 4..11 ql[lane-4]
  auto pair_of = [](int e, int& p, int& q) {
    int pp = 0, rem = e;
    while (rem >= 8 - pp) {
      rem -= 8 - pp;
      ++pp;
    }
    p = pp;
    q = pp + rem;
  };
  int pA, qA, pB = 0, qB = 0;
  pair_of(lane, pA, qA);
  if (lane < 4) pair_of(32 + lane, pB, qB);
  int fo2A[6], fo2B[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    fo2A[r] = f_off(r, pA);
    fo2B[r] = f_off(r, pB);
  }
  const int hoA = h_off(pA, qA), hoB = h_off(pB, qB);
  const int qL = lane - 4;                                   // ql index on lanes 4..11
  const int joL = qL < 6 ? LJX + qL : LJU + qL - 6;          // Jh offset (valid on lanes 4..11)
  // S4: entry (vi, vj), vi <= vj <= 6, (6,6) excluded -> 27 lanes
  int vi = 0, vj = 0;
  {
    int rem = lane < 27 ? lane : 0, ii = 0;
    while (rem >= 7 - ii) {
      rem -= 7 - ii;
      ++ii;
    }
    vi = ii;
    vj = ii + rem;
  }
  // y = B k: component offset / which k (0 -> k0, 1 -> k1); rows 0,1 of B are zero
  auto yoff = [](int i) { return i == 2 ? LB21 : i == 3 ? LB30 : (i == 4 || i == 5) ? LDT : LZ; };
  const int yoi = yoff(vi), yoj = vj < 6 ? yoff(vj) : LZ;
  const int yki = (vi == 3 || vi == 4) ? 0 : 1, ykj = (vj == 3 || vj == 4) ? 0 : 1;
  const double wsym = vi == vj ? 0.5 : 1.0;  // 0.5 * (1 or 2) y_i V_ij y_j

  double acc0 = 0.0, acc1 = 0.0;
  // Vx = cost_Jx.back(), Vxx = cost_Hx.back()    (:343-344): the terminal record straight from the context
  {
    const double* rec = R + (size_t)N * kRecStride;
    for (int e = lane; e < 42; e += 32) {
      const int i = e / 6, j = e % 6;
      M[e] = i == 6 ? rec[LJX + j] : rec[h_off(i < j ? i : j, i < j ? j : i)];
    }
  }
  constexpr int kStageD = kBackChunk * kRecStride;
  auto prefetch = [&](int q, int s) {  // records of knots [4q, 4q+4) -> stage s (the context pads to whole chunks)
    const double* src = R + (size_t)q * kStageD;
    double* dst = ring + s * kStageD;
#pragma unroll 1
    for (int i = lane; i < kStageD / 2; i += 32) cp_async16(dst + i * 2, src + i * 2);
    cp_async_commit();
  };
  const int q_last = (N - 1) / kBackChunk;
  prefetch(q_last, 0);
  int stage = 0;
  for (int q = q_last; q >= 0; --q, stage ^= 1) {
    if (q > 0) {
      prefetch(q - 1, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    const int k0 = q * kBackChunk;
    int kend = k0 + kBackChunk - 1;
    if (kend > N - 1) kend = N - 1;
    const double* lin = ring + stage * kStageD;
    for (int kn = kend; kn >= k0; --kn) {
      const double* rec = lin + (kn - k0) * kRecStride;
      // ---- S1: G = M F
      {
        double f[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) f[r] = rec[fo1[r]];
        const double* m1 = M + s1i * 6;
        double g1 = m1[0] * f[0];
#pragma unroll
        for (int r = 1; r < 6; ++r) g1 = fma(m1[r], f[r], g1);
        G[s1i * 8 + s1c] = g1;
        if (lane < 24) {
          const double* m2 = M + (4 + s1i) * 6;
          double g2 = m2[0] * f[0];
#pragma unroll
          for (int r = 1; r < 6; ++r) g2 = fma(m2[r], f[r], g2);
          G[(4 + s1i) * 8 + s1c] = g2;
        }
      }
      __syncwarp();
      // ---- S2: Qh = Hh + F^T G[0:6], ql = Jh + G[6]
      {
        double q = rec[hoA];
#pragma unroll
        for (int r = 0; r < 6; ++r) q = fma(rec[fo2A[r]], G[r * 8 + qA], q);
        Qh[pA * 8 + qA] = q;
        Qh[qA * 8 + pA] = q;
        if (lane < 4) {
          double q2 = rec[hoB];
#pragma unroll
          for (int r = 0; r < 6; ++r) q2 = fma(rec[fo2B[r]], G[r * 8 + qB], q2);
          Qh[pB * 8 + qB] = q2;
          Qh[qB * 8 + pB] = q2;
        } else if (lane < 12) {
          ql[qL] = rec[joL] + G[6 * 8 + qL];
        }
      }
      __syncwarp();
      // ---- S3: Kh = -(Quu + lambda I)^-1 [Qux | Qu]   (closed-form 2x2 inverse, :361-366)
      const double q00 = Qh[6 * 8 + 6], q01 = Qh[6 * 8 + 7], q11 = Qh[7 * 8 + 7];
      if (lane < 14) {
        const double t00 = q00 + lambda, t11 = q11 + lambda;
        const double det = t00 * t11 - q01 * q01;
        const double invdet = 1.0 / det;
        const int rr = lane / 7, j = lane - rr * 7;
        const double n0 = rr == 0 ? -(t11 * invdet) : q01 * invdet;   // -inv[rr][0]
        const double n1 = rr == 0 ? q01 * invdet : -(t00 * invdet);   // -inv[rr][1]
        const double b0 = j < 6 ? Qh[j * 8 + 6] : ql[6];
        const double b1 = j < 6 ? Qh[j * 8 + 7] : ql[7];
        const double kv = fma(n1, b1, n0 * b0);
        Kh[lane] = kv;
        gains[kn * kGainStride + (j < 6 ? rr * 6 + j : 12 + rr)] = kv;
      }
      __syncwarp();
      // ---- S4: M' and the delta_V contributions
      if (lane < 27) {
        const double ki0 = Kh[vi], ki1 = Kh[7 + vi], kj0 = Kh[vj], kj1 = Kh[7 + vj];
        const double ui0 = Qh[vi * 8 + 6], ui1 = Qh[vi * 8 + 7];             // Qux[a][vi]
        const double uj0 = vj < 6 ? Qh[vj * 8 + 6] : ql[6];                  // Qux[a][vj] or Qu[a]
        const double uj1 = vj < 6 ? Qh[vj * 8 + 7] : ql[7];
        const double t0 = fma(q01, kj1, fma(q00, kj0, uj0));                 // T[0][vj]
        const double t1 = fma(q11, kj1, fma(q01, kj0, uj1));                 // T[1][vj]
        double v = vj < 6 ? Qh[vi * 8 + vj] : ql[vi];
        v += fma(ki1, t1, ki0 * t0);
        v += fma(ui1, kj1, ui0 * kj0);
        const double k0v = Kh[6], k1v = Kh[13];
        const double yi = (yki == 0 ? k0v : k1v) * rec[yoi];
        if (vj < 6) {
          M[vi * 6 + vj] = v;
          M[vj * 6 + vi] = v;
          const double yj = (ykj == 0 ? k0v : k1v) * rec[yoj];
          acc1 = fma(wsym * yi, v * yj, acc1);
        } else {
          M[36 + vi] = v;
          acc0 = fma(yi, v, acc0);
        }
      } else if (lane == 27) {
        const double k0v = Kh[6], k1v = Kh[13];
        acc0 += fma(k1v, rec[LJU + 1], k0v * rec[LJU]);
        acc1 += 0.5 * fma(k1v * k1v, rec[LHU + 1], (k0v * k0v) * rec[LHU]);
      }
      __syncwarp();
    }
  }
  dV[0] = warp_sum(acc0);
  dV[1] = warp_sum(acc1);
}

#endif  // CILQR_STRICT

// iqr, ilqr_optimizer.cc:793-824: time-varying LQR about the goals (zero control), Q = diag(1e-3, 1e-3,
// 1e-3, 1e-3, 1e-2, 5e-3), R = diag(0.2, 0.05) (off-diagonals 0, quirk Q4).  It is the Riccati recursion of
// Backward with Jx = Ju = 0, Hx = Q, Hu = R, lambda = 0 -- K_k = (R + B'PB)^-1 B'PA, P = Q + A'P(A - BK) is
// algebraically the same value update -- so INIT only writes the "linearisation records" of that LQ problem
// (A_k, B_k about goal_k with zero control) and the BACK phase does the sweep; its gains -(Quu)^-1 Qux are
// the -K_lqr the rollout adds, k = 0 because the gradient stays zero.  (xbar, ubar) = (goals, 0) go to slot Xs.
__device__ void iqr_records(const Ctx& c, double* Xs) {
  const KernelArgs& a = c.a;
  const DevParams& P = a.P;
  const int N = a.N, Kc = a.Kc, lane = c.lane;
  const double dt = P.dt;
  double* R = c.linrec();
  const double Qd[6] = {0.001, 0.001, 0.001, 0.001, 0.01, 0.005};
  for (int k = lane; k <= N; k += 32) {
    double* rec = R + (size_t)k * kRecStride;
    double g[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) g[i] = c.goal(k, i);
    if (k < N) {
      double A11[11], b21;
      dynamics_jacobian(P, g, 0.0, A11, &b21);
#pragma unroll
      for (int i = 0; i < 11; ++i) rec[LA + i] = A11[i];
      rec[LB21] = b21;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) rec[LJX + i] = 0.0;
    rec[LJU] = 0.0;
    rec[LJU + 1] = 0.0;
    rec[LHX + 0] = Qd[0];
    rec[LHX + 1] = 0.0;
    rec[LHX + 2] = 0.0;
    rec[LHX + 3] = Qd[1];
    rec[LHX + 4] = 0.0;
    rec[LHX + 5] = Qd[2];
    rec[LHX + 6] = Qd[3];
    rec[LHX + 7] = Qd[4];
    rec[LHX + 8] = Qd[5];
    rec[LHU] = 0.2;
    rec[LHU + 1] = 0.05;
#if CILQR_STRICT
    rec[LH10] = 0.0;
    rec[LH10 + 1] = 0.0;
    rec[LH10 + 2] = 0.0;
#endif
    rec[LZ] = 0.0;
    rec[LO] = 1.0;
    rec[LDT] = dt;
    rec[LB30] = 0.5 * dt * dt;
    // (xbar, ubar) = (goals, 0)
#pragma unroll
    for (int i = 0; i < 6; ++i) Xs[i * Kc + k] = g[i];
    Xs[6 * Kc + k] = 0.0;
    Xs[7 * Kc + k] = 0.0;
  }
  __syncwarp();
}

__device__ __forceinline__ unsigned fnv1a(unsigned h, unsigned byte) { return (h ^ (byte & 0xffu)) * 16777619u; }

// trajectory slot ([8][Kc], component-major) -> states [K][6] and controls [N][2] (knot-major)
__device__ __noinline__ void copy_traj(const Ctx& c_ref, const double* Xs, double* states, double* controls) {
  const Ctx c = c_ref;  // a private copy (registers): the caller's Ctx lives in local memory and every char store below could alias it
  const int N = c.a.N, Kc = c.a.Kc;
#pragma unroll 1
  for (int k = c.lane; k <= N; k += 32) {
    if (states) {
#pragma unroll
      for (int i = 0; i < 6; ++i) states[k * 6 + i] = Xs[i * Kc + k];
    }
    if (controls && k < N) {
      controls[k * 2] = Xs[6 * Kc + k];
      controls[k * 2 + 1] = Xs[7 * Kc + k];
    }
  }
}

// lane segments + group circles of the context -> this warp's shared-memory stage
__device__ __noinline__ void stage_segments(Ctx& c) {
  const KernelArgs& a = c.a;
  if (c.seg_staged) return;
  c.seg_staged = true;
  const int n16 = ((a.S_left + a.S_right) * kSegStride + 1) / 2;
  const double* src = c.gseg();
  double* dst = c.smp() + a.sm.seg;
  for (int i = c.lane; i < n16; i += 32) cp_async16(dst + i * 2, src + i * 2);
  const int ng = (a.S_left + kGroup - 1) / kGroup + (a.S_right + kGroup - 1) / kGroup;
  const int g16 = (ng * 3 + a.S_left + a.S_right + 1) / 2;  // group circles + certificate radii
  const double* gs = c.ggrp();
  double* gd = c.smp() + a.sm.grp;
  for (int i = c.lane; i < g16; i += 32) cp_async16(gd + i * 2, gs + i * 2);
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();
}

// ---- exits of Optimize (ilqr_optimizer.cc:225,238,285,303,319): results of the context -> outputs
__device__ __noinline__ void finish_scenario(const Ctx& c) {
  const KernelArgs& a = c.a;
  const DevParams& P = a.P;
  const int N = a.N, K = N + 1, lane = c.lane;
  const CtxHdr* h = c.h;
  const size_t b = h->b;
  const double* Xs = c.slot(h->cur);
  copy_traj(c, Xs, a.states + b * K * 6, a.controls + b * N * 2);
  if (lane == 0 && a.stats) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    const unsigned long long t0 = *(volatile unsigned long long*)(a.stats + 8);
    unsigned long long bk = (now - t0) / 2000000ull;
    atomicAdd(a.stats + 10 + (bk < 255 ? bk : 255), 1ull);
  }
  if (lane == 0) {
    atomicAdd(a.ticket + 2, 1u);
    double* st = a.status + b * 8;
    st[0] = h->status;
    st[1] = h->iter;
    for (int i = 0; i < 5; ++i) st[2 + i] = h->cost_acc[i];
    st[7] = (double)h->ahash;
    if (a.hist_len) {
      a.hist_len[b * 2] = h->n_cost;
      a.hist_len[b * 2 + 1] = h->n_iter_traj;
    }
  }
  if (a.result) {
    // TrajectoryPlanner::Plan's post-processing of opt_trajectory (trajectory_planner.cpp:103-125): the same
    // record with the station accumulated point by point (sequentially, like the reference's running sum)
#pragma unroll 1
    for (int k = lane; k < K; k += 32) {
      double* tp = a.result + (b * K + k) * 13;
      const double de = Xs[5 * a.Kc + k];
      tp[0] = P.dt * k;
      tp[1] = k > 0 ? nt_hypot(Xs[k] - Xs[k - 1], Xs[a.Kc + k] - Xs[a.Kc + k - 1]) : 0.0;  // segment length for now
      tp[2] = Xs[k];
      tp[3] = Xs[a.Kc + k];
      tp[4] = Xs[2 * a.Kc + k];
      tp[5] = nt_tan(de) / P.L;
      tp[6] = Xs[3 * a.Kc + k];
      tp[7] = Xs[4 * a.Kc + k];
      tp[8] = k < N ? Xs[6 * a.Kc + k] : 0.0;
      tp[9] = de;
      tp[10] = k < N ? Xs[7 * a.Kc + k] : 0.0;
      tp[11] = 0.0;
      tp[12] = 0.0;
    }
    __syncwarp();
    if (lane == 0) {
      double acc = 0.0;
      double* tp = a.result + b * K * 13;
      for (int k = 0; k < K; ++k) {
        acc += k > 0 ? tp[k * 13 + 1] : 0.0;
        tp[k * 13 + 1] = acc;
      }
    }
    __syncwarp();
  }
  if (a.trajectory) {
    // TransformToTrajectory (:771-791): time, s, x, y, theta, kappa, velocity, a, jerk, delta, delta_rate, lb, rb
#pragma unroll 1
    for (int k = lane; k < K; k += 32) {
      double* tp = a.trajectory + (b * K + k) * 13;
      const double de = Xs[5 * a.Kc + k];
      tp[0] = k * P.dt;
      tp[1] = 0.0;
      tp[2] = Xs[k];
      tp[3] = Xs[a.Kc + k];
      tp[4] = Xs[2 * a.Kc + k];
      tp[5] = nt_tan(de) / P.L;
      tp[6] = Xs[3 * a.Kc + k];
      tp[7] = Xs[4 * a.Kc + k];
      tp[8] = k < N ? Xs[6 * a.Kc + k] : 0.0;
      tp[9] = de;
      tp[10] = k < N ? Xs[7 * a.Kc + k] : 0.0;
      tp[11] = 0.0;
      tp[12] = 0.0;
    }
  }
}

// iter_trajs.push_back (ilqr_optimizer.cc:170,294) and cost_.push_back (:173,283,296) when requested
__device__ __noinline__ void push_traj(const Ctx& c, const double* Xs) {
  const KernelArgs& a = c.a;
  CtxHdr* h = c.h;
  const int n = h->n_iter_traj;
  if (a.iter_states && n < a.hist_cap) {
    const size_t o = (size_t)h->b * a.hist_cap + n;
    copy_traj(c, Xs, a.iter_states + o * (a.N + 1) * 6, a.iter_controls ? a.iter_controls + o * a.N * 2 : nullptr);
  }
  __syncwarp();
  if (c.lane == 0) h->n_iter_traj = n + 1;
}
__device__ __noinline__ void push_cost(const Ctx& c, const double cost5[5]) {
  const KernelArgs& a = c.a;
  CtxHdr* h = c.h;
  const int n = h->n_cost;
  __syncwarp();
  if (c.lane == 0) {
    if (a.cost_hist && n < a.hist_cap)
      for (int i = 0; i < 5; ++i) a.cost_hist[((size_t)h->b * a.hist_cap + n) * 5 + i] = cost5[i];
    for (int i = 0; i < 5; ++i) h->cost_acc[i] = cost5[i];
    h->n_cost = n + 1;
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// INIT: next scenario of the batch -> context.  TransformGoals (:141-152), ShrinkConstraints +
// NormalizeHalfPlane (:438-495), LineSegment2d construction (line_segment2d.cpp:40-49), group circles,
// LQR gains of the initial guess (:793-824).
__device__ __noinline__ int phase_init(Ctx& c) {
  const KernelArgs& a = c.a;
  const DevParams& P = a.P;
  const int N = a.N, K = N + 1, lane = c.lane;
  const DebugPtrs* dbg = a.debug ? &a.dbg : nullptr;
  unsigned int b = 0;
  if (lane == 0) b = atomicAdd(a.ticket, 1u);
  b = __shfl_sync(kFull, b, 0);
  if (b >= (unsigned)a.B) return PH_DONE;
  if (a.ready) {
    // the host path streams the inputs in behind the running kernel: wait for this scenario's chunk
    // Watchdog: the transfer died or was never enqueued.  The scenario stays unsolved (its status row keeps
    // the sentinel the host wrote before the launch) and the launch is flagged: cilqr_plan_batch returns
    // CILQR_E_TIMEOUT instead of handing back a trajectory nobody computed (the reference never returns
    // an un-filled trajectory as success, trajectory_planner.cpp:91-94).
    const unsigned long long t_wait = global_ns();
    while (b >= *(const volatile unsigned int*)a.ready) {
      __nanosleep(2000);
      if (global_ns() - t_wait > a.watchdog_ns) {
        if (lane == 0) raise_error(a, kErrStarved);
        return PH_DONE;
      }
    }
    __threadfence();
  }
  c.bind(b);
  CtxHdr* h = c.h;
  if (lane == 0) {
    const double* st = a.start + (size_t)b * 4;
    h->g0[0] = st[0];
    h->g0[1] = st[1];
    h->g0[2] = st[2];
    h->g0[3] = st[3];
    h->g0[4] = 0.0;
    h->g0[5] = 0.0;
    h->b = b;
    h->ahash = 2166136261u;
    h->lambda = 1.0;
    h->dlambda = 1.0;
    h->iter = 0;
    h->status = 4;
    h->cur = 0;
    h->nflip = 0;
    h->n_cost = 0;
    h->n_iter_traj = 0;
    h->rmode = 0;
    h->emode = 0;
    h->imode = 1;
    h->retired = 0;
    h->deferred = 0;
  }
  __syncwarp();
  {
    // launch constants as values, loaded once (see bar_add)
    const int M_max = a.M_max, Kc = a.Kc, S_left = a.S_left, S_right = a.S_right;
    const double shrink_corr = P.shrink_corr, shrink_lane = P.shrink_lane;
    const int32_t* cnt = c.cnt;
    double* dbg_corridor = dbg ? dbg->corridor : nullptr;
    double* dbg_lanes = dbg ? dbg->lanes : nullptr;
    const double* raw = a.corridor + (size_t)b * K * M_max * 3;
    double* planes = c.planes();
    const int total = K * M_max;
#pragma unroll 1
    for (int idx = lane; idx < total; idx += 32) {
      const int k = idx / M_max, m = idx - k * M_max;
      if (m < cnt[k]) {
        const double e0 = raw[idx * 3], e1 = raw[idx * 3 + 1];
        double e2 = raw[idx * 3 + 2];
        e2 = e2 - shrink_corr * (e0 * e0 + e1 * e1) / nt_hypot(e0, e1);
        const double nrm = nt_hypot(nt_hypot(e0, e1), e2);
        planes[(m * 3 + 0) * Kc + k] = e0 / nrm;
        planes[(m * 3 + 1) * Kc + k] = e1 / nrm;
        planes[(m * 3 + 2) * Kc + k] = e2 / nrm;
        if (dbg_corridor) {
          double* o = dbg_corridor + ((size_t)b * total + idx) * 3;
          o[0] = e0 / nrm;
          o[1] = e1 / nrm;
          o[2] = e2 / nrm;
        }
      }
    }
    const int ST_ = S_left + S_right;
    double* seg = c.smp() + a.sm.seg;
    double* gsg = c.gseg();
    const double* lane_left = a.lane_left + (size_t)b * S_left * 7;
    const double* lane_right = a.lane_right + (size_t)b * S_right * 7;
#pragma unroll 1
    for (int s = lane; s < ST_; s += 32) {
      const double* ln = s < S_left ? lane_left + s * 7 : lane_right + (s - S_left) * 7;
      const double e0 = ln[0], e1 = ln[1];
      double e2 = ln[2];
      e2 = e2 - shrink_lane * (e0 * e0 + e1 * e1) / nt_hypot(e0, e1);
      const double nrm = nt_hypot(nt_hypot(e0, e1), e2);
      const double dx = ln[5] - ln[3], dy = ln[6] - ln[4];
      const double len = nt_hypot(dx, dy);
      double r[kSegStride];
      r[0] = ln[3];
      r[1] = ln[4];
      r[2] = ln[5];
      r[3] = ln[6];
      r[4] = len <= 1e-10 ? 0.0 : dx / len;  // line_segment2d.cpp:40-49
      r[5] = len <= 1e-10 ? 0.0 : dy / len;
      r[6] = len;
      r[7] = e0 / nrm;
      r[8] = e1 / nrm;
      r[9] = e2 / nrm;
#pragma unroll
      for (int q = 0; q < kSegStride; ++q) {
        seg[s * kSegStride + q] = r[q];
        gsg[s * kSegStride + q] = r[q];
      }
      if (dbg_lanes) {
        double* o = dbg_lanes + ((size_t)b * ST_ + s) * 3;
        o[0] = r[7];
        o[1] = r[8];
        o[2] = r[9];
      }
    }
    __syncwarp();
    // bounding circles of the segment groups (see nearest_segment)
    const int ngl = (S_left + kGroup - 1) / kGroup, ngr = (S_right + kGroup - 1) / kGroup;
    double* grp = c.ggrp();
#pragma unroll 1
    for (int g = lane; g < ngl + ngr; g += 32) {
      const int side = g < ngl ? 0 : 1;
      const int S = side == 0 ? S_left : S_right;
      const int s_lo = (side == 0 ? g : g - ngl) * kGroup;
      const int s_hi = s_lo + kGroup < S ? s_lo + kGroup : S;
      const double* sg0 = seg + (side == 0 ? 0 : S_left) * kSegStride;
      double xmin = 1.7976931348623157e308, xmax = -xmin, ymin = xmin, ymax = -xmin;
#pragma unroll 1
      for (int s2 = s_lo; s2 < s_hi; ++s2) {
        const double* sg = sg0 + s2 * kSegStride;
        xmin = fmin(xmin, fmin(sg[0], sg[2]));
        xmax = fmax(xmax, fmax(sg[0], sg[2]));
        ymin = fmin(ymin, fmin(sg[1], sg[3]));
        ymax = fmax(ymax, fmax(sg[1], sg[3]));
      }
      const double cx = 0.5 * (xmin + xmax), cy = 0.5 * (ymin + ymax);
      double r2 = 0.0;
#pragma unroll 1
      for (int s2 = s_lo; s2 < s_hi; ++s2) {
        const double* sg = sg0 + s2 * kSegStride;
        const double ax = sg[0] - cx, ay = sg[1] - cy, bx = sg[2] - cx, by = sg[3] - cy;
        r2 = fmax(r2, fmax(ax * ax + ay * ay, bx * bx + by * by));
      }
      grp[g * 3] = cx;
      grp[g * 3 + 1] = cy;
      // radius, inflated so that rounding can only make the pruning test more conservative
      grp[g * 3 + 2] = sqrt(r2) * (1.0 + 1e-9) + 1e-9;
    }
    // certificate radii of the nearest-segment fast path (see nearest_segment): for segment s, half of
    // min over |j - s| > kNearWin of (|mid_s - mid_j| - (len_s + len_j) / 2) -- a lower bound on the distance
    // between the two segments --, deflated, squared
    double* cert = grp + (ngl + ngr) * 3;
#pragma unroll 1
    for (int s = lane; s < ST_; s += 32) {
      const int side = s < S_left ? 0 : 1;
      const int s0 = side == 0 ? 0 : S_left, S = side == 0 ? S_left : S_right;
      const double* me = seg + s * kSegStride;
      const double mx = 0.5 * (me[0] + me[2]), my = 0.5 * (me[1] + me[3]);
      double D = 1.7976931348623157e308;
#pragma unroll 1
      for (int j = 0; j < S; ++j) {
        const int dj = j - (s - s0);
        if (dj >= -kNearWin && dj <= kNearWin) continue;
        const double* o = seg + (s0 + j) * kSegStride;
        const double ex = mx - 0.5 * (o[0] + o[2]), ey = my - 0.5 * (o[1] + o[3]);
        D = fmin(D, sqrt(ex * ex + ey * ey) - 0.5 * (me[6] + o[6]));
      }
      const double h = 0.5 * D * (1.0 - 1e-9) - 1e-9;
      cert[s] = h > 0.0 ? (h < 1e150 ? h * h * (1.0 - 1e-9) : 1.7976931348623157e308) : 0.0;
    }
  }
  __syncwarp();
  // no previous iterate: any valid index is an upper bound for the nearest-segment search
  unsigned char* n0 = c.nidx(0);
#pragma unroll 1
  for (int i = lane; i < a.cl.nidx_bytes; i += 32) n0[i] = 0;
  if (a.init_mode != 0) {
    // the caller's initial guess instead of iqr (ilqr_optimizer.cc:168): slot 0 <- (X, U) as given (mode 2: the
    // InitGuess copy, :107-139), or <- (0, U) with zero gains so that the rollout phase reproduces
    // x_{k+1} = Dynamics(x_k, u_k) from goals_[0] (mode 1: OpenLoopRollout, slover/ilqr.h:362-370)
    double* Xs = c.slot(0);
    const double* gu = a.guess_controls + (size_t)b * N * 2;
    const double* gx = a.init_mode == 2 ? a.guess_states + (size_t)b * K * 6 : nullptr;
    double* gains = c.gains();
#pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
#pragma unroll
      for (int i = 0; i < 6; ++i) Xs[i * a.Kc + k] = gx ? gx[k * 6 + i] : 0.0;
      Xs[6 * a.Kc + k] = k < N ? gu[k * 2] : 0.0;
      Xs[7 * a.Kc + k] = k < N ? gu[k * 2 + 1] : 0.0;
      if (k < N && !gx)
        for (int i = 0; i < kGainStride; ++i) gains[k * kGainStride + i] = 0.0;
    }
    if (gx && (a.init_states || a.init_controls)) {  // iter_trajs[0] for the adapter (:170)
#pragma unroll 1
      for (int k = lane; k <= N; k += 32) {
        if (a.init_states)
          for (int i = 0; i < 6; ++i) a.init_states[((size_t)b * K + k) * 6 + i] = gx[k * 6 + i];
        if (a.init_controls && k < N) {
          a.init_controls[((size_t)b * N + k) * 2] = gu[k * 2];
          a.init_controls[((size_t)b * N + k) * 2 + 1] = gu[k * 2 + 1];
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      h->imode = 0;
      h->emode = 0;
      h->rmode = gx ? 0 : 3;
    }
    __syncwarp();
    return gx ? PH_EVAL : PH_ROLL;
  }
  iqr_records(c, c.slot(0));
  __syncwarp();
  return PH_BACK;  // the iqr sweep
}

// ROLL: rollout of the closed loop u_k = ubar_k + K_k (x - xbar_k) + alpha k_k (Forward,
// ilqr_optimizer.cc:392-415) for up to EIGHT contexts per warp, four lanes each: lane a of a group rolls
// out step size alpha_{gb+a} of its context's current line-search group and writes the candidate to
// trajectory slot cand_slot(cur, a) (lanes without a wanted step size shadow the highest wanted lane of
// their group and store nothing).  The rollout is a serial chain in which only the step sizes are
// parallel, so one context alone would leave 28 lanes idle.  Every group streams its nominal trajectory
// (slot `cur`) and gains from its context through its own two-stage cp.async ring of kRollChunk knots;
// the serial loop only reads shared memory.  The mode comes from the context header:
//   rmode 0  the LQR initial guess (iqr, :830-841): one candidate, controls clamped to their bounds
//            instead of angle-wrapped
//   rmode 1  speculative rollout of a group of four step sizes (:246-252); a lane that would need the
//            general (fmod) branch of NormalizeAngle -- a blown-up rollout -- is DEFERRED instead of
//            dragging the warp through that branch at every step
//   rmode 2  faithful repeat of the deferred step sizes, requested by EVAL only if the search reaches one
// A lane whose state stops being finite is RETIRED: every later state of that rollout would be non-finite
// too, its cost NaN/inf, and the reference rejects such a step (the comparisons of :258 are false).
constexpr int kRollGroups = 8;
__device__ __forceinline__ double* ctx_base(const KernelArgs& a, const int* gidx, int id) {
  return a.ws + (size_t)gidx[id] * a.cl.stride;  // gidx: the CTA's slot -> global context index table (shared memory)
}
__device__ __noinline__ void roll_multi(const KernelArgs& a, int sm_off, const int* gidx, int my_id, int lane) {
  double* sm = smem_cta + sm_off;
  const DevParams& P = a.P;
  const int g = lane >> 2, ai = lane & 3;
  const bool valid = my_id >= 0;
  const int id0 = __shfl_sync(kFull, my_id, 0);  // group 0 is always valid: benign operands for empty groups
  double* cx = ctx_base(a, gidx, valid ? my_id : id0);
  CtxHdr* h = reinterpret_cast<CtxHdr*>(cx + a.cl.hdr);
  const int cur = h->cur, rmode = h->rmode;
  const int gb = (rmode == 0 || rmode == 3) ? 0 : h->gb;
  unsigned want = 1u;
  if (rmode == 1) want = (1u << (kNAlpha - gb < kSpec ? kNAlpha - gb : kSpec)) - 1u;
  else if (rmode == 2) want = (h->deferred >> gb) & ((1u << kSpec) - 1u);
  if (want == 0) want = 1u;  // (cannot happen; keeps the shadow index valid)
  // a scenario far beyond the mean iteration count usually blows up its largest step sizes: take the
  // general wrap at once instead of deferring it to a second pass (it runs alone in this warp anyway)
  // rmode 3: open-loop rollout of the caller's controls (no clamp, no wrap: OpenLoopRollout, slover/ilqr.h:362-370)
  const bool iqr = rmode == 0 || rmode == 3, clamp_u = rmode == 0, allow_general = rmode != 1 || h->iter >= a.hot_iter;
  const bool wanted = (want >> ai) & 1u;
  const bool owner = valid && wanted;
  const int ca = wanted ? ai : 31 - __clz(want);
  const double alpha = kAlphaList[gb + ca];
  // launch constants as values, loaded once (see bar_add): below they would be re-read after every store to `out`
  const int N = a.N, Kc = a.Kc;
  const double dt = P.dt, lq = CILQR_LQ(P), jmin = P.jmin, jmax = P.jmax, drmin = P.drmin, drmax = P.drmax;
  const double* Xs = cx + a.cl.slots + cur * 8 * Kc;
  const double* gains = cx + a.cl.gains;
  double* out = cx + a.cl.slots + cand_slot(cur, ca) * 8 * Kc;
  double* ring = sm + g * kRingDoubles;
  constexpr int kStage = 8 * kRollChunk + kGainStride * kRollChunk;  // doubles per ring stage
  double x[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) x[i] = h->g0[i];
  if (owner) {
#pragma unroll
    for (int i = 0; i < 6; ++i) out[i * Kc] = x[i];
  }
  // stage s of a group's ring: [8][kRollChunk] nominal x0..x5,u0,u1 then [kRollChunk][16] gain records;
  // 16 + 32 sixteen-byte pieces per chunk, twelve per lane of the group
  auto prefetch = [&](int k0, int s) {
    double* st = ring + s * kStage;
#pragma unroll
    for (int p = ai; p < 4 * kRollChunk + kGainStride * kRollChunk / 2; p += 4) {
      if (p < 4 * kRollChunk) {
        const int comp = p / (kRollChunk / 2), part = p - comp * (kRollChunk / 2);
        cp_async16(st + comp * kRollChunk + part * 2, Xs + comp * Kc + k0 + part * 2);
      } else {
        const int q = p - 4 * kRollChunk;
        cp_async16(st + 8 * kRollChunk + q * 2, gains + (size_t)k0 * kGainStride + q * 2);
      }
    }
    cp_async_commit();
  };
  bool dead = false, defer = false, slow = false;
  prefetch(0, 0);
  int stage = 0;
  for (int k0 = 0; k0 < N; k0 += kRollChunk, stage ^= 1) {
    if (k0 + kRollChunk < N) {
      prefetch(k0 + kRollChunk, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    const double* st = ring + stage * kStage;
    const int kn = N - k0 < kRollChunk ? N - k0 : kRollChunk;
    for (int kk = 0; kk < kn; ++kk) {
      const int k = k0 + kk;
      const double* Kk = st + 8 * kRollChunk + kk * kGainStride;
      double dx[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) dx[i] = x[i] - st[i * kRollChunk + kk];
      double s0 = Kk[0] * dx[0], s1 = Kk[6] * dx[0];
#pragma unroll
      for (int i = 1; i < 6; ++i) {
        s0 = CILQR_FMA(Kk[i], dx[i], s0);
        s1 = CILQR_FMA(Kk[6 + i], dx[i], s1);
      }
      double u0 = st[6 * kRollChunk + kk] + s0 + alpha * Kk[12];
      double u1 = st[7 * kRollChunk + kk] + s1 + alpha * Kk[13];
      if (clamp_u) {
        u0 = fmin(jmax, fmax(u0, jmin));     // clamp, ilqr_optimizer.cc:826-836
        u1 = fmin(drmax, fmax(u1, drmin));
      } else if (iqr) {
        // open loop: the caller's controls as they are
      } else {
        u1 = wrap_angle(u1, allow_general, slow);  // ilqr_optimizer.cc:408
      }
      rollout_step(dt, lq, x, u0, u1, allow_general, slow);
      if (!iqr && !dead && !defer) {
        const double chk = ((x[0] + x[1]) + (x[2] + x[3])) + (x[4] + x[5]);
        if (slow) defer = true;  // this step was not computed faithfully: the whole rollout is repeated later
        else if (!(fabs(chk) <= 1.7976931348623157e308)) dead = true;  // NaN or inf somewhere in x
      }
      if (dead || defer) {
        // park the lane on the nominal trajectory: benign operands for the remaining steps
#pragma unroll
        for (int i = 0; i < 6; ++i) x[i] = Xs[i * Kc + k + 1];
        slow = false;
      } else if (owner) {
        out[6 * Kc + k] = u0;
        out[7 * Kc + k] = u1;
#pragma unroll
        for (int i = 0; i < 6; ++i) out[i * Kc + k + 1] = x[i];
      }
    }
    __syncwarp();
  }
  const unsigned db = __ballot_sync(kFull, dead), fb = __ballot_sync(kFull, defer);
  const unsigned ret = (db >> (4 * g)) & want, def = (fb >> (4 * g)) & want;
  __syncwarp();
  if (valid && ai == 0) {
    if (rmode == 0 || rmode == 3) {
      h->cur = cand_slot(cur, 0);
      h->emode = 0;
    } else if (rmode == 1) {
      h->retired |= ret << gb;
      h->deferred |= def << gb;
      h->ai = gb;
      h->emode = 1;
    } else {
      h->retired |= ret << gb;
      h->deferred &= ~(((1u << kSpec) - 1u) << gb);
    }
  }
  if (a.init_states || a.init_controls) {
    // iter_trajs[0] for the adapter (:170): copy the initial guesses out, one context at a time
    __syncwarp();
    for (int gg = 0; gg < kRollGroups; ++gg) {
      const int id = __shfl_sync(kFull, my_id, gg * 4), rm = __shfl_sync(kFull, rmode, gg * 4);
      if (id < 0 || (rm != 0 && rm != 3)) continue;
      double* cg = ctx_base(a, gidx, id);
      const CtxHdr* hg = reinterpret_cast<const CtxHdr*>(cg + a.cl.hdr);
      const int curg = __shfl_sync(kFull, cur, gg * 4);
      const double* src = cg + a.cl.slots + cand_slot(curg, 0) * 8 * a.Kc;
      const size_t bb = hg->b;
      for (int k = lane; k <= a.N; k += 32) {
        if (a.init_states)
          for (int i = 0; i < 6; ++i) a.init_states[(bb * (a.N + 1) + k) * 6 + i] = src[i * a.Kc + k];
        if (a.init_controls && k < a.N) {
          a.init_controls[(bb * a.N + k) * 2] = src[6 * a.Kc + k];
          a.init_controls[(bb * a.N + k) * 2 + 1] = src[7 * a.Kc + k];
        }
      }
    }
  }
}

// LIN: linearise + quadratise the iterate (:203-214) -> records in the context.
__device__ __noinline__ int phase_lin(Ctx& c) {
  const KernelArgs& a = c.a;
  CtxHdr* h = c.h;
  const unsigned b = h->b;
  c.bind(b);
  const DebugPtrs* dbg = (a.debug && h->iter == 0) ? &a.dbg : nullptr;
  linearize_all(c, c.slot(h->cur), c.nidx(h->cur), dbg, (int)b);
  return PH_BACK;
}

// BACK: Riccati sweep (:216-233), gradient-norm exit (:235-241).
__device__ __noinline__ int phase_back(Ctx& c) {
  const KernelArgs& a = c.a;
  const int N = a.N, lane = c.lane;
  CtxHdr* h = c.h;
  const unsigned b = h->b;
  c.bind(b);
  const DebugPtrs* dbg = (a.debug && h->iter == 0) ? &a.dbg : nullptr;
  const bool iqr = h->imode != 0;
  const double lambda = iqr ? 0.0 : h->lambda;
  const double* Xs = c.slot(h->cur);
  double dV[2];
#if CILQR_STRICT
  if (iqr) s_iqr_sweep(c);
  else backward_pass(c, lambda, dV);
#else
  backward_pass(c, lambda, dV);
#endif
  __syncwarp();
  if (iqr) {  // gains of the LQR initial guess (:793-824) are in place: roll it out (:830-841)
    if (lane == 0) {
      h->imode = 0;
      h->rmode = 0;
    }
    return PH_ROLL;
  }
  const double* gains = c.gains();
  if (dbg) {
    for (int k = lane; k < N; k += 32) {
      if (dbg->Kg) for (int i = 0; i < 12; ++i) dbg->Kg[((size_t)b * N + k) * 12 + i] = gains[k * kGainStride + i];
      if (dbg->kg) for (int i = 0; i < 2; ++i) dbg->kg[((size_t)b * N + k) * 2 + i] = gains[k * kGainStride + 12 + i];
    }
    if (dbg->dV && lane == 0) {
      dbg->dV[(size_t)b * 2] = dV[0];
      dbg->dV[(size_t)b * 2 + 1] = dV[1];
    }
  }
  // CalGradientNorm (:322-332)
#if CILQR_STRICT
  const double gnorm = s_gradient_norm(c, Xs);
#else
  double acc = 0.0;
#pragma unroll 1
  for (int k = lane; k < N; k += 32) {
    const double v0 = fabs(gains[k * kGainStride + 12]) / (fabs(Xs[6 * a.Kc + k]) + 1.0);
    const double v1 = fabs(gains[k * kGainStride + 13]) / (fabs(Xs[7 * a.Kc + k]) + 1.0);
    acc += fmax(v0, v1);
  }
  const double gnorm = warp_sum(acc) / N;
#endif
  if (dbg && dbg->gnorm && lane == 0) dbg->gnorm[b] = gnorm;
  if (gnorm < 1e-6 && lambda < 1e-5) {
    if (lane == 0) h->status = 2;
    __syncwarp();
    finish_scenario(c);
    return PH_INIT;
  }
  if (lane == 0) {
    h->dV0 = dV[0];
    h->dV1 = dV[1];
    h->gb = 0;
    h->rmode = 1;
    h->retired = 0;
    h->deferred = 0;
  }
  return PH_ROLL;
}

// EVAL: TotalCost of the initial guess (:172) or of ONE line-search candidate, followed by the
// accept / reject / lambda / convergence logic of Optimize (:253-309).  The nearest-segment index table
// of a trajectory lives with its slot: nidx(s) belongs to the trajectory in slot s.
//
// eval_initial   cost of the initial guess; iter_trajs[0], cost_[0]
// eval_prepare   which candidate is next (skips retired ones); PH_ROLL when the next group of step sizes
//                or a deferred repeat must be rolled out first; -2 when every step size was rejected
// eval_decide    accept (the candidate becomes the iterate, exits) or reject (advance to the next one)
__device__ __noinline__ int eval_initial(Ctx& c) {
  const KernelArgs& a = c.a;
  const int N = a.N, K = N + 1, lane = c.lane;
  CtxHdr* h = c.h;
  const unsigned b = h->b;
  const DebugPtrs* dbg = a.debug ? &a.dbg : nullptr;
  double cost5[5];
  const int cur = h->cur;
  stage_segments(c);
  const double* Xs = c.slot(cur);
  eval_cost(c, Xs, c.nidx(0), c.nidx(cur), cost5);  // nidx(0) is all zero: no previous iterate
  __syncwarp();
  push_cost(c, cost5);
  push_traj(c, Xs);
  if (lane == 0) h->cost_old = cost5[0];
  if (dbg) {
    copy_traj(c, Xs, dbg->X0 ? dbg->X0 + (size_t)b * K * 6 : nullptr, dbg->U0 ? dbg->U0 + (size_t)b * N * 2 : nullptr);
    if (dbg->cost0 && lane == 0) for (int i = 0; i < 5; ++i) dbg->cost0[(size_t)b * 5 + i] = cost5[i];
    if (dbg->nearest) {
      const unsigned char* nn = c.nidx(cur);
#pragma unroll 1
      for (int i = lane; i < K * 10; i += 32) dbg->nearest[(size_t)b * K * 10 + i] = nn[i];
    }
  }
  return PH_LIN;
}

__device__ __forceinline__ int eval_prepare(Ctx& c, int& ai_out) {
  CtxHdr* h = c.h;
  int ai = h->ai;
  const int gb = h->gb;
  const unsigned retired = h->retired, deferred = h->deferred;
  while (ai < kNAlpha && ai < gb + kSpec && ((retired >> ai) & 1u)) ++ai;  // non-finite rollout: rejected
  ai_out = ai;
  if (ai >= kNAlpha) return -2;
  if (ai >= gb + kSpec) {  // next group of step sizes
    __syncwarp();
    if (c.lane == 0) {
      h->gb = gb + kSpec;
      h->rmode = 1;
    }
    return PH_ROLL;
  }
  if ((deferred >> ai) & 1u) {
    __syncwarp();
    if (c.lane == 0) {
      h->ai = ai;
      h->rmode = 2;
    }
    return PH_ROLL;
  }
  return -1;
}

// every step size rejected (:297-308)
__device__ __noinline__ int eval_all_rejected(Ctx& c) {
  const DevParams& P = c.a.P;
  CtxHdr* h = c.h;
  const double reg_ratio = 1.6, reg_min = 1e-8, reg_max = 1e11;
  const double dl = fmax(h->dlambda * reg_ratio, reg_ratio);
  const double lam = fmax(h->lambda * dl, reg_min);
  const int iter = h->iter;
  const bool overflow = lam > reg_max;
  const bool exhausted = iter + 1 >= P.max_iter;
  __syncwarp();
  if (c.lane == 0) {
    h->ahash = fnv1a(h->ahash, (unsigned)kNAlpha);
    h->dlambda = dl;
    h->lambda = lam;
    if (overflow) h->status = 3;
    else h->iter = iter + 1;
  }
  __syncwarp();
  if (overflow || exhausted) {
    finish_scenario(c);
    return PH_INIT;
  }
  return PH_LIN;
}

__device__ __noinline__ int eval_decide(Ctx& c, int ai, const double cost5[5]) {
  const KernelArgs& a = c.a;
  const DevParams& P = a.P;
  const int lane = c.lane;
  CtxHdr* h = c.h;
  const double reg_ratio = 1.6, reg_min = 1e-8, beta_min = 1e-4, beta_max = 10.0;
  const int cur = h->cur, gb = h->gb;
  const unsigned retired = h->retired;
  const double* cd = c.slot(cand_slot(cur, ai));
  const double alpha = kAlphaList[ai];
  const double cost_old = h->cost_old;
  const double dcost = cost_old - cost5[0];
  const double expected = -alpha * (h->dV0 + alpha * h->dV1);
  const double z = dcost / expected;
  if ((z > beta_min && z < beta_max) && dcost > 0.0) {
    // ---- accepted (:266-296): the candidate becomes the iterate
    const double dl = fmin(h->dlambda / reg_ratio, 1.0 / reg_ratio);
    const double lam = h->lambda;
    int status = -1;
    if (dcost < P.abs_tol || dcost / cost_old < P.rel_tol) status = dcost < P.abs_tol ? 0 : 1;
    const int iter = h->iter;
    __syncwarp();
    if (lane == 0) {
      h->ahash = fnv1a(h->ahash, (unsigned)ai);
      h->cur = cand_slot(cur, ai);
      h->dlambda = dl;
      h->lambda = lam * dl * (lam > reg_min ? 1.0 : 0.0);
      h->cost_old = cost5[0];
      if (status >= 0) h->status = status;
    }
    push_cost(c, cost5);
    if (status >= 0) {
      finish_scenario(c);
      return PH_INIT;
    }
    push_traj(c, cd);
    if (iter + 1 >= P.max_iter) {  // loop exhausted (:312-319)
      if (lane == 0) h->iter = iter + 1;
      __syncwarp();
      finish_scenario(c);
      return PH_INIT;
    }
    if (lane == 0) h->iter = iter + 1;
    return PH_LIN;
  }
  // rejected: next step size
  ++ai;
  while (ai < kNAlpha && ai < gb + kSpec && ((retired >> ai) & 1u)) ++ai;
  if (ai < kNAlpha) {
    __syncwarp();
    if (lane == 0) h->ai = ai;
    return PH_EVAL;  // (a group change / deferred repeat is resolved by the next eval_prepare)
  }
  return eval_all_rejected(c);
}

// cost of line-search candidate ai of the context (any warp may do this: the result only depends on the
// context), nearest-segment indices -> nidx of the candidate's slot
__device__ __forceinline__ void eval_candidate(Ctx& c, int cur, int ai, double cost5[5]) {
  const int s = cand_slot(cur, ai);
  stage_segments(c);
  eval_cost(c, c.slot(s), c.nidx(cur), c.nidx(s), cost5);
  __syncwarp();
}

__device__ __noinline__ int phase_eval(Ctx& c) {
  const KernelArgs& a = c.a;
  const int N = a.N, K = N + 1, lane = c.lane;
  CtxHdr* h = c.h;
  const unsigned b = h->b;
  c.bind(b);
  if (h->emode == 0) return eval_initial(c);
  // ---- line search (:246-265): candidates in the reference's order until the first accept
  int ai;
  const int r = eval_prepare(c, ai);
  if (r == -2) return eval_all_rejected(c);
  if (r >= 0) return r;
  double cost5[5];
  eval_candidate(c, h->cur, ai, cost5);
  if (a.debug) {  // stage dump mode: one backward + one forward only; the initial guess is returned
    const DebugPtrs* dbg = &a.dbg;
    if (h->iter == 0 && ai == 0) {
      const double* cd = c.slot(cand_slot(h->cur, ai));
      copy_traj(c, cd, dbg->Xn ? dbg->Xn + (size_t)b * K * 6 : nullptr, dbg->Un ? dbg->Un + (size_t)b * N * 2 : nullptr);
      if (dbg->costn && lane == 0) for (int i = 0; i < 5; ++i) dbg->costn[(size_t)b * 5 + i] = cost5[i];
    }
    __syncwarp();
    finish_scenario(c);
    return PH_INIT;
  }
  return eval_decide(c, ai, cost5);
}

// ---- Help requests: a warp that runs a hot context alone (see the scheduler) lets idle warps of the CTA
// evaluate the other candidates of the current line-search group at the same time.  One request per CTA.
constexpr int kHelpClosed = 1 << 20;
struct HelpBoard {
  int owner;             // context index of the requester, -1: board free
  int next;              // next list position to claim (atomic); >= kHelpClosed: closed
  int n;                 // list length (kind 0) / number of linearisation windows (kind 1)
  int kind;              // 0: line-search candidates, 1: windows of the linearisation
  int done_cnt;          // kind 1: windows finished (atomic)
  int cur;               // the iterate's slot when the request was opened (an accept changes the header)
  int list[kSpec];       // candidate (step-size) indices in search order; position 0 is the owner's
  volatile int done[kSpec];
  double cost[kSpec][5];
};

// Line search of a hot context with helpers.  Returns the next phase (never PH_EVAL).
__device__ __noinline__ int gang_eval(Ctx& c, HelpBoard* hb, int mine) {
  const int lane = c.lane;
  CtxHdr* h = c.h;
  c.bind(h->b);
  for (;;) {
    __syncwarp();
    int ai;
    const int r = eval_prepare(c, ai);
    if (r == -2) return eval_all_rejected(c);
    if (r >= 0) return r;
    // candidates of this group that can be evaluated now, in search order
    const int gb = h->gb;
    const unsigned retired = h->retired, deferred = h->deferred;
    int list[kSpec], n = 0;
    for (int q = ai; q < kNAlpha && q < gb + kSpec; ++q) {
      if ((retired >> q) & 1u) continue;
      if ((deferred >> q) & 1u) break;
      list[n++] = q;
    }
    // open the board (if it is free) for positions 1..n-1
    int have_board = 0;
    if (n > 1) {
      if (lane == 0) have_board = atomicCAS(&hb->owner, -1, mine) == -1;
      have_board = __shfl_sync(kFull, have_board, 0);
      if (have_board) {
        if (lane == 0) {
          hb->n = n;
          hb->kind = 0;
          hb->cur = h->cur;
          for (int i = 0; i < kSpec; ++i) {
            hb->list[i] = i < n ? list[i] : 0;
            hb->done[i] = 0;
          }
          __threadfence_block();
          atomicExch(&hb->next, 1);
        }
        __syncwarp();
      }
    }
    double cost5[5];
    const int cur0 = h->cur;
    eval_candidate(c, cur0, list[0], cost5);
    int next = PH_EVAL;
    for (int i = 0;; ) {
      next = eval_decide(c, list[i], cost5);
      if (next != PH_EVAL) break;  // accepted, or exits
      if (++i >= n) break;         // rejected all of the list: prepare again (next group / deferred / all rejected)
      // cost of list[i]: from a helper if one took it, else computed here
      int mine_too = 1;
      if (have_board) {
        int k = 0;
        if (lane == 0) k = atomicCAS(&hb->next, i, i + 1);  // claim position i if nobody has
        k = __shfl_sync(kFull, k, 0);
        mine_too = k == i;
      }
      if (mine_too) {
        if (have_board && lane == 0) hb->done[i] = 1;  // (the close below only has to wait for helpers)
        eval_candidate(c, cur0, list[i], cost5);
      } else {
        unsigned spins = 0;
        while (hb->done[i] == 0 && ++spins < (1u << 24)) __nanosleep(200);
        if (spins >= (1u << 24) && lane == 0) raise_error(c.a, kErrHelp);
        __threadfence_block();
#pragma unroll
        for (int q = 0; q < 5; ++q) cost5[q] = hb->cost[i][q];
      }
    }
    if (have_board) {
      // close: nobody may claim any more; wait for the helpers that already did (they read the candidate
      // slots and write the nidx tables the next rollout / evaluation will overwrite)
      int claimed = 0;
      if (lane == 0) claimed = atomicExch(&hb->next, kHelpClosed);
      claimed = __shfl_sync(kFull, claimed, 0);
      claimed = claimed < n ? claimed : n;
      for (int i = 1; i < claimed; ++i) {
        unsigned spins = 0;
        while (hb->done[i] == 0 && ++spins < (1u << 24)) __nanosleep(200);
        if (spins >= (1u << 24) && lane == 0) raise_error(c.a, kErrHelp);
      }
      __threadfence_block();
      __syncwarp();
      if (lane == 0) atomicExch(&hb->owner, -1);
      __syncwarp();
    }
    if (next != PH_EVAL) return next;
  }
}

// Linearisation of a hot context with helpers: the windows of the horizon are independent.
__device__ __noinline__ int gang_lin(Ctx& c, HelpBoard* hb, int mine) {
  const KernelArgs& a = c.a;
  const int lane = c.lane;
  CtxHdr* h = c.h;
  c.bind(h->b);
  int have_board = 0;
  if (lane == 0) have_board = atomicCAS(&hb->owner, -1, mine) == -1;
  have_board = __shfl_sync(kFull, have_board, 0);
  if (!have_board || a.debug) {
    if (have_board && lane == 0) atomicExch(&hb->owner, -1);
    return phase_lin(c);
  }
  const int cur = h->cur, n = (a.N + kWin) / kWin;
  if (lane == 0) {
    hb->n = n;
    hb->kind = 1;
    hb->cur = cur;
    hb->done_cnt = 0;
    __threadfence_block();
    atomicExch(&hb->next, 0);
  }
  __syncwarp();
  for (;;) {
    int k = 0;
    if (lane == 0) k = atomicAdd(&hb->next, 1);
    k = __shfl_sync(kFull, k, 0);
    if (k >= n) break;
    linearize_window(c, k * kWin, c.slot(cur), c.nidx(cur), nullptr, 0);
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicAdd(&hb->done_cnt, 1);
  }
  // every window has been claimed; wait for the helpers' ones
  if (lane == 0) atomicExch(&hb->next, kHelpClosed);
  unsigned spins = 0;
  while (*(volatile int*)&hb->done_cnt < n && ++spins < (1u << 24)) __nanosleep(200);
  if (spins >= (1u << 24) && lane == 0) raise_error(a, kErrHelp);
  __threadfence();
  __syncwarp();
  if (lane == 0) atomicExch(&hb->owner, -1);
  __syncwarp();
  return PH_BACK;
}

// An idle warp looks at the board; returns true if it evaluated a candidate for somebody.
__device__ __noinline__ bool help_once(const KernelArgs& a, int smem, const int* gidx, HelpBoard* hb, int lane) {
  const int owner = *(volatile int*)&hb->owner;
  const int nx = *(volatile int*)&hb->next;
  if (owner < 0 || nx >= *(volatile int*)&hb->n || nx >= kHelpClosed) return false;
  int k = 0;
  if (lane == 0) k = atomicAdd(&hb->next, 1);
  k = __shfl_sync(kFull, k, 0);
  if (k >= kHelpClosed || k >= *(volatile int*)&hb->n) return false;
  __threadfence_block();
  // the request is stable from here on: its owner waits for done[k] before it changes anything
  const int id = *(volatile int*)&hb->owner;
  const int cur = hb->cur;
  Ctx c(a, smem, ctx_base(a, gidx, id), lane);
  c.bind(c.h->b);
  if (*(volatile int*)&hb->kind == 1) {
    linearize_window(c, k * kWin, c.slot(cur), c.nidx(cur), nullptr, 0);
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicAdd(&hb->done_cnt, 1);
    return true;
  }
  const int ai = hb->list[k];
  double cost5[5];
  eval_candidate(c, cur, ai, cost5);
  if (lane == 0) {
    for (int q = 0; q < 5; ++q) hb->cost[k][q] = cost5[q];
    __threadfence();
    hb->done[k] = 1;
  }
  __syncwarp();
  return true;
}

// ------------------------------------------------------------------------------------------
// Persistent kernel: one CTA per SM, up to kCtaWarps warps, ctx_per_cta contexts (> warps).
//
// Scheduling is CTA-local and barrier-free.  s_state[c] holds the phase context c waits for (or BUSY
// while a warp runs it, DONE after the last scenario); s_type is the phase type the CTA currently
// runs.  A warp looking for work claims (shared-memory CAS) a free context that waits for s_type; only
// when none is left does it move s_type to the type with the most waiting work, so at any moment the
// SM executes one phase body (two while stragglers of the previous type finish) and its hot code stays
// inside the ~32 KB the instruction cache can feed to unaligned warps.  A context whose next phase is
// again s_type (EVAL -> EVAL: the next step size of the line search) keeps its warp.  Warps never wait
// for each other; they only nap when every live context is being run by another warp.
constexpr int ST_BUSY = 6;
__global__ void __launch_bounds__(32 * kCtaWarps, 1) cilqr_solve_kernel(const __grid_constant__ KernelArgs a) {
  __shared__ int s_state[kMaxCtx];
  __shared__ int s_iter[kMaxCtx];  // iterations run so far by the scenario in each context (claim priority)
  __shared__ HelpBoard s_help;
  __shared__ int s_type;
  __shared__ int s_gidx[kMaxCtx];  // slot -> global context index (contexts live at a.ws + index * stride)
  __shared__ int s_drained;        // the batch's ticket is exhausted: no INIT will produce work any more
  __shared__ int s_donating;       // this CTA is handing its unfinished contexts to the next launch of the relay
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int C = a.ctx_per_cta;
  const int smem = warp * (a.sm.total_bytes / 8);  // this warp's stage, as an offset into smem_cta
  if (a.resume) {
    // a later launch of the drain relay: adopt up to resume_per_cta contexts that the previous launch left
    // unfinished, wherever in the workspace they live; they resume at the phase they were waiting for
    const unsigned n = a.resume[0];
    const unsigned first = (unsigned)blockIdx.x * (unsigned)a.resume_per_cta;
    if (first >= n) return;  // (uniform for the CTA)
    for (int i = threadIdx.x; i < kMaxCtx; i += blockDim.x) {
      int st = PH_DONE, it = 0, g = 0;
      if (i < a.resume_per_cta && first + i < n) {
        g = (int)a.resume[1 + first + i];
        const CtxHdr* hd = reinterpret_cast<const CtxHdr*>(a.ws + (size_t)g * a.cl.stride + a.cl.hdr);
        st = hd->wait;
        it = hd->iter;
      }
      s_gidx[i] = g;
      s_state[i] = st;
      s_iter[i] = it;
    }
  } else {
    for (int i = threadIdx.x; i < kMaxCtx; i += blockDim.x) {
      s_gidx[i] = blockIdx.x * C + i;
      s_state[i] = i < C ? PH_INIT : PH_DONE;
      s_iter[i] = 0;
    }
  }
  if (threadIdx.x == 0) {
    s_type = a.resume ? PH_BACK : PH_INIT;
    s_drained = a.resume ? 1 : 0;
    s_donating = 0;
    s_help.owner = -1;
    s_help.next = kHelpClosed;
    s_help.n = 0;
  }
  if (threadIdx.x == 0 && a.stats) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    atomicMin(a.stats + 8, now);
  }
  __syncthreads();
  // typical duration of the phases relative to one another: INIT 5, BACK 2, ROLL 5, EVAL 3, LIN 3
  const int wt[kNumTypes] = {5, 2, 5, 3, 3};
  unsigned naps = 0;
  unsigned st_pass = 0, st_poll = 0, st_fail = 0, st_ph[kNumTypes] = {0, 0, 0, 0, 0}, st_sw = 0;
  for (;;) {
    ++st_pass;
    // ---- snapshot of the context table (all lanes, identical result)
    const volatile int* st = s_state;
    int sl[kCtxWords];
#pragma unroll
    for (int w = 0; w < kCtxWords; ++w) sl[w] = st[lane + 32 * w];
    int cnt[kNumTypes];
#pragma unroll
    for (int p = 0; p < kNumTypes; ++p) {
      int n = 0;
#pragma unroll
      for (int w = 0; w < kCtxWords; ++w) n += __popc(__ballot_sync(kFull, sl[w] == p));
      cnt[p] = n;
    }
    if (cnt[0] + cnt[1] + cnt[2] + cnt[3] + cnt[4] == 0) {
      bool b = false;
#pragma unroll
      for (int w = 0; w < kCtxWords; ++w) b |= sl[w] == ST_BUSY;
      if (!__any_sync(kFull, b)) break;  // every context is DONE
      // the remaining contexts are all being run by other warps: help one of them if it asks ...
      if (help_once(a, smem, s_gidx, &s_help, lane)) {
        naps = 0;
        continue;
      }
      // ... else back off (up to ~4 us between polls)
      ++naps;
      ++st_poll;
      __nanosleep(naps < 16 ? 250 : 4000);
      if (naps > (1u << 22)) {  // (watchdog: never spin forever)
        if (lane == 0) raise_error(a, kErrIdle);
        break;
      }
      continue;
    }
    naps = 0;
    // ---- drain relay: the ticket is exhausted (contexts waiting for INIT are dead) and this CTA is down to its
    // last few contexts -- too few to keep sixteen warps busy, yet they would hold the SM (all of its registers)
    // for as long as the slowest of them runs.  Hand them to the next launch of the relay, which packs the
    // leftovers of all CTAs into a few CTAs, and leave the SM to the next batch's kernel.  A context carries its
    // whole state in the workspace; the phase it waits for goes into its header.
    if (a.donate_thr > 0 && *(const volatile int*)&s_drained) {
      int nb = 0;
#pragma unroll
      for (int w = 0; w < kCtxWords; ++w) nb += __popc(__ballot_sync(kFull, sl[w] == ST_BUSY));
      const int live_d = nb + cnt[1] + cnt[2] + cnt[3] + cnt[4];
      if (live_d <= a.donate_thr || *(const volatile int*)&s_donating) {
        if (lane == 0) s_donating = 1;
#pragma unroll
        for (int w = 0; w < kCtxWords; ++w) {
          const int idx = lane + 32 * w;
          const int stw = sl[w];
          if (stw == PH_INIT) {
            atomicCAS(&s_state[idx], PH_INIT, PH_DONE);
          } else if (stw >= PH_BACK && stw <= PH_LIN) {
            if (atomicCAS(&s_state[idx], stw, ST_BUSY) == stw) {
              __threadfence();  // acquire the context, then publish it with its waiting phase
              CtxHdr* hd = reinterpret_cast<CtxHdr*>(ctx_base(a, s_gidx, idx) + a.cl.hdr);
              hd->wait = stw;
              __threadfence();
              const unsigned pos = atomicAdd(a.donate, 1u);
              a.donate[1 + pos] = (unsigned)s_gidx[idx];
              *(volatile int*)&s_state[idx] = PH_DONE;
            }
          }
        }
        __syncwarp();
        __nanosleep(500);
        continue;  // busy contexts are donated when their warps release them; then every slot is DONE
      }
    }
    // ---- hot contexts first: a scenario far beyond the mean iteration count (the batch has a few with 10x)
    // would otherwise advance one phase per epoch and finish long after everything else; it is taken
    // whatever it waits for and keeps its warp, phase after phase, until it exits
    {
      // (once no more contexts are alive than the CTA has warps -- the end of the batch -- every context is
      // treated as hot: each gets its own warp for good, and the idle warps help)
      int nbusy = 0;
#pragma unroll
      for (int w = 0; w < kCtxWords; ++w) nbusy += __popc(__ballot_sync(kFull, sl[w] == ST_BUSY));
      const int live = nbusy + cnt[0] + cnt[1] + cnt[2] + cnt[3] + cnt[4];
      const int hot_thr = live <= a.hot_live ? 0 : a.hot_iter;
      int hk = -1;
#pragma unroll
      for (int w = 0; w < kCtxWords; ++w) {
        const int idx = lane + 32 * w;
        if (sl[w] < PH_DONE && sl[w] != PH_INIT && s_iter[idx] >= hot_thr) {
          const int k1 = (s_iter[idx] << 8) | idx;
          hk = k1 > hk ? k1 : hk;
        }
      }
      hk = __reduce_max_sync(kFull, hk);
      if (hk >= 0) {
        const int mine = hk & 255;
        int want_state = 0;
#pragma unroll
        for (int w = 0; w < kCtxWords; ++w) {
          const int v = __shfl_sync(kFull, sl[w], mine & 31);
          if ((mine >> 5) == w) want_state = v;
        }
        int ok = 0;
        if (lane == 0) ok = atomicCAS(&s_state[mine], want_state, ST_BUSY) == want_state;
        ok = __shfl_sync(kFull, ok, 0);
        if (ok) {
          __threadfence();  // acquire
          Ctx c(a, smem, ctx_base(a, s_gidx, mine), lane);
          int next = want_state;
          do {
            c.seg_staged = c.seg_staged && next == PH_EVAL;
            ++st_ph[next];
#ifdef CILQR_HOT_TIMING
            const long long tq0 = clock64();
            const int ph_now = next;
#endif
            if (next == PH_BACK) next = phase_back(c);
            else if (next == PH_LIN) next = gang_lin(c, &s_help, mine);
            else if (next == PH_ROLL) {
              roll_multi(a, smem, s_gidx, lane < 4 ? mine : -1, lane);
              __threadfence_block();
              next = PH_EVAL;
            } else if (c.h->emode == 1 && !a.debug) {
              next = gang_eval(c, &s_help, mine);
            } else {
              next = phase_eval(c);
            }
            __syncwarp();
#ifdef CILQR_HOT_TIMING
            if (lane == 0 && a.stats) {
              atomicAdd(a.stats + 266 + ph_now, (unsigned long long)(clock64() - tq0));
              atomicAdd(a.stats + 272 + ph_now, 1ull);
            }
#endif
          } while (next != PH_INIT && next < PH_DONE && !*(const volatile int*)&s_donating);
          __threadfence();  // release
          if (lane == 0) {
            s_iter[mine] = (next == PH_INIT || next >= PH_DONE) ? 0 : c.h->iter;
            if (next == PH_DONE) s_drained = 1;
            *(volatile int*)&s_state[mine] = next;
          }
          __syncwarp();
          naps = 0;
          continue;
        }
      }
    }
    int type = *(const volatile int*)&s_type;
    if (cnt[type] == 0) {
      int best = 0, best_score = -1;
#pragma unroll
      for (int p = 0; p < kNumTypes; ++p) {
        const int score = cnt[p] ? cnt[p] * wt[p] : -1;
        if (score > best_score) {
          best_score = score;
          best = p;
        }
      }
      type = best;
      ++st_sw;
      if (lane == 0) s_type = type;
    }
    // claim the waiting context of that type that has run the most iterations: scenarios with long
    // iteration counts (the batch has a few with 10x the mean) must not idle in the pool, or they finish
    // long after everything else and the SM drains.  Priority in buckets of four iterations; inside a
    // bucket every warp prefers a different context, so warps that look for work at the same moment do
    // not all go for the same one.
    const int rot = warp * 5;
    int key = -1;
#pragma unroll
    for (int w = 0; w < kCtxWords; ++w) {
      if (sl[w] == type) {
        const int idx = lane + 32 * w;
        const int k1 = ((s_iter[idx] >> 2) << 8) | ((idx + rot) & (kMaxCtx - 1));
        key = k1 > key ? k1 : key;
      }
    }
    key = __reduce_max_sync(kFull, key);
    int mine = ((key & 255) - rot) & (kMaxCtx - 1);
    if (type == PH_ROLL) {
      // rollouts use four lanes per context: claim up to eight waiting contexts for this warp
      int my_id = -1, got_n = 0;
      for (int r = 0; r < kRollGroups && key >= 0; ++r) {
        int ok = 0;
        if (lane == 0) ok = atomicCAS(&s_state[mine], type, ST_BUSY) == type;
        ok = __shfl_sync(kFull, ok, 0);
        if (ok) {
          if ((lane >> 2) == got_n) my_id = mine;
          ++got_n;
        }
        // next candidate (this one is taken either way)
        key = -1;
#pragma unroll
        for (int w = 0; w < kCtxWords; ++w) {
          const int idx = lane + 32 * w;
          if (idx == mine) sl[w] = ST_BUSY;
          if (sl[w] == type) {
            const int k1 = ((s_iter[idx] >> 2) << 8) | ((idx + rot) & (kMaxCtx - 1));
            key = k1 > key ? k1 : key;
          }
        }
        key = __reduce_max_sync(kFull, key);
        mine = ((key & 255) - rot) & (kMaxCtx - 1);
      }
      if (got_n == 0) {
        ++st_fail;
        __nanosleep(100);
        continue;
      }
      __threadfence();  // acquire
      st_ph[PH_ROLL] += got_n;
      roll_multi(a, smem, s_gidx, my_id, lane);
      __threadfence();  // release
      __syncwarp();
      if ((lane & 3) == 0 && my_id >= 0) *(volatile int*)&s_state[my_id] = PH_EVAL;
      __syncwarp();
      continue;
    }
    int got = 0;
    if (lane == 0) got = atomicCAS(&s_state[mine], type, ST_BUSY) == type;
    got = __shfl_sync(kFull, got, 0);
    if (!got) {
      ++st_fail;
      __nanosleep(100);
      continue;
    }
    __threadfence();  // acquire: the context was last written by another warp (plain stores, read back by cp.async too)
    Ctx c(a, smem, ctx_base(a, s_gidx, mine), lane);
    int next = type;
    do {
      if (next == PH_ROLL) break;  // rollouts are batched eight contexts per warp: back to the scheduler
      if (next != PH_EVAL) c.seg_staged = false;  // the other phases reuse the segment region of the stage
      ++st_ph[next];
      if (next == PH_INIT) next = phase_init(c);
      else if (next == PH_BACK) next = phase_back(c);
      else if (next == PH_LIN) next = phase_lin(c);
      else next = phase_eval(c);
      __syncwarp();
    } while (next < PH_DONE && next == *(const volatile int*)&s_type);
    __threadfence();  // release
    if (lane == 0) {
      s_iter[mine] = next == PH_INIT || next == PH_DONE ? 0 : c.h->iter;
      if (next == PH_DONE) s_drained = 1;  // phase_init found the ticket exhausted
      *(volatile int*)&s_state[mine] = next;
    }
    __syncwarp();
  }
  if (a.stats && lane == 0) {
    atomicAdd(a.stats + 0, (unsigned long long)st_pass);
    atomicAdd(a.stats + 1, (unsigned long long)st_poll);
    atomicAdd(a.stats + 2, (unsigned long long)st_fail);
    for (int p = 0; p < 4; ++p) atomicAdd(a.stats + 3 + p, (unsigned long long)st_ph[p]);  // (LIN count == BACK count)
    atomicAdd(a.stats + 7, (unsigned long long)st_sw);
  }
}

}  // namespace cilqr
