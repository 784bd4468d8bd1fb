// cilqr_capi.cu -- extern "C" shim (include/cilqr_b200.h) over the sm_100a solve kernel.
//
// Replaces, for the host program, IlqrOptimizer's constructor/Init (ilqr_optimizer.cc:13-51) with
// cilqr_create and IlqrOptimizer::Plan (ilqr_optimizer.cc:53-95) with cilqr_plan_batch.  There is
// no CPU path in this library: without a CUDA device every entry point fails with
// CILQR_E_NO_DEVICE / CILQR_E_CUDA.
#include "cilqr_kernel.cuh"
#include "corridor_kernel.cuh"
#include "cilqr_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "../../include/cilqr_b200.h"

using cilqr::DevParams;
using cilqr::KernelArgs;
using cilqr::SmemLayout;
using cilqr::CtxLayout;

namespace {

constexpr int kSlots = 2;           // double buffering of the host path
constexpr int kStatsWords = 8 + 2 + 256 + 12;  // scheduler counters, start time, completion histogram (2 ms buckets)
constexpr int kDefaultChunk = 4096; // scenarios per H2D chunk (watermark granularity) on the host path
constexpr int kRelayCap = 8192;     // entries per hand-over list of the drain relay (>= 1 + SMs * first threshold)

struct Slot {
  cudaStream_t stream = nullptr;
  unsigned int* ticket = nullptr;       // device [4]: scenario counter, error bits, scenarios finished, pad
  unsigned int* flags_host = nullptr;   // pinned [4]: the same words after the launch (host path)
  unsigned int* relay = nullptr;        // device [2][kRelayCap]: the two hand-over lists of the drain relay
  int launched_B = -1;                  // batch size of the last launch on this slot (-1: none to check)
  unsigned long long* stats = nullptr;  // scheduler counters of the last launch (cilqr_debug_stats)
  double* ws = nullptr;
  size_t ws_bytes = 0;
  // device staging for the host API
  char* in_buf = nullptr;
  char* out_buf = nullptr;
  size_t in_bytes = 0, out_bytes = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // host path: H2D copies run on copy_stream ahead of the solve kernel, which only takes scenarios
  // below the watermark *ready (written by the copy stream after every chunk)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_reset = nullptr;
  unsigned int* ready = nullptr;       // device
  unsigned int* ready_host = nullptr;  // pinned, one entry per chunk
  int ready_cap = 0;
};

}  // namespace

struct cilqr_handle {
  int device = 0;
  int num_sms = 0;
  int smem_optin = 0;
  int solve_smem_max = 0;  // dynamic shared memory the solve kernel may use: opt-in size minus its static part
  CilqrParams params;
  DevParams dev;
  int N_max = 0, M_max = 0, S_max = 0, B_max = 0;
  int chunk = kDefaultChunk;
  bool no_zero_copy = false;  // CILQR_NO_ZERO_COPY=1: always stage outputs on the device (development knob)
  int watchdog_ms = 4000;     // how long the kernel waits for the host path's watermark (cilqr_debug_host_path)
  int starve_after = -1;      // test hook: the host path never raises the watermark beyond this many scenarios
  Slot slots[kSlots];
  int64_t launches = 0;
  int last_slot = 0;
  bool timed = false;
  std::string cuda_err;
  // corridor builder: staging for the host path and launch timing
  char* corr_buf = nullptr;
  size_t corr_bytes = 0;
  cudaEvent_t corr_ev0 = nullptr, corr_ev1 = nullptr;
  bool corr_timed = false;
  int64_t corr_launches = 0;
  // scratch / events of the other translation units (cilqr_internal.h)
  char* aux_buf[3] = {nullptr, nullptr, nullptr};
  size_t aux_bytes[3] = {0, 0, 0};
  cudaEvent_t aux_ev[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
};

namespace {

int fail_cuda(cilqr_handle* h, cudaError_t e, const char* where) {
  if (h) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s: %s", where, cudaGetErrorString(e));
    h->cuda_err = buf;
  }
  return CILQR_E_CUDA;
}
#define CK(call)                                            \
  do {                                                      \
    cudaError_t e_ = (call);                                \
    if (e_ != cudaSuccess) return fail_cuda(h, e_, #call);  \
  } while (0)

void make_dev_params(const CilqrParams& p, DevParams* d) {
  d->dt = p.delta_t;
  d->L = p.wheel_base;
  d->inv_L = 1.0 / p.wheel_base;
  d->rt = 1.0 / p.barrier_t;
  d->eps = p.barrier_eps;
  d->inv_eps = 1.0 / p.barrier_eps;
  d->inv_eps2 = 1.0 / (p.barrier_eps * p.barrier_eps);
#if CILQR_STRICT
  // the strict build shares its libm with oracle/libcilqr_oracle_pm.so on the host side too
#define CILQR_HOST_LOG pm_log
#define CILQR_HOST_HYPOT pm_hypot
#else
#define CILQR_HOST_LOG log
#define CILQR_HOST_HYPOT hypot
#endif
  d->relax_c = -0.5 * d->rt - d->rt * CILQR_HOST_LOG(p.barrier_eps);
  d->rt_log_eps = d->rt * CILQR_HOST_LOG(p.barrier_eps);
  d->vmax = p.max_velocity;
  d->amin = p.min_acceleration;
  d->amax = p.max_acceleration;
  d->dmin = p.delta_min;
  d->dmax = p.delta_max;
  d->jmin = p.jerk_min;
  d->jmax = p.jerk_max;
  d->drmin = p.delta_rate_min;
  d->drmax = p.delta_rate_max;
  d->wx = p.w_x_target;
  d->wy = p.w_y_target;
  d->wth = p.w_theta;
  d->wv = p.w_v;
  d->wa = p.w_a;
  d->wd = p.w_delta;
  d->wj = p.w_jerk;
  d->wdr = p.w_delta_rate;
  d->abs_tol = p.abs_cost_tol;
  d->rel_tol = p.rel_cost_tol;
  // disc centres, ilqr_optimizer.cc:556-565
  const double Ld = (p.rear_hang_length + p.wheel_base + p.front_hang_length) / p.num_of_disc;
  for (int j = 0; j < cilqr::kDisc; ++j) d->off[j] = (Ld * (j - 0.5) - p.rear_hang_length);
  // CalculateDiscRadius, ilqr_optimizer.cc:97-104
  const double length = p.front_hang_length + p.wheel_base + p.rear_hang_length;
  const double r = CILQR_HOST_HYPOT(p.width / 2.0, length / 2.0 / p.num_of_disc);
  d->shrink_corr = r + p.safe_margin;
  d->shrink_lane = r;
  d->max_iter = p.max_iter_num;
}

SmemLayout make_layout(int N, int M_max, int S_left, int S_right) {
  SmemLayout L;
  const int K = N + 1, S = S_left + S_right;
  const int ng = (S_left + cilqr::kGroup - 1) / cilqr::kGroup + (S_right + cilqr::kGroup - 1) / cilqr::kGroup;
  auto even = [](int x) { return (x + 1) & ~1; };
  // EVAL: seg, grp, trig
  L.seg = 0;
  L.grp = S * cilqr::kSegStride;
  L.trig = even(L.grp + ng * 3 + S);  // group circles, then the certificate radii of the nearest-segment search
  L.pl_e = even(L.trig + 2 * K);
#if CILQR_STRICT
  const int e_end = L.pl_e + 32 * cilqr::strict_row_width(M_max);  // the term buffer of the ordered cost sums
#else
  const int e_end = L.pl_e + cilqr::kTileBufs * M_max * cilqr::kPlaneTile;
#endif
  // LIN: lin window, plane tiles.  BACK: record ring, scr.  INIT builds seg/grp while iqr uses scr, so scr
  // lies behind both
  L.lin = 0;
  L.pl_b = even(cilqr::kWin * cilqr::kLinStride);
  L.red = L.pl_b + cilqr::kTileBufs * M_max * cilqr::kPlaneTile;  // [9][32] disc-lane partial sums
  const int l_end = L.red + 9 * 32;
  L.bring = 0;
  L.scr = std::max(even(2 * cilqr::kBackChunk * cilqr::kRecStride), L.trig);
  const int b_end = std::max(L.scr + cilqr::kScratch, l_end);
  // ROLL: ring
  L.ring = 0;
  const int total = std::max(std::max(e_end, b_end), cilqr::kRollGroups * cilqr::kRingDoubles);
  L.total_bytes = (total * 8 + 15) / 16 * 16;  // per-warp stages are packed back to back in the CTA
  return L;
}

CtxLayout make_ctx_layout(int N, int M_max, int S_left, int S_right) {
  CtxLayout C;
  const int K = N + 1, Kc = (K + 3) / 4 * 4, S = S_left + S_right;
  const int ng = (S_left + cilqr::kGroup - 1) / cilqr::kGroup + (S_right + cilqr::kGroup - 1) / cilqr::kGroup;
  auto even = [](int x) { return (x + 1) & ~1; };
  C.planes = 0;
  C.slots = M_max * 3 * Kc;
  C.gains = C.slots + cilqr::kTrajSlots * 8 * Kc;
  C.lin = C.gains + (N + cilqr::kRollChunk - 1) / cilqr::kRollChunk * cilqr::kRollChunk * cilqr::kGainStride;
  C.seg = C.lin + (K + cilqr::kBackChunk - 1) / cilqr::kBackChunk * cilqr::kBackChunk * cilqr::kRecStride;
  C.grp = C.seg + S * cilqr::kSegStride;
  C.nidx = even(C.grp + ng * 3 + S);
  C.nidx_bytes = (K * 10 + 7) / 8 * 8;
  C.hdr = C.nidx + cilqr::kTrajSlots * C.nidx_bytes / 8;
  C.stride = (C.hdr + cilqr::kHdrDoubles + 15) / 16 * 16;
  return C;
}

struct Launch {
  SmemLayout sm;
  CtxLayout cl;
  int grid = 0;
  int warps = 0;
  int ctx = 0;
  int Kc = 0;
};

int plan_launch(cilqr_handle* h, int B, int N, int M_max, int S_left, int S_right, Launch* out) {
  Launch L;
  L.sm = make_layout(N, M_max, S_left, S_right);
  L.cl = make_ctx_layout(N, M_max, S_left, S_right);
  L.Kc = (N + 1 + 3) / 4 * 4;
  // warps per CTA: as many per-warp stages as fit the SM's shared memory, at most kCtaWarps
  const int W = std::min(cilqr::kCtaWarps, h->solve_smem_max / L.sm.total_bytes);
  if (W < 1) return CILQR_E_SMEM;
  L.warps = W;  // (the kernel's dynamic shared-memory limit was raised to smem_optin once, in cilqr_create)
  // persistent grid: one CTA (W warps) per SM, each owning `ctx` scenario contexts
  L.grid = std::max(1, std::min(B, h->num_sms));
  int ctx = cilqr::kMaxCtx;  // measured best at N = 100 (profiles/): more waiting contexts -> fewer idle warps
  if (const char* e = getenv("CILQR_B200_CTX")) ctx = atoi(e);  // development knob
  ctx = std::max(1, std::min(ctx, cilqr::kMaxCtx));
  L.ctx = std::max(1, std::min(ctx, (B + L.grid - 1) / L.grid));
  *out = L;
  return CILQR_OK;
}

int ensure_ws(cilqr_handle* h, Slot* s, size_t bytes) {
  if (s->ws_bytes >= bytes) return CILQR_OK;
  CK(cudaStreamSynchronize(s->stream));
  if (s->ws) CK(cudaFree(s->ws));
  s->ws = nullptr;
  s->ws_bytes = 0;
  CK(cudaMalloc(&s->ws, bytes));
  s->ws_bytes = bytes;
  return CILQR_OK;
}

int validate(const cilqr_handle* h, const CilqrBatchIn* in, const CilqrBatchOut* out) {
  if (!h || !in || !out) return CILQR_E_INVALID;
  // guards of IlqrOptimizer::Plan, ilqr_optimizer.cc:64-78
  if (!out->states || !out->controls || !out->status) return CILQR_E_INVALID;
  if (!in->start || !in->coarse || !in->corridor || !in->corridor_cnt || !in->lane_left || !in->lane_right)
    return CILQR_E_INVALID;
  if (in->B < 0 || in->N < 1 || in->M_max < 1 || in->S_left < 1 || in->S_right < 1) return CILQR_E_INVALID;
  if (in->S_left > 255 || in->S_right > 255) return CILQR_E_CAPACITY;  // nearest index is cached as a byte
  if (in->N > h->N_max || in->M_max > h->M_max || in->S_left > h->S_max || in->S_right > h->S_max)
    return CILQR_E_CAPACITY;
  if ((out->iter_states || out->cost_hist) && out->hist_cap < 1) return CILQR_E_INVALID;
  if (in->init_mode < CILQR_INIT_IQR || in->init_mode > CILQR_INIT_GUESS) return CILQR_E_INVALID;
  if (in->init_mode != CILQR_INIT_IQR && !in->init_controls) return CILQR_E_INVALID;
  if (in->init_mode == CILQR_INIT_GUESS && !in->init_states) return CILQR_E_INVALID;
  return CILQR_OK;
}

int launch_solve(cilqr_handle* h, Slot* s, cudaStream_t stream, const CilqrBatchIn* in, const CilqrBatchOut* out,
                 const CilqrDebugOut* dbg, const unsigned int* ready = nullptr) {
  if (in->B == 0) return CILQR_OK;
  Launch L;
  int rc = plan_launch(h, in->B, in->N, in->M_max, in->S_left, in->S_right, &L);
  if (rc != CILQR_OK) return rc;
  rc = ensure_ws(h, s, (size_t)L.grid * L.ctx * L.cl.stride * sizeof(double));
  if (rc != CILQR_OK) return rc;
  KernelArgs a;
  memset(&a, 0, sizeof(a));
  a.P = h->dev;
  a.sm = L.sm;
  a.B = in->B;
  a.N = in->N;
  a.M_max = in->M_max;
  a.S_left = in->S_left;
  a.S_right = in->S_right;
  a.cl = L.cl;
  a.Kc = L.Kc;
  a.ctx_per_cta = L.ctx;
  a.hot_iter = 16;
  a.hot_live = L.warps;
  if (const char* e = getenv("CILQR_B200_HOT_LIVE")) a.hot_live = atoi(e);  // development knob
  if (const char* e = getenv("CILQR_B200_HOT")) a.hot_iter = atoi(e);  // development knob
  a.start = in->start;
  a.coarse = in->coarse;
  a.corridor = in->corridor;
  a.corridor_cnt = in->corridor_cnt;
  a.lane_left = in->lane_left;
  a.lane_right = in->lane_right;
  a.init_mode = in->init_mode;
  a.guess_states = in->init_states;
  a.guess_controls = in->init_controls;
  a.states = out->states;
  a.controls = out->controls;
  a.status = out->status;
  a.trajectory = out->trajectory;
  a.result = out->result;
  a.init_states = out->init_states;
  a.init_controls = out->init_controls;
  a.cost_hist = out->cost_hist;
  a.iter_states = out->iter_states;
  a.iter_controls = out->iter_controls;
  a.hist_len = out->hist_len;
  a.hist_cap = out->hist_cap;
  a.ws = s->ws;
  a.ticket = s->ticket;
  a.ready = ready;
  a.watchdog_ns = (unsigned long long)std::max(1, h->watchdog_ms) * 1000000ull;
  a.stats = s->stats;
  // an unsolved scenario must be recognisable: every status row starts as the sentinel (all bits set = NaN)
  CK(cudaMemsetAsync(out->status, 0xff, (size_t)in->B * CILQR_STATUS_DOUBLES * sizeof(double), stream));
  CK(cudaMemsetAsync(s->stats, 0, kStatsWords * sizeof(unsigned long long), stream));
  CK(cudaMemsetAsync(s->stats + 8, 0xff, sizeof(unsigned long long), stream));  // start time: atomicMin
  a.debug = 0;
  if (dbg) {
    a.debug = 1;
    a.dbg.corridor = dbg->corridor;
    a.dbg.lanes = dbg->lanes;
    a.dbg.X0 = dbg->X0;
    a.dbg.U0 = dbg->U0;
    a.dbg.cost0 = dbg->cost0;
    a.dbg.A11 = dbg->A11;
    a.dbg.Jx = dbg->Jx;
    a.dbg.Ju = dbg->Ju;
    a.dbg.Hx = dbg->Hx;
    a.dbg.Hu = dbg->Hu;
    a.dbg.Kg = dbg->Kg;
    a.dbg.kg = dbg->kg;
    a.dbg.dV = dbg->dV;
    a.dbg.Xn = dbg->Xn;
    a.dbg.Un = dbg->Un;
    a.dbg.costn = dbg->costn;
    a.dbg.nearest = dbg->nearest;
    a.dbg.gnorm = dbg->gnorm;
  }
  CK(cudaMemsetAsync(s->ticket, 0, 4 * sizeof(unsigned int), stream));
  // Drain relay (optional, full grids only; OFF by default -- see DESIGN.md section 2.5 for the measurements).
  // Stage 1 solves the batch until a CTA is down to thr[0] unfinished contexts and hands those over; every later
  // stage packs the leftovers of the previous one kMaxCtx per CTA into a smaller grid and hands over again at its
  // own threshold; the last stage runs to the end.  It trades single-batch latency (a stage starts when the
  // previous one has ended everywhere) for SMs that are free for the next batch's kernel during the drain.
  // Results do not depend on it: a context is self-contained and every phase of it is run by exactly one warp
  // with the same code.  CILQR_B200_RELAY="t1,t2,..." (development knob).
  int thr[6] = {0, 0, 0, 0, 0, 0}, n_thr = 0;
  if (const char* e = getenv("CILQR_B200_RELAY")) {
    const char* p = e;
    while (n_thr < 6 && *p) {
      thr[n_thr++] = atoi(p);
      while (*p && *p != ',') ++p;
      if (*p == ',') ++p;
    }
  }
  bool relay = !dbg && n_thr > 0 && thr[0] > 0 && L.grid == h->num_sms && L.ctx > thr[0] &&
               1 + (long long)L.grid * thr[0] <= kRelayCap;
  for (int i = 1; relay && i < n_thr; ++i) relay = thr[i] < thr[i - 1];
  unsigned int* lists[2] = {s->relay, s->relay + kRelayCap};
  if (relay) CK(cudaMemsetAsync(s->relay, 0, 2 * kRelayCap * sizeof(unsigned int), stream));
  a.donate_thr = relay ? thr[0] : 0;
  a.donate = relay ? lists[0] : nullptr;
  a.resume = nullptr;
  a.resume_per_cta = 0;
  CK(cudaEventRecord(s->ev0, stream));
  cilqr::cilqr_solve_kernel<<<L.grid, 32 * L.warps, L.sm.total_bytes * L.warps, stream>>>(a);
  CK(cudaGetLastError());
  h->launches += 1;
  if (relay) {
    KernelArgs r = a;
    int prev_grid = L.grid;
    for (int st = 0; st < n_thr && thr[st] > 0; ++st) {
      const int per = cilqr::kMaxCtx;
      const int grid = (prev_grid * thr[st] + per - 1) / per;  // upper bound of the entries / per
      const int next_thr = st + 1 < n_thr ? thr[st + 1] : 0;
      r.ctx_per_cta = per;
      r.resume = lists[st & 1];
      r.resume_per_cta = per;
      r.donate_thr = next_thr;
      r.donate = next_thr > 0 ? lists[(st + 1) & 1] : nullptr;
      if (next_thr > 0) CK(cudaMemsetAsync(lists[(st + 1) & 1], 0, sizeof(unsigned int), stream));
      cilqr::cilqr_solve_kernel<<<grid, 32 * L.warps, L.sm.total_bytes * L.warps, stream>>>(r);
      CK(cudaGetLastError());
      h->launches += 1;
      prev_grid = grid;
    }
  }
  CK(cudaEventRecord(s->ev1, stream));
  h->last_slot = (int)(s - h->slots);
  h->timed = true;
  s->launched_B = in->B;
  return CILQR_OK;
}

// After the launch's stream has been synchronised: did every scenario finish, did a watchdog fire?
int check_launch(cilqr_handle* h, Slot* s, const unsigned int flags[4]) {
  const int B = s->launched_B;
  s->launched_B = -1;
  if (B < 0) return CILQR_OK;
  if (flags[1] == 0 && flags[2] == (unsigned)B) return CILQR_OK;
  char buf[200];
  snprintf(buf, sizeof(buf), "solve launch incomplete: %u of %d scenarios finished, error bits 0x%x (1: input transfer "
           "starved, 2: idle watchdog, 4: help-board watchdog)", flags[2], B, flags[1]);
  h->cuda_err = buf;
  return CILQR_E_TIMEOUT;
}

}  // namespace

extern "C" {

int cilqr_abi_version(void) { return CILQR_ABI_VERSION; }

void cilqr_default_params(CilqrParams* p) {
  if (!p) return;
  // vehicle_param.h:26-64
  p->front_hang_length = 0.96;
  p->wheel_base = 1.0;
  p->rear_hang_length = 0.929;
  p->width = 1.942;
  p->max_velocity = 20.0;
  p->min_acceleration = -5.0;
  p->max_acceleration = 5.0;
  p->jerk_min = -10.0;
  p->jerk_max = 10.0;
  p->delta_min = -40.0 / 180 * M_PI;
  p->delta_max = 40.0 / 180 * M_PI;
  p->delta_rate_min = p->delta_min / 3.0;
  p->delta_rate_max = p->delta_max / 3.0;
  // planner_config.h:45-66
  p->safe_margin = 0.2;
  p->w_jerk = 1;
  p->w_delta_rate = 1;
  p->w_x_target = 0.5;
  p->w_y_target = 0.5;
  p->w_theta = 1e-3;
  p->w_v = 0.0;
  p->w_a = 0.0;
  p->w_delta = 0.0;
  p->abs_cost_tol = 1e-2;
  p->rel_cost_tol = 1e-2;
  // barrier_function.h:143-146
  p->barrier_t = 5.0;
  p->barrier_eps = 0.01;
  p->delta_t = 0.1;  // planner_config.h:94
  p->num_of_disc = 5;
  p->max_iter_num = 200;
}

const char* cilqr_strerror(int code) {
  switch (code) {
    case CILQR_OK: return "ok";
    case CILQR_E_INVALID: return "invalid argument (null pointer, empty constraint set or bad size)";
    case CILQR_E_CUDA: return "CUDA runtime error";
    case CILQR_E_NO_DEVICE: return "no CUDA device of compute capability 10.x (this library has no CPU fallback)";
    case CILQR_E_CAPACITY: return "request exceeds the capacity the handle was created with";
    case CILQR_E_SMEM: return "horizon does not fit the per-warp shared-memory stage";
    case CILQR_E_NCCL: return "NCCL is missing or failed (only the gathered copy of cilqr_plan_sharded needs it)";
    case CILQR_E_TIMEOUT: return "the solve kernel did not finish every scenario (see cilqr_last_cuda_error); "
                                 "unsolved scenarios carry the NaN sentinel in their status row";
    default: return "unknown error";
  }
}

const char* cilqr_last_cuda_error(const cilqr_handle* h) { return h ? h->cuda_err.c_str() : ""; }

int cilqr_create(const CilqrParams* params, int device, int N_max, int M_max, int S_max, int B_max,
                 cilqr_handle** out) {
  if (!out) return CILQR_E_INVALID;
  *out = nullptr;
  if (!params || N_max < 1 || M_max < 1 || S_max < 1 || B_max < 1) return CILQR_E_INVALID;
  if (params->num_of_disc != cilqr::kDisc) return CILQR_E_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1 || device < 0 || device >= count) {
    cudaGetLastError();
    return CILQR_E_NO_DEVICE;
  }
  cilqr_handle* h = new (std::nothrow) cilqr_handle();
  if (!h) return CILQR_E_INVALID;
  h->device = device;
  h->params = *params;
  make_dev_params(*params, &h->dev);
  h->N_max = N_max;
  h->M_max = M_max;
  h->S_max = S_max;
  h->B_max = B_max;
  const char* env_chunk = getenv("CILQR_CHUNK");
  if (env_chunk && atoi(env_chunk) > 0) h->chunk = atoi(env_chunk);
  if (const char* e = getenv("CILQR_NO_ZERO_COPY")) h->no_zero_copy = atoi(e) != 0;
  auto bail = [&](int rc) {
    cilqr_destroy(h);
    return rc;
  };
  cudaDeviceProp prop;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    cudaGetLastError();
    return bail(CILQR_E_CUDA);
  }
  if (prop.major != 10) return bail(CILQR_E_NO_DEVICE);
  h->num_sms = prop.multiProcessorCount;
  h->smem_optin = (int)prop.sharedMemPerBlockOptin;
  {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, (const void*)cilqr::cilqr_solve_kernel) != cudaSuccess) return bail(CILQR_E_CUDA);
    h->solve_smem_max = h->smem_optin - (int)fa.sharedSizeBytes;
  }
  if (make_layout(N_max, M_max, S_max, S_max).total_bytes > h->solve_smem_max) return bail(CILQR_E_SMEM);
  // function attributes are per-device state shared by every handle: set them once, to the maximum, so that
  // handles of different shapes on different host threads cannot shrink each other's limit between a
  // set-attribute and a launch
  // (the dynamic limit is the opt-in size minus the kernel's static shared memory)
  auto raise_limit = [&](const void* fn) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, fn);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin - (int)fa.sharedSizeBytes);
    if (e != cudaSuccess) fail_cuda(h, e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    return e == cudaSuccess;
  };
  if (!raise_limit((const void*)cilqr::cilqr_solve_kernel) || !raise_limit((const void*)corridor::corridor_build_kernel) ||
      cilqr_internal_dp_set_smem(h->smem_optin) != CILQR_OK || cilqr_internal_tracker_set_smem(h->smem_optin) != CILQR_OK) {
    fprintf(stderr, "cilqr_b200: %s\n", h->cuda_err.c_str());
    return bail(CILQR_E_CUDA);
  }
  if (const char* e = getenv("CILQR_WATCHDOG_MS")) h->watchdog_ms = std::max(1, atoi(e));
  for (int i = 0; i < kSlots; ++i) {
    Slot* s = &h->slots[i];
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(CILQR_E_CUDA);
    if (cudaMalloc(&s->ticket, 4 * sizeof(unsigned int)) != cudaSuccess) return bail(CILQR_E_CUDA);
    if (cudaHostAlloc((void**)&s->flags_host, 4 * sizeof(unsigned int), cudaHostAllocDefault) != cudaSuccess)
      return bail(CILQR_E_CUDA);
    if (cudaMalloc(&s->relay, 2 * kRelayCap * sizeof(unsigned int)) != cudaSuccess) return bail(CILQR_E_CUDA);
    if (cudaMalloc(&s->stats, kStatsWords * sizeof(unsigned long long)) != cudaSuccess) return bail(CILQR_E_CUDA);
    if (cudaEventCreate(&s->ev0) != cudaSuccess || cudaEventCreate(&s->ev1) != cudaSuccess) return bail(CILQR_E_CUDA);
    if (cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(CILQR_E_CUDA);
    if (cudaEventCreateWithFlags(&s->ev_reset, cudaEventDisableTiming) != cudaSuccess) return bail(CILQR_E_CUDA);
    if (cudaMalloc(&s->ready, sizeof(unsigned int)) != cudaSuccess) return bail(CILQR_E_CUDA);
  }
  *out = h;
  return CILQR_OK;
}

void cilqr_destroy(cilqr_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (int i = 0; i < kSlots; ++i) {
    Slot* s = &h->slots[i];
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->ticket) cudaFree(s->ticket);
    if (s->flags_host) cudaFreeHost(s->flags_host);
    if (s->relay) cudaFree(s->relay);
    if (s->stats) cudaFree(s->stats);
    if (s->ws) cudaFree(s->ws);
    if (s->in_buf) cudaFree(s->in_buf);
    if (s->out_buf) cudaFree(s->out_buf);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->copy_stream) {
      cudaStreamSynchronize(s->copy_stream);
      cudaStreamDestroy(s->copy_stream);
    }
    if (s->ev_reset) cudaEventDestroy(s->ev_reset);
    if (s->ready) cudaFree(s->ready);
    if (s->ready_host) cudaFreeHost(s->ready_host);
    if (s->stream) cudaStreamDestroy(s->stream);
  }
  for (int i = 0; i < 3; ++i) {
    if (h->aux_buf[i]) cudaFree(h->aux_buf[i]);
    for (int j = 0; j < 2; ++j)
      if (h->aux_ev[i][j]) cudaEventDestroy(h->aux_ev[i][j]);
  }
  if (h->corr_buf) cudaFree(h->corr_buf);
  if (h->corr_ev0) cudaEventDestroy(h->corr_ev0);
  if (h->corr_ev1) cudaEventDestroy(h->corr_ev1);
  delete h;
}

int cilqr_plan_batch_device(cilqr_handle* h, const CilqrBatchIn* in, const CilqrBatchOut* out, void* cuda_stream) {
  int rc = validate(h, in, out);
  if (rc != CILQR_OK) return rc;
  CK(cudaSetDevice(h->device));
  Slot* s = &h->slots[0];
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : s->stream;
  return launch_solve(h, s, st, in, out, nullptr);
}

int cilqr_debug_first_iteration(cilqr_handle* h, const CilqrBatchIn* in, const CilqrDebugOut* dbg) {
  if (!h || !in || !dbg) return CILQR_E_INVALID;
  CK(cudaSetDevice(h->device));
  // the regular outputs are still produced; park them in scratch device memory
  const size_t K = in->N + 1;
  double* tmp = nullptr;
  const size_t n = (size_t)in->B * (K * 6 + (size_t)in->N * 2 + 8);
  CK(cudaMalloc(&tmp, n * sizeof(double)));
  CilqrBatchOut out;
  memset(&out, 0, sizeof(out));
  out.states = tmp;
  out.controls = tmp + (size_t)in->B * K * 6;
  out.status = out.controls + (size_t)in->B * in->N * 2;
  int rc = validate(h, in, &out);
  if (rc == CILQR_OK) rc = launch_solve(h, &h->slots[0], h->slots[0].stream, in, &out, dbg);
  cudaError_t e = cudaStreamSynchronize(h->slots[0].stream);
  cudaFree(tmp);
  if (rc != CILQR_OK) return rc;
  if (e != cudaSuccess) return fail_cuda(h, e, "debug kernel");
  CK(cudaMemcpy(h->slots[0].flags_host, h->slots[0].ticket, 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost));
  return check_launch(h, &h->slots[0], h->slots[0].flags_host);
}

int cilqr_synchronize(cilqr_handle* h) {
  if (!h) return CILQR_E_INVALID;
  CK(cudaSetDevice(h->device));
  int rc = CILQR_OK;
  for (int i = 0; i < kSlots; ++i) {
    Slot* s = &h->slots[i];
    CK(cudaStreamSynchronize(s->stream));
    if (s->launched_B >= 0) {
      // (a launch enqueued on a caller stream: the copy below is ordered after it only if the caller has
      // synchronised that stream, which is the documented contract of cilqr_plan_batch_device)
      CK(cudaMemcpy(s->flags_host, s->ticket, 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost));
      const int r = check_launch(h, s, s->flags_host);
      if (r != CILQR_OK) rc = r;
    }
  }
  return rc;
}

int cilqr_debug_host_path(cilqr_handle* h, int watchdog_ms, int starve_after) {
  if (!h) return CILQR_E_INVALID;
  if (watchdog_ms > 0) h->watchdog_ms = watchdog_ms;
  h->starve_after = starve_after;
  return CILQR_OK;
}

// Host path: ONE solve launch for the whole batch.  The inputs travel host -> device in chunks on a copy
// stream; after every chunk the copy stream bumps a device watermark, and the kernel's INIT phase only
// takes scenarios below it, so the solve starts as soon as the first chunk has landed and the rest of the
// transfer hides behind it (a persistent kernel fed by a stream, instead of one launch -- and one drain
// tail -- per chunk).  The outputs return in one D2H pass after the kernel.
int cilqr_plan_batch(cilqr_handle* h, const CilqrBatchIn* in, const CilqrBatchOut* out) {
  int rc = validate(h, in, out);
  if (rc != CILQR_OK) return rc;
  if (in->B > h->B_max) return CILQR_E_CAPACITY;
  if (in->B == 0) return CILQR_OK;
  CK(cudaSetDevice(h->device));
  const size_t K = in->N + 1, N = in->N, M = in->M_max, B = in->B;
  // chunk schedule: small first chunks (the kernel starts after the first one), doubling up to h->chunk
  std::vector<size_t> chunk_end;
  {
    size_t done = 0, c = std::min<size_t>(512, (size_t)h->chunk);
    while (done < (size_t)in->B) {
      done = std::min<size_t>(done + c, (size_t)in->B);
      chunk_end.push_back(done);
      c = std::min<size_t>(c * 2, (size_t)h->chunk);
    }
  }
  const int n_chunks = (int)chunk_end.size();
  const int H = out->hist_cap;
  // per-scenario byte counts
  const size_t b_start = 4 * 8, b_coarse = K * 6 * 8, b_corr = K * M * 3 * 8, b_cnt = K * 4;
  const size_t b_ll = (size_t)in->S_left * 7 * 8, b_lr = (size_t)in->S_right * 7 * 8;
  const size_t b_st = K * 6 * 8, b_ct = N * 2 * 8, b_status = 8 * 8, b_traj = K * 13 * 8;
  const size_t b_ch = (size_t)H * 5 * 8, b_is = (size_t)H * K * 6 * 8, b_ic = (size_t)H * N * 2 * 8, b_hl = 2 * 4;
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  const bool has_gx = in->init_mode == CILQR_INIT_GUESS, has_gu = in->init_mode != CILQR_INIT_IQR;
  const size_t in_need = up(b_start * B) + up(b_coarse * B) + up(b_corr * B) + up(b_cnt * B) + up(b_ll * B) + up(b_lr * B) +
                         (has_gx ? up(b_st * B) : 0) + (has_gu ? up(b_ct * B) : 0);
  size_t out_need = up(b_st * B) + up(b_ct * B) + up(b_status * B);
  if (out->trajectory) out_need += up(b_traj * B);
  if (out->result) out_need += up(b_traj * B);
  if (out->init_states) out_need += up(b_st * B);
  if (out->init_controls) out_need += up(b_ct * B);
  if (out->cost_hist) out_need += up(b_ch * B);
  if (out->iter_states) out_need += up(b_is * B);
  if (out->iter_controls) out_need += up(b_ic * B);
  if (out->hist_len) out_need += up(b_hl * B);
  Slot* s = &h->slots[0];
  if (s->in_bytes < in_need) {
    CK(cudaStreamSynchronize(s->stream));
    if (s->in_buf) CK(cudaFree(s->in_buf));
    s->in_buf = nullptr;
    s->in_bytes = 0;
    CK(cudaMalloc(&s->in_buf, in_need));
    s->in_bytes = in_need;
  }
  if (s->out_bytes < out_need) {
    CK(cudaStreamSynchronize(s->stream));
    if (s->out_buf) CK(cudaFree(s->out_buf));
    s->out_buf = nullptr;
    s->out_bytes = 0;
    CK(cudaMalloc(&s->out_buf, out_need));
    s->out_bytes = out_need;
  }
  if (s->ready_cap < n_chunks) {
    if (s->ready_host) CK(cudaFreeHost(s->ready_host));
    s->ready_host = nullptr;
    s->ready_cap = 0;
    CK(cudaHostAlloc((void**)&s->ready_host, sizeof(unsigned int) * n_chunks, cudaHostAllocDefault));
    s->ready_cap = n_chunks;
  }
  // device arrays of the whole batch
  char* p = s->in_buf;
  auto carve_in = [&](size_t per) {
    char* r = p;
    p += up(per * B);
    return r;
  };
  char* d_start = carve_in(b_start);
  char* d_coarse = carve_in(b_coarse);
  char* d_corr = carve_in(b_corr);
  char* d_cnt = carve_in(b_cnt);
  char* d_ll = carve_in(b_ll);
  char* d_lr = carve_in(b_lr);
  char* d_gx = has_gx ? carve_in(b_st) : nullptr;
  char* d_gu = has_gu ? carve_in(b_ct) : nullptr;
  CilqrBatchIn din = *in;
  din.start = (const double*)d_start;
  din.coarse = (const double*)d_coarse;
  din.corridor = (const double*)d_corr;
  din.corridor_cnt = (const int32_t*)d_cnt;
  din.lane_left = (const double*)d_ll;
  din.lane_right = (const double*)d_lr;
  din.init_states = (const double*)d_gx;
  din.init_controls = (const double*)d_gu;
  // Outputs: a caller buffer in pinned (page-locked, device-mapped) host memory is written by the kernel
  // directly as scenarios finish -- the results cross PCIe during the solve and need no D2H pass; any other
  // buffer gets a device staging area and one copy after the kernel.
  char* q = s->out_buf;
  auto zero_copy = [&](void* host) -> void* {
    if (!host || h->no_zero_copy) return nullptr;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, host) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return attr.type == cudaMemoryTypeHost ? attr.devicePointer : nullptr;
  };
  bool staged[11];
  int n_out = 0;
  auto carve = [&](void* host, size_t per) -> char* {
    staged[n_out] = false;
    if (!host) {
      ++n_out;
      return nullptr;
    }
    if (void* d = zero_copy(host)) {
      ++n_out;
      return (char*)d;
    }
    staged[n_out++] = true;
    char* r = q;
    q += up(per * B);
    return r;
  };
  CilqrBatchOut dout;
  memset(&dout, 0, sizeof(dout));
  dout.hist_cap = H;
  dout.states = (double*)carve(out->states, b_st);
  dout.controls = (double*)carve(out->controls, b_ct);
  dout.status = (double*)carve(out->status, b_status);
  dout.trajectory = (double*)carve(out->trajectory, b_traj);
  dout.init_states = (double*)carve(out->init_states, b_st);
  dout.init_controls = (double*)carve(out->init_controls, b_ct);
  dout.cost_hist = (double*)carve(out->cost_hist, b_ch);
  dout.iter_states = (double*)carve(out->iter_states, b_is);
  dout.iter_controls = (double*)carve(out->iter_controls, b_ic);
  dout.hist_len = (int32_t*)carve(out->hist_len, b_hl);
  dout.result = (double*)carve(out->result, b_traj);
  // Order of enqueueing: watermark = 0, then EVERY input chunk with its watermark update on the copy stream,
  // and only then the kernel on the solve stream.  The kernel therefore never depends on work that is
  // enqueued after its launch: when launches are synchronous (ncu, CUDA_LAUNCH_BLOCKING=1, a debugger) the
  // copy stream still runs to completion underneath it.  With pinned inputs the copies are asynchronous, the
  // kernel starts as soon as the first chunk has landed and the rest of the transfer hides behind the solve;
  // with pageable inputs cudaMemcpyAsync stages synchronously, so the call degrades to copy-then-solve.
  CK(cudaMemsetAsync(s->ready, 0, sizeof(unsigned int), s->copy_stream));
  CK(cudaEventRecord(s->ev_reset, s->copy_stream));
  CK(cudaStreamWaitEvent(s->stream, s->ev_reset, 0));
  cudaError_t ce = cudaSuccess;
  for (int ci = 0; ci < n_chunks && ce == cudaSuccess; ++ci) {
    const size_t b0 = ci ? chunk_end[ci - 1] : 0, nb = chunk_end[ci] - b0;
    auto h2d = [&](char* dev, const void* src, size_t per) {
      if (ce == cudaSuccess)
        ce = cudaMemcpyAsync(dev + per * b0, (const char*)src + per * b0, per * nb, cudaMemcpyHostToDevice, s->copy_stream);
    };
    h2d(d_start, in->start, b_start);
    h2d(d_coarse, in->coarse, b_coarse);
    h2d(d_corr, in->corridor, b_corr);
    h2d(d_cnt, in->corridor_cnt, b_cnt);
    h2d(d_ll, in->lane_left, b_ll);
    h2d(d_lr, in->lane_right, b_lr);
    if (d_gx) h2d(d_gx, in->init_states, b_st);
    if (d_gu) h2d(d_gu, in->init_controls, b_ct);
    size_t mark = b0 + nb;
    if (h->starve_after >= 0) mark = std::min<size_t>(mark, (size_t)h->starve_after);  // test hook
    s->ready_host[ci] = (unsigned int)mark;
    if (ce == cudaSuccess)
      ce = cudaMemcpyAsync(s->ready, &s->ready_host[ci], sizeof(unsigned int), cudaMemcpyHostToDevice, s->copy_stream);
  }
  if (ce != cudaSuccess) {
    cudaStreamSynchronize(s->copy_stream);
    return fail_cuda(h, ce, "host-to-device copy");
  }
  if (dout.cost_hist) CK(cudaMemsetAsync(dout.cost_hist, 0, b_ch * B, s->stream));
  rc = launch_solve(h, s, s->stream, &din, &dout, nullptr, s->ready);
  if (rc != CILQR_OK) {
    cudaStreamSynchronize(s->copy_stream);
    return rc;
  }
  CK(cudaMemcpyAsync(s->flags_host, s->ticket, 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost, s->stream));
  int i_out = 0;
  auto d2h = [&](void* dst, const void* dev, size_t per) -> cudaError_t {
    const bool need = staged[i_out++];
    if (!dst || !need) return cudaSuccess;
    return cudaMemcpyAsync(dst, dev, per * B, cudaMemcpyDeviceToHost, s->stream);
  };
  CK(d2h(out->states, dout.states, b_st));
  CK(d2h(out->controls, dout.controls, b_ct));
  CK(d2h(out->status, dout.status, b_status));
  CK(d2h(out->trajectory, dout.trajectory, b_traj));
  CK(d2h(out->init_states, dout.init_states, b_st));
  CK(d2h(out->init_controls, dout.init_controls, b_ct));
  CK(d2h(out->cost_hist, dout.cost_hist, b_ch));
  CK(d2h(out->iter_states, dout.iter_states, b_is));
  CK(d2h(out->iter_controls, dout.iter_controls, b_ic));
  CK(d2h(out->hist_len, dout.hist_len, b_hl));
  CK(d2h(out->result, dout.result, b_traj));
  CK(cudaStreamSynchronize(s->copy_stream));
  CK(cudaStreamSynchronize(s->stream));
  return check_launch(h, s, s->flags_host);
}

int cilqr_debug_completion_histogram(cilqr_handle* h, uint64_t out[256]) {
  if (!h || !out || !h->timed) return CILQR_E_INVALID;
  Slot* s = &h->slots[h->last_slot];
  CK(cudaSetDevice(h->device));
  CK(cudaEventSynchronize(s->ev1));
  CK(cudaMemcpy(out, s->stats + 10, 256 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  if (getenv("CILQR_HOT_TIMING_DUMP")) {  // development: cycles / calls per phase type of hot contexts
    uint64_t t[12];
    CK(cudaMemcpy(t, s->stats + 266, 12 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    for (int p = 0; p < 6; ++p)
      fprintf(stderr, "hot phase %d: %llu calls, %.1f us each\n", p, (unsigned long long)t[6 + p],
              t[6 + p] ? (double)t[p] / t[6 + p] / 1965.0 : 0.0);
  }
  return CILQR_OK;
}

int cilqr_debug_stats(cilqr_handle* h, uint64_t out[8]) {
  if (!h || !out || !h->timed) return CILQR_E_INVALID;
  Slot* s = &h->slots[h->last_slot];
  CK(cudaSetDevice(h->device));
  CK(cudaEventSynchronize(s->ev1));
  CK(cudaMemcpy(out, s->stats, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return CILQR_OK;
}

int cilqr_kernel_launches(const cilqr_handle* h, int64_t* solve_launches) {
  if (!h || !solve_launches) return CILQR_E_INVALID;
  *solve_launches = h->launches;
  return CILQR_OK;
}

int cilqr_last_kernel_ms(cilqr_handle* h, float* ms) {
  if (!h || !ms || !h->timed) return CILQR_E_INVALID;
  Slot* s = &h->slots[h->last_slot];
  CK(cudaEventSynchronize(s->ev1));
  CK(cudaEventElapsedTime(ms, s->ev0, s->ev1));
  return CILQR_OK;
}

int cilqr_occupancy(const cilqr_handle* h, int N, int S_left, int S_right, int* warps_per_sm,
                    int* smem_bytes_per_warp) {
  if (!h) return CILQR_E_INVALID;
  Launch L;
  int rc = plan_launch(const_cast<cilqr_handle*>(h), 1 << 30, N, h->M_max, S_left, S_right, &L);
  if (rc != CILQR_OK) return rc;
  if (warps_per_sm) *warps_per_sm = L.warps;
  if (smem_bytes_per_warp) *smem_bytes_per_warp = L.sm.total_bytes;
  return CILQR_OK;
}

// ---- corridor builder -------------------------------------------------------------------------------
void cilqr_corridor_default_config(CilqrCorridorConfig* c) {
  if (!c) return;
  c->max_diff_x = 25.0;  // planner_config.h:77-85
  c->max_diff_y = 25.0;
  c->radius = 150.0;
  c->max_axis_x = 10.0;
  c->max_axis_y = 10.0;
  c->lane_segment_length = 5.0;
  c->point_cap = 0;
}

static int corridor_launch(cilqr_handle* h, const CilqrCorridorConfig* cfg, const CilqrCorridorIn* in,
                           const CilqrCorridorOut* out, cudaStream_t st) {
  corridor::Args a;
  a.B = in->B;
  a.K = in->K;
  a.P_max = in->P_max;
  a.M_max = in->M_max;
  a.cap = (cfg->point_cap > 0 ? cfg->point_cap : 64) + 1;  // + the knot itself (flipData's zero slot)
  a.max_diff_x = cfg->max_diff_x;
  a.max_diff_y = cfg->max_diff_y;
  a.radius = cfg->radius;
  a.max_axis_x = cfg->max_axis_x;
  a.max_axis_y = cfg->max_axis_y;
  a.traj = in->traj;
  a.obs_points = in->obs_points;
  a.obs_cnt = in->obs_cnt;
  a.corridor = out->corridor;
  a.corridor_cnt = out->corridor_cnt;
  a.polygon = out->polygon;
  a.code = out->code;
  // threads per CTA: as many warps as the per-thread shared-memory slice allows, 1 CTA per SM above
  // 113 KB, 2 below
  const size_t per_thread = corridor::smem_bytes_per_thread(a.cap);
  int threads = (int)((size_t)(h->smem_optin - 1024) / per_thread) / 32 * 32;
  if (threads > 256) threads = 256;
  if (threads < 32) return CILQR_E_SMEM;
  const size_t smem = per_thread * threads;
  static_assert(sizeof(float) == 4, "");
  const long long items = (long long)in->B * in->K;
  int per_sm = 1;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, corridor::corridor_build_kernel, threads, smem));
  if (per_sm < 1) per_sm = 1;
  long long blocks = (items + threads - 1) / threads;
  const long long resident = (long long)h->num_sms * per_sm;
  if (blocks > resident) blocks = resident;  // persistent: grid-stride over the knots
  if (blocks < 1) blocks = 1;
  if (!h->corr_ev0) {
    CK(cudaEventCreate(&h->corr_ev0));
    CK(cudaEventCreate(&h->corr_ev1));
  }
  CK(cudaEventRecord(h->corr_ev0, st));
  corridor::corridor_build_kernel<<<(unsigned)blocks, threads, smem, st>>>(a);
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->corr_ev1, st));
  h->corr_timed = true;
  h->corr_launches++;
  return CILQR_OK;
}

static int corridor_validate(const cilqr_handle* h, const CilqrCorridorConfig* cfg, const CilqrCorridorIn* in,
                             const CilqrCorridorOut* out) {
  if (!h || !cfg || !in || !out) return CILQR_E_INVALID;
  if (in->B < 0 || in->K < 1 || in->P_max < 0 || in->M_max < 1) return CILQR_E_INVALID;
  if (in->P_max + 8 > 250 || cfg->point_cap < 0 || cfg->point_cap > 250) return CILQR_E_CAPACITY;
  if (in->B > 0 && (!in->traj || !in->obs_cnt || (in->P_max > 0 && !in->obs_points) || !out->corridor ||
                    !out->corridor_cnt || !out->code))
    return CILQR_E_INVALID;
  return CILQR_OK;
}

int cilqr_corridor_batch_device(cilqr_handle* h, const CilqrCorridorConfig* cfg, const CilqrCorridorIn* in,
                                const CilqrCorridorOut* out, void* cuda_stream) {
  int rc = corridor_validate(h, cfg, in, out);
  if (rc != CILQR_OK) return rc;
  if (in->B == 0) return CILQR_OK;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->slots[0].stream;
  return corridor_launch(h, cfg, in, out, st);
}

static int corr_ensure(cilqr_handle* h, size_t bytes) {
  if (h->corr_bytes >= bytes) return CILQR_OK;
  CK(cudaStreamSynchronize(h->slots[0].stream));  // nothing enqueued may still read the old buffer
  if (h->corr_buf) cudaFree(h->corr_buf);
  h->corr_buf = nullptr;
  h->corr_bytes = 0;
  CK(cudaMalloc(&h->corr_buf, bytes));
  h->corr_bytes = bytes;
  return CILQR_OK;
}

int cilqr_corridor_batch(cilqr_handle* h, const CilqrCorridorConfig* cfg, const CilqrCorridorIn* in,
                         const CilqrCorridorOut* out) {
  int rc = corridor_validate(h, cfg, in, out);
  if (rc != CILQR_OK) return rc;
  if (in->B == 0) return CILQR_OK;
  CK(cudaSetDevice(h->device));
  const size_t items = (size_t)in->B * in->K;
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t b_traj = up(items * 3 * 8), b_pts = up(items * in->P_max * 2 * 8), b_cnt = up(items * 4);
  const size_t b_cor = up(items * in->M_max * 3 * 8), b_ccnt = up(items * 4);
  const size_t b_poly = out->polygon ? up(items * in->M_max * 2 * 8) : 0, b_code = up(items * 4);
  rc = corr_ensure(h, b_traj + b_pts + b_cnt + b_cor + b_ccnt + b_poly + b_code);
  if (rc != CILQR_OK) return rc;
  char* p = h->corr_buf;
  CilqrCorridorIn din = *in;
  CilqrCorridorOut dout;
  din.traj = (const double*)p; p += b_traj;
  din.obs_points = (const double*)p; p += b_pts;
  din.obs_cnt = (const int32_t*)p; p += b_cnt;
  dout.corridor = (double*)p; p += b_cor;
  dout.corridor_cnt = (int32_t*)p; p += b_ccnt;
  dout.polygon = out->polygon ? (double*)p : nullptr; p += b_poly;
  dout.code = (int32_t*)p;
  cudaStream_t st = h->slots[0].stream;
  CK(cudaMemcpyAsync((void*)din.traj, in->traj, items * 3 * 8, cudaMemcpyHostToDevice, st));
  if (in->P_max > 0)
    CK(cudaMemcpyAsync((void*)din.obs_points, in->obs_points, items * in->P_max * 2 * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync((void*)din.obs_cnt, in->obs_cnt, items * 4, cudaMemcpyHostToDevice, st));
  // slots beyond corridor_cnt are not written by the kernel: return them as zeros, not as stale staging data
  CK(cudaMemsetAsync(dout.corridor, 0, b_cor + b_ccnt + b_poly + b_code, st));
  rc = corridor_launch(h, cfg, &din, &dout, st);
  if (rc != CILQR_OK) return rc;
  CK(cudaMemcpyAsync(out->corridor, dout.corridor, items * in->M_max * 3 * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(out->corridor_cnt, dout.corridor_cnt, items * 4, cudaMemcpyDeviceToHost, st));
  if (out->polygon)
    CK(cudaMemcpyAsync(out->polygon, dout.polygon, items * in->M_max * 2 * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(out->code, dout.code, items * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return CILQR_OK;
}

int cilqr_lane_constraints_device(cilqr_handle* h, const CilqrCorridorConfig* cfg, int B, int n, int S_max,
                                  int is_left, const double* boundary, double* out, int32_t* count,
                                  void* cuda_stream) {
  if (!h || !cfg || B < 0 || n < 1 || S_max < 1) return CILQR_E_INVALID;
  if (B == 0) return CILQR_OK;
  if (!boundary || !out || !count) return CILQR_E_INVALID;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->slots[0].stream;
  corridor::LaneArgs a;
  a.B = B;
  a.n = n;
  a.S_max = S_max;
  a.is_left = is_left ? 1 : 0;
  a.seg_len = cfg->lane_segment_length;
  a.boundary = boundary;
  a.out = out;
  a.count = count;
  const int threads = 128;
  const unsigned blocks = (unsigned)(((long long)B * 32 + threads - 1) / threads);
  corridor::lane_constraints_kernel<<<blocks, threads, 0, st>>>(a);
  CK(cudaGetLastError());
  return CILQR_OK;
}

int cilqr_lane_constraints(cilqr_handle* h, const CilqrCorridorConfig* cfg, int B, int n, int S_max, int is_left,
                           const double* boundary, double* out, int32_t* count) {
  if (!h || !cfg || B < 0 || n < 1 || S_max < 1) return CILQR_E_INVALID;
  if (B == 0) return CILQR_OK;
  if (!boundary || !out || !count) return CILQR_E_INVALID;
  CK(cudaSetDevice(h->device));
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t b_in = up((size_t)B * n * 2 * 8), b_out = up((size_t)B * S_max * 7 * 8), b_cnt = up((size_t)B * 4);
  int rc = corr_ensure(h, b_in + b_out + b_cnt);
  if (rc != CILQR_OK) return rc;
  char* p = h->corr_buf;
  double* d_in = (double*)p;
  double* d_out = (double*)(p + b_in);
  int32_t* d_cnt = (int32_t*)(p + b_in + b_out);
  cudaStream_t st = h->slots[0].stream;
  CK(cudaMemcpyAsync(d_in, boundary, (size_t)B * n * 2 * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(d_out, 0, (size_t)B * S_max * 7 * 8, st));
  rc = cilqr_lane_constraints_device(h, cfg, B, n, S_max, is_left, d_in, d_out, d_cnt, st);
  if (rc != CILQR_OK) return rc;
  CK(cudaMemcpyAsync(out, d_out, (size_t)B * S_max * 7 * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(count, d_cnt, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return CILQR_OK;
}

int cilqr_corridor_last_kernel_ms(cilqr_handle* h, float* ms) {
  if (!h || !ms || !h->corr_timed) return CILQR_E_INVALID;
  CK(cudaSetDevice(h->device));
  CK(cudaEventSynchronize(h->corr_ev1));
  CK(cudaEventElapsedTime(ms, h->corr_ev0, h->corr_ev1));
  return CILQR_OK;
}

// ---- cilqr_internal.h -------------------------------------------------------------------------------
int cilqr_internal_device(const cilqr_handle* h) { return h->device; }
cudaStream_t cilqr_internal_stream(cilqr_handle* h) { return h->slots[0].stream; }
int cilqr_internal_num_sms(const cilqr_handle* h) { return h->num_sms; }
int cilqr_internal_smem_optin(const cilqr_handle* h) { return h->smem_optin; }
int cilqr_internal_fail(cilqr_handle* h, cudaError_t e, const char* where) { return fail_cuda(h, e, where); }
int cilqr_internal_scratch(cilqr_handle* h, int slot, size_t bytes, char** out) {
  if (h->aux_bytes[slot] < bytes) {
    if (h->aux_buf[slot]) cudaFree(h->aux_buf[slot]);
    h->aux_buf[slot] = nullptr;
    h->aux_bytes[slot] = 0;
    CK(cudaMalloc(&h->aux_buf[slot], bytes));
    h->aux_bytes[slot] = bytes;
  }
  *out = h->aux_buf[slot];
  return CILQR_OK;
}
int cilqr_internal_events(cilqr_handle* h, int slot, cudaEvent_t* e0, cudaEvent_t* e1) {
  for (int j = 0; j < 2; ++j)
    if (!h->aux_ev[slot][j]) CK(cudaEventCreate(&h->aux_ev[slot][j]));
  *e0 = h->aux_ev[slot][0];
  *e1 = h->aux_ev[slot][1];
  return CILQR_OK;
}

}  // extern "C"


