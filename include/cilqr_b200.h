/*
 * cilqr_b200.h -- C ABI of the B200-native batched constrained-iLQR solver.
 *
 * Drop-in boundary for the hot path of mpt0816/Cilqr:
 *     planning::IlqrOptimizer::Plan(start_state, coarse_traj, corridor, left_lane_cons,
 *                                   right_lane_cons, opt_trajectory*, iter_trajs*)
 *     (reference: algorithm/ilqr/ilqr_optimizer.h:41-48, algorithm/ilqr/ilqr_optimizer.cc:53-95)
 * The reference has no FFI layer; the seam is that C++ class.  include/cilqr/ilqr_optimizer_b200.h
 * provides a header-compatible planning::IlqrOptimizer that packs its arguments into the flat
 * arrays below (B = 1) and calls this ABI.  The same ABI solves B independent scenarios per call.
 *
 * Plain pointers and sizes only; no C++/Eigen/torch types.  All functions return 0 on success or a
 * negative CILQR_E_* code and never throw.  A handle owns its device buffers and CUDA streams; it
 * is thread-compatible (distinct handles may be used from distinct threads) but not re-entrant.
 *
 * Wire format (all floating point is IEEE double -- the reference's arithmetic type; scenario
 * major, C order, K = N + 1 knots):
 *   start        [B][4]              x, y, theta, v            trajectory_planner.cpp:73-75
 *   coarse       [B][K][6]           x, y, theta, v, a, delta  fields TransformGoals reads,
 *                                                              ilqr_optimizer.cc:141-152
 *   corridor     [B][K][M_max][3]    raw half-planes (a,b,c), a*x + b*y < c   corridor.h:20-22
 *   corridor_cnt [B][K] int32        planes used at each knot (slots >= cnt are never read)
 *   lane_left    [B][S_left][7]      a, b, c, x0, y0, x1, y1: half-plane + LineSegment2d
 *   lane_right   [B][S_right][7]     start/end as constructed in corridor.cc:279,300
 *   states       [B][K][6]   out     x, y, theta, v, a, delta
 *   controls     [B][N][2]   out     jerk, delta_rate
 *   status       [B][8]      out     see CILQR_ST_* below
 *   trajectory   [B][K][13]  out     TrajectoryPoint records (discretized_trajectory.h:26-43) as
 *                                    TransformToTrajectory fills them, ilqr_optimizer.cc:771-791
 *   result       [B][K][13]  out     the same records after TrajectoryPlanner::Plan's post-processing
 *                                    (station accumulated, trajectory_planner.cpp:103-125)
 * Shrinking/normalising the constraints (ilqr_optimizer.cc:438-495) is part of the solve.
 *
 * Two further sections below widen the boundary to the stages that feed the solve inside
 * TrajectoryPlanner::Plan (trajectory_planner.cpp:28-86), each with the same conventions and with
 * outputs laid out as the next stage's inputs, so the three chain on the device:
 *     DpPlanner::Plan   -> cilqr_dp_plan_batch(_device)        (drop-in: cilqr/dp_planner_b200.h)
 *     Corridor::Plan    -> cilqr_corridor_batch(_device), cilqr_lane_constraints(_device)
 *                                                              (drop-in: cilqr/corridor_b200.h)
 */
#ifndef CILQR_B200_H_
#define CILQR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CILQR_ABI_VERSION 4

/* error codes */
#define CILQR_OK 0
#define CILQR_E_INVALID (-1)   /* null pointer / bad size: the reference's guards, ilqr_optimizer.cc:64-78 */
#define CILQR_E_CUDA (-2)      /* CUDA runtime error (see cilqr_last_cuda_error) */
#define CILQR_E_NO_DEVICE (-3) /* no sm_100 device: there is no CPU fallback */
#define CILQR_E_CAPACITY (-4)  /* batch/horizon exceeds what the handle was created for */
#define CILQR_E_SMEM (-5)      /* horizon does not fit the per-warp shared-memory stage */
#define CILQR_E_TIMEOUT (-6)   /* the solve kernel gave up on some scenarios (input transfer starved / watchdog); \
                                  their status rows keep the NaN sentinel -- never returned as success, cf. \
                                  trajectory_planner.cpp:91-94 */

#define CILQR_E_NCCL (-7)      /* the gathered copy of cilqr_plan_sharded was requested but NCCL failed / is missing */

/* status[b][CILQR_ST_FLAG] values: which exit of Optimize() was taken */
#define CILQR_CONVERGED_ABS 0    /* dcost < abs_cost_tol         ilqr_optimizer.cc:281,287 */
#define CILQR_CONVERGED_REL 1    /* dcost/cost_old < rel_tol     ilqr_optimizer.cc:282,289 */
#define CILQR_CONVERGED_GRAD 2   /* gradient norm exit           ilqr_optimizer.cc:236-241 */
#define CILQR_LAMBDA_OVERFLOW 3  /* "kUnsolved"                  ilqr_optimizer.cc:302-307 */
#define CILQR_MAX_ITER 4         /* loop exhausted               ilqr_optimizer.cc:312-319 */

/* layout of one status record (8 doubles).  Every row is pre-filled with all-ones bytes (a NaN) before the
 * launch: a scenario the kernel did not finish is recognisable by status[b][CILQR_ST_FLAG] != itself. */
#define CILQR_ST_FLAG 0       /* one of the values above */
#define CILQR_ST_ITERS 1      /* loop index `iter` at exit */
#define CILQR_ST_COST 2       /* total cost of the returned trajectory ... */
#define CILQR_ST_COST_TARGET 3
#define CILQR_ST_COST_DYNAMIC 4
#define CILQR_ST_COST_CORRIDOR 5
#define CILQR_ST_COST_LANE 6  /* ... and its breakdown, struct Cost ilqr_optimizer.h:14-27 */
#define CILQR_ST_ALPHA_HASH 7 /* FNV-1a over the accepted line-search index per iteration (11 = rejected) */
#define CILQR_STATUS_DOUBLES 8
#define CILQR_TRAJPOINT_DOUBLES 13

/* POD mirror of VehicleParam (vehicle_param.h:21-64), IlqrConfig/Weights (planner_config.h:45-73),
 * the RelaxBarrierFunction members (barrier_function.h:143-146) and delta_t (planner_config.h:94). */
typedef struct CilqrParams {
  double front_hang_length, wheel_base, rear_hang_length, width;
  double max_velocity, min_acceleration, max_acceleration;
  double jerk_min, jerk_max, delta_min, delta_max, delta_rate_min, delta_rate_max;
  double safe_margin;
  double w_jerk, w_delta_rate, w_x_target, w_y_target, w_theta, w_v, w_a, w_delta;
  double abs_cost_tol, rel_cost_tol;
  double barrier_t, barrier_eps;
  double delta_t;
  int32_t num_of_disc; /* must be 5 (planner_config.h:58); other values -> CILQR_E_INVALID */
  int32_t max_iter_num;
} CilqrParams;

typedef struct CilqrBatchIn {
  int32_t B, N, M_max, S_left, S_right;
  const double* start;
  const double* coarse;
  const double* corridor;
  const int32_t* corridor_cnt;
  const double* lane_left;
  const double* lane_right;
  /* Initial guess of Optimize (ilqr_optimizer.cc:168-169).  CILQR_INIT_IQR (0, default): the live line :169, the
   * LQR tracking guess `iqr` (:793-842).  CILQR_INIT_OPEN_LOOP: roll init_controls [B][N][2] out from the start
   * state (OpenLoopRollout, algorithm/slover/ilqr.h:362-370).  CILQR_INIT_GUESS: init_states [B][K][6] and
   * init_controls [B][N][2] as given -- what InitGuess (:107-139; the commented-out alternative on line :168)
   * copies out of the Tracker's trajectory; the tracker itself (algorithm/ilqr/tracker.cc) stays with the caller. */
  int32_t init_mode;
  const double* init_states;
  const double* init_controls;
} CilqrBatchIn;
#define CILQR_INIT_IQR 0
#define CILQR_INIT_OPEN_LOOP 1
#define CILQR_INIT_GUESS 2

typedef struct CilqrBatchOut {
  double* states;        /* required */
  double* controls;      /* required */
  double* status;        /* required */
  double* trajectory;    /* optional: [B][K][13] */
  double* init_states;   /* optional: [B][K][6]  the initial guess = iter_trajs[0], ilqr_optimizer.cc:170 */
  double* init_controls; /* optional: [B][N][2] */
  /* optional per-accept history (debug / adapter use, small B): cost_hist [B][hist_cap][5] mirrors
   * cost_ (ilqr_optimizer.h:50-52, pushes at ilqr_optimizer.cc:173,283,296); iter_states
   * [B][hist_cap][K][6] and iter_controls [B][hist_cap][N][2] mirror iter_trajs (:170,294);
   * hist_len [B][2] int32 = {entries pushed to cost_, entries pushed to iter_trajs} */
  double* cost_hist;
  double* iter_states;
  double* iter_controls;
  int32_t* hist_len;
  int32_t hist_cap;
  /* optional: [B][K][13] the records TrajectoryPlanner::Plan builds from opt_trajectory AFTER the solve
   * (trajectory_planner.cpp:103-125): like `trajectory`, with s = accumulated hypot of consecutive (x, y)
   * (sequential sum, :110-112) and kappa = tan(delta) / wheel_base (:118); time = delta_t * k.  This is the
   * planner's published result (PlanningNode::PlanCallback, planning_node.cc:82-88). */
  double* result;
} CilqrBatchOut;

typedef struct cilqr_handle cilqr_handle;

/* Defaults of the reference's parameter structs. */
void cilqr_default_params(CilqrParams* p);

/* Creates a solver bound to CUDA device `device` for horizons up to N_max steps, up to M_max planes
 * per knot, up to S_max lane segments per side and host batches up to B_max scenarios.
 * Replaces IlqrOptimizer::IlqrOptimizer / Init (ilqr_optimizer.cc:13-51). */
int cilqr_create(const CilqrParams* params, int device, int N_max, int M_max, int S_max, int B_max,
                 cilqr_handle** out);
void cilqr_destroy(cilqr_handle* h);

/* Solve B scenarios given HOST pointers: chunks the batch, overlaps H2D / solve / D2H on separate
 * streams, blocks until the outputs are in host memory.  Replaces IlqrOptimizer::Plan. */
int cilqr_plan_batch(cilqr_handle* h, const CilqrBatchIn* in, const CilqrBatchOut* out);

/* Same with DEVICE pointers; enqueues on `cuda_stream` (a cudaStream_t passed as void*, NULL =
 * the handle's own stream) and returns without synchronising. */
int cilqr_plan_batch_device(cilqr_handle* h, const CilqrBatchIn* in, const CilqrBatchOut* out,
                            void* cuda_stream);
/* Waits for the handle's streams and checks the last launch: CILQR_E_TIMEOUT when it left scenarios
 * unsolved.  After cilqr_plan_batch_device on a caller stream, synchronise that stream first. */
int cilqr_synchronize(cilqr_handle* h);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU: the batch of independent scenarios sharded over the GPUs of one box from ONE host
 * process (the reference's host is one C++ process, planning_node.cc).  Device r of G solves the contiguous ids
 * [r * per, (r + 1) * per), per = cilqr_multi_shard_size(B) = ceil(B / G): no exchange inside the solve.
 * cilqr_plan_sharded takes HOST pointers for the whole batch (states / controls / status and optionally `result`;
 * the other optional outputs must be NULL) and blocks until they are filled.  gathered_dev, when not NULL, is an
 * array of G DEVICE pointers, one per GPU, each with room for G * per * (6K + 2N + 8) doubles: after the solve ONE
 * ncclAllGather over NVLink leaves every shard's result block [states per x K x 6 | controls per x N x 2 |
 * status per x 8] on every GPU (block r at offset r * per * (6K + 2N + 8); rows beyond B are zero).  NCCL is
 * loaded at run time (libnccl.so.2); without it only the gathered copy is unavailable (CILQR_E_NCCL). */
typedef struct cilqr_multi cilqr_multi;
int cilqr_multi_create(const CilqrParams* params, int n_devices, const int* devices /* NULL: 0..n-1 */, int N_max,
                       int M_max, int S_max, int B_max_per_device, cilqr_multi** out);
void cilqr_multi_destroy(cilqr_multi* m);
int cilqr_multi_devices(const cilqr_multi* m);
int cilqr_multi_shard_size(const cilqr_multi* m, int B);
const char* cilqr_multi_last_error(const cilqr_multi* m);
int cilqr_plan_sharded(cilqr_multi* m, const CilqrBatchIn* in, const CilqrBatchOut* out, double* const* gathered_dev);

/* Introspection used by bench.py / tests. */
int cilqr_kernel_launches(const cilqr_handle* h, int64_t* solve_launches);
int cilqr_last_kernel_ms(cilqr_handle* h, float* ms);      /* CUDA-event time of the last solve kernel */
int cilqr_occupancy(const cilqr_handle* h, int N, int S_left, int S_right, int* warps_per_sm,
                    int* smem_bytes_per_warp);
/* Scheduler counters of the last solve launch (development aid): scheduler passes, idle polls, failed
 * context claims, phases run {INIT, BACK, ROLL, EVAL}, phase-type switches -- summed over all warps. */
int cilqr_debug_stats(cilqr_handle* h, uint64_t out[8]);
/* Scenarios completed per 2 ms bucket since the start of the last solve launch (development aid). */
int cilqr_debug_completion_histogram(cilqr_handle* h, uint64_t out[256]);
const char* cilqr_strerror(int code);
const char* cilqr_last_cuda_error(const cilqr_handle* h);
int cilqr_abi_version(void);

/* Stage-level dump of the first iteration (test hook; device pointers, B scenarios):
 *   constraints [B][K][M_max][3] + lanes [B][S_left+S_right][3]  shrunk+normalised (a4)
 *   init X/U, cost5 of the initial guess, linearisation at the initial guess
 *   (A11 [B][N][12]: A02,A03,A04,A05,A12,A13,A14,A15,A23,A24,A25,B21; Jx [B][K][6]; Ju [B][N][2];
 *   Hx [B][K][9]: H00,H01,H02,H11,H12,H22,H33,H44,H55; Hu [B][N][2]),
 *   gains of Backward(lambda=1) (Kg [B][N][12], kg [B][N][2], dV [B][2]) and Forward(alpha=1)
 *   (Xn [B][K][6], Un [B][N][2], cost5n [B][5]).  Any pointer may be NULL. */
typedef struct CilqrDebugOut {
  double *corridor, *lanes, *X0, *U0, *cost0, *A11, *Jx, *Ju, *Hx, *Hu, *Kg, *kg, *dV, *Xn, *Un, *costn;
  int32_t* nearest; /* [B][K][5][2] nearest lane segment per disc/side at the initial guess */
  double* gnorm;    /* [B] CalGradientNorm (ilqr_optimizer.cc:322-332) of the first Backward */
} CilqrDebugOut;
int cilqr_debug_first_iteration(cilqr_handle* h, const CilqrBatchIn* in_dev, const CilqrDebugOut* out_dev);
/* Test hook of the host path: watchdog_ms (> 0) = how long the kernel waits for an input chunk before it
 * flags the launch; starve_after >= 0 = never raise the input watermark beyond that many scenarios (the
 * rest of the batch then times out and cilqr_plan_batch must return CILQR_E_TIMEOUT); -1 = normal. */
int cilqr_debug_host_path(cilqr_handle* h, int watchdog_ms, int starve_after);

/* ------------------------------------------------------------------------------------------------
 * Safe-corridor builder: the step immediately before the solve.
 * Replaces Corridor::Plan (algorithm/ilqr/corridor.h:31-36, corridor.cc:17-54) for B trajectories of
 * K knots at once: BuildCorridorConstraints (:56-87) = per knot AddCorridorPoints (:89-120) +
 * BuildCorridor (:122-263, three cv::convexHull calls on CV_32F points), and
 * CalLeft/RightLaneConstraints (:265-307) with LaneBoundarySample (:309-322) and HalfPlaneConstraint
 * (:324-331).  The environment queries (Environment::QueryStatic/DynamicObstaclesPoints,
 * utils/environment.cpp:163-194) stay with the caller: it passes, per knot, the obstacle points it
 * would hand to BuildCorridor (static obstacle points first, then the dynamic ones at pt.time).
 * The outputs have exactly the layout cilqr_plan_batch(_device) reads (corridor / corridor_cnt /
 * lane_left / lane_right), so on the device path the two calls chain without touching the host.
 *
 *   traj        [B][K][3]            x, y, theta of the coarse trajectory (TrajectoryPoint fields the
 *                                    corridor reads, corridor.cc:76-79,96-97)
 *   obs_points  [B][K][P_max][2]     obstacle points at knot k; obs_cnt [B][K] int32 of them are valid
 *   corridor    [B][K][M_max][3] out half-planes (a,b,c), a*x + b*y < c, un-normalised polygon edges
 *   corridor_cnt[B][K] int32     out
 *   polygon     [B][K][M_max][2] out (optional) ConvexPolygons vertices, corridor.cc:245-250
 *   code        [B][K] int32     out CILQR_CORR_* per knot; Corridor::Plan returns false iff any is != 0
 * A knot with a non-zero code has corridor_cnt = 0.
 */
#define CILQR_CORR_OK 0
#define CILQR_CORR_NO_POINTS 1      /* corridor.cc:127-130 */
#define CILQR_CORR_FEW_POINTS 2     /* fewer than 4 flipped points, corridor.cc:179-182 */
#define CILQR_CORR_ORIGIN 3         /* knot on the flipped hull: undefined in the reference (:193-209) */
#define CILQR_CORR_CAPACITY 4       /* more than M_max planes at a knot */
#define CILQR_CORR_POINT_CAPACITY 5 /* more points inside the +-max_diff window than point_cap */

/* CorridorConfig, algorithm/params/planner_config.h:75-86 (is_multiple_sample = false). */
typedef struct CilqrCorridorConfig {
  double max_diff_x, max_diff_y, radius, max_axis_x, max_axis_y, lane_segment_length;
  int32_t point_cap; /* capacity for the points that pass the +-max_diff filter at one knot (<= 250);
                        sets the shared memory per thread.  0 = default (64). */
} CilqrCorridorConfig;
void cilqr_corridor_default_config(CilqrCorridorConfig* c);

typedef struct CilqrCorridorIn {
  int32_t B, K, P_max, M_max;
  const double* traj;
  const double* obs_points;
  const int32_t* obs_cnt;
} CilqrCorridorIn;

typedef struct CilqrCorridorOut {
  double* corridor;      /* required */
  int32_t* corridor_cnt; /* required */
  double* polygon;       /* optional */
  int32_t* code;         /* required */
} CilqrCorridorOut;

/* HOST pointers; blocks until the outputs are in host memory. */
int cilqr_corridor_batch(cilqr_handle* h, const CilqrCorridorConfig* cfg, const CilqrCorridorIn* in,
                         const CilqrCorridorOut* out);
/* DEVICE pointers; enqueues on `cuda_stream` (NULL = the handle's stream) and returns. */
int cilqr_corridor_batch_device(cilqr_handle* h, const CilqrCorridorConfig* cfg, const CilqrCorridorIn* in,
                                const CilqrCorridorOut* out, void* cuda_stream);

/* Lane constraints of B boundary polylines of n points each (boundary [B][n][2]):
 * out [B][S_max][7] = a,b,c,x0,y0,x1,y1 per segment, count [B] int32 = number of segments, or -1 when
 * fewer than two points were sampled (the reference returns false, corridor.cc:275-277), -2 when more
 * than S_max.  is_left selects CalLeftLaneConstraints' segment direction (corridor.cc:279 vs :300). */
int cilqr_lane_constraints(cilqr_handle* h, const CilqrCorridorConfig* cfg, int B, int n, int S_max, int is_left,
                           const double* boundary, double* out, int32_t* count);
int cilqr_lane_constraints_device(cilqr_handle* h, const CilqrCorridorConfig* cfg, int B, int n, int S_max,
                                  int is_left, const double* boundary, double* out, int32_t* count,
                                  void* cuda_stream);
/* CUDA-event time of the last corridor kernel enqueued through this handle. */
int cilqr_corridor_last_kernel_ms(cilqr_handle* h, float* ms);

/* ------------------------------------------------------------------------------------------------
 * Coarse DP planner: the step before the corridor.
 * Replaces DpPlanner::Plan (algorithm/planner/dp_planner.h:31-37, dp_planner.cpp:135-281) for B scenarios at
 * once: the 5 x 7 x 10 lattice search (GetCost :87-133, GetCollisionCost :40-85, InterpolateLinearly :283-320)
 * with the collision checks of Environment (utils/environment.cpp:51-141) and the trajectory profile
 * (ComputePathProfile, utils/discrete_points_math.cc:27-176).  The Environment is passed as flat arrays:
 *
 *   ref        [R][7]                 s, x, y, theta, kappa, left_bound, right_bound: the centre line
 *                                     (CenterLinePoint.msg -> Environment::set_reference); shared by the batch
 *   barrier    [NB][2]                Environment::road_barrier_ (environment.cpp:24-49), sorted by x; shared
 *   start      [B][3]                 x, y, theta of the planning start (dp_planner.cpp:135-141)
 *   static_poly[B][n_static][V][2]    static obstacle polygons, static_nv [B][n_static] vertices used
 *   dyn_time   [B][n_dyn][T]          sample times of every dynamic obstacle (ascending), dyn_samples [B][n_dyn]
 *   dyn_poly   [B][n_dyn][T][V][2]    its polygon at every sample, dyn_nv [B][n_dyn]
 *   trajectory [B][K][13]  out (opt.) TrajectoryPoint records (time, s, x, y, theta, kappa, velocity, a, jerk,
 *                                     delta, delta_rate, left_bound, right_bound), K = cilqr_dp_num_knots()
 *   coarse     [B][K][6]   out (opt.) x, y, theta, velocity, a, delta  = CilqrBatchIn::coarse
 *   xytheta    [B][K][3]   out (opt.) x, y, theta                      = CilqrCorridorIn::traj
 *   ok         [B] int32   out        DpPlanner::Plan's return value (min_cost < dp_w_obstacle)
 *   cost       [B]         out (opt.) min_cost;  waypoints [B][5][3] out (opt.) s index, l index, current_s
 */
typedef struct CilqrDpConfig { /* PlannerConfig, planner_config.h:88-141 + VehicleParam fields the planner reads */
  double tf, delta_t, dp_nominal_velocity, dp_w_obstacle, dp_w_lateral, dp_w_lateral_change,
      dp_w_lateral_velocity_change, dp_w_longitudinal_velocity_bias, dp_w_longitudinal_velocity_change;
  double max_velocity, width, wheel_base, front_hang_length, rear_hang_length;
} CilqrDpConfig;
void cilqr_dp_default_config(CilqrDpConfig* c);
int cilqr_dp_num_knots(const CilqrDpConfig* c);

typedef struct CilqrDpIn {
  int32_t B, R, NB, V, n_static, n_dyn, T;
  const double* ref;
  const double* barrier;
  const double* start;
  const double* static_poly;
  const int32_t* static_nv;
  const double* dyn_time;
  const int32_t* dyn_samples;
  const double* dyn_poly;
  const int32_t* dyn_nv;
} CilqrDpIn;

typedef struct CilqrDpOut {
  double* trajectory;
  double* coarse;
  double* xytheta;
  int32_t* ok; /* required */
  double* cost;
  double* waypoints;
} CilqrDpOut;

/* DEVICE pointers; enqueues on `cuda_stream` (NULL = the handle's stream) and returns. */
int cilqr_dp_plan_batch_device(cilqr_handle* h, const CilqrDpConfig* cfg, const CilqrDpIn* in, const CilqrDpOut* out,
                               void* cuda_stream);
/* HOST pointers; blocks until the outputs are in host memory. */
int cilqr_dp_plan_batch(cilqr_handle* h, const CilqrDpConfig* cfg, const CilqrDpIn* in, const CilqrDpOut* out);
int cilqr_dp_last_kernel_ms(cilqr_handle* h, float* ms);

/* ------------------------------------------------------------------------------------------------
 * Tracker initial guess: the alternative to iqr that README.md:61 recommends.
 * Replaces Tracker::Plan (algorithm/ilqr/tracker.h:48-51, tracker.cc:11-17,169-215) -- a 10 ms closed-loop simulation
 * along the coarse trajectory with lateral and longitudinal discrete LQR controllers whose gains come from a DARE
 * fixed-point iteration at every step (math::SolveLQRProblem, algorithm/math/linear_quadratic_regulator.cc:30-70) --
 * and the copy IlqrOptimizer::InitGuess (ilqr_optimizer.cc:107-139) makes of its result, for B scenarios at once.
 *
 *   start          [B][4]            x, y, theta, v of start_state (trajectory_planner.cpp:73-75)
 *   coarse_traj    [B][K][13]        TrajectoryPoint records of the coarse trajectory (time, s, x, y, theta, kappa,
 *                                    velocity, ...): exactly cilqr_dp_plan_batch's `trajectory` output
 *   traj           [B][K][13]  out (opt.) the tracker's trajectory (opt_trajectory of Tracker::Plan)
 *   guess_states   [B][K][6]   out (opt.) x, y, theta, velocity, a, delta      = CilqrBatchIn::init_states
 *   guess_controls [B][K-1][2] out (opt.) jerk, delta_rate                     = CilqrBatchIn::init_controls
 *   ok             [B] int32   out        Tracker::Plan's return value (the simulation reached every knot)
 * With init_mode = CILQR_INIT_GUESS the solve starts from (guess_states, guess_controls): DP -> tracker -> solve chain
 * on the device.
 */
typedef struct CilqrTrackerConfig { /* TrackerConfig, planner_config.h:18-43 + the VehicleParam fields the tracker reads */
  double sumulation_dt, dt, tolerance;
  double lat_weight_l, lat_weight_theta, lat_weight_delta, lat_weight_delta_rate, lat_preview_time;
  double lon_weight_s, lon_weight_v, lon_weight_a, lon_weight_j;
  double wheel_base, delta_min, delta_max, min_acceleration, max_acceleration, delta_rate_min, delta_rate_max, jerk_min,
      jerk_max;
  int32_t max_num_iteration;
} CilqrTrackerConfig;
void cilqr_tracker_default_config(CilqrTrackerConfig* c);
/* DEVICE pointers; enqueues on `cuda_stream` (NULL = the handle's stream) and returns. */
int cilqr_tracker_batch_device(cilqr_handle* h, const CilqrTrackerConfig* cfg, int B, int K, const double* start,
                               const double* coarse_traj, double* traj, double* guess_states, double* guess_controls,
                               int32_t* ok, void* cuda_stream);
/* HOST pointers; blocks until the outputs are in host memory. */
int cilqr_tracker_batch(cilqr_handle* h, const CilqrTrackerConfig* cfg, int B, int K, const double* start,
                        const double* coarse_traj, double* traj, double* guess_states, double* guess_controls,
                        int32_t* ok);
int cilqr_tracker_last_kernel_ms(cilqr_handle* h, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* CILQR_B200_H_ */
