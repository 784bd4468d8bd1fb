/*
 * cilqr_oracle.h -- CPU restatement (plain C99, IEEE double) of the CILQR solve path of
 * mpt0816/Cilqr:  IlqrOptimizer::Plan  (algorithm/ilqr/ilqr_optimizer.cc:53-95) and everything it
 * executes per solve.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product path
 * (cilqr_b200/csrc) never links, includes or calls anything in oracle/.
 *
 * PARITY PINNED AGAINST THE REFERENCE'S OWN SOURCE (with one stand-in): Eigen, ROS and OpenCV are not installed,
 * but ilqr_optimizer.cc, vehicle_model.cc and barrier_function.h compile UNMODIFIED against the stand-ins of
 * oracle/ref_stubs (an "Eigen-lite" header reproducing the Eigen 3.4 semantics the solver depends on -- lazy `auto`
 * expressions, coefficient order of products, assignment aliasing rules, closed-form 2x2 inverse --, no-op ROS
 * logging macros, empty OpenCV headers): oracle/Makefile target `_ref`, wrapper oracle/ref_solver_wrapper.cc.
 * tests/test_reference_pins.py runs the reference's IlqrOptimizer::Optimize and this restatement on the same
 * scenarios: the iqr initial guess, the returned states and controls and the five-component cost of every accepted
 * iterate are BIT-IDENTICAL (hundreds of solves, horizons 30-200, both roads, lambda-overflow exits included), and
 * so are Dynamics, DynamicsJacbian, the barrier value, NormalizeAngle and LineSegment2d::DistanceTo.  What the pin
 * cannot cover is Eigen itself: the stand-in's semantics come from knowledge of the Eigen 3.4 sources.
 * Further cross-checks: an independent NumPy restatement (oracle/cilqr_numpy.py) and analytic known-answer tests
 * (tests/test_oracle_*.py).
 *
 * The one deliberate deviation: `iqr` declares R uninitialised and sets only its diagonal
 * (ilqr_optimizer.cc:811-813); the off-diagonals are indeterminate in the reference, 0 here.
 */
#ifndef CILQR_ORACLE_H_
#define CILQR_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define CILQR_ORACLE_NX 6
#define CILQR_ORACLE_NU 2
#define CILQR_ORACLE_NALPHA 11

/* Mirrors VehicleParam (algorithm/params/vehicle_param.h:21-64), IlqrConfig + Weights
 * (algorithm/params/planner_config.h:45-73) and the hard-wired RelaxBarrierFunction members
 * (algorithm/ilqr/barrier_function.h:143-146).  Field order is the wire order used by the ctypes
 * mirror in tests. */
typedef struct cilqr_oracle_params {
  double front_hang_length, wheel_base, rear_hang_length, width;
  double max_velocity, min_acceleration, max_acceleration;
  double jerk_min, jerk_max, delta_min, delta_max, delta_rate_min, delta_rate_max;
  double safe_margin;
  double w_jerk, w_delta_rate, w_x_target, w_y_target, w_theta, w_v, w_a, w_delta;
  double abs_cost_tol, rel_cost_tol;
  double barrier_t, barrier_eps;
  double delta_t;
  int num_of_disc;
  int max_iter_num;
} cilqr_oracle_params;

void cilqr_oracle_default_params(cilqr_oracle_params* p);

/* One scenario in the wire format of include/cilqr_b200.h (doubles, scenario-major). */
typedef struct cilqr_oracle_problem {
  int N;                     /* steps; K = N + 1 knots */
  int M_max;                 /* row pitch of corridor */
  int S_left, S_right;       /* lane segments per side */
  const double* start;       /* [4]  x, y, theta, v                          */
  const double* coarse;      /* [K][6] x, y, theta, v, a, delta              */
  const double* corridor;    /* [K][M_max][3] raw (a,b,c): a x + b y < c     */
  const int* corridor_cnt;   /* [K] planes used at knot k                    */
  const double* lane_left;   /* [S_left][7]  a, b, c, x0, y0, x1, y1         */
  const double* lane_right;  /* [S_right][7]                                 */
  /* initial guess of Optimize (ilqr_optimizer.cc:168-169): 0 = iqr (the live line :169);
   * 1 = open-loop rollout of init_controls from goals_[0] (OpenLoopRollout, algorithm/slover/ilqr.h:362-370);
   * 2 = init_states / init_controls as given -- what InitGuess (:107-139, the commented-out line :168) copies out
   *     of the tracker's trajectory */
  int init_mode;
  const double* init_states;   /* [K][6] (mode 2) */
  const double* init_controls; /* [N][2] (modes 1, 2) */
} cilqr_oracle_problem;

enum {
  CILQR_ORACLE_CONVERGED_ABS = 0,  /* dcost < abs_cost_tol      ilqr_optimizer.cc:281,287 */
  CILQR_ORACLE_CONVERGED_REL = 1,  /* dcost/cost_old < rel tol  ilqr_optimizer.cc:282,289 */
  CILQR_ORACLE_CONVERGED_GRAD = 2, /* gnorm exit                ilqr_optimizer.cc:236-241 */
  CILQR_ORACLE_LAMBDA_OVERFLOW = 3,/* kUnsolved                 ilqr_optimizer.cc:302-307 */
  CILQR_ORACLE_MAX_ITER = 4        /* loop ran out              ilqr_optimizer.cc:312-319 */
};

typedef struct cilqr_oracle_result {
  double* states;        /* [K][6] out */
  double* controls;      /* [N][2] out */
  double* init_states;   /* [K][6] out or NULL (iter_trajs[0]) */
  double* init_controls; /* [N][2] out or NULL */
  int status;
  int iters;             /* value of `iter` at exit (loop index) */
  int accepted;          /* number of accepted forward passes */
  unsigned int alpha_hash; /* FNV-1a over the per-iteration alpha index (11 = all rejected) */
  double cost[5];        /* total,target,dynamic,corridor,lane of the returned trajectory */
  double cost_init[5];
  double lambda;
  /* optional per-iteration trace, 8 doubles per backward pass:
   * iter, alpha_idx (11 = rejected), cost_new, dcost, z, lambda (before update), dV0, dV1 */
  double* trace;
  int trace_cap;
  int trace_len;
  /* optional history of cost_ (ilqr_optimizer.h:50-52): 5 doubles per entry */
  double* cost_hist;
  int cost_hist_cap;
  int cost_hist_len;
} cilqr_oracle_result;

/* Full solve: Plan -> TransformGoals -> Optimize.  Returns 0, or -1 on the reference's guard
 * failures (ilqr_optimizer.cc:64-78). */
int cilqr_oracle_solve(const cilqr_oracle_params* p, const cilqr_oracle_problem* pb,
                       cilqr_oracle_result* out);

/* Batch driver used as the CPU baseline: scenario-major arrays exactly as CilqrBatchIn lays them
 * out; `nthreads` worker threads over disjoint id ranges.  status_out is [B][8] doubles:
 * status, iters, cost[5], alpha_hash. Returns number of scenarios that ended through a success
 * exit (status <= 2). */
int cilqr_oracle_solve_batch(const cilqr_oracle_params* p, int B, int N, int M_max, int S_left,
                             int S_right, const double* start, const double* coarse,
                             const double* corridor, const int* corridor_cnt,
                             const double* lane_left, const double* lane_right, double* states,
                             double* controls, double* status_out, int nthreads);
/* the same with the initial guess of cilqr_oracle_problem::init_mode for every scenario: init_states [B][K][6],
 * init_controls [B][N][2] */
int cilqr_oracle_solve_batch_init(const cilqr_oracle_params* p, int B, int N, int M_max, int S_left,
                                  int S_right, const double* start, const double* coarse,
                                  const double* corridor, const int* corridor_cnt,
                                  const double* lane_left, const double* lane_right, int init_mode,
                                  const double* init_states, const double* init_controls, double* states,
                                  double* controls, double* status_out, int nthreads);

/* ---- primitives, exported for unit tests and stage-level GPU parity ---- */
double cilqr_oracle_normalize_angle(double a);                       /* math_utils.cpp:53-59 */
void cilqr_oracle_dynamics(const cilqr_oracle_params* p, const double x[6], const double u[2],
                           double xn[6]);                            /* vehicle_model.cc:88-121 */
void cilqr_oracle_dynamics_jacobian(const cilqr_oracle_params* p, const double x[6],
                                    const double u[2], double A[36], double B[12]); /* :21-86 */
double cilqr_oracle_barrier_value(const cilqr_oracle_params* p, double g); /* barrier_function.h:104-113 */
double cilqr_oracle_barrier_dcoef(const cilqr_oracle_params* p, double g); /* :115-125 (coefficient of dx) */
void cilqr_oracle_barrier_hcoef(const cilqr_oracle_params* p, double g, double* c_outer,
                                double* c_ddx);                      /* :127-140 */
double cilqr_oracle_segment_distance(const double seg[7], double x, double y); /* line_segment2d.cpp:40-49,61-75 */
double cilqr_oracle_disc_radius(const cilqr_oracle_params* p);       /* ilqr_optimizer.cc:97-104 */

/* Stateful context for stage-level checks (mirrors the members of IlqrOptimizer). */
typedef struct cilqr_oracle_ctx cilqr_oracle_ctx;
cilqr_oracle_ctx* cilqr_oracle_ctx_create(const cilqr_oracle_params* p,
                                          const cilqr_oracle_problem* pb);
void cilqr_oracle_ctx_destroy(cilqr_oracle_ctx* c);
/* shrunk+normalised constraints (ilqr_optimizer.cc:438-495): corridor [K][M_max][3], lanes [S][3] */
void cilqr_oracle_ctx_constraints(const cilqr_oracle_ctx* c, double* corridor, double* lane_left,
                                  double* lane_right);
void cilqr_oracle_ctx_iqr(cilqr_oracle_ctx* c, double* states, double* controls);    /* :793-842 */
double cilqr_oracle_ctx_total_cost(cilqr_oracle_ctx* c, const double* states,
                                   const double* controls, double cost5[5]);         /* :417-436 */
/* linearise + quadratise at (states, controls): fills As[N][36] Bs[N][12] Jx[K][6] Ju[N][2]
 * Hx[K][36] Hu[N][4] (any may be NULL)                                      :203-214 */
void cilqr_oracle_ctx_linearize(cilqr_oracle_ctx* c, const double* states, const double* controls,
                                double* As, double* Bs, double* Jx, double* Ju, double* Hx,
                                double* Hu);
/* Backward on the last linearisation: Ks[N][12] (2x6 row-major) ks[N][2] dV[2]       :334-390 */
void cilqr_oracle_ctx_backward(cilqr_oracle_ctx* c, double lambda, double* Ks, double* ks,
                               double dV[2]);
/* Forward(alpha) from (states, controls) with the last gains -> new_states,new_controls :392-415 */
void cilqr_oracle_ctx_forward(cilqr_oracle_ctx* c, double alpha, const double* states,
                              const double* controls, double* new_states, double* new_controls);
int cilqr_oracle_ctx_nearest(const cilqr_oracle_ctx* c, int side, double x, double y); /* :605-618 */

#ifdef __cplusplus
}
#endif
#endif /* CILQR_ORACLE_H_ */
