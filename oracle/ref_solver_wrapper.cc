// ref_solver_wrapper.cc -- extern "C" access to the REFERENCE'S OWN CILQR solver, compiled unmodified from
// /root/reference (algorithm/ilqr/ilqr_optimizer.cc, vehicle_model.cc, barrier_function.h + the geometry sources)
// against the stand-ins in oracle/ref_stubs: an Eigen-lite header that reproduces the Eigen semantics the solver
// depends on (lazy `auto` expressions, coefficient order of products, assignment aliasing rules, closed-form 2x2
// inverse -- see its header comment), no-op ROS logging macros and empty OpenCV headers.  So every line of solver
// control flow and arithmetic is the reference's; only the linear-algebra primitives underneath are a
// restatement.  TEST INFRASTRUCTURE ONLY.  This file contains no reference code: it only calls it.
//
// IlqrOptimizer::Plan itself falls off the end of a non-void function (ilqr_optimizer.cc:53-95), which GCC >= 8
// compiles to a fall-through at -O2; the wrapper therefore performs Plan's three statements itself
// (start_state_ = ..., cost_.clear(), TransformGoals, Optimize) through the private members.
#include <iostream>
#include <vector>

#define private public
#include "algorithm/ilqr/ilqr_optimizer.h"
#undef private

using namespace planning;

namespace planning {
// tracker.cc (dynamic-size Eigen, not on the live path: its call site is commented out, ilqr_optimizer.cc:168) is
// not compiled; the two symbols the solver's translation unit references are defined empty
void Tracker::InitMatrix() {}
bool Tracker::Plan(const TrajectoryPoint&, const DiscretizedTrajectory&, DiscretizedTrajectory* const) { return false; }
}  // namespace planning

namespace {

struct Problem {
  TrajectoryPoint start;
  DiscretizedTrajectory coarse;
  CorridorConstraints corridor;
  LaneConstraints left, right;
};

Problem make_problem(int N, int M_max, int S_left, int S_right, const double* start, const double* coarse,
                     const double* corridor, const int* cnt, const double* lane_left, const double* lane_right) {
  Problem p;
  const int K = N + 1;
  p.start.x = start[0]; p.start.y = start[1]; p.start.theta = start[2]; p.start.velocity = start[3];
  std::vector<TrajectoryPoint> pts(K);
  for (int k = 0; k < K; ++k) {
    const double* c = coarse + (size_t)k * 6;
    pts[k].x = c[0]; pts[k].y = c[1]; pts[k].theta = c[2]; pts[k].velocity = c[3]; pts[k].a = c[4]; pts[k].delta = c[5];
  }
  p.coarse = DiscretizedTrajectory(pts);
  p.corridor.resize(K);
  for (int k = 0; k < K; ++k)
    for (int m = 0; m < cnt[k]; ++m) {
      const double* q = corridor + ((size_t)k * M_max + m) * 3;
      p.corridor[k].push_back(Eigen::Vector3d(q[0], q[1], q[2]));
    }
  auto lanes = [](int S, const double* l) {
    LaneConstraints out;
    for (int s = 0; s < S; ++s) {
      const double* q = l + (size_t)s * 7;
      out.push_back(std::make_pair(Eigen::Vector3d(q[0], q[1], q[2]),
                                   math::LineSegment2d(math::Vec2d(q[3], q[4]), math::Vec2d(q[5], q[6]))));
    }
    return out;
  };
  p.left = lanes(S_left, lane_left);
  p.right = lanes(S_right, lane_right);
  return p;
}

struct Quiet {  // the reference prints every iteration to std::cout
  std::ios_base::iostate old;
  Quiet() : old(std::cout.rdstate()) { std::cout.setstate(std::ios_base::failbit); }
  ~Quiet() { std::cout.clear(old); }
};

}  // namespace

extern "C" {

// The whole solve.  states [K][6], controls [N][2] of the returned trajectory; init_states / init_controls of
// iter_trajs[0] (the iqr initial guess); cost_hist [cap][5] = cost(); returns 0.
// overrides: NULL, or {max_iter_num, abs_cost_tol, rel_cost_tol} (IlqrConfig, planner_config.h:63-66) to reach the
// exits the default tolerances make rare.
int ref_ilqr_solve_cfg(double dt, int N, int M_max, int S_left, int S_right, const double* start, const double* coarse,
                       const double* corridor, const int* cnt, const double* lane_left, const double* lane_right,
                       double* states, double* controls, double* init_states, double* init_controls, double* cost_hist,
                       int cap, int* n_cost, int* n_iter_trajs, const double* overrides) {
  Quiet q;
  const int K = N + 1;
  Problem p = make_problem(N, M_max, S_left, S_right, start, coarse, corridor, cnt, lane_left, lane_right);
  IlqrConfig config;
  if (overrides) {
    config.max_iter_num = (int)overrides[0];
    config.abs_cost_tol = overrides[1];
    config.rel_cost_tol = overrides[2];
  }
  VehicleParam vehicle;
  IlqrOptimizer opt(config, vehicle, N * dt + 0.5 * dt, dt);  // num_of_knots_ = floor(horizon / dt + 1) = N + 1
  if (opt.num_of_knots_ != K) return -1;
  DiscretizedTrajectory result;
  std::vector<DiscretizedTrajectory> iters;
  // IlqrOptimizer::Plan, ilqr_optimizer.cc:61-63,84-92
  opt.start_state_ = p.start;
  opt.cost_.clear();
  opt.TransformGoals(p.coarse);
  opt.Optimize(p.start, p.coarse, p.corridor, p.left, p.right, &result, &iters);
  auto unpack = [&](const DiscretizedTrajectory& t, double* X, double* U) {
    for (int k = 0; k < K; ++k) {
      const TrajectoryPoint& tp = t.trajectory()[k];
      if (X) {
        double* x = X + (size_t)k * 6;
        x[0] = tp.x; x[1] = tp.y; x[2] = tp.theta; x[3] = tp.velocity; x[4] = tp.a; x[5] = tp.delta;
      }
      if (U && k < N) {
        U[(size_t)k * 2] = tp.jerk;
        U[(size_t)k * 2 + 1] = tp.delta_rate;
      }
    }
  };
  if ((int)result.trajectory().size() != K || iters.empty()) return -2;
  unpack(result, states, controls);
  unpack(iters[0], init_states, init_controls);
  const std::vector<Cost> ch = opt.cost();
  *n_cost = (int)ch.size();
  *n_iter_trajs = (int)iters.size();
  for (int i = 0; i < (int)ch.size() && i < cap; ++i) {
    double* c = cost_hist + (size_t)i * 5;
    c[0] = ch[i].total_cost; c[1] = ch[i].target_cost; c[2] = ch[i].dynamic_cost; c[3] = ch[i].corridor_cost;
    c[4] = ch[i].lane_boundary_cost;
  }
  return 0;
}

int ref_ilqr_solve(double dt, int N, int M_max, int S_left, int S_right, const double* start, const double* coarse,
                   const double* corridor, const int* cnt, const double* lane_left, const double* lane_right,
                   double* states, double* controls, double* init_states, double* init_controls, double* cost_hist,
                   int cap, int* n_cost, int* n_iter_trajs) {
  return ref_ilqr_solve_cfg(dt, N, M_max, S_left, S_right, start, coarse, corridor, cnt, lane_left, lane_right, states,
                            controls, init_states, init_controls, cost_hist, cap, n_cost, n_iter_trajs, nullptr);
}

// VehicleModel::Dynamics / DynamicsJacbian (vehicle_model.cc:21-121)
void ref_dynamics(double dt, const double* x, const double* u, double* next) {
  IlqrConfig config;
  VehicleParam vehicle;
  VehicleModel m(config, vehicle, 8.0, dt);
  State s, n;
  Control c;
  for (int i = 0; i < 6; ++i) s(i, 0) = x[i];
  c(0, 0) = u[0]; c(1, 0) = u[1];
  m.Dynamics(s, c, &n);
  for (int i = 0; i < 6; ++i) next[i] = n(i, 0);
}
void ref_dynamics_jacobian(double dt, const double* x, const double* u, double* A /*[6][6] row-major*/,
                           double* B /*[6][2]*/) {
  IlqrConfig config;
  VehicleParam vehicle;
  VehicleModel m(config, vehicle, 8.0, dt);
  State s;
  Control c;
  for (int i = 0; i < 6; ++i) s(i, 0) = x[i];
  c(0, 0) = u[0]; c(1, 0) = u[1];
  SystemMatrix Am;
  InputMatrix Bm;
  m.DynamicsJacbian(s, c, &Am, &Bm);
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j < 6; ++j) A[i * 6 + j] = Am(i, j);
    for (int j = 0; j < 2; ++j) B[i * 2 + j] = Bm(i, j);
  }
}

// RelaxBarrierFunction<6> as the optimizer configures it (barrier_function.h:81-147)
double ref_barrier_value(double g) {
  RelaxBarrierFunction<kStateNum> b;
  IlqrConfig config;
  IlqrOptimizer opt(config, VehicleParam(), 8.05, 0.1);
  return opt.state_barrier_.value(g);
}

}  // extern "C"
