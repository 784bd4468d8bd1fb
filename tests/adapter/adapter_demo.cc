// adapter_demo.cc -- drives planning::IlqrOptimizer (include/cilqr/ilqr_optimizer_b200.h) exactly the
// way the reference's TrajectoryPlanner does (algorithm/planner/trajectory_planner.cpp:26,73-97,131-137):
// construct with (IlqrConfig, VehicleParam, tf, dt), call Plan once, read opt_trajectory,
// iter_trajs[0] and cost().
//
//   adapter_demo <scenario.bin> <result.bin>
//   adapter_demo --latency <repeat> <scenario.bin>...   one IlqrOptimizer, `repeat` timed Plan calls per scenario
//                                                       (after one warm-up call); prints "LAT <file#> <ms>" lines
// scenario.bin (doubles): N, M_max, S_left, S_right, start[4], coarse[K][6], cnt[K], corridor[K][M_max][3],
//                         lane_left[S_left][7], lane_right[S_right][7]
// result.bin   (doubles): ok, status, iters, n_cost, n_iter_trajs, opt[K][13], iter0[K][13], cost[n_cost][5]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cilqr/ilqr_optimizer_b200.h"

using namespace planning;

struct Scenario {
  int N = 0, M_max = 0, K = 0;
  TrajectoryPoint start_state;
  std::vector<TrajectoryPoint> coarse;
  CorridorConstraints corridor;
  LaneConstraints left, right;
};

static bool load_scenario(const char* path, Scenario* sc) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return false;
  std::vector<double> d;
  double buf[1024];
  size_t n;
  while ((n = std::fread(buf, sizeof(double), 1024, f)) > 0) d.insert(d.end(), buf, buf + n);
  std::fclose(f);
  size_t o = 0;
  const int N = (int)d[o++], M_max = (int)d[o++], S_left = (int)d[o++], S_right = (int)d[o++];
  const int K = N + 1;
  sc->N = N;
  sc->M_max = M_max;
  sc->K = K;
  TrajectoryPoint& start_state = sc->start_state;
  start_state.x = d[o]; start_state.y = d[o + 1]; start_state.theta = d[o + 2]; start_state.velocity = d[o + 3];
  o += 4;
  std::vector<TrajectoryPoint>& coarse = sc->coarse;
  coarse.resize(K);
  for (int k = 0; k < K; ++k, o += 6) {
    coarse[k].x = d[o]; coarse[k].y = d[o + 1]; coarse[k].theta = d[o + 2];
    coarse[k].velocity = d[o + 3]; coarse[k].a = d[o + 4]; coarse[k].delta = d[o + 5];
  }
  std::vector<int> cnt(K);
  for (int k = 0; k < K; ++k) cnt[k] = (int)d[o++];
  sc->corridor = CorridorConstraints(K);
  for (int k = 0; k < K; ++k)
    for (int m = 0; m < M_max; ++m, o += 3)
      if (m < cnt[k]) sc->corridor[k].push_back(Eigen::Vector3d(d[o], d[o + 1], d[o + 2]));
  auto read_lane = [&](int S) {
    LaneConstraints lane;
    for (int s = 0; s < S; ++s, o += 7)
      lane.emplace_back(Eigen::Vector3d(d[o], d[o + 1], d[o + 2]),
                        math::LineSegment2d(math::Vec2d(d[o + 3], d[o + 4]), math::Vec2d(d[o + 5], d[o + 6])));
    return lane;
  };
  sc->left = read_lane(S_left);
  sc->right = read_lane(S_right);
  return true;
}

// config 0 of BASELINE.json is the reference's actual use: one ego, one Plan call per click
// (planning_node.cc:82-88).  Wall-clock latency of that call through the drop-in class.
static int latency_mode(int repeat, int nfiles, char** files) {
  std::vector<Scenario> scs(nfiles);
  for (int i = 0; i < nfiles; ++i)
    if (!load_scenario(files[i], &scs[i])) return 2;
  IlqrConfig config;
  VehicleParam vehicle;
  const double dt = 0.1;
  IlqrOptimizer opt(config, vehicle, scs[0].N * dt, dt);
  DiscretizedTrajectory out;
  std::vector<DiscretizedTrajectory> iters;
  if (!opt.Plan(scs[0].start_state, DiscretizedTrajectory(scs[0].coarse), scs[0].corridor, scs[0].left, scs[0].right,
                &out, &iters))
    return 1;  // warm-up: creates the handle, allocates the staging buffers
  for (int i = 0; i < nfiles; ++i) {
    const Scenario& sc = scs[i];
    const DiscretizedTrajectory coarse(sc.coarse);
    for (int r = 0; r < repeat; ++r) {
      iters.clear();
      const auto t0 = std::chrono::steady_clock::now();
      const bool ok = opt.Plan(sc.start_state, coarse, sc.corridor, sc.left, sc.right, &out, &iters);
      const auto t1 = std::chrono::steady_clock::now();
      if (!ok) return 1;
      std::printf("LAT %d %.6f %d %d\n", i, std::chrono::duration<double, std::milli>(t1 - t0).count(),
                  opt.last_status(), opt.last_iterations());
    }
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 4 && std::strcmp(argv[1], "--latency") == 0) return latency_mode(std::atoi(argv[2]), argc - 3, argv + 3);
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s scenario.bin result.bin\n", argv[0]);
    return 2;
  }
  Scenario sc;
  if (!load_scenario(argv[1], &sc)) return 2;
  const int N = sc.N, K = sc.K;
  const TrajectoryPoint& start_state = sc.start_state;
  const std::vector<TrajectoryPoint>& coarse = sc.coarse;
  const CorridorConstraints& corridor = sc.corridor;
  const LaneConstraints &left = sc.left, &right = sc.right;

  IlqrConfig config;
  VehicleParam vehicle;
  const double dt = 0.1, tf = N * dt;
  IlqrOptimizer ilqr_optimizer;
  ilqr_optimizer = IlqrOptimizer(config, vehicle, tf, dt);  // trajectory_planner.cpp:26

  DiscretizedTrajectory opt_trajectory;
  std::vector<DiscretizedTrajectory> iter_trajs;
  // guard checks first (ilqr_optimizer.cc:64-78)
  int guards_ok = 1;
  guards_ok &= !ilqr_optimizer.Plan(start_state, DiscretizedTrajectory(coarse), corridor, left, right, nullptr, &iter_trajs);
  guards_ok &= !ilqr_optimizer.Plan(start_state, DiscretizedTrajectory(coarse), CorridorConstraints(), left, right, &opt_trajectory, &iter_trajs);
  guards_ok &= !ilqr_optimizer.Plan(start_state, DiscretizedTrajectory(coarse), corridor, LaneConstraints(), right, &opt_trajectory, &iter_trajs);
  std::vector<TrajectoryPoint> shorter(coarse.begin(), coarse.end() - 1);
  guards_ok &= !ilqr_optimizer.Plan(start_state, DiscretizedTrajectory(shorter), corridor, left, right, &opt_trajectory, &iter_trajs);
  guards_ok &= opt_trajectory.empty() && iter_trajs.empty();

  const bool ok = ilqr_optimizer.Plan(start_state, DiscretizedTrajectory(coarse), corridor, left, right,
                                      &opt_trajectory, &iter_trajs);
  std::vector<Cost> cost = ilqr_optimizer.cost();
  std::printf("Plan -> %d (guards %d), status %d, iterations %d, knots %zu, iter_trajs %zu, cost entries %zu\n", (int)ok,
              guards_ok, ilqr_optimizer.last_status(), ilqr_optimizer.last_iterations(), opt_trajectory.trajectory().size(),
              iter_trajs.size(), cost.size());
  std::vector<double> out;
  out.push_back(ok && guards_ok ? 1.0 : 0.0);
  out.push_back(ilqr_optimizer.last_status());
  out.push_back(ilqr_optimizer.last_iterations());
  out.push_back((double)cost.size());
  out.push_back((double)iter_trajs.size());
  auto dump = [&](const DiscretizedTrajectory& t) {
    for (int k = 0; k < K; ++k) {
      TrajectoryPoint p = k < (int)t.trajectory().size() ? t.trajectory()[k] : TrajectoryPoint();
      const double r[13] = {p.time, p.s, p.x, p.y, p.theta, p.kappa, p.velocity, p.a, p.jerk, p.delta, p.delta_rate,
                            p.left_bound, p.right_bound};
      out.insert(out.end(), r, r + 13);
    }
  };
  dump(opt_trajectory);
  dump(iter_trajs.empty() ? DiscretizedTrajectory() : iter_trajs[0]);  // trajectory_planner.cpp:133-135 reads it blindly
  for (const Cost& c : cost) {
    const double r[5] = {c.total_cost, c.target_cost, c.dynamic_cost, c.corridor_cost, c.lane_boundary_cost};
    out.insert(out.end(), r, r + 5);
  }
  FILE* g = std::fopen(argv[2], "wb");
  if (!g) return 2;
  std::fwrite(out.data(), sizeof(double), out.size(), g);
  std::fclose(g);
  return ok ? 0 : 1;
}
