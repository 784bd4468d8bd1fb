// dp_demo.cc -- drives planning::DpPlanner (include/cilqr/dp_planner_b200.h) the way the reference's
// TrajectoryPlanner does (algorithm/planner/trajectory_planner.cpp:24,32): construct with (PlannerConfig, Env),
// call Plan(start_x, start_y, start_theta, result).
//
//   dp_demo <scene.bin> <result.bin>
// scene.bin  (doubles): R, n_static, n_dyn, T, start[3], ref[R][7], static[n_static][4][2],
//                       per dynamic obstacle and sample: time, 4 corner points
// result.bin (doubles): ok, K, min_cost, trajectory[K][11] (time, s, x, y, theta, kappa, velocity, a, jerk, delta,
//                       delta_rate)
#include <cstdio>
#include <vector>

#include "cilqr/dp_planner_b200.h"

using namespace planning;

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  std::vector<double> d;
  double buf[1024];
  size_t n;
  while ((n = std::fread(buf, sizeof(double), 1024, f)) > 0) d.insert(d.end(), buf, buf + n);
  std::fclose(f);
  size_t o = 0;
  const int R = (int)d[o++], n_static = (int)d[o++], n_dyn = (int)d[o++], T = (int)d[o++];
  const double sx = d[o], sy = d[o + 1], sth = d[o + 2];
  o += 3;
  Env env = std::make_shared<Environment>();
  std::vector<TrajectoryPoint> line(R);
  for (int i = 0; i < R; ++i, o += 7) {
    line[i].s = d[o]; line[i].x = d[o + 1]; line[i].y = d[o + 2]; line[i].theta = d[o + 3]; line[i].kappa = d[o + 4];
    line[i].left_bound = d[o + 5]; line[i].right_bound = d[o + 6];
  }
  env->reference_ = DiscretizedTrajectory(line);
  // road barrier as Environment::set_reference leaves it (left / right lists); the demo receives it ready-made
  const int nL = (int)d[o++];
  for (int i = 0; i < nL; ++i, o += 2) env->left_.emplace_back(d[o], d[o + 1]);
  const int nR = (int)d[o++];
  for (int i = 0; i < nR; ++i, o += 2) env->right_.emplace_back(d[o], d[o + 1]);
  for (int j = 0; j < n_static; ++j) {
    std::vector<math::Vec2d> p;
    for (int v = 0; v < 4; ++v, o += 2) p.emplace_back(d[o], d[o + 1]);
    env->obstacles().emplace_back(p);
  }
  for (int j = 0; j < n_dyn; ++j) {
    Environment::DynamicObstacle ob;
    for (int t = 0; t < T; ++t) {
      const double time = d[o++];
      std::vector<math::Vec2d> p;
      for (int v = 0; v < 4; ++v, o += 2) p.emplace_back(d[o], d[o + 1]);
      ob.emplace_back(time, math::Polygon2d(p));
    }
    env->dynamic_obstacles().push_back(ob);
  }
  PlannerConfig config;
  DpPlanner dp(config, env);  // trajectory_planner.cpp:24
  DiscretizedTrajectory result;
  const bool ok = dp.Plan(sx, sy, sth, result);  // :32
  std::vector<double> r;
  r.push_back(ok ? 1.0 : 0.0);
  r.push_back((double)result.trajectory().size());
  r.push_back(dp.min_cost());
  for (const auto& p : result.trajectory()) {
    const double row[11] = {p.time, p.s, p.x, p.y, p.theta, p.kappa, p.velocity, p.a, p.jerk, p.delta, p.delta_rate};
    r.insert(r.end(), row, row + 11);
  }
  f = std::fopen(argv[2], "wb");
  if (!f) return 2;
  std::fwrite(r.data(), sizeof(double), r.size(), f);
  std::fclose(f);
  std::printf("dp_demo: ok=%d knots=%zu\n", (int)ok, result.trajectory().size());
  return result.trajectory().empty() ? 1 : 0;
}
