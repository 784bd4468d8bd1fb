// corridor_demo.cc -- drives planning::Corridor (include/cilqr/corridor_b200.h) the way the reference's
// TrajectoryPlanner does (algorithm/planner/trajectory_planner.cpp:25,49-57, planning_node.cc:87-103): construct with
// (CorridorConfig, Env), call Plan once, read the constraints, polygons, lanes and points_for_corridors().
//
//   corridor_demo <scene.bin> <result.bin>
// scene.bin  (doubles): K, n_static, n_dyn, nL, nR, traj[K][4] (x, y, theta, time), static[n_static][2],
//                       per dynamic obstacle: samples, then per sample: time, 4 corner points;
//                       left[nL][2], right[nR][2]
// result.bin (doubles): ok, K, then per knot: m, planes[m][3], polygon[m][2], n_points; SL, left[SL][7], SR, right[SR][7]
#include <cstdio>
#include <vector>

#include "cilqr/corridor_b200.h"

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
  const int K = (int)d[o++], n_static = (int)d[o++], n_dyn = (int)d[o++], nL = (int)d[o++], nR = (int)d[o++];
  std::vector<TrajectoryPoint> pts(K);
  for (int k = 0; k < K; ++k, o += 4) {
    pts[k].x = d[o]; pts[k].y = d[o + 1]; pts[k].theta = d[o + 2]; pts[k].time = d[o + 3];
  }
  Env env = std::make_shared<Environment>();
  for (int i = 0; i < n_static; ++i, o += 2) env->static_.emplace_back(d[o], d[o + 1]);
  for (int j = 0; j < n_dyn; ++j) {
    const int samples = (int)d[o++];
    Environment::DynamicObstaclePoints ob;
    for (int s = 0; s < samples; ++s) {
      const double t = d[o++];
      std::vector<math::Vec2d> c;
      for (int q = 0; q < 4; ++q, o += 2) c.emplace_back(d[o], d[o + 1]);
      ob.emplace_back(t, c);
    }
    env->dynamic_.push_back(ob);
  }
  for (int i = 0; i < nL; ++i, o += 2) env->left_.emplace_back(d[o], d[o + 1]);
  for (int i = 0; i < nR; ++i, o += 2) env->right_.emplace_back(d[o], d[o + 1]);

  CorridorConfig config;
  Corridor corridor;
  corridor = Corridor(config, env);  // trajectory_planner.cpp:25 (there: member initialiser)
  CorridorConstraints cc;
  ConvexPolygons polys;
  LaneConstraints left, right;
  // guards first (corridor.cc:23-36)
  const bool g1 = corridor.Plan(DiscretizedTrajectory(), &cc, &polys, &left, &right);
  const bool g2 = corridor.Plan(DiscretizedTrajectory(pts), nullptr, &polys, &left, &right);
  const bool ok = corridor.Plan(DiscretizedTrajectory(pts), &cc, &polys, &left, &right);  // :49-57
  std::vector<double> r;
  r.push_back(ok && !g1 && !g2 ? 1.0 : 0.0);
  r.push_back(K);
  const auto pfc = corridor.points_for_corridors();
  for (int k = 0; k < K && ok; ++k) {
    r.push_back((double)cc[k].size());
    for (const auto& c : cc[k]) { r.push_back(c[0]); r.push_back(c[1]); r.push_back(c[2]); }
    for (const auto& p : polys[k]) { r.push_back(p[0]); r.push_back(p[1]); }
    r.push_back((double)pfc[k].size());
  }
  for (const LaneConstraints* lane : {&left, &right}) {
    if (!ok) break;
    r.push_back((double)lane->size());
    for (const auto& s : *lane) {
      r.push_back(s.first[0]); r.push_back(s.first[1]); r.push_back(s.first[2]);
      r.push_back(s.second.start().x()); r.push_back(s.second.start().y());
      r.push_back(s.second.end().x()); r.push_back(s.second.end().y());
    }
  }
  f = std::fopen(argv[2], "wb");
  if (!f) return 2;
  std::fwrite(r.data(), sizeof(double), r.size(), f);
  std::fclose(f);
  std::printf("corridor_demo: ok=%d knots=%d\n", (int)ok, K);
  return ok ? 0 : 1;
}
