// Compile check of the drop-in class against the REFERENCE'S OWN headers (tests/test_adapters_real_headers.py):
// the call sites below are the reference's (file:line in the comments).
#include "cilqr/corridor_b200.h"
using namespace planning;
bool drive(const PlannerConfig& config, const Env& env, const DiscretizedTrajectory& coarse) {
  Corridor corridor(config.corridor_config, env);  // trajectory_planner.cpp:25
  CorridorConstraints cc;
  ConvexPolygons polys;
  LaneConstraints left, right;
  const bool ok = corridor.Plan(coarse, &cc, &polys, &left, &right);  // :49-57
  auto pts = corridor.points_for_corridors();                         // planning_node.cc:88
  return ok && !pts.empty();
}
int main() { return 0; }
