// Compile check of the drop-in class against the REFERENCE'S OWN headers (tests/test_adapters_real_headers.py):
// the call sites below are the reference's (file:line in the comments).
#include "cilqr/dp_planner_b200.h"
using namespace planning;
bool drive(const PlannerConfig& config, const Env& env, DiscretizedTrajectory& out) {
  DpPlanner dp(config, env);               // trajectory_planner.cpp:24
  return dp.Plan(0.0, 0.0, 0.0, out);      // :32
}
int main() { return 0; }
