// Stand-in for algorithm/params/planner_config.h:45-73 (Weights, IlqrConfig without the tracker).
#pragma once
#include "algorithm/params/vehicle_param.h"
namespace planning {
struct Weights {
  double jerk = 1;
  double delta_rate = 1;
  double x_target = 0.5;
  double y_target = 0.5;
  double theta = 1e-3;
  double v = 0.0;
  double a = 0.0;
  double delta = 0.0;
};
struct IlqrConfig {
  int num_of_disc = 5;
  double safe_margin = 0.2;
  double t = 100.0;
  double t_rate = 10.0;
  Weights weights;
  int max_iter_num = 200;
  double abs_cost_tol = 1e-2;
  double rel_cost_tol = 1e-2;
  double alpha = 1.0;
  double gamma = 0.5;
  double rho = 1e-9;
};
}  // namespace planning
namespace planning {
// Stand-in for algorithm/params/planner_config.h:75-86.
struct CorridorConfig {
  bool is_multiple_sample = false;
  double max_diff_x = 25.0;
  double max_diff_y = 25.0;
  double radius = 150.0;
  double max_axis_x = 10.0;
  double max_axis_y = 10.0;
  double lane_segment_length = 5.0;
};
}  // namespace planning
namespace planning {
// Stand-in for algorithm/params/planner_config.h:88-141 (the fields planning::DpPlanner reads).
struct PlannerConfig {
  double delta_t = 0.1;
  double tf = 8;
  double dp_nominal_velocity = 10.0;
  double dp_w_obstacle = 1000;
  double dp_w_lateral = 0.1;
  double dp_w_lateral_change = 0.5;
  double dp_w_lateral_velocity_change = 1.0;
  double dp_w_longitudinal_velocity_bias = 10.0;
  double dp_w_longitudinal_velocity_change = 1.0;
  VehicleParam vehicle;
  CorridorConfig corridor_config;
  IlqrConfig ilqr_config;
};
}  // namespace planning
