// Stand-in for algorithm/params/vehicle_param.h:21-64 (the fields the solver reads).
#pragma once
#include <cmath>
namespace planning {
class VehicleParam {
 public:
  double front_hang_length = 0.96;
  double wheel_base = 1.0;
  double rear_hang_length = 0.929;
  double width = 1.942;
  double max_velocity = 20.0;
  double min_acceleration = -5.0;
  double max_acceleration = 5.0;
  double jerk_min = -10.0;
  double jerk_max = 10.0;
  double delta_min = -40.0 / 180 * M_PI;
  double delta_max = 40.0 / 180 * M_PI;
  double delta_rate_min = delta_min / 3.0;
  double delta_rate_max = delta_max / 3.0;
};
}  // namespace planning
