// No-op stand-in for <ros/ros.h>: the reference's solver uses only the logging macros.  TEST INFRASTRUCTURE ONLY.
#pragma once
#define ROS_ERROR(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_INFO(...) ((void)0)
namespace ros {
struct Duration {
  explicit Duration(double) {}
  void sleep() {}
};
}  // namespace ros
