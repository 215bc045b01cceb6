// trajectory_generators::TrajectoryGeneratorSharedData — the per-cycle inputs Local_Planner copies in before
// initializeTheories_wi_Shared_data() (local_planner.cpp:528-533). Mirror of
// src/dddmr_local_planner/trajectory_generators/include/trajectory_generators/trajectory_shared_data.h:58-105 minus the
// tf2 buffer (no tf lookups happen on the hot path) and updateGoalatRobotFrame() (not used by any shipped theory).
#ifndef B200LP_TRAJECTORY_SHARED_DATA_H_
#define B200LP_TRAJECTORY_SHARED_DATA_H_
#include <string>

#include "b200lp/ros_compat.hpp"

namespace trajectory_generators {
class TrajectoryGeneratorSharedData {
 public:
  TrajectoryGeneratorSharedData() : current_allowed_max_linear_speed_(-1.0) {}
  geometry_msgs::msg::TransformStamped robot_pose_;
  nav_msgs::msg::Odometry robot_state_;
  nav_msgs::msg::Path prune_plan_;
  double current_allowed_max_linear_speed_;
  std::string global_frame_, base_frame_;
};
}  // namespace trajectory_generators
#endif
