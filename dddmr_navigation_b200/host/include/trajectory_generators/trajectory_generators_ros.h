// trajectory_generators::Trajectory_Generators_ROS — the node that loads generator plugins from `plugins:` and
// forwards to the StackedGenerator (reference: include/trajectory_generators/trajectory_generators_ros.h:45-49,
// src/trajectory_generators_ros.cpp:45-111). pluginlib::ClassLoader is replaced by b200lp::PluginFactory in builds
// without ROS 2; the YAML keys (`plugins`, `<name>.plugin`) are unchanged.
#ifndef B200LP_TRAJECTORY_GENERATORS_ROS_H_
#define B200LP_TRAJECTORY_GENERATORS_ROS_H_
#include <string>
#include <vector>

#include "trajectory_generators/stacked_generator.h"

namespace trajectory_generators {
class Trajectory_Generators_ROS : public rclcpp::Node {
 public:
  explicit Trajectory_Generators_ROS(std::string name) : rclcpp::Node(std::move(name)) {}
  void initial();
  bool hasMoreTrajectories(std::string pname) { return stacked_generator_.hasMoreTrajectories(pname); }
  bool nextTrajectory(std::string pname, base_trajectory::Trajectory& comp_traj) {
    return stacked_generator_.nextTrajectory(pname, comp_traj);
  }
  void initializeTheories_wi_Shared_data() { stacked_generator_.initializeTheories_wi_Shared_data(); }
  std::shared_ptr<TrajectoryGeneratorSharedData> getSharedDataPtr() { return stacked_generator_.getSharedDataPtr(); }
  StackedGenerator* getStackedGeneratorPtr() { return &stacked_generator_; }

 private:
  std::vector<std::string> plugins_;
  StackedGenerator stacked_generator_;
};
}  // namespace trajectory_generators
#endif
