// mpc_critics::MPC_Critics_ROS — loads critic plugins from `plugins:` and binds each to its generator
// (reference: include/mpc_critics/mpc_critics_ros.h:46-54, src/mpc_critics_ros.cpp:45-97).
#ifndef B200LP_MPC_CRITICS_ROS_H_
#define B200LP_MPC_CRITICS_ROS_H_
#include <string>
#include <vector>

#include "mpc_critics/stacked_scoring_model.h"

namespace mpc_critics {
class MPC_Critics_ROS : public rclcpp::Node {
 public:
  explicit MPC_Critics_ROS(std::string name) : rclcpp::Node(std::move(name)) {}
  void initial();
  void scoreTrajectory(std::string traj_gen_name, base_trajectory::Trajectory& one_traj) {
    stacked_scoring_model_.scoreTrajectory(traj_gen_name, one_traj);
  }
  void updateSharedData() { stacked_scoring_model_.getSharedDataPtr()->updateData(); }
  std::shared_ptr<ModelSharedData> getSharedDataPtr() { return stacked_scoring_model_.getSharedDataPtr(); }
  StackedScoringModel* getStackedScoringModelPtr() { return &stacked_scoring_model_; }

 private:
  std::vector<std::string> plugins_;
  StackedScoringModel stacked_scoring_model_;
};
}  // namespace mpc_critics
#endif
