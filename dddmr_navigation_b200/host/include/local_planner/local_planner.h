// local_planner::Local_Planner — the CALLER of the hot path, reduced to what the path needs: the cycle of
// Local_Planner::computeVelocityCommand (src/dddmr_local_planner/local_planner/src/local_planner.cpp:482-621) and
// getBestTrajectory (:447-480), driving the generator and critic plugin stacks through their reference interfaces.
// Everything ROS supplies in the reference (tf pose, odometry, the perception stack's aggregated cloud, the pruned
// global plan) is set by the embedding code through the setters below; publishers, tf, prunePlan() and the
// perception opinions are outside the path (SURVEY.md §8f lists prunePlan / path-blocked as the next rows).
#ifndef B200LP_LOCAL_PLANNER_H_
#define B200LP_LOCAL_PLANNER_H_
#include <memory>
#include <string>
#include <vector>

#include "dddmr_sys_core/dddmr_enum_states.h"
#include "mpc_critics/mpc_critics_ros.h"
#include "trajectory_generators/trajectory_generators_ros.h"

namespace perception_3d {
// the two members of perception_3d::SharedData the cycle reads (perception_3d/include/perception_3d/shared_data.h:79,85)
struct SharedData {
  pcl::PointCloud<pcl::PointXYZI>::Ptr aggregate_observation_;
  double current_allowed_max_linear_speed_ = -1.0;
};
}  // namespace perception_3d

namespace local_planner {

class Local_Planner {
 public:
  explicit Local_Planner(const std::string& name) : name_(name) {}
  void initial(const std::shared_ptr<perception_3d::SharedData>& perception_3d,
               const std::shared_ptr<mpc_critics::MPC_Critics_ROS>& mpc_critics,
               const std::shared_ptr<trajectory_generators::Trajectory_Generators_ROS>& trajectory_generators);

  // inputs the reference pulls from ROS at the top of the cycle
  void setGlobalPose(const geometry_msgs::msg::TransformStamped& trans_gbl2b) { trans_gbl2b_ = trans_gbl2b; got_pose_ = true; }
  void cbOdom(const nav_msgs::msg::Odometry& msg) { robot_state_ = msg; got_odom_ = true; }
  void setPrunePlan(const nav_msgs::msg::Path& prune_plan) { prune_plan_ = prune_plan; }  // prunePlan()'s output

  dddmr_sys_core::PlannerState computeVelocityCommand(std::string traj_gen_name, base_trajectory::Trajectory& best_traj);
  void getBestTrajectory(std::string traj_gen_name, base_trajectory::Trajectory& best_traj);

  // true: hand the cloud to the generator's session before initialise() (one launch per cycle);
  // false: behave exactly like the unpatched reference caller (the critics' shared data is the only cloud hand-over)
  void setEarlyObservationHandOver(bool on) { early_observation_ = on; }

  std::shared_ptr<std::vector<base_trajectory::Trajectory>> trajectories_;

 private:
  std::string name_;
  std::shared_ptr<perception_3d::SharedData> perception_3d_;
  std::shared_ptr<mpc_critics::MPC_Critics_ROS> mpc_critics_ros_;
  std::shared_ptr<trajectory_generators::Trajectory_Generators_ROS> trajectory_generators_ros_;
  geometry_msgs::msg::TransformStamped trans_gbl2b_;
  nav_msgs::msg::Odometry robot_state_;
  nav_msgs::msg::Path prune_plan_;
  bool got_odom_ = false, got_pose_ = false, early_observation_ = true;
};

}  // namespace local_planner
#endif
