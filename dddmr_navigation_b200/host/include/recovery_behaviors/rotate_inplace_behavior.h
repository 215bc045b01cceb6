// recovery_behaviors::RotateInPlaceBehavior — the OTHER caller of the hot path: the recovery
// behaviour's control loop (src/dddmr_local_planner/recovery_behaviors/behaviors/rotate_inplace_behavior.cpp:137-305) drives
// the same generator / critic plugin stacks as Local_Planner, with the rotate-in-place theory, once per control period
// until the robot has turned a full circle. The ROS action server, the rate object, tf and the publishers of the reference
// stay outside: the embedding code hands in what they supply (pose, odometry, time) and gets the velocity command back.
#ifndef B200LP_ROTATE_INPLACE_BEHAVIOR_H_
#define B200LP_ROTATE_INPLACE_BEHAVIOR_H_
#include <memory>
#include <string>
#include <vector>

#include "dddmr_sys_core/dddmr_enum_states.h"
#include "local_planner/local_planner.h"

namespace recovery_behaviors {

class RotateInPlaceBehavior {
 public:
  // what one pass of the `while` loop of runBehavior() leaves behind
  struct Step {
    bool finished = false;                       // the loop ended in this pass (`break`) or its condition is now false
    dddmr_sys_core::RecoveryState result = dddmr_sys_core::RECOVERY_DONE;  // valid when finished
    double cmd_linear_x = 0.0, cmd_angular_z = 0.0;  // the Twist published in this pass
    int best_id = -1;
    double best_cost = -1.0;
    bool got_180 = false;
    double dist_left = 0.0;
  };

  explicit RotateInPlaceBehavior(const std::string& name) : name_(name) {}
  // rotate_inplace_behavior.cpp:52-76: `tolerance`, `frequency`, `trajectory_generator_name` parameters
  void initial(const std::shared_ptr<perception_3d::SharedData>& perception_3d,
               const std::shared_ptr<mpc_critics::MPC_Critics_ROS>& mpc_critics,
               const std::shared_ptr<trajectory_generators::Trajectory_Generators_ROS>& trajectory_generators,
               const std::string& trajectory_generator_name, double tolerance = 0.3, double frequency = 10.0);
  // :125-135 — the part of runBehavior() in front of the loop: the start heading
  void begin(const geometry_msgs::msg::TransformStamped& trans_gbl2b, double now_s);
  // :140-305 — one pass of the loop (the loop condition is evaluated first, like `while` does)
  Step step(const geometry_msgs::msg::TransformStamped& trans_gbl2b, const nav_msgs::msg::Odometry& robot_state, double now_s);
  void getBestTrajectory(const std::string& traj_gen_name, base_trajectory::Trajectory& best_traj);  // :78-110

  double frequency() const { return frequency_; }
  std::shared_ptr<std::vector<base_trajectory::Trajectory>> trajectories_;

 private:
  std::string name_, trajectory_generator_name_;
  double tolerance_ = 0.3, frequency_ = 10.0;
  std::shared_ptr<perception_3d::SharedData> perception_3d_;
  std::shared_ptr<mpc_critics::MPC_Critics_ROS> mpc_critics_ros_;
  std::shared_ptr<trajectory_generators::Trajectory_Generators_ROS> trajectory_generators_ros_;
  double current_angle_ = 0.0, start_angle_ = 0.0, last_valid_control_ = 0.0;
  bool got_180_ = false;
};

// tf2::impl::getYaw and angles::shortest_angular_distance, as the loop uses them
double yaw_of(double qx, double qy, double qz, double qw);
double shortest_angular_distance(double from, double to);

}  // namespace recovery_behaviors
#endif
