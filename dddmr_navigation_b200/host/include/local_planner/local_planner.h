// local_planner::Local_Planner — the CALLER of the hot path, reduced to what the path needs: the cycle of
// Local_Planner::computeVelocityCommand (src/dddmr_local_planner/local_planner/src/local_planner.cpp:482-621) and
// getBestTrajectory (:447-480), driving the generator and critic plugin stacks through their reference interfaces.
// Everything ROS supplies in the reference (tf pose, odometry, the perception stack's aggregated cloud, the pruned
// global plan) is set by the embedding code through the setters below; publishers and tf are outside the path.
// setPlan() / prunePlan() (:322-343, :374-445) and the path-blocked opinion of perception_3d::PathBlockedStrategy
// (dddmr_perception_3d/plugins/path_blocked_strategy.cpp:56-100, consumed at local_planner.cpp:597-607) — SURVEY.md
// §8(f) rows 1 and 2 — run on the device through the same session.
#ifndef B200LP_LOCAL_PLANNER_H_
#define B200LP_LOCAL_PLANNER_H_
#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "dddmr_sys_core/dddmr_enum_states.h"
#include "mpc_critics/mpc_critics_ros.h"
#include "trajectory_generators/trajectory_generators_ros.h"

namespace perception_3d {
// the members of perception_3d::SharedData the cycle reads (perception_3d/include/perception_3d/shared_data.h:59,79,85)
struct SharedData {
  pcl::PointCloud<pcl::PointXYZI>::Ptr aggregate_observation_;
  pcl::PointCloud<pcl::PointXYZI> pcl_prune_plan_;
  double current_allowed_max_linear_speed_ = -1.0;
};
enum PerceptionOpinion { PASS = 0, PATH_BLOCKED_WAIT = 1, PATH_BLOCKED_REPLANNING = 2 };  // perception_3d/sensor.h

// perception_3d::MultiLayerSpinningLidar reduced to the producer of the local planner's observation: the filter chain of
// cbSensor (dddmr_perception_3d/plugins/multilayer_spinning_lidar.cpp:232-269; what pcl::fromROSMsg / the stitcher leave in
// pcl_msg and the two tf lookups are handed in) runs on the device of the generator's session, and the observation stays
// there; getObservation() (Sensor::getObservation) reads it back for host-side consumers. SURVEY.md §8(f) row 4.
class MultiLayerSpinningLidar {
 public:
  // stitcher_num: the `stitcher_num` parameter (:118): the last stitcher_num scans are concatenated before filtering
  MultiLayerSpinningLidar(const std::string& name, const std::string& traj_gen_name, int slot, double perception_window_size,
                          double marking_height, bool is_local_planner, int stitcher_num = 0)
      : name_(name), traj_gen_name_(traj_gen_name), slot_(slot), perception_window_size_(perception_window_size),
        marking_height_(marking_height), is_local_planner_(is_local_planner), stitcher_num_(stitcher_num),
        sensor_current_observation_(new pcl::PointCloud<pcl::PointXYZI>) {}
  void cbSensor(const pcl::PointCloud<pcl::PointXYZ>& pcl_msg, const geometry_msgs::msg::TransformStamped& trans_b2s,
                const geometry_msgs::msg::TransformStamped& trans_gbl2b);
  pcl::PointCloud<pcl::PointXYZI>::Ptr getObservation();
  const std::string& getName() const { return name_; }
  const std::string& generatorName() const { return traj_gen_name_; }
  int slot() const { return slot_; }
  const b200lp_observation_info& lastInfo() const { return last_info_; }

 private:
  std::string name_, traj_gen_name_;
  int slot_;
  double perception_window_size_, marking_height_;
  bool is_local_planner_, observation_stale_ = false;
  int stitcher_num_ = 0;
  std::deque<pcl::PointCloud<pcl::PointXYZ>> pcl_stitcher_;  // multilayer_spinning_lidar.cpp:186-199
  pcl::PointCloud<pcl::PointXYZI>::Ptr sensor_current_observation_;
  b200lp_observation_info last_info_{};
};

// perception_3d::StackedPerception::aggregateObservations (src/stacked_perception.cpp:128-140): the plugins' observations
// are concatenated in plugin order — on the device, where the result is the critics' cloud of the next cycle; the host
// copy lands in SharedData::aggregate_observation_ for the reference's other readers.
class StackedPerception {
 public:
  explicit StackedPerception(const std::shared_ptr<SharedData>& shared_data) : shared_data_(shared_data) {}
  void addPluginToVector(const std::shared_ptr<MultiLayerSpinningLidar>& plugin) { plugins_.push_back(plugin); }
  void aggregateObservations();

 private:
  std::shared_ptr<SharedData> shared_data_;
  std::vector<std::shared_ptr<MultiLayerSpinningLidar>> plugins_;
};

// perception_3d::PathBlockedStrategy (plugins/path_blocked_strategy.cpp): the `check_radius` parameter and selfMark(),
// answered by the device against the voxel grid of the cloud the critics query.
class PathBlockedStrategy {
 public:
  explicit PathBlockedStrategy(double check_radius) : check_radius_(check_radius) {}
  void selfMark(const std::string& traj_gen_name);
  PerceptionOpinion getOpinion() const { return opinion_; }
  double getBlockedRatio() const { return prune_plan_blocked_ratio_; }

 private:
  double check_radius_ = 0.0, prune_plan_blocked_ratio_ = 0.0;
  PerceptionOpinion opinion_ = PASS;
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
  void setPrunePlan(const nav_msgs::msg::Path& prune_plan) { prune_plan_ = prune_plan; }  // a prune plan made elsewhere

  // Local_Planner::setPlan (:322-343) and prunePlan (:374-445): the global plan lives on the device, pruning runs
  // there, prune_plan_ / pcl_prune_plan_ are read back for the reference's other consumers (publishers, perception).
  void setPlan(const std::vector<geometry_msgs::msg::PoseStamped>& orig_global_plan, const std::string& traj_gen_name);
  void prunePlan(double forward_distance, double backward_distance, const std::string& traj_gen_name);
  const nav_msgs::msg::Path& getPrunePlan() const { return prune_plan_; }
  const pcl::PointCloud<pcl::PointXYZI>& getPCLPrunePlan() const { return pcl_prune_plan_; }
  // optional path-blocked strategy: its opinion is looped after scoring like local_planner.cpp:597-607
  void setPathBlockedStrategy(const std::shared_ptr<perception_3d::PathBlockedStrategy>& s) { path_blocked_ = s; }

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
  pcl::PointCloud<pcl::PointXYZI> pcl_prune_plan_;
  std::vector<geometry_msgs::msg::PoseStamped> global_plan_;
  std::shared_ptr<perception_3d::PathBlockedStrategy> path_blocked_;
  bool got_odom_ = false, got_pose_ = false, early_observation_ = true;
};

}  // namespace local_planner
#endif
