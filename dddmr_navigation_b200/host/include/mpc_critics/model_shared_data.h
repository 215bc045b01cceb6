// mpc_critics::ModelSharedData — what Local_Planner hands the critics each cycle (local_planner.cpp:579-586). Mirror of
// src/dddmr_local_planner/mpc_critics/include/mpc_critics/model_shared_data.h:67-116. updateData() keeps its job of
// deriving the float-cast prune-plan cloud, but builds NO kd-tree: the obstacle index is the voxel grid the session
// builds in HBM when it is handed pcl_perception_.
#ifndef B200LP_MODEL_SHARED_DATA_H_
#define B200LP_MODEL_SHARED_DATA_H_
#include <string>

#include "b200lp/ros_compat.hpp"

namespace mpc_critics {
class ModelSharedData {
 public:
  ModelSharedData() : heading_deviation_(0.0) {}
  void updateData() {
    global_frame_ = robot_pose_.header.frame_id;
    base_frame_ = robot_pose_.child_frame_id;
    pcl_prune_plan_.reset(new pcl::PointCloud<pcl::PointXYZI>);
    for (const auto& ps : prune_plan_.poses) {
      pcl::PointXYZI ipt;
      ipt.x = (float)ps.pose.position.x;
      ipt.y = (float)ps.pose.position.y;
      ipt.z = (float)ps.pose.position.z;
      ipt.intensity = 0.f;
      pcl_prune_plan_->push_back(ipt);
    }
  }
  pcl::PointCloud<pcl::PointXYZI>::Ptr pcl_perception_;
  pcl::PointCloud<pcl::PointXYZI>::Ptr pcl_prune_plan_;
  nav_msgs::msg::Path prune_plan_;
  geometry_msgs::msg::TransformStamped robot_pose_;
  std::string global_frame_, base_frame_;
  nav_msgs::msg::Odometry robot_state_;
  double heading_deviation_;
};
}  // namespace mpc_critics
#endif
