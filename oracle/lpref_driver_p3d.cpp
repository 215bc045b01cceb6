// lpref_driver_p3d.cpp — runs the REFERENCE's own perception_3d::PathBlockedStrategy::selfMark
// (src/dddmr_perception_3d/plugins/path_blocked_strategy.cpp, compiled where it lies against oracle/ref_shims/) for
// tests/test_reference_sources.py. TEST INFRASTRUCTURE ONLY. The strategy keeps its ratio private; its opinion is what the
// local planner consumes (local_planner.cpp:597-607), so that is what is reported.
#include <memory>
#include <string>

#include <perception_3d/sensor.h>

#include "../include/b200lp.h"

extern "C" int lpref_path_blocked(const void* pts, size_t n, size_t stride, const float* pcl_xyzi, size_t m,
                                  double check_radius, int32_t* opinion_out) {
  if ((n && !pts) || (m && !pcl_xyzi) || !opinion_out || stride < 12) return B200LP_E_INVALID;
  try {
    auto node = std::make_shared<rclcpp::Node>("perception_3d");
    node->set_parameter_override("path_blocked.check_radius", check_radius);
    auto shared = std::make_shared<perception_3d::SharedData>();
    shared->aggregate_observation_.reset(new pcl::PointCloud<pcl::PointXYZI>);
    shared->aggregate_observation_->points.resize(n);
    for (size_t i = 0; i < n; ++i) {
      const float* p = (const float*)((const char*)pts + i * stride);
      pcl::PointXYZI q;
      q.x = p[0]; q.y = p[1]; q.z = p[2];
      shared->aggregate_observation_->points[i] = q;
    }
    for (size_t i = 0; i < m; ++i) {  // what Local_Planner::prunePlan leaves in pcl_prune_plan_ (local_planner.cpp:518)
      pcl::PointXYZI q;
      q.x = pcl_xyzi[4 * i]; q.y = pcl_xyzi[4 * i + 1]; q.z = pcl_xyzi[4 * i + 2]; q.intensity = pcl_xyzi[4 * i + 3];
      shared->pcl_prune_plan_.points.push_back(q);
    }
    auto tfbuf = std::make_shared<tf2_ros::Buffer>();
    auto utils = std::make_shared<perception_3d::GlobalUtils>("map", "base_link", 10.0, 0.3, 1.0, 1.0, tfbuf);
    auto s = lpref::create<perception_3d::Sensor>("perception_3d::PathBlockedStrategy");
    s->setSharedData(shared);
    s->initialize("path_blocked", node, utils);
    s->selfMark();
    *opinion_out = (int32_t)s->getOpinion();
  } catch (const std::exception& e) {
    return B200LP_E_INVALID;
  }
  return B200LP_OK;
}
