// base_trajectory::Trajectory — host-side mirror of the reference's trajectory container
// (src/dddmr_local_planner/base_trajectory/include/base_trajectory/trajectory.h:47-126, src/trajectory.cpp:34-94).
// Same public members and accessors; the per-pose payload is held in flat arrays that the B200 generator adapters
// fill straight from one device read-back (b200lp_read_pose_batch), and the PoseStamped / PointCloud objects the
// reference API returns are built on demand. `id_` is the only addition: the trajectory's index in the cycle's
// generated list, which is how the critic adapters find the device-side scores of a trajectory they are handed.
#ifndef B200LP_BASE_TRAJECTORY_H_
#define B200LP_BASE_TRAJECTORY_H_

#include <utility>
#include <vector>

#include "b200lp/ros_compat.hpp"

namespace base_trajectory {

typedef std::pair<pcl::PointXYZ, pcl::PointXYZ> cuboid_min_max_t;

class Trajectory {
 public:
  Trajectory() : xv_(0.0), yv_(0.0), thetav_(0.0), cost_(-1.0), time_delta_(0.0) {}
  Trajectory(double xv, double yv, double thetav, double time_delta, unsigned int num_pts)
      : xv_(xv), yv_(yv), thetav_(thetav), cost_(-1.0), time_delta_(time_delta) {
    pose7_.reserve(7u * num_pts);
  }

  double xv_, yv_, thetav_;  ///< velocities that seeded the trajectory
  double cost_;              ///< StackedScoringModel result (negative = rejected)
  double time_delta_;        ///< time gap between points
  int id_ = -1;              ///< index in the cycle's generated-trajectory list (B200 addition)

  geometry_msgs::msg::PoseStamped getPoint(unsigned int index) const {
    geometry_msgs::msg::PoseStamped p;
    const double* s = &pose7_[7u * index];
    p.pose.position.x = s[0]; p.pose.position.y = s[1]; p.pose.position.z = s[2];
    p.pose.orientation.x = s[3]; p.pose.orientation.y = s[4]; p.pose.orientation.z = s[5]; p.pose.orientation.w = s[6];
    return p;
  }
  pcl::PointXYZI getPCLPoint(unsigned int index) const {
    pcl::PointXYZI p;
    p.x = pcl3_[3u * index]; p.y = pcl3_[3u * index + 1]; p.z = pcl3_[3u * index + 2];
    p.intensity = 0.f;
    return p;
  }
  void setPoint(unsigned int, double, double, double) {}  // a no-op in the reference too (trajectory.cpp:60-62)

  bool addPoint(const geometry_msgs::msg::PoseStamped& pos, const pcl::PointCloud<pcl::PointXYZ>& cuboid,
                const cuboid_min_max_t& mm) {
    const auto& q = pos.pose;
    const double row[7] = {q.position.x, q.position.y, q.position.z, q.orientation.x, q.orientation.y, q.orientation.z,
                           q.orientation.w};
    pose7_.insert(pose7_.end(), row, row + 7);
    const float p3[3] = {(float)q.position.x, (float)q.position.y, (float)q.position.z};
    pcl3_.insert(pcl3_.end(), p3, p3 + 3);
    for (std::size_t k = 0; k < 8; ++k) {
      const pcl::PointXYZ v = k < cuboid.size() ? cuboid[k] : pcl::PointXYZ();
      cuboid24_.push_back(v.x); cuboid24_.push_back(v.y); cuboid24_.push_back(v.z);
    }
    const float a[6] = {mm.first.x, mm.first.y, mm.first.z, mm.second.x, mm.second.y, mm.second.z};
    aabb6_.insert(aabb6_.end(), a, a + 6);
    return true;
  }
  // bulk form used by the generator adapters: rows come straight from the device read-back
  void assignPoints(const double* pose7, const float* pcl3, const float* cuboid24, const float* aabb6, unsigned int n) {
    pose7_.assign(pose7, pose7 + 7u * n);
    pcl3_.assign(pcl3, pcl3 + 3u * n);
    cuboid24_.assign(cuboid24, cuboid24 + 24u * n);
    aabb6_.assign(aabb6, aabb6 + 6u * n);
  }

  pcl::PointCloud<pcl::PointXYZ> getCuboid(unsigned int index) const {
    pcl::PointCloud<pcl::PointXYZ> c;
    const float* s = &cuboid24_[24u * index];
    for (int k = 0; k < 8; ++k) c.push_back(pcl::PointXYZ(s[3 * k], s[3 * k + 1], s[3 * k + 2]));
    return c;
  }
  cuboid_min_max_t getCuboidMinMax(unsigned int index) const {
    const float* s = &aabb6_[6u * index];
    return cuboid_min_max_t(pcl::PointXYZ(s[0], s[1], s[2]), pcl::PointXYZ(s[3], s[4], s[5]));
  }
  void getEndpoint(double&, double&, double&) const {}  // empty in the reference (trajectory.cpp:85-87)
  void resetPoints() {
    pose7_.clear(); pcl3_.clear(); cuboid24_.clear(); aabb6_.clear();
  }
  unsigned int getPointsSize() const { return (unsigned int)(pose7_.size() / 7u); }

 private:
  std::vector<double> pose7_;    // position xyz + orientation xyzw per pose
  std::vector<float> pcl3_;      // float-cast position
  std::vector<float> cuboid24_;  // 8 vertices, order blb,brb,blt,flb,brt,frt,flt,frb
  std::vector<float> aabb6_;     // min xyz, max xyz
};

}  // namespace base_trajectory
#endif
