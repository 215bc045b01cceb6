// b200lp::Session — the C++ owner of one b200lp_ctx, shared by a generator plugin and the critic plugins bound to it.
//
// The reference runs a cycle as: generator initialise() -> nextTrajectory() x S -> critics' shared data update ->
// scoreTrajectory() x S x critics (local_planner.cpp:528-587). The device runs the whole cycle in ONE b200lp_plan call (two kernels back to back), so
// the plugins are adapters around a session: the generator opens the cycle, the first consumer of results triggers
// the launch, everyone else reads cached read-backs. A session is looked up by GENERATOR NAME, which is the key the
// reference binds critics to generators with (`<critic>.trajectory_generator`, mpc_critics_ros.cpp:71-79), so neither
// shared-data struct needs a new member.
//
// Errors: C-ABI failures surface as b200lp::Error (message from b200lp_last_error). There is no CPU path: without
// libb200lp.so + a CUDA device the first cycle throws.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../../include/b200lp.h"
#include "b200lp/ros_compat.hpp"
#include "base_trajectory/trajectory.h"

namespace b200lp {

class Error : public std::runtime_error {
 public:
  Error(int code, const std::string& what) : std::runtime_error(what), code_(code) {}
  int code() const { return code_; }

 private:
  int code_;
};

struct TheoryConfig {
  b200lp_limits limits{};
  b200lp_params params{};
  float cuboid[8][3]{};  // order blb,brb,blt,flb,brt,frt,flt,frb (dd_simple…cpp:211-218)
};

class Session {
 public:
  static std::shared_ptr<Session> forGenerator(const std::string& generator_name);
  static void resetAll();              // drop every session (tests, node shutdown)
  static void setDevice(int device);   // CUDA device new contexts are created on (default: $B200LP_DEVICE or 0)

  ~Session();

  // ---- configuration (plugin onInitialize) ----
  void configureTheory(const TheoryConfig& cfg);
  int addCritic(const b200lp_critic& critic);  // -> position in the generator's ordered stack
  void setGridConfig(const b200lp_grid_config& g);

  // ---- per-cycle inputs ----
  // Optional fast path for a patched Local_Planner: hand the aggregated observation over BEFORE initialise(), so the
  // single launch of the cycle already scores against it.
  void setObservation(const pcl::PointCloud<pcl::PointXYZI>::ConstPtr& cloud);
  void beginCycle(const geometry_msgs::msg::TransformStamped& robot_pose, const nav_msgs::msg::Odometry& robot_state,
                  const nav_msgs::msg::Path& prune_plan, double current_allowed_max_linear_speed);

  // ---- the steps either side of the cycle ----
  void setGlobalPlan(const std::vector<double>& poses7);  // Local_Planner::setPlan
  // Local_Planner::prunePlan on the device; fills poses7 (prune_plan_.poses) and pcl_xyzi (pcl_prune_plan_, x y z tag)
  b200lp_prune_info prunePlan(const double robot_xyz[3], double forward_distance, double backward_distance,
                              std::vector<double>& poses7, std::vector<float>& pcl_xyzi);
  b200lp_blocked pathBlocked(double check_radius);  // PathBlockedStrategy::selfMark

  // ---- the observation producer in front of the cycle ----
  // MultiLayerSpinningLidar::cbSensor's filter chain on the device; the result stays there as sensor `sensor`'s observation
  b200lp_observation_info sensorObservation(int sensor, const pcl::PointCloud<pcl::PointXYZ>& scan,
                                            const geometry_msgs::msg::TransformStamped& trans_b2s,
                                            const geometry_msgs::msg::TransformStamped& trans_gbl2b, const b200lp_sensor_params& sp);
  void readObservation(int sensor, pcl::PointCloud<pcl::PointXYZI>& out);  // Sensor::getObservation for host-side consumers
  // StackedPerception::aggregateObservations: the sensors' observations are concatenated ON THE DEVICE and become the
  // critics' cloud; `aggregate` receives the host copy other consumers read and is this cloud's identity from here on
  // (handing the same object to setObservation / the critics' shared data does not upload anything again).
  void aggregateObservations(const std::vector<int>& sensors, const pcl::PointCloud<pcl::PointXYZI>::Ptr& aggregate);

  // ---- generator side ----
  int trajectoryCount();
  void fillTrajectory(int id, base_trajectory::Trajectory& traj, bool with_points);

  // ---- critic side ----
  // Value critic `critic_index` of the stack returned for `traj` (NaN if an earlier critic of the stack already
  // rejected it — StackedScoringModel never asks in that case). Re-launches when the cloud / heading deviation the
  // critics' shared data holds differ from what the cycle was launched with.
  double criticScore(int critic_index, const base_trajectory::Trajectory& traj,
                     const pcl::PointCloud<pcl::PointXYZI>::ConstPtr& pcl_perception, double heading_deviation);

  // ---- results of the last launch ----
  const b200lp_result& result();
  int launchesThisCycle() const { return launches_this_cycle_; }
  std::uint64_t cycle() const { return cycle_; }
  b200lp_ctx* ctx() { return ctx_; }

 private:
  Session() = default;
  void ensureContext();
  void ensureLaunched();
  void launch();
  void uploadCloud(const pcl::PointCloud<pcl::PointXYZI>::ConstPtr& cloud);
  void loadPoints();
  [[noreturn]] void raise(int code, const char* where);

  std::mutex mu_;
  b200lp_ctx* ctx_ = nullptr;
  bool have_theory_ = false, config_dirty_ = true;
  TheoryConfig theory_{};
  std::vector<b200lp_critic> critics_;
  b200lp_grid_config grid_{};

  // Is the device's grid still the caller's cloud? Pointer identity alone misses a caller that refills the SAME cloud object
  // in place: the token also carries the header stamp / seq, the size, the address of the point storage and the bits of
  // three points (first, middle, last).
  struct CloudToken {
    const void* object = nullptr;
    const void* storage = nullptr;
    std::size_t size = 0;
    std::uint64_t stamp = 0;
    std::uint32_t seq = 0;
    float probe[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool operator==(const CloudToken& o) const;
  };
  static CloudToken tokenOf(const pcl::PointCloud<pcl::PointXYZI>::ConstPtr& cloud);
  pcl::PointCloud<pcl::PointXYZI>::ConstPtr cloud_;  // kept alive
  CloudToken cloud_token_;
  bool cloud_uploaded_ = false;
  bool cloud_from_device_ = false;  // cloud_ is the host copy of a device-side aggregate: the device already holds it
  b200lp_query query_{};
  std::vector<double> plan7_;
  std::vector<double> plan7_device_;  // what the device-side prune plan holds; launch() skips the upload when equal
  bool plan_resident_ = false;
  bool in_cycle_ = false, launched_ = false, points_loaded_ = false;
  std::uint64_t cycle_ = 0;
  int launches_this_cycle_ = 0;

  b200lp_result result_{};
  std::vector<float> vel_;
  std::vector<double> dt_, cost_, scores_;
  std::vector<int32_t> steps_;
  std::vector<int64_t> pose_off_;
  std::vector<double> pose7_;
  std::vector<float> pcl3_, cuboid24_, aabb6_;
};

}  // namespace b200lp
