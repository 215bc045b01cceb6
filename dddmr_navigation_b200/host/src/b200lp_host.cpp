// b200lp_host.cpp — implementation of the host layer above the C ABI: parameter store, session, generator and critic
// adapters, plugin loaders and the cycle driver. Compiled with g++ into libb200lp_host.so, linked against libb200lp.so.
// Nothing here computes a rollout, a distance or a score: every number comes back from the device.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <sstream>

#include "b200lp/plugin_factory.hpp"
#include "b200lp/session.hpp"
#include "local_planner/local_planner.h"
#include "recovery_behaviors/rotate_inplace_behavior.h"
#include "mpc_critics/b200_models.h"
#include "trajectory_generators/b200_theories.h"

// =====================================================================================================
// rclcpp stand-in: parameter store + the YAML subset of the reference's config files
// =====================================================================================================
#ifndef B200LP_HAVE_ROS2
namespace rclcpp {
namespace {
[[noreturn]] void type_error(const std::string& name, const char* want) {
  throw std::runtime_error("parameter '" + name + "' is not " + want);
}
std::string trim(const std::string& s) {
  const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
std::string unquote(std::string s) {
  s = trim(s);
  if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) return s.substr(1, s.size() - 2);
  return s;
}
bool parse_number(const std::string& s, double& out, bool& is_int) {
  if (s.empty()) return false;
  char* end = nullptr;
  out = std::strtod(s.c_str(), &end);
  if (end == s.c_str() || *end != '\0') return false;
  is_int = s.find_first_of(".eEnN") == std::string::npos;
  return true;
}
ParameterVariant parse_scalar_or_list(const std::string& raw) {
  const std::string s = trim(raw);
  if (!s.empty() && s.front() == '[') {
    std::vector<std::string> items;
    std::string body = s.substr(1, s.rfind(']') - 1), cur;
    std::stringstream ss(body);
    while (std::getline(ss, cur, ',')) {
      if (!trim(cur).empty()) items.push_back(unquote(cur));
    }
    std::vector<double> nums;
    bool all_num = !items.empty();
    for (const auto& it : items) {
      double v; bool is_int;
      if (parse_number(it, v, is_int)) nums.push_back(v);
      else all_num = false;
    }
    if (all_num) return nums;
    return items;
  }
  const std::string u = unquote(s);
  if (u.size() != s.size()) return u;  // quoted => string
  if (u == "true" || u == "True") return true;
  if (u == "false" || u == "False") return false;
  double v; bool is_int;
  if (parse_number(u, v, is_int)) {
    if (is_int) return (int64_t)v;
    return v;
  }
  return u;
}
}  // namespace

double Parameter::as_double() const {
  if (auto p = std::get_if<double>(&v_)) return *p;
  if (auto p = std::get_if<int64_t>(&v_)) return (double)*p;
  type_error(name_, "a double");
}
bool Parameter::as_bool() const {
  if (auto p = std::get_if<bool>(&v_)) return *p;
  type_error(name_, "a bool");
}
std::string Parameter::as_string() const {
  if (auto p = std::get_if<std::string>(&v_)) return *p;
  type_error(name_, "a string");
}
std::vector<double> Parameter::as_double_array() const {
  if (auto p = std::get_if<std::vector<double>>(&v_)) return *p;
  type_error(name_, "a double array");
}
std::vector<std::string> Parameter::as_string_array() const {
  if (auto p = std::get_if<std::vector<std::string>>(&v_)) return *p;
  type_error(name_, "a string array");
}

void Node::load_parameters_yaml(const std::string& text) {
  // indentation-scoped key stack; only the subtree <name_>/ros__parameters is kept
  struct Level { int indent; std::string key; };
  std::vector<Level> stack;
  std::stringstream ss(text);
  std::string line;
  while (std::getline(ss, line)) {
    // strip comments outside quotes
    bool in_q = false; char qc = 0;
    for (size_t i = 0; i < line.size(); ++i) {
      const char c = line[i];
      if (in_q) { if (c == qc) in_q = false; }
      else if (c == '"' || c == '\'') { in_q = true; qc = c; }
      else if (c == '#') { line.erase(i); break; }
    }
    if (trim(line).empty()) continue;
    const int indent = (int)line.find_first_not_of(' ');
    const size_t colon = line.find(':');
    if (colon == std::string::npos) continue;
    const std::string key = unquote(line.substr(indent, colon - indent));
    const std::string val = trim(line.substr(colon + 1));
    while (!stack.empty() && stack.back().indent >= indent) stack.pop_back();
    if (val.empty()) {
      stack.push_back({indent, key});
      continue;
    }
    // path: [node, ros__parameters, a, b, ...] + key
    if (stack.size() < 2) continue;
    std::string node = stack[0].key;
    if (!node.empty() && node.front() == '/') node.erase(0, 1);
    if (node != name_ || stack[1].key != "ros__parameters") continue;
    std::string full;
    for (size_t i = 2; i < stack.size(); ++i) full += stack[i].key + ".";
    full += key;
    overrides_[full] = parse_scalar_or_list(val);
  }
}

void Node::declare_parameter(const std::string& name, const ParameterValue& def) {
  if (declared_.count(name)) throw std::runtime_error("parameter '" + name + "' has already been declared");
  auto it = overrides_.find(name);
  ParameterVariant v = def.get();
  if (it != overrides_.end()) {
    // an integer literal overriding a double default is a double (rclcpp would reject it; YAML authors write 1 for 1.0)
    if (std::holds_alternative<double>(v) && std::holds_alternative<int64_t>(it->second)) v = (double)std::get<int64_t>(it->second);
    else v = it->second;
  }
  declared_[name] = v;
  has_value_[name] = true;
}
void Node::declare_parameter(const std::string& name, ParameterType type) {
  if (declared_.count(name)) throw std::runtime_error("parameter '" + name + "' has already been declared");
  auto it = overrides_.find(name);
  if (it != overrides_.end()) {
    declared_[name] = it->second;
    has_value_[name] = true;
  } else {
    declared_[name] = type == PARAMETER_DOUBLE_ARRAY ? ParameterVariant(std::vector<double>()) : ParameterVariant(std::vector<std::string>());
    has_value_[name] = false;
  }
}
Parameter Node::get_parameter(const std::string& name) const {
  auto it = declared_.find(name);
  if (it == declared_.end()) throw std::runtime_error("parameter '" + name + "' has not been declared");
  if (!has_value_.at(name)) throw std::runtime_error("parameter '" + name + "' is not set");
  return Parameter(name, it->second, true);
}
bool Node::get_parameter(const std::string& name, double& out) const {
  auto it = declared_.find(name);
  if (it == declared_.end()) return false;
  out = Parameter(name, it->second, true).as_double();
  return true;
}
bool Node::get_parameter(const std::string& name, bool& out) const {
  auto it = declared_.find(name);
  if (it == declared_.end()) return false;
  out = Parameter(name, it->second, true).as_bool();
  return true;
}
bool Node::get_parameter(const std::string& name, std::string& out) const {
  auto it = declared_.find(name);
  if (it == declared_.end()) return false;
  out = Parameter(name, it->second, true).as_string();
  return true;
}
}  // namespace rclcpp
#endif  // !B200LP_HAVE_ROS2

// =====================================================================================================
// b200lp::Session
// =====================================================================================================
namespace b200lp {
namespace {
std::mutex g_registry_mu;
std::map<std::string, std::shared_ptr<Session>> g_registry;
int g_device = -1;
int default_device() {
  if (g_device >= 0) return g_device;
  if (const char* e = std::getenv("B200LP_DEVICE")) return std::atoi(e);
  return 0;
}
}  // namespace

std::shared_ptr<Session> Session::forGenerator(const std::string& generator_name) {
  std::lock_guard<std::mutex> lk(g_registry_mu);
  auto& slot = g_registry[generator_name];
  if (!slot) slot.reset(new Session());
  return slot;
}
void Session::resetAll() {
  std::lock_guard<std::mutex> lk(g_registry_mu);
  g_registry.clear();
}
void Session::setDevice(int device) { g_device = device; }

Session::~Session() {
  if (ctx_) b200lp_destroy(ctx_);
}

void Session::raise(int code, const char* where) {
  const char* msg = b200lp_last_error(ctx_);
  throw Error(code, std::string(where) + ": " + (msg ? msg : "") + " (b200lp code " + std::to_string(code) + ")");
}

void Session::configureTheory(const TheoryConfig& cfg) {
  std::lock_guard<std::mutex> lk(mu_);
  theory_ = cfg;
  have_theory_ = true;
  config_dirty_ = true;
}
int Session::addCritic(const b200lp_critic& critic) {
  std::lock_guard<std::mutex> lk(mu_);
  if ((int)critics_.size() >= B200LP_MAX_CRITICS) throw Error(B200LP_E_INVALID, "more than B200LP_MAX_CRITICS critics bound to one generator");
  critics_.push_back(critic);
  config_dirty_ = true;
  return (int)critics_.size() - 1;
}
void Session::setGridConfig(const b200lp_grid_config& g) {
  std::lock_guard<std::mutex> lk(mu_);
  grid_ = g;
  config_dirty_ = true;
}

void Session::ensureContext() {
  if (ctx_ && !config_dirty_) return;
  if (!have_theory_) throw Error(B200LP_E_STATE, "b200lp::Session: no generator plugin has configured this session");
  if (ctx_) {
    b200lp_destroy(ctx_);
    ctx_ = nullptr;
  }
  const int rc = b200lp_create(&ctx_, default_device(), &theory_.limits, &theory_.params, &theory_.cuboid[0][0],
                               critics_.empty() ? nullptr : critics_.data(), (int)critics_.size(), &grid_);
  if (rc != B200LP_OK) {
    ctx_ = nullptr;
    raise(rc, "b200lp_create");
  }
  config_dirty_ = false;
  cloud_uploaded_ = false;
  plan_resident_ = false;
}

bool Session::CloudToken::operator==(const CloudToken& o) const {
  return object == o.object && storage == o.storage && size == o.size && stamp == o.stamp && seq == o.seq &&
         std::memcmp(probe, o.probe, sizeof(probe)) == 0;
}

Session::CloudToken Session::tokenOf(const pcl::PointCloud<pcl::PointXYZI>::ConstPtr& cloud) {
  CloudToken t;
  if (!cloud) return t;
  t.object = cloud.get();
  t.size = cloud->points.size();
  t.storage = t.size ? (const void*)cloud->points.data() : nullptr;
  t.stamp = cloud->header.stamp;
  t.seq = cloud->header.seq;
  if (t.size) {
    const std::size_t pick[3] = {0, t.size / 2, t.size - 1};
    for (int k = 0; k < 3; ++k) {
      const auto& p = cloud->points[pick[k]];
      t.probe[3 * k] = p.x; t.probe[3 * k + 1] = p.y; t.probe[3 * k + 2] = p.z;
    }
  }
  return t;
}

void Session::uploadCloud(const pcl::PointCloud<pcl::PointXYZI>::ConstPtr& cloud) {
  cloud_ = cloud;
  cloud_token_ = tokenOf(cloud);
  const size_t n = cloud ? cloud->points.size() : 0;
  const int rc = b200lp_set_cloud(ctx_, n ? (const void*)cloud->points.data() : nullptr, n, sizeof(pcl::PointXYZI));
  if (rc != B200LP_OK) raise(rc, "b200lp_set_cloud");
  cloud_uploaded_ = true;
  cloud_from_device_ = false;
}

void Session::setObservation(const pcl::PointCloud<pcl::PointXYZI>::ConstPtr& cloud) {
  std::lock_guard<std::mutex> lk(mu_);
  ensureContext();
  // the host copy of a device-side aggregate (aggregateObservations) is already where the kernels read it
  if (!(cloud_from_device_ && cloud_uploaded_ && cloud && tokenOf(cloud) == cloud_token_)) uploadCloud(cloud);
  launched_ = false;
}

namespace {
void transform_to_array(const geometry_msgs::msg::TransformStamped& t, double out[7]) {
  const auto& tr = t.transform;
  const double v[7] = {tr.translation.x, tr.translation.y, tr.translation.z, tr.rotation.x, tr.rotation.y, tr.rotation.z, tr.rotation.w};
  std::memcpy(out, v, sizeof(v));
}
}  // namespace

b200lp_observation_info Session::sensorObservation(int sensor, const pcl::PointCloud<pcl::PointXYZ>& scan,
                                                   const geometry_msgs::msg::TransformStamped& trans_b2s,
                                                   const geometry_msgs::msg::TransformStamped& trans_gbl2b,
                                                   const b200lp_sensor_params& sp) {
  std::lock_guard<std::mutex> lk(mu_);
  ensureContext();
  double b2s[7], g2b[7];
  transform_to_array(trans_b2s, b2s);
  transform_to_array(trans_gbl2b, g2b);
  b200lp_observation_info info{};
  const int rc = b200lp_sensor_observation(ctx_, sensor, scan.points.empty() ? nullptr : (const void*)scan.points.data(),
                                           scan.points.size(), sizeof(pcl::PointXYZ), b2s, g2b, &sp, &info);
  if (rc != B200LP_OK) raise(rc, "b200lp_sensor_observation");
  return info;
}

void Session::readObservation(int sensor, pcl::PointCloud<pcl::PointXYZI>& out) {
  std::lock_guard<std::mutex> lk(mu_);
  ensureContext();
  size_t n = 0;
  int rc = b200lp_read_observation(ctx_, sensor, nullptr, 0, sizeof(pcl::PointXYZI), &n);
  if (rc != B200LP_OK && n == 0) raise(rc, "b200lp_read_observation");
  out.points.assign(n, pcl::PointXYZI());
  out.width = (uint32_t)n;
  if (n) {
    rc = b200lp_read_observation(ctx_, sensor, out.points.data(), n, sizeof(pcl::PointXYZI), &n);
    if (rc != B200LP_OK) raise(rc, "b200lp_read_observation");
  }
}

void Session::aggregateObservations(const std::vector<int>& sensors, const pcl::PointCloud<pcl::PointXYZI>::Ptr& aggregate) {
  if (!aggregate) throw Error(B200LP_E_INVALID, "b200lp::Session::aggregateObservations: null aggregate cloud");
  aggregate->points.clear();
  for (int s : sensors) {  // the host copy, sensor by sensor (`*aggregate += *plugin->getObservation()`)
    pcl::PointCloud<pcl::PointXYZI> one;
    readObservation(s, one);
    *aggregate += one;
  }
  std::lock_guard<std::mutex> lk(mu_);
  std::vector<int32_t> ids(sensors.begin(), sensors.end());
  size_t total = 0;
  const int rc = b200lp_aggregate_observations(ctx_, ids.data(), (int)ids.size(), &total);
  if (rc != B200LP_OK) raise(rc, "b200lp_aggregate_observations");
  if (total != aggregate->points.size()) throw Error(B200LP_E_STATE, "b200lp::Session::aggregateObservations: host copy and device aggregate differ in size");
  cloud_ = aggregate;
  cloud_token_ = tokenOf(aggregate);
  cloud_uploaded_ = true;
  cloud_from_device_ = true;
  launched_ = false;
}

void Session::beginCycle(const geometry_msgs::msg::TransformStamped& robot_pose, const nav_msgs::msg::Odometry& robot_state,
                         const nav_msgs::msg::Path& prune_plan, double current_allowed_max_linear_speed) {
  std::lock_guard<std::mutex> lk(mu_);
  const auto& t = robot_pose.transform;
  const double pose[7] = {t.translation.x, t.translation.y, t.translation.z, t.rotation.x, t.rotation.y, t.rotation.z, t.rotation.w};
  std::memcpy(query_.pose, pose, sizeof(pose));
  query_.twist[0] = robot_state.twist.twist.linear.x;
  query_.twist[1] = robot_state.twist.twist.linear.y;
  query_.twist[2] = robot_state.twist.twist.angular.z;
  query_.max_speed_override = current_allowed_max_linear_speed;
  // heading_deviation stays what the critics' shared data last said; criticScore() re-launches if it changed
  plan7_.clear();
  plan7_.reserve(prune_plan.poses.size() * 7);
  for (const auto& ps : prune_plan.poses) {
    const auto& p = ps.pose;
    const double row[7] = {p.position.x, p.position.y, p.position.z, p.orientation.x, p.orientation.y, p.orientation.z, p.orientation.w};
    plan7_.insert(plan7_.end(), row, row + 7);
  }
  in_cycle_ = true;
  launched_ = false;
  points_loaded_ = false;
  launches_this_cycle_ = 0;
  ++cycle_;
}

void Session::setGlobalPlan(const std::vector<double>& poses7) {
  std::lock_guard<std::mutex> lk(mu_);
  ensureContext();
  const int rc = b200lp_set_global_plan(ctx_, poses7.data(), poses7.size() / 7);
  if (rc != B200LP_OK) raise(rc, "b200lp_set_global_plan");
}

b200lp_prune_info Session::prunePlan(const double robot_xyz[3], double forward_distance, double backward_distance,
                                     std::vector<double>& poses7, std::vector<float>& pcl_xyzi) {
  std::lock_guard<std::mutex> lk(mu_);
  ensureContext();
  b200lp_prune_info info{};
  int rc = b200lp_prune_plan(ctx_, robot_xyz, forward_distance, backward_distance, &info);
  if (rc != B200LP_OK) raise(rc, "b200lp_prune_plan");
  if (info.status == 1) return info;  // the reference returns before touching the prune plan
  poses7.assign((size_t)info.n_prune * 7, 0.0);
  pcl_xyzi.assign((size_t)info.n_prune * 4, 0.f);
  if (info.n_prune) {
    rc = b200lp_read_prune_plan(ctx_, poses7.data(), pcl_xyzi.data(), (size_t)info.n_prune);
    if (rc != B200LP_OK) raise(rc, "b200lp_read_prune_plan");
  }
  plan7_device_ = poses7;
  plan_resident_ = true;
  launched_ = false;
  return info;
}

b200lp_blocked Session::pathBlocked(double check_radius) {
  std::lock_guard<std::mutex> lk(mu_);
  ensureContext();
  if (!cloud_uploaded_) uploadCloud(cloud_);
  b200lp_blocked b{};
  const int rc = b200lp_path_blocked(ctx_, check_radius, &b);
  if (rc != B200LP_OK) raise(rc, "b200lp_path_blocked");
  return b;
}

void Session::launch() {
  ensureContext();
  if (!cloud_uploaded_) uploadCloud(cloud_);  // after a context rebuild (or never set: the empty cloud)
  int rc = B200LP_OK;
  // the device-side prune plan stays where it is when the generator was handed exactly that plan
  if (!(plan_resident_ && plan7_ == plan7_device_)) {
    rc = b200lp_set_plan(ctx_, plan7_.empty() ? nullptr : plan7_.data(), plan7_.size() / 7);
    if (rc != B200LP_OK) raise(rc, "b200lp_set_plan");
    plan_resident_ = false;
  }
  rc = b200lp_plan(ctx_, &query_, &result_);
  if (rc != B200LP_OK) raise(rc, "b200lp_plan");
  const size_t n = (size_t)result_.n_traj, nc = critics_.size();
  vel_.assign(n * 3, 0.f);
  steps_.assign(n, 0);
  dt_.assign(n, 0.0);
  cost_.assign(n, 0.0);
  scores_.assign(n * std::max<size_t>(nc, 1), std::numeric_limits<double>::quiet_NaN());
  if (n) {
    b200lp_traj_view v{};
    v.vel = vel_.data();
    v.num_steps = steps_.data();
    v.time_delta = dt_.data();
    v.cost = cost_.data();
    v.critic_scores = nc ? scores_.data() : nullptr;
    rc = b200lp_read_trajectories(ctx_, 0, &v);
    if (rc != B200LP_OK) raise(rc, "b200lp_read_trajectories");
  }
  launched_ = true;
  ++launches_this_cycle_;
}

void Session::ensureLaunched() {
  if (!in_cycle_) throw Error(B200LP_E_STATE, "b200lp::Session: no cycle is open (the generator's initialise() has not run)");
  if (!launched_) launch();
}

void Session::loadPoints() {
  if (points_loaded_) return;
  const size_t n = (size_t)result_.n_traj;
  pose_off_.assign(n + 1, 0);
  size_t total = 0;
  for (size_t i = 0; i < n; ++i) total += (size_t)steps_[i];
  pose7_.assign(total * 7, 0.0);
  pcl3_.assign(total * 3, 0.f);
  cuboid24_.assign(total * 24, 0.f);
  aabb6_.assign(total * 6, 0.f);
  if (n) {
    b200lp_pose_view v{};
    v.pose = pose7_.data();
    v.pcl_pose = pcl3_.data();
    v.cuboid = cuboid24_.data();
    v.aabb = aabb6_.data();
    const int rc = b200lp_read_pose_batch(ctx_, 0, 0, (int32_t)n, pose_off_.data(), &v, total);
    if (rc != B200LP_OK) raise(rc, "b200lp_read_pose_batch");
  }
  points_loaded_ = true;
}

int Session::trajectoryCount() {
  std::lock_guard<std::mutex> lk(mu_);
  ensureLaunched();
  return result_.n_traj;
}

void Session::fillTrajectory(int id, base_trajectory::Trajectory& traj, bool with_points) {
  std::lock_guard<std::mutex> lk(mu_);
  ensureLaunched();
  if (id < 0 || id >= result_.n_traj) throw Error(B200LP_E_INVALID, "b200lp::Session: trajectory id out of range");
  // what generateTrajectory() leaves in the object (dd_simple…cpp:358,401-402): velocities from the float sample, cost 0
  traj.xv_ = (double)vel_[3 * id];
  traj.yv_ = (theory_.params.theory == B200LP_THEORY_OMNI_SIMPLE) ? (double)vel_[3 * id + 1] : 0.0;
  traj.thetav_ = (double)vel_[3 * id + 2];
  traj.time_delta_ = dt_[id];
  traj.cost_ = 0.0;
  traj.id_ = id;
  traj.resetPoints();
  if (with_points) {
    loadPoints();
    const size_t r = (size_t)pose_off_[id];
    traj.assignPoints(&pose7_[r * 7], &pcl3_[r * 3], &cuboid24_[r * 24], &aabb6_[r * 6], (unsigned int)steps_[id]);
  }
}

double Session::criticScore(int critic_index, const base_trajectory::Trajectory& traj,
                            const pcl::PointCloud<pcl::PointXYZI>::ConstPtr& pcl_perception, double heading_deviation) {
  std::lock_guard<std::mutex> lk(mu_);
  if (!in_cycle_) throw Error(B200LP_E_STATE, "b200lp critic: scoreTrajectory before the generator's initialise()");
  // the critics' view of the world is authoritative: (re)launch if the cycle ran against something else
  if (pcl_perception && !(tokenOf(pcl_perception) == cloud_token_)) {
    ensureContext();
    uploadCloud(pcl_perception);
    launched_ = false;
  }
  if (heading_deviation != query_.heading_deviation) {
    query_.heading_deviation = heading_deviation;
    launched_ = false;
  }
  if (!launched_) launch();
  const int id = traj.id_;
  if (id < 0 || id >= result_.n_traj || (float)traj.xv_ != vel_[3 * id] || (float)traj.thetav_ != vel_[3 * id + 2])
    throw Error(B200LP_E_INVALID, "b200lp critic: the trajectory was not produced by this cycle's B200 generator (id_ mismatch)");
  if (critic_index < 0 || critic_index >= (int)critics_.size()) throw Error(B200LP_E_INVALID, "b200lp critic: bad stack index");
  return scores_[(size_t)id * critics_.size() + critic_index];
}

const b200lp_result& Session::result() {
  std::lock_guard<std::mutex> lk(mu_);
  ensureLaunched();
  return result_;
}
}  // namespace b200lp

// =====================================================================================================
// generator adapters
// =====================================================================================================
namespace trajectory_generators {

void B200TheoryBase::readParameters(int theory) {
  b200lp::TheoryConfig c;
  c.params.theory = theory;
  auto dbl = [&](const char* key, double def, double& out) {
    node_->declare_parameter(name_ + key, rclcpp::ParameterValue(def));
    node_->get_parameter(name_ + key, out);
  };
  b200lp_limits& L = c.limits;
  b200lp_params& P = c.params;
  // names and defaults: dd_simple…cpp:47-133, omni_simple…cpp:47-157, dd_rotate_inplace_theory.cpp:47-127
  dbl(".min_vel_x", 0.01, L.min_vel_x);
  dbl(".max_vel_x", 0.1, L.max_vel_x);
  if (theory == B200LP_THEORY_OMNI_SIMPLE) {
    dbl(".min_vel_y", 0.01, L.min_vel_y);
    dbl(".max_vel_y", 0.1, L.max_vel_y);
    dbl(".min_vel_trans", 0.01, L.min_vel_trans);
    dbl(".max_vel_trans", 0.1, L.max_vel_trans);
  }
  dbl(".min_vel_theta", 0.1, L.min_vel_theta);
  dbl(".max_vel_theta", 0.1, L.max_vel_theta);
  dbl(".acc_lim_x", 0.3, L.acc_lim_x);
  if (theory == B200LP_THEORY_OMNI_SIMPLE) dbl(".acc_lim_y", 0.3, L.acc_lim_y);
  dbl(".acc_lim_theta", 0.5, L.acc_lim_theta);
  double prune_forward, prune_backward;  // read by the caller's prunePlan(), declared here like the reference does
  dbl(".prune_forward", 3.0, prune_forward);
  dbl(".prune_backward", 1.0, prune_backward);
  L.deceleration_ratio = 2.0;
  if (theory != B200LP_THEORY_DD_ROTATE_INPLACE) {
    dbl(".deceleration_ratio", 2.0, L.deceleration_ratio);
    bool umc = false;
    node_->declare_parameter(name_ + ".use_motor_constraint", rclcpp::ParameterValue(false));
    node_->get_parameter(name_ + ".use_motor_constraint", umc);
    L.use_motor_constraint = umc ? 1 : 0;
  } else {
    L.use_motor_constraint = 1;  // the rotate theory always applies the wheel-rpm test (dd_rotate_inplace_theory.cpp:259-268)
  }
  dbl(".max_motor_shaft_rpm", 3000.0, L.max_motor_shaft_rpm);
  dbl(".wheel_diameter", 0.15, L.wheel_diameter);
  dbl(".gear_ratio", 30.0, L.gear_ratio);
  dbl(".robot_radius", 0.25, L.robot_radius);
  dbl(".controller_frequency", 10.0, P.controller_frequency);
  dbl(".sim_time", 2.0, P.sim_time);
  dbl(".linear_x_sample", 10.0, P.linear_x_sample);
  P.linear_y_sample = 0.0;
  if (theory == B200LP_THEORY_OMNI_SIMPLE) dbl(".linear_y_sample", 10.0, P.linear_y_sample);
  dbl(".angular_z_sample", 10.0, P.angular_z_sample);
  dbl(".sim_granularity", 0.1, P.sim_granularity);
  dbl(".angular_sim_granularity", 0.05, P.angular_sim_granularity);
  if (theory == B200LP_THEORY_DD_ROTATE_INPLACE) dbl(".rotation_speed", 0.4, L.rotation_speed);
  // cuboid: 8 named vertices, stored in the load-bearing order blb,brb,blt,flb,brt,frt,flt,frb (dd_simple…cpp:211-218)
  static const char* kOrder[8] = {"blb", "brb", "blt", "flb", "brt", "frt", "flt", "frb"};
  for (int k = 0; k < 8; ++k) {
    const std::string key = name_ + ".cuboid." + kOrder[k];
    node_->declare_parameter(key, rclcpp::PARAMETER_DOUBLE_ARRAY);
    const std::vector<double> v = node_->get_parameter(key).as_double_array();
    if (v.size() != 3) throw std::runtime_error("parameter '" + key + "' must hold 3 numbers");
    for (int a = 0; a < 3; ++a) c.cuboid[k][a] = (float)v[a];
  }
  node_->declare_parameter(name_ + ".b200_materialize_points", rclcpp::ParameterValue(true));
  node_->get_parameter(name_ + ".b200_materialize_points", materialize_points_);
  session_ = b200lp::Session::forGenerator(name_);
  session_->configureTheory(c);
}

void B200TheoryBase::initialise() {
  next_ = 0;
  session_->beginCycle(shared_data_->robot_pose_, shared_data_->robot_state_, shared_data_->prune_plan_,
                       shared_data_->current_allowed_max_linear_speed_);
}
bool B200TheoryBase::hasMoreTrajectories() { return next_ < session_->trajectoryCount(); }
bool B200TheoryBase::nextTrajectory(base_trajectory::Trajectory& _traj) {
  if (!hasMoreTrajectories()) return false;
  session_->fillTrajectory(next_, _traj, materialize_points_);
  ++next_;
  return true;
}

void Trajectory_Generators_ROS::initial() {
  auto& factory = b200lp::PluginFactory<TrajectoryGeneratorTheory>::instance();
  this->declare_parameter("plugins", rclcpp::PARAMETER_STRING_ARRAY);
  plugins_ = this->get_parameter("plugins").as_string_array();
  for (const auto& pname : plugins_) {
    const std::string key = pname + ".plugin";
    this->declare_parameter(key, rclcpp::ParameterValue(""));
    const std::string type = this->get_parameter(key).as_string();
    std::shared_ptr<TrajectoryGeneratorTheory> plugin = factory.createSharedInstance(type);
    stacked_generator_.addPlugin(pname, plugin);
    plugin->initialize(pname, shared_from_this());
  }
}

}  // namespace trajectory_generators

// =====================================================================================================
// critic adapters
// =====================================================================================================
namespace mpc_critics {

void B200ModelBase::bind(int kind) {
  b200lp_critic c{};
  c.kind = kind;
  node_->declare_parameter(name_ + ".weight", rclcpp::ParameterValue(1.0));
  node_->get_parameter(name_ + ".weight", weight_);
  c.weight = weight_;
  if (kind == B200LP_CRITIC_PURE_PURSUIT) {  // pure_pursuit_model.cpp:49-55
    node_->declare_parameter(name_ + ".translation_weight", rclcpp::ParameterValue(0.5));
    node_->get_parameter(name_ + ".translation_weight", c.translation_weight);
    node_->declare_parameter(name_ + ".orientation_weight", rclcpp::ParameterValue(0.5));
    node_->get_parameter(name_ + ".orientation_weight", c.orientation_weight);
  }
  // the loader has declared `<name>.trajectory_generator` already (mpc_critics_ros.cpp:71-73)
  std::string generator;
  if (!node_->has_parameter(name_ + ".trajectory_generator"))
    node_->declare_parameter(name_ + ".trajectory_generator", rclcpp::ParameterValue(""));
  node_->get_parameter(name_ + ".trajectory_generator", generator);
  session_ = b200lp::Session::forGenerator(generator);
  index_ = session_->addCritic(c);
}

double B200ModelBase::scoreTrajectory(base_trajectory::Trajectory& traj) {
  return session_->criticScore(index_, traj, shared_data_->pcl_perception_, shared_data_->heading_deviation_);
}

void MPC_Critics_ROS::initial() {
  auto& factory = b200lp::PluginFactory<ScoringModel>::instance();
  this->declare_parameter("plugins", rclcpp::PARAMETER_STRING_ARRAY);
  plugins_ = this->get_parameter("plugins").as_string_array();
  for (const auto& pname : plugins_) {
    const std::string key = pname + ".plugin";
    this->declare_parameter(key, rclcpp::ParameterValue(""));
    const std::string type = this->get_parameter(key).as_string();
    const std::string gkey = pname + ".trajectory_generator";
    this->declare_parameter(gkey, rclcpp::ParameterValue(""));
    const std::string generator = this->get_parameter(gkey).as_string();
    std::shared_ptr<ScoringModel> plugin = factory.createSharedInstance(type);
    stacked_scoring_model_.addPluginByTraj(generator, plugin);
    plugin->initialize(pname, shared_from_this());
  }
}

}  // namespace mpc_critics

// =====================================================================================================
// plugin registration under the reference's type strings
// =====================================================================================================
namespace {
struct RegisterPlugins {
  RegisterPlugins() {
    using namespace trajectory_generators;
    using namespace mpc_critics;
    auto& tg = b200lp::PluginFactory<TrajectoryGeneratorTheory>::instance();
    tg.add("trajectory_generators::DDSimpleTrajectoryGeneratorTheory", [] { return std::make_shared<DDSimpleTrajectoryGeneratorTheory>(); });
    tg.add("trajectory_generators::OmniSimpleTrajectoryGeneratorTheory", [] { return std::make_shared<OmniSimpleTrajectoryGeneratorTheory>(); });
    tg.add("trajectory_generators::DDRotateInplaceTheory", [] { return std::make_shared<DDRotateInplaceTheory>(); });
    auto& mc = b200lp::PluginFactory<ScoringModel>::instance();
    mc.add("mpc_critics::CollisionModel", [] { return std::make_shared<CollisionModel>(); });
    mc.add("mpc_critics::CollisionMinMaxModel", [] { return std::make_shared<CollisionMinMaxModel>(); });
    mc.add("mpc_critics::StickPathModel", [] { return std::make_shared<StickPathModel>(); });
    mc.add("mpc_critics::PurePursuitModel", [] { return std::make_shared<PurePursuitModel>(); });
    mc.add("mpc_critics::TowardGlobalPlanModel", [] { return std::make_shared<TowardGlobalPlanModel>(); });
    mc.add("mpc_critics::ShortestAngleModel", [] { return std::make_shared<ShortestAngleModel>(); });
    mc.add("mpc_critics::TwirlingModel", [] { return std::make_shared<TwirlingModel>(); });
  }
} g_register_plugins;
}  // namespace

// =====================================================================================================
// the cycle driver (the caller of the path)
// =====================================================================================================
namespace perception_3d {
void MultiLayerSpinningLidar::cbSensor(const pcl::PointCloud<pcl::PointXYZ>& pcl_msg,
                                       const geometry_msgs::msg::TransformStamped& trans_b2s,
                                       const geometry_msgs::msg::TransformStamped& trans_gbl2b) {
  b200lp_sensor_params sp{};
  sp.perception_window_size = perception_window_size_;
  sp.marking_height = marking_height_;
  sp.leaf_size = 0.1f;  // sor.setLeafSize(0.1f, 0.1f, 0.1f) (:254)
  sp.is_local_planner = is_local_planner_ ? 1 : 0;
  if (stitcher_num_ <= 0) {  // "if not stitch, save copy time" (:180-183)
    last_info_ = b200lp::Session::forGenerator(traj_gen_name_)->sensorObservation(slot_, pcl_msg, trans_b2s, trans_gbl2b, sp);
  } else {  // :184-199 — keep the last stitcher_num scans, filter their concatenation
    if (pcl_stitcher_.size() >= (size_t)stitcher_num_) pcl_stitcher_.pop_front();
    pcl_stitcher_.push_back(pcl_msg);
    pcl::PointCloud<pcl::PointXYZ> stitched;
    for (const auto& scan : pcl_stitcher_) stitched += scan;
    last_info_ = b200lp::Session::forGenerator(traj_gen_name_)->sensorObservation(slot_, stitched, trans_b2s, trans_gbl2b, sp);
  }
  observation_stale_ = true;
}

pcl::PointCloud<pcl::PointXYZI>::Ptr MultiLayerSpinningLidar::getObservation() {
  if (observation_stale_) {
    b200lp::Session::forGenerator(traj_gen_name_)->readObservation(slot_, *sensor_current_observation_);
    observation_stale_ = false;
  }
  return sensor_current_observation_;
}

void StackedPerception::aggregateObservations() {
  shared_data_->aggregate_observation_.reset(new pcl::PointCloud<pcl::PointXYZI>);  // :130
  if (plugins_.empty()) return;
  std::vector<int> slots;
  for (const auto& p : plugins_) slots.push_back(p->slot());
  b200lp::Session::forGenerator(plugins_.front()->generatorName())->aggregateObservations(slots, shared_data_->aggregate_observation_);
}

void PathBlockedStrategy::selfMark(const std::string& traj_gen_name) {
  const b200lp_blocked b = b200lp::Session::forGenerator(traj_gen_name)->pathBlocked(check_radius_);
  prune_plan_blocked_ratio_ = b.ratio;
  opinion_ = b.opinion ? PATH_BLOCKED_WAIT : PASS;
}
}  // namespace perception_3d

namespace local_planner {

void Local_Planner::initial(const std::shared_ptr<perception_3d::SharedData>& perception_3d,
                            const std::shared_ptr<mpc_critics::MPC_Critics_ROS>& mpc_critics,
                            const std::shared_ptr<trajectory_generators::Trajectory_Generators_ROS>& trajectory_generators) {
  perception_3d_ = perception_3d;
  mpc_critics_ros_ = mpc_critics;
  trajectory_generators_ros_ = trajectory_generators;
}

void Local_Planner::setPlan(const std::vector<geometry_msgs::msg::PoseStamped>& orig_global_plan, const std::string& traj_gen_name) {
  if (orig_global_plan.size() < 3) return;  // "Size of global plan is smaller than 3." (:324-327)
  global_plan_ = orig_global_plan;
  std::vector<double> g7;
  g7.reserve(global_plan_.size() * 7);
  for (const auto& ps : global_plan_) {
    const auto& p = ps.pose;
    const double row[7] = {p.position.x, p.position.y, p.position.z, p.orientation.x, p.orientation.y, p.orientation.z, p.orientation.w};
    g7.insert(g7.end(), row, row + 7);
  }
  b200lp::Session::forGenerator(traj_gen_name)->setGlobalPlan(g7);
}

void Local_Planner::prunePlan(double forward_distance, double backward_distance, const std::string& traj_gen_name) {
  if (global_plan_.size() < 3) return;  // :376-377
  const double xyz[3] = {trans_gbl2b_.transform.translation.x, trans_gbl2b_.transform.translation.y, trans_gbl2b_.transform.translation.z};
  std::vector<double> p7;
  std::vector<float> pcl4;
  const b200lp_prune_info info = b200lp::Session::forGenerator(traj_gen_name)->prunePlan(xyz, forward_distance, backward_distance, p7, pcl4);
  if (info.status == 1) return;
  prune_plan_.poses.clear();  // :379-380
  pcl_prune_plan_.points.clear();
  for (int i = 0; i < info.n_prune; ++i) {
    geometry_msgs::msg::PoseStamped ps;
    ps.pose.position.x = p7[i * 7]; ps.pose.position.y = p7[i * 7 + 1]; ps.pose.position.z = p7[i * 7 + 2];
    ps.pose.orientation.x = p7[i * 7 + 3]; ps.pose.orientation.y = p7[i * 7 + 4]; ps.pose.orientation.z = p7[i * 7 + 5]; ps.pose.orientation.w = p7[i * 7 + 6];
    prune_plan_.poses.push_back(ps);
    pcl::PointXYZI pt;
    pt.x = pcl4[i * 4]; pt.y = pcl4[i * 4 + 1]; pt.z = pcl4[i * 4 + 2]; pt.intensity = pcl4[i * 4 + 3];
    pcl_prune_plan_.points.push_back(pt);
  }
  if (perception_3d_) perception_3d_->pcl_prune_plan_ = pcl_prune_plan_;  // :518
}

void Local_Planner::getBestTrajectory(std::string traj_gen_name, base_trajectory::Trajectory& best_traj) {
  best_traj.cost_ = -1;  // in case every trajectory is rejected
  double minimum_cost = 9999999;
  for (auto& traj : *trajectories_) {
    mpc_critics_ros_->scoreTrajectory(traj_gen_name, traj);
    if (traj.cost_ >= 0 && traj.cost_ <= minimum_cost) {  // `<=`: ties go to the LAST trajectory (local_planner.cpp:460)
      best_traj = traj;
      minimum_cost = traj.cost_;
    }
  }
}

dddmr_sys_core::PlannerState Local_Planner::computeVelocityCommand(std::string traj_gen_name, base_trajectory::Trajectory& best_traj) {
  if (!got_odom_ || !got_pose_) return dddmr_sys_core::TF_FAIL;
  if (!perception_3d_ || !perception_3d_->aggregate_observation_) return dddmr_sys_core::PERCEPTION_MALFUNCTION;

  // local_planner.cpp:528-535 — seed the generators' shared data, open the cycle
  auto tg = trajectory_generators_ros_->getSharedDataPtr();
  tg->robot_pose_ = trans_gbl2b_;
  tg->robot_state_ = robot_state_;
  tg->prune_plan_ = prune_plan_;
  tg->current_allowed_max_linear_speed_ = perception_3d_->current_allowed_max_linear_speed_;
  if (early_observation_) b200lp::Session::forGenerator(traj_gen_name)->setObservation(perception_3d_->aggregate_observation_);
  trajectory_generators_ros_->initializeTheories_wi_Shared_data();

  // :549-557 — queue every trajectory
  trajectories_ = std::make_shared<std::vector<base_trajectory::Trajectory>>();
  while (trajectory_generators_ros_->hasMoreTrajectories(traj_gen_name)) {
    base_trajectory::Trajectory a_traj;
    if (trajectory_generators_ros_->nextTrajectory(traj_gen_name, a_traj)) trajectories_->push_back(a_traj);
  }

  // :577-587 — seed the critics' shared data, score, pick
  {
    std::unique_lock<mpc_critics::StackedScoringModel::model_mutex_t> critics_lock(*(mpc_critics_ros_->getStackedScoringModelPtr()->getMutex()));
    auto mc = mpc_critics_ros_->getSharedDataPtr();
    mc->robot_pose_ = trans_gbl2b_;
    mc->robot_state_ = robot_state_;
    mc->pcl_perception_ = perception_3d_->aggregate_observation_;
    mc->prune_plan_ = prune_plan_;
    mpc_critics_ros_->updateSharedData();
    getBestTrajectory(traj_gen_name, best_traj);
  }
  // :597-607 — loop the perception opinions (here: the optional path-blocked strategy)
  if (path_blocked_) {
    path_blocked_->selfMark(traj_gen_name);
    if (path_blocked_->getOpinion() == perception_3d::PATH_BLOCKED_WAIT) return dddmr_sys_core::PATH_BLOCKED_WAIT;
    if (path_blocked_->getOpinion() == perception_3d::PATH_BLOCKED_REPLANNING) return dddmr_sys_core::PATH_BLOCKED_REPLANNING;
  }
  return best_traj.cost_ < 0 ? dddmr_sys_core::ALL_TRAJECTORIES_FAIL : dddmr_sys_core::TRAJECTORY_FOUND;
}

}  // namespace local_planner

// =============================================================================================
// recovery_behaviors::RotateInPlaceBehavior (rotate_inplace_behavior.cpp:137-305)
// =============================================================================================
namespace recovery_behaviors {

double yaw_of(double x, double y, double z, double w) {  // tf2::impl::getYaw (tf2/impl/utils.h)
  const double sqx = x * x, sqy = y * y, sqz = z * z, sqw = w * w;
  const double sarg = -2 * (x * z - w * y) / (sqx + sqy + sqz + sqw);  // normalization added from urdfom_headers
  if (sarg <= -0.99999) return -2 * std::atan2(y, x);
  if (sarg >= 0.99999) return 2 * std::atan2(y, x);
  return std::atan2(2 * (x * y + w * z), sqw + sqx - sqy - sqz);
}

double shortest_angular_distance(double from, double to) {  // angles/angles.h: normalize_angle(to - from)
  const double result = std::fmod((to - from) + M_PI, 2.0 * M_PI);
  if (result <= 0.0) return result + M_PI;
  return result - M_PI;
}

void RotateInPlaceBehavior::initial(const std::shared_ptr<perception_3d::SharedData>& perception_3d,
                                    const std::shared_ptr<mpc_critics::MPC_Critics_ROS>& mpc_critics,
                                    const std::shared_ptr<trajectory_generators::Trajectory_Generators_ROS>& trajectory_generators,
                                    const std::string& trajectory_generator_name, double tolerance, double frequency) {
  perception_3d_ = perception_3d;
  mpc_critics_ros_ = mpc_critics;
  trajectory_generators_ros_ = trajectory_generators;
  trajectory_generator_name_ = trajectory_generator_name;
  tolerance_ = tolerance;
  frequency_ = frequency;
}

void RotateInPlaceBehavior::getBestTrajectory(const std::string& traj_gen_name, base_trajectory::Trajectory& best_traj) {
  best_traj.cost_ = -1;  // :82
  double minimum_cost = 9999999;
  for (auto& traj : *trajectories_) {
    mpc_critics_ros_->scoreTrajectory(traj_gen_name, traj);
    if (traj.cost_ >= 0 && traj.cost_ <= minimum_cost) {  // :92
      best_traj = traj;
      minimum_cost = traj.cost_;
    }
  }
}

void RotateInPlaceBehavior::begin(const geometry_msgs::msg::TransformStamped& t, double now_s) {
  current_angle_ = yaw_of(t.transform.rotation.x, t.transform.rotation.y, t.transform.rotation.z, t.transform.rotation.w);
  start_angle_ = current_angle_;
  got_180_ = false;
  last_valid_control_ = now_s;
}

RotateInPlaceBehavior::Step RotateInPlaceBehavior::step(const geometry_msgs::msg::TransformStamped& trans_gbl2b,
                                                        const nav_msgs::msg::Odometry& robot_state, double now_s) {
  Step s;
  // :140-142 — the loop condition
  if (!(!got_180_ || std::fabs(shortest_angular_distance(current_angle_, start_angle_)) > tolerance_)) {
    s.finished = true;
    s.got_180 = got_180_;
    return s;
  }
  // :186-193 — aggregateObservations() has run (the embedding code's perception loop); the pose is read
  current_angle_ = yaw_of(trans_gbl2b.transform.rotation.x, trans_gbl2b.transform.rotation.y, trans_gbl2b.transform.rotation.z,
                          trans_gbl2b.transform.rotation.w);
  // :203-219 — the distance left to rotate
  if (!got_180_) {
    const double distance_to_180 = std::fabs(shortest_angular_distance(current_angle_, start_angle_ + M_PI));
    s.dist_left = M_PI + distance_to_180;
    if (distance_to_180 < tolerance_) got_180_ = true;
  } else {
    s.dist_left = std::fabs(shortest_angular_distance(current_angle_, start_angle_));
  }
  // :223-239 — open the cycle, queue every trajectory
  auto tg = trajectory_generators_ros_->getSharedDataPtr();
  tg->robot_pose_ = trans_gbl2b;
  tg->robot_state_ = robot_state;
  trajectory_generators_ros_->initializeTheories_wi_Shared_data();
  trajectories_ = std::make_shared<std::vector<base_trajectory::Trajectory>>();
  while (trajectory_generators_ros_->hasMoreTrajectories(trajectory_generator_name_)) {
    base_trajectory::Trajectory a_traj;
    if (trajectory_generators_ros_->nextTrajectory(trajectory_generator_name_, a_traj)) trajectories_->push_back(a_traj);
  }
  // :246-256 — the critics' shared data, the scores, and the RESET of the critics' cloud (it is a shared_ptr copied from
  // the perception stack: the behaviour must not keep it alive, and the next pass brings a new one)
  base_trajectory::Trajectory best_traj;
  {
    std::unique_lock<mpc_critics::StackedScoringModel::model_mutex_t> critics_lock(*(mpc_critics_ros_->getStackedScoringModelPtr()->getMutex()));
    auto mc = mpc_critics_ros_->getSharedDataPtr();
    mc->robot_pose_ = trans_gbl2b;
    mc->robot_state_ = robot_state;
    mc->pcl_perception_ = perception_3d_->aggregate_observation_;
    mpc_critics_ros_->updateSharedData();
    getBestTrajectory(trajectory_generator_name_, best_traj);
    mc->pcl_perception_.reset(new pcl::PointCloud<pcl::PointXYZI>);
  }
  s.best_id = best_traj.cost_ < 0 ? -1 : best_traj.id_;
  s.best_cost = best_traj.cost_;
  s.got_180 = got_180_;
  if (got_180_) {  // :258-268 — half a circle behind us and back within tolerance of 180: stop, succeed
    s.finished = true;
    s.result = dddmr_sys_core::RECOVERY_DONE;
    return s;
  }
  if (best_traj.cost_ < 0) {  // :270-289 — every trajectory rejected: stand still, give up after 5 s
    if (now_s - last_valid_control_ > 5.0) {
      s.finished = true;
      s.result = dddmr_sys_core::RECOVERY_FAIL;
    }
    return s;
  }
  s.cmd_linear_x = best_traj.xv_;      // :291-296
  s.cmd_angular_z = best_traj.thetav_;
  last_valid_control_ = now_s;
  return s;
}

}  // namespace recovery_behaviors

