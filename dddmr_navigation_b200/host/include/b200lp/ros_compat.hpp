// ros_compat.hpp — stand-ins for the handful of ROS 2 / PCL types the local-planner plugin interfaces name.
//
// The reference's plugin API (trajectory_generators::TrajectoryGeneratorTheory, mpc_critics::ScoringModel,
// base_trajectory::Trajectory) is written against rclcpp, geometry_msgs, nav_msgs and pcl. None of those exist
// in this build image, so the host layer compiles against the minimal look-alikes below: same namespaces, same
// member names, same memory layout where layout matters (pcl::PointXYZI = 32 bytes, pcl::PointXYZ = 16 bytes,
// SURVEY.md A6). With ROS 2 present define B200LP_HAVE_ROS2 and the real headers are used instead; nothing in
// the adapters depends on more than what is declared here.
#pragma once
#ifdef B200LP_HAVE_ROS2
#include <geometry_msgs/msg/pose_stamped.hpp>
#include <geometry_msgs/msg/transform_stamped.hpp>
#include <nav_msgs/msg/odometry.hpp>
#include <nav_msgs/msg/path.hpp>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <rclcpp/rclcpp.hpp>
#else
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

namespace builtin_interfaces::msg {
struct Time {
  int32_t sec = 0;
  uint32_t nanosec = 0;
};
}  // namespace builtin_interfaces::msg

namespace std_msgs::msg {
struct Header {
  builtin_interfaces::msg::Time stamp;
  std::string frame_id;
};
}  // namespace std_msgs::msg

namespace geometry_msgs::msg {
struct Point { double x = 0, y = 0, z = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { std_msgs::msg::Header header; Pose pose; };
struct PoseArray { std_msgs::msg::Header header; std::vector<Pose> poses; };
struct Transform { Vector3 translation; Quaternion rotation; };
struct TransformStamped { std_msgs::msg::Header header; std::string child_frame_id; Transform transform; };
struct Twist { Vector3 linear, angular; };
struct TwistWithCovariance { Twist twist; };
struct PoseWithCovariance { Pose pose; };
}  // namespace geometry_msgs::msg

namespace nav_msgs::msg {
struct Path { std_msgs::msg::Header header; std::vector<geometry_msgs::msg::PoseStamped> poses; };
struct Odometry {
  std_msgs::msg::Header header;
  std::string child_frame_id;
  geometry_msgs::msg::PoseWithCovariance pose;
  geometry_msgs::msg::TwistWithCovariance twist;
};
}  // namespace nav_msgs::msg

namespace pcl {
struct alignas(16) PointXYZ {
  float x = 0.f, y = 0.f, z = 0.f, pad_ = 1.f;
  PointXYZ() = default;
  PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) PointXYZI {
  float x = 0.f, y = 0.f, z = 0.f, pad_ = 1.f;
  float intensity = 0.f, pad2_[3] = {0.f, 0.f, 0.f};
};
static_assert(sizeof(PointXYZ) == 16 && sizeof(PointXYZI) == 32, "pcl point layout (SURVEY.md A6)");

struct PCLHeader {
  uint32_t seq = 0;
  uint64_t stamp = 0;
  std::string frame_id;
};

template <class PointT>
class PointCloud {
 public:
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  PCLHeader header;
  std::vector<PointT> points;
  uint32_t width = 0, height = 1;
  bool is_dense = true;
  void push_back(const PointT& p) {
    points.push_back(p);
    width = (uint32_t)points.size();
  }
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() {
    points.clear();
    width = 0;
  }
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
  PointCloud& operator+=(const PointCloud& o) {
    points.insert(points.end(), o.points.begin(), o.points.end());
    width = (uint32_t)points.size();
    return *this;
  }
};
}  // namespace pcl

namespace rclcpp {
// Parameter store with the declare/get surface the plugins' onInitialize() uses.
using ParameterVariant = std::variant<bool, int64_t, double, std::string, std::vector<double>, std::vector<std::string>>;

class ParameterValue {
 public:
  ParameterValue() = default;
  explicit ParameterValue(bool v) : v_(v) {}
  explicit ParameterValue(int v) : v_((int64_t)v) {}
  explicit ParameterValue(int64_t v) : v_(v) {}
  explicit ParameterValue(double v) : v_(v) {}
  explicit ParameterValue(const char* v) : v_(std::string(v)) {}
  explicit ParameterValue(const std::string& v) : v_(v) {}
  explicit ParameterValue(const std::vector<double>& v) : v_(v) {}
  explicit ParameterValue(const std::vector<std::string>& v) : v_(v) {}
  const ParameterVariant& get() const { return v_; }

 private:
  ParameterVariant v_ = false;
};

enum ParameterType { PARAMETER_DOUBLE_ARRAY, PARAMETER_STRING_ARRAY };

class Parameter {
 public:
  Parameter() = default;
  Parameter(std::string name, ParameterVariant v, bool set) : name_(std::move(name)), v_(std::move(v)), set_(set) {}
  bool is_set() const { return set_; }
  double as_double() const;
  bool as_bool() const;
  std::string as_string() const;
  std::vector<double> as_double_array() const;
  std::vector<std::string> as_string_array() const;

 private:
  std::string name_;
  ParameterVariant v_ = false;
  bool set_ = false;
};

class Node : public std::enable_shared_from_this<Node> {
 public:
  using SharedPtr = std::shared_ptr<Node>;
  using WeakPtr = std::weak_ptr<Node>;
  explicit Node(std::string name) : name_(std::move(name)) {}
  const std::string& get_name() const { return name_; }

  // YAML overrides (what `--params-file` provides): set programmatically or loaded with load_parameters_yaml()
  void set_parameter_override(const std::string& name, ParameterVariant v) { overrides_[name] = std::move(v); }
  // Minimal YAML subset: nested maps by indentation, scalars, inline lists `[a, b]`, `#` comments. Reads the subtree
  // `<node name>: ros__parameters:` of the text, flattening nested keys with '.' exactly like rclcpp does.
  void load_parameters_yaml(const std::string& yaml_text);

  void declare_parameter(const std::string& name, const ParameterValue& def);
  void declare_parameter(const std::string& name, ParameterType type);
  template <class T>
  void declare_parameter(const std::string& name, const T& def) { declare_parameter(name, ParameterValue(def)); }
  Parameter get_parameter(const std::string& name) const;
  bool get_parameter(const std::string& name, double& out) const;
  bool get_parameter(const std::string& name, bool& out) const;
  bool get_parameter(const std::string& name, std::string& out) const;
  bool has_parameter(const std::string& name) const { return declared_.count(name) != 0; }

 private:
  std::string name_;
  std::map<std::string, ParameterVariant> overrides_, declared_;
  std::map<std::string, bool> has_value_;
};
}  // namespace rclcpp
#endif  // B200LP_HAVE_ROS2
