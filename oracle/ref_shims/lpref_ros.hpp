// lpref_ros.hpp — just enough of rclcpp / geometry_msgs / nav_msgs for the reference's theory and critic sources to
// compile unchanged (TEST INFRASTRUCTURE, see lpref_eigen.hpp). Parameters come from overrides the driver sets.
#pragma once
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

namespace builtin_interfaces::msg { struct Time { int32_t sec = 0; uint32_t nanosec = 0; }; }
namespace std_msgs::msg { struct Header { builtin_interfaces::msg::Time stamp; std::string frame_id; }; }
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

namespace rclcpp {
using ParameterVariant = std::variant<bool, int64_t, double, std::string, std::vector<double>, std::vector<std::string>>;
class ParameterValue {
 public:
  ParameterValue() = default;
  explicit ParameterValue(bool v) : v_(v) {}
  explicit ParameterValue(int v) : v_((int64_t)v) {}
  explicit ParameterValue(double v) : v_(v) {}
  explicit ParameterValue(const char* v) : v_(std::string(v)) {}
  explicit ParameterValue(const std::string& v) : v_(v) {}
  const ParameterVariant& get() const { return v_; }
 private:
  ParameterVariant v_ = false;
};
enum ParameterType { PARAMETER_DOUBLE_ARRAY, PARAMETER_STRING_ARRAY };
class Parameter {
 public:
  Parameter() = default;
  Parameter(std::string n, ParameterVariant v) : name_(std::move(n)), v_(std::move(v)) {}
  double as_double() const { return std::holds_alternative<int64_t>(v_) ? (double)std::get<int64_t>(v_) : std::get<double>(v_); }
  bool as_bool() const { return std::get<bool>(v_); }
  std::string as_string() const { return std::get<std::string>(v_); }
  std::vector<double> as_double_array() const { return std::get<std::vector<double>>(v_); }
  std::vector<std::string> as_string_array() const { return std::get<std::vector<std::string>>(v_); }
 private:
  std::string name_;
  ParameterVariant v_ = false;
};
struct Logger {
  Logger get_child(const std::string&) const { return *this; }
  Logger get_logger() const { return *this; }
};
namespace node_interfaces {
struct NodeLoggingInterface {
  using SharedPtr = std::shared_ptr<NodeLoggingInterface>;
  Logger get_logger() const { return Logger(); }
};
}  // namespace node_interfaces
struct CallbackGroup { using SharedPtr = std::shared_ptr<CallbackGroup>; };
struct Time { double seconds() const { return 0.0; } };
struct Clock { using SharedPtr = std::shared_ptr<Clock>; Time now() const { return Time(); } };

class Node : public std::enable_shared_from_this<Node> {
 public:
  using SharedPtr = std::shared_ptr<Node>;
  using WeakPtr = std::weak_ptr<Node>;
  explicit Node(std::string name) : name_(std::move(name)) {}
  Logger get_logger() const { return Logger(); }
  Clock::SharedPtr get_clock() const { return std::make_shared<Clock>(); }
  void set_parameter_override(const std::string& n, ParameterVariant v) { overrides_[n] = std::move(v); }
  void declare_parameter(const std::string& n, const ParameterValue& def) {
    auto it = overrides_.find(n);
    ParameterVariant v = def.get();
    if (it != overrides_.end()) {
      if (std::holds_alternative<double>(v) && std::holds_alternative<int64_t>(it->second)) v = (double)std::get<int64_t>(it->second);
      else v = it->second;
    }
    declared_[n] = v;
  }
  void declare_parameter(const std::string& n, ParameterType) {
    auto it = overrides_.find(n);
    if (it == overrides_.end()) throw std::runtime_error("parameter '" + n + "' is not set");
    declared_[n] = it->second;
  }
  Parameter get_parameter(const std::string& n) const {
    auto it = declared_.find(n);
    if (it == declared_.end()) throw std::runtime_error("parameter '" + n + "' has not been declared");
    return Parameter(n, it->second);
  }
  bool get_parameter(const std::string& n, double& out) const { auto it = declared_.find(n); if (it == declared_.end()) return false; out = Parameter(n, it->second).as_double(); return true; }
  bool get_parameter(const std::string& n, bool& out) const { auto it = declared_.find(n); if (it == declared_.end()) return false; out = Parameter(n, it->second).as_bool(); return true; }
  bool get_parameter(const std::string& n, std::string& out) const { auto it = declared_.find(n); if (it == declared_.end()) return false; out = Parameter(n, it->second).as_string(); return true; }
  bool get_parameter(const std::string& n, int& out) const { auto it = declared_.find(n); if (it == declared_.end()) return false; out = (int)Parameter(n, it->second).as_double(); return true; }
 private:
  std::string name_;
  std::map<std::string, ParameterVariant> overrides_, declared_;
};
}  // namespace rclcpp

// logging: swallowed (the format arguments are still evaluated for side-effect parity, none have any)
#define LPREF_LOG(...) do { } while (0)
#define RCLCPP_DEBUG(...) LPREF_LOG(__VA_ARGS__)
#define RCLCPP_INFO(...) LPREF_LOG(__VA_ARGS__)
#define RCLCPP_WARN(...) LPREF_LOG(__VA_ARGS__)
#define RCLCPP_ERROR(...) LPREF_LOG(__VA_ARGS__)
#define RCLCPP_FATAL(...) LPREF_LOG(__VA_ARGS__)
#define RCLCPP_INFO_THROTTLE(...) LPREF_LOG(__VA_ARGS__)
#define RCLCPP_WARN_THROTTLE(...) LPREF_LOG(__VA_ARGS__)
#define RCLCPP_DEBUG_THROTTLE(...) LPREF_LOG(__VA_ARGS__)
