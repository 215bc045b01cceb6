// lpref_tf2.hpp — the slice of tf2 / tf2_eigen / tf2_geometry_msgs (ROS 2 Humble) the reference's path uses
// (TEST INFRASTRUCTURE, see lpref_eigen.hpp; arithmetic per SURVEY.md A5).
#pragma once
#include <math.h>  // tf2/LinearMath/Scalar.h pulls <math.h>: unqualified sqrt/fabs on floats pick the float overloads (A1)

#include <memory>

#include "lpref_eigen.hpp"
#include "lpref_ros.hpp"

namespace tf2_ros {
class Buffer {};
class TransformListener {};
}  // namespace tf2_ros

namespace tf2 {
using TimePoint = double;
class Quaternion {
 public:
  Quaternion() : x_(0), y_(0), z_(0), w_(1) {}
  Quaternion(double x, double y, double z, double w) : x_(x), y_(y), z_(z), w_(w) {}
  double x() const { return x_; }
  double y() const { return y_; }
  double z() const { return z_; }
  double w() const { return w_; }
  double length2() const { return x_ * x_ + y_ * y_ + z_ * z_ + w_ * w_; }
  void setRPY(double roll, double pitch, double yaw) {
    const double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
    const double cy = cos(hy), sy = sin(hy), cp = cos(hp), sp = sin(hp), cr = cos(hr), sr = sin(hr);
    x_ = sr * cp * cy - cr * sp * sy;
    y_ = cr * sp * cy + sr * cp * sy;
    z_ = cr * cp * sy - sr * sp * cy;
    w_ = cr * cp * cy + sr * sp * sy;
  }
  Quaternion& normalize() {
    const double n = sqrt(length2());
    x_ /= n; y_ /= n; z_ /= n; w_ /= n;
    return *this;
  }
 private:
  double x_, y_, z_, w_;
};

// tf2::Matrix3x3(q) (setRotation) and getEulerYPR, non-gimbal branch as upstream
class Matrix3x3 {
 public:
  explicit Matrix3x3(const Quaternion& q) {
    const double d = q.length2();
    const double s = 2.0 / d;
    const double xs = q.x() * s, ys = q.y() * s, zs = q.z() * s;
    const double wx = q.w() * xs, wy = q.w() * ys, wz = q.w() * zs;
    const double xx = q.x() * xs, xy = q.x() * ys, xz = q.x() * zs;
    const double yy = q.y() * ys, yz = q.y() * zs, zz = q.z() * zs;
    m[0][0] = 1.0 - (yy + zz); m[0][1] = xy - wz;         m[0][2] = xz + wy;
    m[1][0] = xy + wz;         m[1][1] = 1.0 - (xx + zz); m[1][2] = yz - wx;
    m[2][0] = xz - wy;         m[2][1] = yz + wx;         m[2][2] = 1.0 - (xx + yy);
  }
  void getEulerYPR(double& yaw, double& pitch, double& roll, unsigned int solution_number = 1) const {
    struct Euler { double yaw, pitch, roll; } out, out2;
    if (fabs(m[2][0]) >= 1) {  // gimbal lock
      out.yaw = 0; out2.yaw = 0;
      const double delta = atan2(m[2][1], m[2][2]);
      if (m[2][0] < 0) { out.pitch = M_PI / 2.0; out2.pitch = M_PI / 2.0; out.roll = delta; out2.roll = delta; }
      else { out.pitch = -M_PI / 2.0; out2.pitch = -M_PI / 2.0; out.roll = delta; out2.roll = delta; }
    } else {
      out.pitch = -asin(m[2][0]);
      out2.pitch = M_PI - out.pitch;
      out.roll = atan2(m[2][1] / cos(out.pitch), m[2][2] / cos(out.pitch));
      out2.roll = atan2(m[2][1] / cos(out2.pitch), m[2][2] / cos(out2.pitch));
      out.yaw = atan2(m[1][0] / cos(out.pitch), m[0][0] / cos(out.pitch));
      out2.yaw = atan2(m[1][0] / cos(out2.pitch), m[0][0] / cos(out2.pitch));
    }
    if (solution_number == 1) { yaw = out.yaw; pitch = out.pitch; roll = out.roll; }
    else { yaw = out2.yaw; pitch = out2.pitch; roll = out2.roll; }
  }
  void getRPY(double& roll, double& pitch, double& yaw, unsigned int solution_number = 1) const { getEulerYPR(yaw, pitch, roll, solution_number); }
 private:
  double m[3][3];
};
using matrix3x3 = Matrix3x3;
class Transform {};

inline void fromMsg(const geometry_msgs::msg::Quaternion& in, Quaternion& out) { out = Quaternion(in.x, in.y, in.z, in.w); }
inline void convert(const geometry_msgs::msg::Quaternion& in, Quaternion& out) { fromMsg(in, out); }

// tf2_eigen: Translation3d * Quaterniond (no normalisation) / translation + Quaterniond(linear)
inline Eigen::Affine3d transformToEigen(const geometry_msgs::msg::Transform& t) {
  return Eigen::Translation3d(t.translation.x, t.translation.y, t.translation.z) *
         Eigen::Quaterniond(t.rotation.w, t.rotation.x, t.rotation.y, t.rotation.z);
}
inline Eigen::Affine3d transformToEigen(const geometry_msgs::msg::TransformStamped& t) { return transformToEigen(t.transform); }
inline geometry_msgs::msg::TransformStamped eigenToTransform(const Eigen::Affine3d& T) {
  geometry_msgs::msg::TransformStamped t;
  t.transform.translation.x = T.translation().x();
  t.transform.translation.y = T.translation().y();
  t.transform.translation.z = T.translation().z();
  const Eigen::Quaterniond q(T.linear());
  t.transform.rotation.x = q.x();
  t.transform.rotation.y = q.y();
  t.transform.rotation.z = q.z();
  t.transform.rotation.w = q.w();
  return t;
}
}  // namespace tf2
