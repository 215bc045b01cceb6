// lpref_eigen.hpp — the slice of Eigen 3.4 that dddmr_local_planner's rollout-and-score sources use, written from the
// published algorithms. TEST INFRASTRUCTURE: lets oracle/Makefile compile the REFERENCE's own
// theory / critic sources (where they lie under /root/reference) without Eigen being installed. Not a copy of Eigen.
#pragma once
#include <math.h>

namespace Eigen {

template <class T>
struct Vec3 {
  T v[3];
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
  T& operator()(int i) { return v[i]; }
  const T& operator()(int i) const { return v[i]; }
  T& x() { return v[0]; }
  T& y() { return v[1]; }
  T& z() { return v[2]; }
  const T& x() const { return v[0]; }
  const T& y() const { return v[1]; }
  const T& z() const { return v[2]; }
  static Vec3 Zero() { return Vec3{{T(0), T(0), T(0)}}; }
  static Vec3 UnitX() { return Vec3{{T(1), T(0), T(0)}}; }
  static Vec3 UnitY() { return Vec3{{T(0), T(1), T(0)}}; }
  static Vec3 UnitZ() { return Vec3{{T(0), T(0), T(1)}}; }
};
using Vector3f = Vec3<float>;
using Vector3d = Vec3<double>;

struct Matrix3d {
  double m[3][3];
  double& operator()(int i, int j) { return m[i][j]; }
  const double& operator()(int i, int j) const { return m[i][j]; }
};

// AngleAxis::toRotationMatrix (Eigen/src/Geometry/AngleAxis.h)
struct AngleAxisd {
  double angle_;
  Vector3d axis_;
  AngleAxisd(double a, const Vector3d& ax) : angle_(a), axis_(ax) {}
  Matrix3d toRotationMatrix() const {
    Matrix3d res;
    const double sin_axis[3] = {sin(angle_) * axis_[0], sin(angle_) * axis_[1], sin(angle_) * axis_[2]};
    const double c = cos(angle_);
    const double cos1_axis[3] = {(1.0 - c) * axis_[0], (1.0 - c) * axis_[1], (1.0 - c) * axis_[2]};
    double tmp;
    tmp = cos1_axis[0] * axis_[1];
    res(0, 1) = tmp - sin_axis[2];
    res(1, 0) = tmp + sin_axis[2];
    tmp = cos1_axis[0] * axis_[2];
    res(0, 2) = tmp + sin_axis[1];
    res(2, 0) = tmp - sin_axis[1];
    tmp = cos1_axis[1] * axis_[2];
    res(1, 2) = tmp - sin_axis[0];
    res(2, 1) = tmp + sin_axis[0];
    res(0, 0) = cos1_axis[0] * axis_[0] + c;
    res(1, 1) = cos1_axis[1] * axis_[1] + c;
    res(2, 2) = cos1_axis[2] * axis_[2] + c;
    return res;
  }
};

// Quaternion <-> rotation matrix (Eigen/src/Geometry/Quaternion.h: toRotationMatrix, quaternionbase_assign_impl)
struct Quaterniond {
  double x_, y_, z_, w_;
  Quaterniond() : x_(0), y_(0), z_(0), w_(1) {}
  Quaterniond(double w, double x, double y, double z) : x_(x), y_(y), z_(z), w_(w) {}
  explicit Quaterniond(const Matrix3d& mat) {
    double t = mat(0, 0) + mat(1, 1) + mat(2, 2);
    double q[3];
    if (t > 0.0) {
      t = sqrt(t + 1.0);
      w_ = 0.5 * t;
      t = 0.5 / t;
      x_ = (mat(2, 1) - mat(1, 2)) * t;
      y_ = (mat(0, 2) - mat(2, 0)) * t;
      z_ = (mat(1, 0) - mat(0, 1)) * t;
    } else {
      int i = 0;
      if (mat(1, 1) > mat(0, 0)) i = 1;
      if (mat(2, 2) > mat(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = sqrt(mat(i, i) - mat(j, j) - mat(k, k) + 1.0);
      q[i] = 0.5 * t;
      t = 0.5 / t;
      w_ = (mat(k, j) - mat(j, k)) * t;
      q[j] = (mat(j, i) + mat(i, j)) * t;
      q[k] = (mat(k, i) + mat(i, k)) * t;
      x_ = q[0]; y_ = q[1]; z_ = q[2];
    }
  }
  double x() const { return x_; }
  double y() const { return y_; }
  double z() const { return z_; }
  double w() const { return w_; }
  Matrix3d toRotationMatrix() const {
    Matrix3d res;
    const double tx = 2.0 * x_, ty = 2.0 * y_, tz = 2.0 * z_;
    const double twx = tx * w_, twy = ty * w_, twz = tz * w_;
    const double txx = tx * x_, txy = ty * x_, txz = tz * x_;
    const double tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    res(0, 0) = 1.0 - (tyy + tzz); res(0, 1) = txy - twz;         res(0, 2) = txz + twy;
    res(1, 0) = txy + twz;         res(1, 1) = 1.0 - (txx + tzz); res(1, 2) = tyz - twx;
    res(2, 0) = txz - twy;         res(2, 1) = tyz + twx;         res(2, 2) = 1.0 - (txx + tyy);
    return res;
  }
};

struct Translation3d {
  Vector3d v;
  Translation3d(double x, double y, double z) : v{{x, y, z}} {}
};

// Transform<double,3,Affine>: linear part L, translation t; product and inverse as Eigen evaluates them for Affine mode
class Affine3d {
 public:
  Affine3d() {}
  explicit Affine3d(const AngleAxisd& aa) : L_(aa.toRotationMatrix()), t_(Vector3d::Zero()) {}
  explicit Affine3d(const Matrix3d& L) : L_(L), t_(Vector3d::Zero()) {}
  Affine3d(const Matrix3d& L, const Vector3d& t) : L_(L), t_(t) {}
  Vector3d& translation() { return t_; }
  const Vector3d& translation() const { return t_; }
  Matrix3d& linear() { return L_; }
  const Matrix3d& linear() const { return L_; }
  Matrix3d rotation() const { return L_; }
  double operator()(int i, int j) const { return j < 3 ? L_(i, j) : t_[i]; }
  Affine3d operator*(const Affine3d& o) const {
    Affine3d r;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) r.L_(i, j) = L_(i, 0) * o.L_(0, j) + L_(i, 1) * o.L_(1, j) + L_(i, 2) * o.L_(2, j);
      r.t_[i] = (L_(i, 0) * o.t_[0] + L_(i, 1) * o.t_[1] + L_(i, 2) * o.t_[2]) + t_[i];
    }
    return r;
  }
  // Transform::inverse(Affine): general 3x3 inverse of the linear part (cofactors / determinant, compute_inverse_size3),
  // translation = -(L^-1 t)
  Affine3d inverse() const {
    auto cof = [&](int i, int j) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      return L_(i1, j1) * L_(i2, j2) - L_(i1, j2) * L_(i2, j1);
    };
    const double c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    const double det = (c00 * L_(0, 0) + c10 * L_(1, 0)) + c20 * L_(2, 0);
    const double invdet = 1.0 / det;
    Affine3d r;
    r.L_(0, 0) = c00 * invdet; r.L_(0, 1) = c10 * invdet; r.L_(0, 2) = c20 * invdet;
    r.L_(1, 0) = cof(0, 1) * invdet; r.L_(1, 1) = cof(1, 1) * invdet; r.L_(1, 2) = cof(2, 1) * invdet;
    r.L_(2, 0) = cof(0, 2) * invdet; r.L_(2, 1) = cof(1, 2) * invdet; r.L_(2, 2) = cof(2, 2) * invdet;
    for (int i = 0; i < 3; ++i) r.t_[i] = -((r.L_(i, 0) * t_[0] + r.L_(i, 1) * t_[1]) + r.L_(i, 2) * t_[2]);
    return r;
  }

 private:
  Matrix3d L_;
  Vector3d t_;
};

inline Affine3d operator*(const Translation3d& t, const Quaterniond& q) { return Affine3d(q.toRotationMatrix(), t.v); }

}  // namespace Eigen
