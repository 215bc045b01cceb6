// lp_oracle.cpp — CPU oracle for the local-planner rollout-and-score path.
//
// TEST INFRASTRUCTURE ONLY. Nothing under dddmr_navigation_b200/ may call into this file; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
//
// PARITY: pinned to the reference's own sources run here (oracle/_ref/liblpref.so, tests/test_reference_sources.py);
// third-party arithmetic (Eigen / PCL / FLANN / tf2) unpinned — see lp_oracle.h. Every function below cites the reference lines it restates
// (paths relative to /root/reference/src/dddmr_local_planner/, shortened to TG/, MC/, BT/, LP/ as in
// SURVEY.md). Arithmetic that lives in un-vendored third-party code (Eigen 3.4, PCL 1.15, FLANN
// 1.9.1, tf2 Humble) follows the numeric contract of SURVEY.md Appendix A.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off (no -march: the reference's x86-64 build has no FMA).
#include "lp_oracle.h"

#include <math.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../dddmr_navigation_b200/csrc/lp_math.h"

#ifdef LPORACLE_WITH_NANOFLANN
#include "nanoflann.hpp"  // -I /root/reference/src/dddmr_global_planner/include/global_planner
#endif

namespace {

// ---------------------------------------------------------------------------------------------
// math dispatch: "shared" = lp_math.h (what the device evaluates), "libm" = glibc.
// ---------------------------------------------------------------------------------------------
struct Math {
  float (*sinf_)(float);
  float (*cosf_)(float);
  double (*sin_)(double);
  double (*cos_)(double);
  double (*asin_)(double);
  double (*atan2_)(double, double);
  double (*fmod_)(double, double);
};
float sh_sinf(float x) { return lpm::sinf(x); }
float sh_cosf(float x) { return lpm::cosf(x); }
double sh_sin(double x) { return lpm::sin(x); }
double sh_cos(double x) { return lpm::cos(x); }
double sh_asin(double x) { return lpm::asin(x); }
double sh_atan2(double y, double x) { return lpm::atan2(y, x); }
double sh_fmod(double x, double y) {
  double r = lpm::fmod_pos(lpm::dabs(x), lpm::dabs(y));
  return x < 0.0 ? -r : r;
}
float lm_sinf(float x) { return ::sinf(x); }
float lm_cosf(float x) { return ::cosf(x); }
double lm_sin(double x) { return ::sin(x); }
double lm_cos(double x) { return ::cos(x); }
double lm_asin(double x) { return ::asin(x); }
double lm_atan2(double y, double x) { return ::atan2(y, x); }
double lm_fmod(double x, double y) { return ::fmod(x, y); }
const Math kShared = {sh_sinf, sh_cosf, sh_sin, sh_cos, sh_asin, sh_atan2, sh_fmod};
const Math kLibm = {lm_sinf, lm_cosf, lm_sin, lm_cos, lm_asin, lm_atan2, lm_fmod};

// ---------------------------------------------------------------------------------------------
// Eigen / tf2 double-precision pieces
// ---------------------------------------------------------------------------------------------
struct Affine {
  double L[3][3];
  double t[3];
};

// Eigen::Quaterniond(w,x,y,z).toRotationMatrix(), no normalisation (A4).
void quat_to_matrix(double x, double y, double z, double w, double R[3][3]) {
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0][0] = 1.0 - (tyy + tzz);
  R[0][1] = txy - twz;
  R[0][2] = txz + twy;
  R[1][0] = txy + twz;
  R[1][1] = 1.0 - (txx + tzz);
  R[1][2] = tyz - twx;
  R[2][0] = txz - twy;
  R[2][1] = tyz + twx;
  R[2][2] = 1.0 - (txx + tyy);
}

// tf2::transformToEigen: Translation3d * Quaterniond (A5). p = x,y,z,qx,qy,qz,qw.
Affine pose_to_affine(const double p[7]) {
  Affine a;
  quat_to_matrix(p[3], p[4], p[5], p[6], a.L);
  a.t[0] = p[0];
  a.t[1] = p[1];
  a.t[2] = p[2];
  return a;
}

// How a 3-term inner product inside Eigen's fixed-size products is associated. The oracle (and the device) evaluate
// (a0*b0 + a1*b1) + a2*b2, which is what Eigen 3.4's scalar reducer of a coefficient-based 3x3 product does as far as its
// source can be read from memory (not re-readable here, Eigen is not vendored). The ALTERNATIVE
// a0*b0 + (a1*b1 + a2*b2) exists only to measure how much of the result depends on that reading
// (tests/test_oracle.py::test_eigen_association_changes_nothing_discrete): test infrastructure, process-wide.
int g_assoc_right = 0;
inline double dot3(double a0, double b0, double a1, double b1, double a2, double b2) {
  return g_assoc_right ? a0 * b0 + (a1 * b1 + a2 * b2) : (a0 * b0 + a1 * b1) + a2 * b2;
}

// Affine3d * Affine3d (A4): L = L1*L2 with each entry a0*b0 + a1*b1 + a2*b2 left to right,
// t = L1*t2 + t1.
Affine affine_mul(const Affine& a, const Affine& b) {
  Affine r;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) r.L[i][j] = dot3(a.L[i][0], b.L[0][j], a.L[i][1], b.L[1][j], a.L[i][2], b.L[2][j]);
    r.t[i] = dot3(a.L[i][0], b.t[0], a.L[i][1], b.t[1], a.L[i][2], b.t[2]) + a.t[i];
  }
  return r;
}

double cofactor3(const double m[3][3], int i, int j) {
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m[i1][j1] * m[i2][j2] - m[i1][j2] * m[i2][j1];
}

// Affine3d::inverse() in Affine mode: cofactor inverse of the linear part, t' = -(L^-1 t) (A4).
Affine affine_inverse(const Affine& a) {
  Affine r;
  const double c00 = cofactor3(a.L, 0, 0), c10 = cofactor3(a.L, 1, 0), c20 = cofactor3(a.L, 2, 0);
  const double det = dot3(c00, a.L[0][0], c10, a.L[1][0], c20, a.L[2][0]);
  const double invdet = 1.0 / det;
  r.L[0][0] = c00 * invdet;
  r.L[0][1] = c10 * invdet;
  r.L[0][2] = c20 * invdet;
  r.L[1][0] = cofactor3(a.L, 0, 1) * invdet;
  r.L[1][1] = cofactor3(a.L, 1, 1) * invdet;
  r.L[1][2] = cofactor3(a.L, 2, 1) * invdet;
  r.L[2][0] = cofactor3(a.L, 0, 2) * invdet;
  r.L[2][1] = cofactor3(a.L, 1, 2) * invdet;
  r.L[2][2] = cofactor3(a.L, 2, 2) * invdet;
  for (int i = 0; i < 3; ++i) r.t[i] = -dot3(r.L[i][0], a.t[0], r.L[i][1], a.t[1], r.L[i][2], a.t[2]);
  return r;
}

// Eigen::Quaterniond(Matrix3d) (A4). out = x,y,z,w.
void matrix_to_quat(const double m[3][3], double q[4]) {
  double t = (m[0][0] + m[1][1]) + m[2][2];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[2][1] - m[1][2]) * t;
    q[1] = (m[0][2] - m[2][0]) * t;
    q[2] = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(((m[i][i] - m[j][j]) - m[k][k]) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m[k][j] - m[j][k]) * t;
    q[j] = (m[j][i] + m[i][j]) * t;
    q[k] = (m[k][i] + m[i][k]) * t;
  }
}

// ---------------------------------------------------------------------------------------------
// containers
// ---------------------------------------------------------------------------------------------
struct Pose {
  double p[7];        // PoseStamped: position + orientation (x,y,z,w)
  float pcl[3];       // Trajectory::pcl_trajectory_path_ entry (BT/src/trajectory.cpp:71-76)
  float cuboid[8][3]; // transformed vertices
  float mn[3], mx[3]; // pcl::getMinMax3D
};

struct Traj {  // base_trajectory::Trajectory (BT/include/base_trajectory/trajectory.h:47-126)
  int sample_index = -1;
  float vel[3] = {0, 0, 0};
  double xv = 0, yv = 0, thetav = 0, cost = -1.0, time_delta = 0;
  std::vector<Pose> poses;
  std::vector<double> critic_scores;  // NaN = not evaluated
  int first_hit_pose = -1;
};

struct Pt {
  float x, y, z;
};

// ---------------------------------------------------------------------------------------------
// radius-search indices. All return exactly {i : fl(d^2(q, p_i)) < r2} with FLANN's L2_Simple
// float accumulation (A2); order is irrelevant to every consumer.
// ---------------------------------------------------------------------------------------------
inline float l2_simple(const float q[3], const Pt& p) {
  float r = 0.0f;
  float d = q[0] - p.x;
  r += d * d;
  d = q[1] - p.y;
  r += d * d;
  d = q[2] - p.z;
  r += d * d;
  return r;
}

struct BucketGrid {  // 1 m buckets, counting-sorted copy of the cloud
  float org[3] = {0, 0, 0};
  int dim[3] = {0, 0, 0};
  std::vector<uint32_t> start;
  std::vector<uint32_t> idx;
  void build(const std::vector<Pt>& pts) {
    start.clear();
    idx.clear();
    dim[0] = dim[1] = dim[2] = 0;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    size_t nfin = 0;
    for (const Pt& p : pts) {
      if (!(std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z))) continue;
      ++nfin;
      mn[0] = std::min(mn[0], p.x); mx[0] = std::max(mx[0], p.x);
      mn[1] = std::min(mn[1], p.y); mx[1] = std::max(mx[1], p.y);
      mn[2] = std::min(mn[2], p.z); mx[2] = std::max(mx[2], p.z);
    }
    if (!nfin) return;
    for (int a = 0; a < 3; ++a) {
      org[a] = floorf(mn[a]);
      dim[a] = (int)(floorf(mx[a]) - org[a]) + 1;
    }
    const size_t ncell = (size_t)dim[0] * dim[1] * dim[2];
    start.assign(ncell + 1, 0);
    auto cell = [&](const Pt& p) {
      int cx = (int)floorf(p.x - org[0]), cy = (int)floorf(p.y - org[1]), cz = (int)floorf(p.z - org[2]);
      cx = std::min(std::max(cx, 0), dim[0] - 1);
      cy = std::min(std::max(cy, 0), dim[1] - 1);
      cz = std::min(std::max(cz, 0), dim[2] - 1);
      return ((size_t)cz * dim[1] + cy) * dim[0] + cx;
    };
    for (const Pt& p : pts)
      if (std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z)) ++start[cell(p) + 1];
    for (size_t c = 0; c < ncell; ++c) start[c + 1] += start[c];
    idx.resize(nfin);
    std::vector<uint32_t> fill(start.begin(), start.end() - 1);
    for (uint32_t i = 0; i < pts.size(); ++i) {
      const Pt& p = pts[i];
      if (std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z)) idx[fill[cell(p)]++] = i;
    }
  }
  template <class F>
  void visit(const float q[3], float r, F&& f) const {
    if (!dim[0]) return;
    int lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
      lo[a] = (int)floorf(q[a] - r - org[a]) - 1;
      hi[a] = (int)floorf(q[a] + r - org[a]) + 1;
      lo[a] = std::max(lo[a], 0);
      hi[a] = std::min(hi[a], dim[a] - 1);
      if (lo[a] > hi[a]) return;
    }
    for (int cz = lo[2]; cz <= hi[2]; ++cz)
      for (int cy = lo[1]; cy <= hi[1]; ++cy) {
        const size_t row = ((size_t)cz * dim[1] + cy) * dim[0];
        for (uint32_t k = start[row + lo[0]]; k < start[row + hi[0] + 1]; ++k)
          if (!f(idx[k])) return;
      }
  }
};

#ifdef LPORACLE_WITH_NANOFLANN
struct NfCloud {
  const std::vector<Pt>* pts = nullptr;
  inline size_t kdtree_get_point_count() const { return pts->size(); }
  inline float kdtree_get_pt(const size_t i, const size_t d) const {
    const Pt& p = (*pts)[i];
    return d == 0 ? p.x : (d == 1 ? p.y : p.z);
  }
  template <class BBOX>
  bool kdtree_get_bbox(BBOX&) const { return false; }
};
using NfTree = nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<float, NfCloud>, NfCloud, 3>;
#endif

}  // namespace

struct lporacle_ctx {
  b200lp_limits lim;
  b200lp_params par;
  float cuboid[8][3];
  std::vector<b200lp_critic> critics;
  const Math* m = &kShared;
  int index_mode = LPORACLE_INDEX_GRID;
  int stride = 1, phase = 0;
  std::string err;

  // ModelSharedData (MC/include/mpc_critics/model_shared_data.h:67-116)
  std::vector<Pt> cloud;            // pcl_perception_
  std::vector<double> plan;         // prune_plan_ (7 doubles per pose)
  std::vector<Pt> pcl_plan;         // pcl_prune_plan_ (float-cast positions)
  BucketGrid grid;
#ifdef LPORACLE_WITH_NANOFLANN
  NfCloud nf_cloud;
  std::unique_ptr<NfTree> nf_tree;
#endif
  bool index_valid = false;
  bool keep_index = false;  // tests only: do not rebuild an index that is still valid (same cloud, same contents)

  // cycle state
  b200lp_query q{};
  std::vector<Traj> trajs;          // Local_Planner::trajectories_
  int n_samples = 0;
};

namespace {

// ---------------------------------------------------------------------------------------------
// velocity sampling
// ---------------------------------------------------------------------------------------------
// trajectory_generators::VelocityIterator (TG/include/trajectory_generators/velocity_iterator.h:44-69)
std::vector<double> velocity_iterator(double mn, double mx, int n) {
  std::vector<double> s;
  if (mn == mx) {
    s.push_back(mn);
    return s;
  }
  n = std::max(2, n);
  const double step = (mx - mn) / double(std::max(1, n - 1));
  double next = mn;
  for (int j = 0; j < n - 1; ++j) {
    const double cur = next;
    next += step;
    s.push_back(cur);
    if (cur < 0 && next > 0) s.push_back(0.0);
  }
  s.push_back(mx);
  return s;
}

// isMotorConstraintSatisfied (TG/theories/dd_simple…cpp:297-312; dd_rotate_inplace_theory.cpp:273-284)
bool motor_ok(const b200lp_limits& L, const float v[3]) {
  const double vr = v[0] + L.robot_radius * v[2];
  const double vl = v[0] - L.robot_radius * v[2];
  const double rpm_r = vr * L.gear_ratio * 60. / 3.1415926 / L.wheel_diameter;
  const double rpm_l = vl * L.gear_ratio * 60. / 3.1415926 / L.wheel_diameter;
  return !(fabs(rpm_r) >= L.max_motor_shaft_rpm || fabs(rpm_l) >= L.max_motor_shaft_rpm);
}

struct Sample {
  float v[3];
};

// initialise() of the three theories:
//   DD simple   TG/theories/dd_simple_trajectory_generator_theory.cpp:236-295
//   omni simple TG/theories/omni_simple_trajectory_generator_theory.cpp:260-330
//   rotate      TG/theories/dd_rotate_inplace_theory.cpp:229-271
std::vector<Sample> make_samples(const lporacle_ctx& c, const b200lp_query& q) {
  std::vector<Sample> out;
  const b200lp_limits& L = c.lim;
  const b200lp_params& P = c.par;
  const double max_vel_th = L.max_vel_theta;
  const double min_vel_th = -1.0 * max_vel_th;
  const float acc[3] = {(float)L.acc_lim_x, (float)L.acc_lim_y, (float)L.acc_lim_theta};
  double min_vel_x = L.min_vel_x, max_vel_x = L.max_vel_x;
  const double min_vel_y = L.min_vel_y, max_vel_y = L.max_vel_y;
  if (!(P.linear_x_sample * P.angular_z_sample > 0)) return out;
  const double sim_period = 1.0 / P.controller_frequency;
  const double tx = q.twist[0], ty = q.twist[1], tw = q.twist[2];
  float max_vel[3] = {0, 0, 0}, min_vel[3] = {0, 0, 0};

  if (P.theory == B200LP_THEORY_DD_SIMPLE) {
    if (q.max_speed_override > 0.0) max_vel_x = std::min(max_vel_x, q.max_speed_override);
    max_vel[0] = (float)std::min(max_vel_x, tx + acc[0] * sim_period);
    max_vel[2] = (float)std::min(max_vel_th, tw + acc[2] * sim_period);
    min_vel[0] = (float)std::max(min_vel_x, tx / L.deceleration_ratio);
    min_vel[2] = (float)std::max(min_vel_th, tw - acc[2] * sim_period);
    if (max_vel[0] < min_vel[0]) {
      min_vel[0] = (float)(tx / L.deceleration_ratio);
      max_vel[0] = (float)(tx / L.deceleration_ratio);
    }
    const std::vector<double> xs = velocity_iterator(min_vel[0], max_vel[0], (int)P.linear_x_sample);
    const std::vector<double> ths = velocity_iterator(min_vel[2], max_vel[2], (int)P.angular_z_sample);
    for (double x : xs)
      for (double th : ths) {
        Sample s;
        s.v[0] = (float)x;
        s.v[1] = 0.0f;
        s.v[2] = (float)th;
        if (!L.use_motor_constraint || motor_ok(L, s.v)) out.push_back(s);
      }
  } else if (P.theory == B200LP_THEORY_OMNI_SIMPLE) {
    max_vel[0] = (float)std::min(max_vel_x, tx + acc[0] * sim_period);
    max_vel[1] = (float)std::min(max_vel_y, ty + acc[1] * sim_period);
    max_vel[2] = (float)std::min(max_vel_th, tw + acc[2] * sim_period);
    min_vel[0] = (float)std::max(min_vel_x, tx - acc[0] * sim_period);
    min_vel[1] = (float)std::max(min_vel_y, ty - acc[1] * sim_period);
    min_vel[2] = (float)std::max(min_vel_th, tw - acc[2] * sim_period);
    if (tx >= max_vel_x / L.deceleration_ratio) min_vel[0] = (float)std::max(min_vel_x, tx / L.deceleration_ratio);
    else if (tx <= min_vel_x / L.deceleration_ratio) max_vel[0] = (float)std::min(max_vel_x, tx / L.deceleration_ratio);
    if (ty >= max_vel_y / L.deceleration_ratio) min_vel[1] = (float)std::max(min_vel_y, ty / L.deceleration_ratio);
    else if (ty <= min_vel_y / L.deceleration_ratio) max_vel[1] = (float)std::min(max_vel_y, ty / L.deceleration_ratio);
    const std::vector<double> xs = velocity_iterator(min_vel[0], max_vel[0], (int)P.linear_x_sample);
    const std::vector<double> ys = velocity_iterator(min_vel[1], max_vel[1], (int)P.linear_y_sample);
    const std::vector<double> ths = velocity_iterator(min_vel[2], max_vel[2], (int)P.angular_z_sample);
    for (double x : xs)
      for (double y : ys)
        for (double th : ths) {
          Sample s;
          s.v[0] = (float)x;
          s.v[1] = (float)y;
          s.v[2] = (float)th;
          out.push_back(s);  // omni isMotorConstraintSatisfied always returns true (omni…cpp:332-341)
        }
  } else {  // rotate in place: exactly (0,0,+w) then (0,0,-w), each under the motor constraint
    Sample sp{{0.0f, 0.0f, (float)L.rotation_speed}};
    Sample sn{{0.0f, 0.0f, (float)(-1.0 * L.rotation_speed)}};
    if (motor_ok(L, sp.v)) out.push_back(sp);
    if (motor_ok(L, sn.v)) out.push_back(sn);
  }
  return out;
}

// ---------------------------------------------------------------------------------------------
// rollout
// ---------------------------------------------------------------------------------------------
// generateTrajectory + computeNewPositions:
//   DD simple   TG/theories/dd_simple…cpp:351-464
//   omni simple TG/theories/omni_simple…cpp:382-505
//   rotate      TG/theories/dd_rotate_inplace_theory.cpp:325-427
bool generate_trajectory(const lporacle_ctx& c, const b200lp_query& q, const float sv[3], Traj& traj) {
  const Math& M = *c.m;
  const b200lp_limits& L = c.lim;
  const b200lp_params& P = c.par;
  const Affine pos_af3 = pose_to_affine(q.pose);
  const double eps = 1e-4;
  traj.cost = 0.0;
  traj.poses.clear();
  double vmag;
  double sim_time = P.sim_time;

  if (P.theory == B200LP_THEORY_DD_SIMPLE) {
    vmag = fabsf(sv[0]);
    if ((L.min_vel_x >= 0 && vmag + eps < L.min_vel_x) && (L.min_vel_theta >= 0 && fabsf(sv[2]) + eps < L.min_vel_theta))
      return false;
    if (L.max_vel_x >= 0 && vmag - eps > L.max_vel_x) return false;
  } else if (P.theory == B200LP_THEORY_OMNI_SIMPLE) {
    // hypot(float,float) resolves to the float overload; glibc hypotf = (float)sqrt(x*x+y*y) in double
    vmag = (double)(float)sqrt((double)sv[0] * (double)sv[0] + (double)sv[1] * (double)sv[1]);
    if ((L.min_vel_trans >= 0 && vmag + eps < L.min_vel_trans) &&
        (L.min_vel_theta >= 0 && fabsf(sv[2]) + eps < L.min_vel_theta))
      return false;
    if (L.max_vel_trans >= 0 && vmag - eps > L.max_vel_trans) return false;
    if (q.max_speed_override > 0.0 && vmag - eps > q.max_speed_override) return false;
  } else {
    vmag = fabsf(sv[0]);
    sim_time = 6.28 / fabsf(sv[2]);  // a_rad_sim_time
  }

  const double sim_time_distance = vmag * sim_time;
  const double sim_time_angle = fabsf(sv[2]) * sim_time;
  const int num_steps = (int)ceil(std::max(sim_time_distance / P.sim_granularity, sim_time_angle / P.angular_sim_granularity));
  if (num_steps == 0) return false;
  const double dt = sim_time / num_steps;
  traj.time_delta = dt;
  traj.xv = sv[0];
  traj.yv = (P.theory == B200LP_THEORY_OMNI_SIMPLE) ? (double)sv[1] : 0.0;
  traj.thetav = sv[2];
  traj.vel[0] = sv[0];
  traj.vel[1] = sv[1];
  traj.vel[2] = sv[2];

  float pos[3] = {0.0f, 0.0f, 0.0f};
  traj.poses.resize(num_steps);
  for (int i = 0; i < num_steps; ++i) {
    float np[3];
    if (P.theory == B200LP_THEORY_OMNI_SIMPLE) {
      np[0] = (float)(pos[0] + (sv[0] * M.cosf_(pos[2]) + sv[1] * M.cos_(M_PI_2 + pos[2])) * dt);
      np[1] = (float)(pos[1] + (sv[0] * M.sinf_(pos[2]) + sv[1] * M.sin_(M_PI_2 + pos[2])) * dt);
    } else {
      np[0] = (float)(pos[0] + (sv[0] * M.cosf_(pos[2])) * dt);
      np[1] = (float)(pos[1] + (sv[0] * M.sinf_(pos[2])) * dt);
    }
    np[2] = (float)(pos[2] + sv[2] * dt);
    pos[0] = np[0];
    pos[1] = np[1];
    pos[2] = np[2];

    // Affine3d(AngleAxisd(pos[2], UnitZ)) with translation (pos[0], pos[1], 0)   (A4)
    Affine b2t;
    const double ang = (double)pos[2];
    const double s = M.sin_(ang), co = M.cos_(ang);
    const double one_minus_c = 1.0 - co;
    b2t.L[0][0] = 0.0 * 0.0 + co;
    b2t.L[1][1] = 0.0 * 0.0 + co;
    b2t.L[2][2] = one_minus_c * 1.0 + co;
    b2t.L[0][1] = 0.0 - s;
    b2t.L[1][0] = 0.0 + s;
    b2t.L[0][2] = 0.0;
    b2t.L[2][0] = 0.0;
    b2t.L[1][2] = 0.0;
    b2t.L[2][1] = 0.0;
    b2t.t[0] = pos[0];
    b2t.t[1] = pos[1];
    b2t.t[2] = 0.0;
    const Affine g = affine_mul(pos_af3, b2t);

    Pose& po = traj.poses[i];
    po.p[0] = g.t[0];
    po.p[1] = g.t[1];
    po.p[2] = g.t[2];
    matrix_to_quat(g.L, &po.p[3]);  // tf2::eigenToTransform (A5)
    // pcl::transformPointCloud(cuboid, out, Affine3d) (A3) + pcl::getMinMax3D
    for (int a = 0; a < 3; ++a) {
      po.mn[a] = std::numeric_limits<float>::max();
      po.mx[a] = -std::numeric_limits<float>::max();
    }
    for (int v = 0; v < 8; ++v) {
      const double x = c.cuboid[v][0], y = c.cuboid[v][1], z = c.cuboid[v][2];
      for (int a = 0; a < 3; ++a) {
        po.cuboid[v][a] = (float)(((g.L[a][0] * x + g.L[a][1] * y) + g.L[a][2] * z) + g.t[a]);
        po.mn[a] = std::min(po.mn[a], po.cuboid[v][a]);
        po.mx[a] = std::max(po.mx[a], po.cuboid[v][a]);
      }
    }
    // Trajectory::addPoint float-casts the position (BT/src/trajectory.cpp:71-76)
    po.pcl[0] = (float)po.p[0];
    po.pcl[1] = (float)po.p[1];
    po.pcl[2] = (float)po.p[2];
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// critics
// ---------------------------------------------------------------------------------------------
struct BoxFrame {  // per-pose quantities of CollisionModel (MC/models/collision_model.cpp:86-119)
  float c[3];
  float ax[3][3];
  double half[3];
};

BoxFrame box_frame(const Pose& po) {
  BoxFrame f;
  f.c[0] = f.c[1] = f.c[2] = 0.0f;
  for (int v = 0; v < 8; ++v)
    for (int a = 0; a < 3; ++a) f.c[a] += po.cuboid[v][a];
  for (int a = 0; a < 3; ++a) f.c[a] /= (float)(size_t)8;
  const int other[3] = {3, 1, 2};  // dx = v3-v0, dy = v1-v0, dz = v2-v0
  for (int e = 0; e < 3; ++e) {
    float d[3];
    for (int a = 0; a < 3; ++a) d[a] = po.cuboid[other[e]][a] - po.cuboid[0][a];
    f.half[e] = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) / 2.;
    for (int a = 0; a < 3; ++a) f.ax[e][a] = (float)(d[a] / (2. * f.half[e]));
  }
  return f;
}

inline bool in_box(const BoxFrame& f, const Pt& p) {  // collision_model.cpp:124-139
  const float dx = p.x - f.c[0], dy = p.y - f.c[1], dz = p.z - f.c[2];
  const double xv = fabsf(dx * f.ax[0][0] + dy * f.ax[0][1] + dz * f.ax[0][2]);
  const double yv = fabsf(dx * f.ax[1][0] + dy * f.ax[1][1] + dz * f.ax[1][2]);
  const double zv = fabsf(dx * f.ax[2][0] + dy * f.ax[2][1] + dz * f.ax[2][2]);
  return xv <= f.half[0] && yv <= f.half[1] && zv <= f.half[2];
}

inline bool in_aabb(const Pose& po, const Pt& p) {  // collision_min_max_model.cpp:74-77
  return p.x >= po.mn[0] && p.x <= po.mx[0] && p.y >= po.mn[1] && p.y <= po.mx[1] && p.z >= po.mn[2] && p.z <= po.mx[2];
}

// radiusSearch(pose, 1.0): calls f(index) for every point with float d^2 < 1.0f until f returns false.
template <class F>
void radius_search(const lporacle_ctx& c, const float q[3], F&& f) {
  const float r2 = (float)(1.0 * 1.0);
  if (c.index_mode == LPORACLE_INDEX_BRUTE) {
    for (uint32_t i = 0; i < c.cloud.size(); ++i)
      if (l2_simple(q, c.cloud[i]) < r2)
        if (!f(i)) return;
    return;
  }
#ifdef LPORACLE_WITH_NANOFLANN
  if (c.index_mode == LPORACLE_INDEX_NANOFLANN) {
    std::vector<nanoflann::ResultItem<uint32_t, float>> res;  // fresh vectors per query, like the reference
    c.nf_tree->radiusSearch(q, r2, res, nanoflann::SearchParameters(0.f, true));
    for (const auto& it : res)
      if (!f(it.first)) return;
    return;
  }
#endif
  c.grid.visit(q, 1.0f, [&](uint32_t i) {
    if (l2_simple(q, c.cloud[i]) < r2) return f(i);
    return true;
  });
}

// CollisionModel::scoreTrajectory (MC/models/collision_model.cpp:51-148) and
// CollisionMinMaxModel::scoreTrajectory (MC/models/collision_min_max_model.cpp:51-88)
double critic_collision(const lporacle_ctx& c, Traj& t, bool minmax) {
  if (c.cloud.size() < 5) return 0.0;
  for (size_t i = 0; i < t.poses.size(); ++i) {
    const Pose& po = t.poses[i];
    bool hit = false;
    if (minmax) {
      radius_search(c, po.pcl, [&](uint32_t k) {
        if (in_aabb(po, c.cloud[k])) { hit = true; return false; }
        return true;
      });
    } else {
      const BoxFrame f = box_frame(po);
      radius_search(c, po.pcl, [&](uint32_t k) {
        if (in_box(f, c.cloud[k])) { hit = true; return false; }
        return true;
      });
    }
    if (hit) {
      if (t.first_hit_pose < 0) t.first_hit_pose = (int)i;
      return -1.0;
    }
  }
  return 0.0;
}

// nearestKSearch(K=1) on the prune-plan cloud: minimum float squared distance.
inline float plan_nn_d2(const lporacle_ctx& c, const float q[3]) {
  float best = std::numeric_limits<float>::max();
  for (const Pt& p : c.pcl_plan) {
    const float d = l2_simple(q, p);
    if (d < best) best = d;
  }
  return best;
}

// StickPathModel::scoreTrajectory (MC/models/stick_path_model.cpp:51-77)
double critic_stick_path(const lporacle_ctx& c, const Traj& t) {
  if (c.pcl_plan.size() < 3) return 10.0;
  double nd = 0.0;
  for (const Pose& po : t.poses) nd += sqrtf(plan_nn_d2(c, po.pcl));
  nd /= c.pcl_plan.size();
  return nd;
}

// TowardGlobalPlanModel::scoreTrajectory (MC/models/toward_global_plan_model.cpp:52-78)
double critic_toward_global_plan(const lporacle_ctx& c, const Traj& t, double weight) {
  if (c.pcl_plan.size() < 3) return 10.0;
  return sqrtf(plan_nn_d2(c, t.poses.back().pcl)) * weight;
}

// PurePursuitModel::scoreTrajectory (MC/models/pure_pursuit_model.cpp:60-114)
double critic_pure_pursuit(const lporacle_ctx& c, const Traj& t, double tw, double ow) {
  const Math& M = *c.m;
  if (c.plan.empty() || t.poses.size() < 2) return -4.0;
  Affine last = pose_to_affine(t.poses.back().p);
  last = affine_inverse(last);
  const Affine goal = pose_to_affine(&c.plan[c.plan.size() - 7]);
  const Affine diff = affine_mul(last, goal);
  double q[4];
  matrix_to_quat(diff.L, q);  // tf2::eigenToTransform
  // tf2::Matrix3x3(q) (A5): only m00, m10, m20 are needed for yaw
  const double d = ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3];
  const double s = 2.0 / d;
  const double ys = q[1] * s, zs = q[2] * s;
  const double wy = q[3] * ys, wz = q[3] * zs;
  const double xy = q[0] * ys, xz = q[0] * zs;
  const double yy = q[1] * ys, zz = q[2] * zs;
  const double m00 = 1.0 - (yy + zz), m10 = xy + wz, m20 = xz - wy;
  double yaw;
  if (fabs(m20) >= 1) {
    yaw = 0;
  } else {
    const double pitch = -M.asin_(m20);
    const double cp = M.cos_(pitch);
    yaw = M.atan2_(m10 / cp, m00 / cp);
  }
  yaw = M.fmod_(yaw + 3.1416, 3.1416);
  const double distance = sqrt((diff.t[0] * diff.t[0] + diff.t[1] * diff.t[1]) + diff.t[2] * diff.t[2]);
  return tw * distance + ow * yaw;
}

// StackedScoringModel::scoreTrajectory (MC/src/stacked_scoring_model.cpp:75-93)
void score_trajectory(const lporacle_ctx& c, Traj& t) {
  const double nan = std::numeric_limits<double>::quiet_NaN();
  t.critic_scores.assign(c.critics.size(), nan);
  t.first_hit_pose = -1;
  for (size_t k = 0; k < c.critics.size(); ++k) {
    const b200lp_critic& cr = c.critics[k];
    double v = 0.0;
    switch (cr.kind) {
      case B200LP_CRITIC_COLLISION: v = critic_collision(c, t, false); break;
      case B200LP_CRITIC_COLLISION_MIN_MAX: v = critic_collision(c, t, true); break;
      case B200LP_CRITIC_STICK_PATH: v = critic_stick_path(c, t); break;
      case B200LP_CRITIC_PURE_PURSUIT: v = critic_pure_pursuit(c, t, cr.translation_weight, cr.orientation_weight); break;
      case B200LP_CRITIC_TOWARD_GLOBAL_PLAN: v = critic_toward_global_plan(c, t, cr.weight); break;
      case B200LP_CRITIC_SHORTEST_ANGLE: {  // MC/models/shortest_angle_model.cpp:51-69
        if (c.q.heading_deviation >= 0) v = (t.thetav >= 0) ? cr.weight : cr.weight * 2;
        else v = (t.thetav >= 0) ? cr.weight * 2 : cr.weight;
        break;
      }
      case B200LP_CRITIC_TWIRLING: v = fabs(t.thetav) * cr.weight; break;  // MC/models/twirling_model.cpp:51-55
    }
    t.critic_scores[k] = v;
    if (v < 0) {
      t.cost = v;
      break;
    } else {
      t.cost += v;
    }
  }
}

void build_index(lporacle_ctx& c) {
  // ModelSharedData::updateData rebuilds the kd-tree every cycle when the cloud has >= 5 points
  // (MC/include/mpc_critics/model_shared_data.h:78-81)
  if (c.cloud.size() < 5) return;
  if (c.keep_index && c.index_valid) return;
  if (c.index_mode == LPORACLE_INDEX_GRID) c.grid.build(c.cloud);
#ifdef LPORACLE_WITH_NANOFLANN
  if (c.index_mode == LPORACLE_INDEX_NANOFLANN) {
    c.nf_cloud.pts = &c.cloud;
    c.nf_tree.reset(new NfTree(3, c.nf_cloud, nanoflann::KDTreeSingleIndexAdaptorParams(15)));
  }
#endif
  c.index_valid = true;
}

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

int lporacle_has_nanoflann(void) {
#ifdef LPORACLE_WITH_NANOFLANN
  return 1;
#else
  return 0;
#endif
}

int lporacle_create(lporacle_ctx** out, const b200lp_limits* limits, const b200lp_params* params,
                    const float* cuboid_xyz, const b200lp_critic* critics, int n_critics, int math_mode,
                    int index_mode) {
  if (!out || !limits || !params || !cuboid_xyz || n_critics < 0 || (n_critics && !critics)) return B200LP_E_INVALID;
  if (index_mode == LPORACLE_INDEX_NANOFLANN && !lporacle_has_nanoflann()) return B200LP_E_INVALID;
  lporacle_ctx* c = new lporacle_ctx();
  c->lim = *limits;
  c->par = *params;
  memcpy(c->cuboid, cuboid_xyz, sizeof(c->cuboid));
  c->critics.assign(critics, critics + n_critics);
  c->m = (math_mode == LPORACLE_MATH_LIBM) ? &kLibm : &kShared;
  c->index_mode = index_mode;
  *out = c;
  return B200LP_OK;
}

void lporacle_destroy(lporacle_ctx* ctx) { delete ctx; }
const char* lporacle_last_error(const lporacle_ctx* ctx) { return ctx ? ctx->err.c_str() : ""; }

int lporacle_set_cloud(lporacle_ctx* c, const void* pts, size_t n, size_t stride) {
  if (!c || (n && !pts) || stride < 12) return B200LP_E_INVALID;
  c->cloud.resize(n);
  const char* b = (const char*)pts;
  for (size_t i = 0; i < n; ++i) memcpy(&c->cloud[i], b + i * stride, 12);
  c->index_valid = false;
  return B200LP_OK;
}

int lporacle_set_plan(lporacle_ctx* c, const double* p, size_t n) {
  if (!c || (n && !p)) return B200LP_E_INVALID;
  c->plan.assign(p, p + n * 7);
  c->pcl_plan.resize(n);  // model_shared_data.h:83-91
  for (size_t i = 0; i < n; ++i) c->pcl_plan[i] = Pt{(float)p[i * 7], (float)p[i * 7 + 1], (float)p[i * 7 + 2]};
  return B200LP_OK;
}

int lporacle_set_eigen_association(int right) {
  g_assoc_right = right != 0;
  return B200LP_OK;
}

int lporacle_set_keep_index(lporacle_ctx* c, int keep) {
  if (!c) return B200LP_E_INVALID;
  c->keep_index = keep != 0;
  return B200LP_OK;
}

int lporacle_set_sample_stride(lporacle_ctx* c, int stride, int phase) {
  if (!c || stride < 1 || phase < 0 || phase >= stride) return B200LP_E_INVALID;
  c->stride = stride;
  c->phase = phase;
  return B200LP_OK;
}

int lporacle_samples(lporacle_ctx* c, const b200lp_query* q, float* out, int cap) {
  if (!c || !q) return B200LP_E_INVALID;
  const std::vector<Sample> s = make_samples(*c, *q);
  for (int i = 0; i < (int)s.size() && i < cap; ++i) memcpy(out + 3 * i, s[i].v, 12);
  return (int)s.size();
}

int lporacle_plan(lporacle_ctx* c, const b200lp_query* q, int n_threads, b200lp_result* out, double* s_index,
                  double* s_rollout, double* s_score) {
  if (!c || !q || !out) return B200LP_E_INVALID;
  c->q = *q;
  double t0 = now_s();
  build_index(*c);
  double t1 = now_s();

  // rollout loop (LP/local_planner/src/local_planner.cpp:549-557)
  const std::vector<Sample> samples = make_samples(*c, *q);
  c->n_samples = (int)samples.size();
  c->trajs.clear();
  for (int si = 0; si < (int)samples.size(); ++si) {
    if (si % c->stride != c->phase) continue;
    Traj t;
    if (generate_trajectory(*c, *q, samples[si].v, t)) {
      t.sample_index = si;
      c->trajs.push_back(std::move(t));
    }
  }
  double t2 = now_s();

  // getBestTrajectory (local_planner.cpp:447-480)
  n_threads = std::max(1, n_threads);
  if (n_threads == 1) {
    for (Traj& t : c->trajs) score_trajectory(*c, t);
  } else {
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (int k = 0; k < n_threads; ++k)
      th.emplace_back([&]() {
        for (;;) {
          const size_t b = next.fetch_add(2);
          if (b >= c->trajs.size()) break;
          for (size_t i = b; i < std::min(b + 2, c->trajs.size()); ++i) score_trajectory(*c, c->trajs[i]);
        }
      });
    for (auto& x : th) x.join();
  }
  memset(out, 0, sizeof(*out));
  out->best_id = -1;
  out->best_cost = -1;
  double minimum_cost = 9999999;
  int64_t n_poses = 0;
  int n_coll = 0;
  for (size_t i = 0; i < c->trajs.size(); ++i) {
    const Traj& t = c->trajs[i];
    n_poses += (int64_t)t.poses.size();
    if (t.first_hit_pose >= 0) ++n_coll;
    if (t.cost >= 0 && t.cost <= minimum_cost) {
      out->best_id = (int32_t)i;
      minimum_cost = t.cost;
      out->best_cost = t.cost;
      out->xv = t.xv;
      out->yv = t.yv;
      out->thetav = t.thetav;
    }
  }
  out->n_samples = c->n_samples;
  out->n_traj = (int32_t)c->trajs.size();
  out->n_poses = n_poses;
  out->n_collided = n_coll;
  double t3 = now_s();
  if (s_index) *s_index = t1 - t0;
  if (s_rollout) *s_rollout = t2 - t1;
  if (s_score) *s_score = t3 - t2;
  return B200LP_OK;
}

int lporacle_read_trajectories(lporacle_ctx* c, const b200lp_traj_view* v) {
  if (!c || !v) return B200LP_E_INVALID;
  const size_t nc = c->critics.size();
  for (size_t i = 0; i < c->trajs.size(); ++i) {
    const Traj& t = c->trajs[i];
    if (v->sample_index) v->sample_index[i] = t.sample_index;
    if (v->vel) memcpy(v->vel + 3 * i, t.vel, 12);
    if (v->num_steps) v->num_steps[i] = (int32_t)t.poses.size();
    if (v->time_delta) v->time_delta[i] = t.time_delta;
    if (v->cost) v->cost[i] = t.cost;
    if (v->critic_scores)
      for (size_t k = 0; k < nc; ++k) v->critic_scores[i * nc + k] = t.critic_scores[k];
    if (v->first_hit_pose) v->first_hit_pose[i] = t.first_hit_pose;
  }
  return B200LP_OK;
}

int lporacle_read_poses(lporacle_ctx* c, int32_t id, const b200lp_pose_view* v) {
  if (!c || !v || id < 0 || (size_t)id >= c->trajs.size()) return B200LP_E_INVALID;
  const Traj& t = c->trajs[id];
  int first_collision_critic = -1;
  for (size_t k = 0; k < c->critics.size(); ++k)
    if (c->critics[k].kind == B200LP_CRITIC_COLLISION || c->critics[k].kind == B200LP_CRITIC_COLLISION_MIN_MAX) {
      first_collision_critic = c->critics[k].kind;
      break;
    }
  for (size_t i = 0; i < t.poses.size(); ++i) {
    const Pose& po = t.poses[i];
    if (v->pose) memcpy(v->pose + 7 * i, po.p, 56);
    if (v->pcl_pose) memcpy(v->pcl_pose + 3 * i, po.pcl, 12);
    if (v->cuboid) memcpy(v->cuboid + 24 * i, po.cuboid, 96);
    if (v->aabb) {
      memcpy(v->aabb + 6 * i, po.mn, 12);
      memcpy(v->aabb + 6 * i + 3, po.mx, 12);
    }
    if (v->collide || v->n_r1) {
      int n = 0;
      bool hit = false;
      if (c->cloud.size() >= 5) {
        const BoxFrame f = box_frame(po);
        radius_search(*c, po.pcl, [&](uint32_t k) {
          ++n;
          if (first_collision_critic == B200LP_CRITIC_COLLISION_MIN_MAX) hit = hit || in_aabb(po, c->cloud[k]);
          else if (first_collision_critic == B200LP_CRITIC_COLLISION) hit = hit || in_box(f, c->cloud[k]);
          return true;
        });
      }
      if (v->collide) v->collide[i] = hit ? 1 : 0;
      if (v->n_r1) v->n_r1[i] = n;
    }
  }
  return B200LP_OK;
}

int lporacle_count_radius(lporacle_ctx* c, int64_t* sum, int64_t* n_poses) {
  if (!c) return B200LP_E_INVALID;
  int64_t s = 0, np = 0;
  for (const Traj& t : c->trajs)
    for (const Pose& po : t.poses) {
      ++np;
      if (c->cloud.size() >= 5) radius_search(*c, po.pcl, [&](uint32_t) { ++s; return true; });
    }
  if (sum) *sum = s;
  if (n_poses) *n_poses = np;
  return B200LP_OK;
}

// ---------------------------------------------------------------------------------------------
// SURVEY.md §8(f) rows next to the path
// ---------------------------------------------------------------------------------------------
// Local_Planner::prunePlan (LP/local_planner/src/local_planner.cpp:374-445) on the plan handed to setPlan (:322-343).
// nearestKSearch(K=1) over pcl::PointXYZ (float-cast positions, L2_Simple); equal distances resolve to the LOWEST index
// (a kd-tree's choice among exact ties is traversal dependent; the restatement fixes it).
int lporacle_prune_plan(const double* g7, size_t n, const double robot_xyz[3], double forward_distance,
                        double backward_distance, double* out_poses7, float* out_pcl_xyzi, size_t capacity,
                        b200lp_prune_info* info) {
  b200lp_prune_info I{};
  I.status = 0; I.nearest_index = -1; I.n_prune = 0; I.n_backward = 0;
  auto done = [&](int rc) { if (info) *info = I; return rc; };
  if (n < 3) { I.status = 1; return done(B200LP_OK); }              // :376-377
  const float q[3] = {(float)robot_xyz[0], (float)robot_xyz[1], (float)robot_xyz[2]};  // :385-388
  float best = 0.f; long nn = -1;
  for (size_t i = 0; i < n; ++i) {
    const Pt p = {(float)g7[i * 7], (float)g7[i * 7 + 1], (float)g7[i * 7 + 2]};       // :333-339
    const float d = l2_simple(q, p);
    if (nn < 0 || d < best) { best = d; nn = (long)i; }
  }
  I.nearest_index = (int32_t)nn;
  // :396-400 (sqrt(float) is sqrtf); the prune plan was cleared at :379-380 and stays empty
  if ((double)sqrtf(best) > 1.0) { I.status = 2; return done(B200LP_OK); }
  auto dist = [&](long a, long b) {  // getDistanceBTWPoseStamp (:346-352)
    const double dx = g7[a * 7] - g7[b * 7], dy = g7[a * 7 + 1] - g7[b * 7 + 1], dz = g7[a * 7 + 2] - g7[b * 7 + 2];
    return sqrt(dx * dx + dy * dy + dz * dz);
  };
  std::vector<long> back, fwd;
  long last = nn;
  for (long i = nn; i >= 0; --i) {            // :403-416
    back.push_back(i);
    if (i < nn) backward_distance -= dist(last, i);
    last = i;
    if (backward_distance < 0) break;
  }
  for (long i = nn; i < (long)n; ++i) {       // :420-438
    fwd.push_back(i);
    if (i > nn) forward_distance -= dist(last, i);
    last = i;
    if (forward_distance < 0) break;
  }
  I.n_backward = (int32_t)back.size();
  I.n_prune = (int32_t)(back.size() + fwd.size());
  if ((size_t)I.n_prune > capacity) return done(B200LP_E_INVALID);
  size_t k = 0;
  if (out_poses7) {  // prune_plan_.poses: backward part reversed (:418), then the forward part
    for (size_t j = back.size(); j-- > 0;) memcpy(out_poses7 + 7 * k++, g7 + 7 * back[j], 56);
    for (long i : fwd) memcpy(out_poses7 + 7 * k++, g7 + 7 * i, 56);
  }
  k = 0;
  if (out_pcl_xyzi) {  // pcl_prune_plan_: push order, NOT reversed; intensity tags :407,:424-429
    for (long i : back) { float* o = out_pcl_xyzi + 4 * k++; o[0] = (float)g7[i * 7]; o[1] = (float)g7[i * 7 + 1]; o[2] = (float)g7[i * 7 + 2]; o[3] = -1.f; }
    for (long i : fwd) { float* o = out_pcl_xyzi + 4 * k++; o[0] = (float)g7[i * 7]; o[1] = (float)g7[i * 7 + 1]; o[2] = (float)g7[i * 7 + 2]; o[3] = (i == 0) ? 0.f : 1.f; }
  }
  return done(B200LP_OK);
}

// perception_3d::PathBlockedStrategy::selfMark (src/dddmr_perception_3d/plugins/path_blocked_strategy.cpp:56-100)
// against the cloud last given to lporacle_set_cloud. Brute force: exact by construction.
int lporacle_path_blocked(lporacle_ctx* c, const float* pcl_xyzi, size_t n, double check_radius, b200lp_blocked* out) {
  if (!c || !out) return B200LP_E_INVALID;
  b200lp_blocked B{};
  B.n_total = (int32_t)n;
  if (!(c->cloud.size() <= 5 || n == 0)) {  // :62-64
    const float r2 = (float)(check_radius * check_radius);  // pcl::KdTreeFLANN::radiusSearch -> FLANN radius^2 as float
    for (size_t i = 0; i < n; ++i) {
      if (pcl_xyzi[4 * i + 3] < 0) continue;  // :79-80 backward poses are skipped
      ++B.n_checked;
      const float q[3] = {pcl_xyzi[4 * i], pcl_xyzi[4 * i + 1], pcl_xyzi[4 * i + 2]};
      for (const Pt& p : c->cloud)
        if (l2_simple(q, p) < r2) { ++B.n_blocked; break; }
    }
    const float orig = (float)n, blocked = (float)B.n_blocked;  // :91-93 float division, then * 100.0 in double
    B.ratio = (blocked) / (orig) * 100.0;
  }
  B.opinion = B.ratio > 0.0 ? 1 : 0;  // PATH_BLOCKED_WAIT : PASS (:96-97)
  *out = B;
  return B200LP_OK;
}

// perception_3d::MultiLayerSpinningLidar::cbSensor from the first transform to the local-planner observation
// (src/dddmr_perception_3d/plugins/multilayer_spinning_lidar.cpp:232-269), with the PCL 1.15 filters it calls restated
// from their published algorithms (PCL is not vendored under /root/reference: this part of the parity is UNPINNED):
//   pcl::transformPointCloud(cloud, cloud, Affine3d)  common/impl/transforms.hpp  
//   pcl::PassThrough<PointXYZ>::applyFilterIndices    filters/impl/passthrough.hpp
//   pcl::VoxelGrid<PointXYZ>::applyFilter             filters/impl/voxel_grid.hpp
//   pcl::CentroidPoint<PointXYZ> / AccumulatorXYZ     common/impl/accumulators.hpp
// order_mode 0: the points of a voxel are added in scan order (a stable sort of the voxel indices) — the order the device
// path defines; 1: in the order libstdc++'s std::sort leaves them (unstable, like PCL's boost spreadsort / std::sort), to
// measure how far an unstable order moves a centroid. Voxel set, voxel order and counts do not depend on the mode.
// out_xyz1: capacity x 4 floats (pcl::PointXYZ layout, data[3] = 1). Returns B200LP_E_INVALID if the capacity is too small.
int lporacle_sensor_observation(const void* scan, size_t n, size_t stride, const double base_from_sensor[7],
                                const double global_from_base[7], const b200lp_sensor_params* sp, int order_mode,
                                float* out_xyz1, size_t capacity, b200lp_observation_info* info) {
  if (!sp || !info || (n && !scan)) return B200LP_E_INVALID;
  struct P3 { float x, y, z; };
  std::vector<P3> cloud(n);
  for (size_t i = 0; i < n; ++i) memcpy(&cloud[i], (const char*)scan + i * stride, 12);
  auto transform = [](std::vector<P3>& c, const Affine& a) {  // :232-233, :266-267
    for (P3& p : c) {
      const double x = p.x, y = p.y, z = p.z;
      p.x = (float)(((a.L[0][0] * x + a.L[0][1] * y) + a.L[0][2] * z) + a.t[0]);
      p.y = (float)(((a.L[1][0] * x + a.L[1][1] * y) + a.L[1][2] * z) + a.t[1]);
      p.z = (float)(((a.L[2][0] * x + a.L[2][1] * y) + a.L[2][2] * z) + a.t[2]);
    }
  };
  transform(cloud, pose_to_affine(base_from_sensor));  // tf2::transformToEigen(trans_b2s_)
  // three pcl::PassThrough runs (:240-251): x and y share the limits set once, z gets its own
  auto pass = [](std::vector<P3>& c, int field, float lo, float hi) {
    std::vector<P3> out;
    out.reserve(c.size());
    for (const P3& p : c) {
      if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
      const float v = field == 0 ? p.x : field == 1 ? p.y : p.z;
      if (!std::isfinite(v)) continue;
      if (v < lo || v > hi) continue;
      out.push_back(p);
    }
    c.swap(out);
  };
  const float wlo = (float)(-sp->perception_window_size), whi = (float)(sp->perception_window_size);
  pass(cloud, 0, wlo, whi);
  pass(cloud, 1, wlo, whi);
  pass(cloud, 2, 0.0f, (float)(sp->marking_height));
  info->n_scan = (int64_t)n;
  info->n_window = (int64_t)cloud.size();
  info->ms_device = 0.f;
  info->n_launches = 0;
  info->ms_upload = 0.f;
  info->reserved_ = 0;
  // pcl::VoxelGrid (:253-256)
  const float leaf = sp->leaf_size > 0.f ? sp->leaf_size : 0.1f;
  const float inv = 1.0f / leaf;  // inverse_leaf_size_ = Ones / leaf_size_
  std::vector<P3> out;
  if (!cloud.empty()) {
    float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
    float mx[3] = {-mn[0], -mn[1], -mn[2]};
    for (const P3& p : cloud) {  // getMinMax3D
      mn[0] = std::min(mn[0], p.x); mx[0] = std::max(mx[0], p.x);
      mn[1] = std::min(mn[1], p.y); mx[1] = std::max(mx[1], p.y);
      mn[2] = std::min(mn[2], p.z); mx[2] = std::max(mx[2], p.z);
    }
    const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1,
                  dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
    if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) {
      out = cloud;  // "Leaf size is too small for the input dataset": output = *input_
    } else {
      int min_b[3], max_b[3], div_b[3];
      for (int a = 0; a < 3; ++a) {
        min_b[a] = (int)std::floor(mn[a] * inv);
        max_b[a] = (int)std::floor(mx[a] * inv);
        div_b[a] = max_b[a] - min_b[a] + 1;
      }
      const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
      struct Ix { unsigned idx; unsigned pt; };
      std::vector<Ix> iv;
      iv.reserve(cloud.size());
      for (size_t i = 0; i < cloud.size(); ++i) {
        const int ijk0 = (int)(std::floor(cloud[i].x * inv) - (float)min_b[0]);
        const int ijk1 = (int)(std::floor(cloud[i].y * inv) - (float)min_b[1]);
        const int ijk2 = (int)(std::floor(cloud[i].z * inv) - (float)min_b[2]);
        iv.push_back({(unsigned)(ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2]), (unsigned)i});
      }
      auto less = [](const Ix& a, const Ix& b) { return a.idx < b.idx; };
      if (order_mode == 1) std::sort(iv.begin(), iv.end(), less);
      else std::stable_sort(iv.begin(), iv.end(), less);
      for (size_t i = 0; i < iv.size();) {
        size_t j = i;
        float sx = 0.0f, sy = 0.0f, sz = 0.0f;  // AccumulatorXYZ: xyz = Zero; xyz += point
        while (j < iv.size() && iv[j].idx == iv[i].idx) {
          sx += cloud[iv[j].pt].x;
          sy += cloud[iv[j].pt].y;
          sz += cloud[iv[j].pt].z;
          ++j;
        }
        const float fn = (float)(j - i);  // xyz / n
        out.push_back({sx / fn, sy / fn, sz / fn});
        i = j;
      }
    }
  }
  if (sp->is_local_planner) transform(out, pose_to_affine(global_from_base));  // :264-268
  info->n_points = (int64_t)out.size();
  if (out.size() > capacity) return B200LP_E_INVALID;
  for (size_t i = 0; i < out.size(); ++i) {
    out_xyz1[4 * i] = out[i].x;
    out_xyz1[4 * i + 1] = out[i].y;
    out_xyz1[4 * i + 2] = out[i].z;
    out_xyz1[4 * i + 3] = 1.0f;
  }
  return B200LP_OK;
}

float lporacle_sinf(int mode, float x) { return (mode ? kLibm : kShared).sinf_(x); }
float lporacle_cosf(int mode, float x) { return (mode ? kLibm : kShared).cosf_(x); }
double lporacle_sin(int mode, double x) { return (mode ? kLibm : kShared).sin_(x); }
double lporacle_cos(int mode, double x) { return (mode ? kLibm : kShared).cos_(x); }
double lporacle_asin(int mode, double x) { return (mode ? kLibm : kShared).asin_(x); }
double lporacle_atan2(int mode, double y, double x) { return (mode ? kLibm : kShared).atan2_(y, x); }
double lporacle_fmod(int mode, double x, double y) { return (mode ? kLibm : kShared).fmod_(x, y); }

int64_t lporacle_sincosf_mismatches(uint32_t lo, uint32_t hi, uint32_t step) {
  int64_t bad = 0;
  if (!step) step = 1;
  for (uint64_t u = lo; u < hi; u += step)
    for (int sg = 0; sg < 2; ++sg) {
      const uint32_t bits = (uint32_t)u | (sg ? 0x80000000u : 0u);
      float y;
      memcpy(&y, &bits, 4);
      const float a = lpm::sinf(y), b = ::sinf(y), cc = lpm::cosf(y), d = ::cosf(y);
      if (memcmp(&a, &b, 4) != 0) ++bad;
      if (memcmp(&cc, &d, 4) != 0) ++bad;
    }
  return bad;
}

}  // extern "C"
