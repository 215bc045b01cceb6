// lp_device.cuh — device-side building blocks of the rollout-and-score path (sm_100a).
//
// Arithmetic contract: this translation unit is compiled with -fmad=false; every float/double
// expression below evaluates exactly like the CPU oracle's (test infrastructure under oracle/), which restates the
// reference. The ONLY fused multiply-adds are the explicit __fmaf_rn calls of the conservative
// pre-test in sweep_points(), whose outcome is re-decided by the exact test before it can matter.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200lp.h"
#include "lp_math.h"

namespace lp {

#ifndef B200LP_WARPS
#define B200LP_WARPS 4
#endif
constexpr int kWarpsPerCta = B200LP_WARPS;
constexpr int kThreads = kWarpsPerCta * 32;
#ifndef B200LP_GROUP
#define B200LP_GROUP 4
#endif
#ifndef B200LP_SC_COMPARE
#define B200LP_SC_COMPARE 0  // 1: short-circuit && in the sweep's pre-test compare (A/B builds)
#endif
constexpr int kGroup = B200LP_GROUP;  // consecutive poses swept against one candidate stream (power of two)
constexpr int kPreStride = 5;   // float4 per pose in the pre-test stash (4 used + 1 pad: conflict-free 80-byte stride)
constexpr int kMaxAxis = 2048;  // cap on samples per velocity axis (incl. the inserted zero)
constexpr int kPlanWarp = 96;   // prune-plan points every warp stages in shared memory (longer plans are scanned in global memory)
constexpr unsigned kFull = 0xffffffffu;

// per-warp pose stash (structure of arrays: [field][lane])
enum Field {
  F_CX = 0, F_CY, F_CZ,                     // cuboid centre (collision_model.cpp:86-95)
  F_AXX, F_AXY, F_AXZ,                      // unit x axis
  F_AYX, F_AYY, F_AYZ,                      // unit y axis
  F_AZX, F_AZY, F_AZZ,                      // unit z axis
  F_HX, F_HY, F_HZ,                         // half extents (exactly representable in float)
  F_PX, F_PY, F_PZ,                         // Trajectory::getPCLPoint
  F_MNX, F_MNY, F_MNZ, F_MXX, F_MXY, F_MXZ, // Trajectory::getCuboidMinMax
  F_COUNT
};

// Cell range of a pose's candidate box = (AABB +- margin) ∩ (pose +- (1+margin)), clamped to the grid.
// Every point that can pass BOTH exact tests of that pose lies in these cells. Empty: x0 > x1.
struct CellBox {
  int x0, x1, y0, y1, z0, z1;
};

struct GridDev {
  const float4* pts;           // cell-sorted points: x,y,z, w = original index bits
  const uint32_t* cell_start;  // n_cells + 1 offsets into pts
  const uint32_t* sat;         // summed-volume table, (nz+1)(ny+1)(nx+1): points with cell < (X,Y,Z) componentwise
  float org[3];
  float inv_xy, inv_z;
  int nx, ny, nz;
  uint32_t n_kept;  // finite points in pts
  uint32_t n_raw;   // raw cloud size: the reference's `points.size() < 5` rule uses this
  float cmax;       // max |coordinate| of the grid bounds (scales the pre-test slack)
};

struct CriticDev {
  int kind;
  int pad;
  double weight, tw, ow;
};

struct Consts {
  b200lp_limits lim;
  b200lp_params par;
  float cuboid[8][3];
  int n_critics;
  int pad;
  CriticDev critics[B200LP_MAX_CRITICS];
  // axis-aligned bounding box of the 8 cuboid vertices in the robot frame (centre, half extents) and the slack that
  // covers the float arithmetic of loose_box(): filled by b200lp_create
  float box_c[3], box_h[3];
  float box_slack, pad2;
};

struct RobotIn {
  double pose[7];
  double twist[3];
  double max_speed_override;
  double heading_deviation;
  int64_t plan_off;
  int32_t plan_n;
  int32_t goal_valid;  // goal[] holds the last pose of the robot's prune plan (the host had the plan when it filled this record)
  double goal[7];      // prune_plan_.poses.back(): what the pure-pursuit critic aims at; lets prep_kernel start before the plan
                       // table itself has reached the device (fleet calls upload it beside prep_kernel)
};

struct RobotMeta {
  int32_t n_samples;  // |sample_params_|
  int32_t n_traj;     // all valid trajectories of the robot (global id space)
  int32_t t_begin, t_end;  // id range this launch scores (sample shard)
  int32_t error;      // 1 = num_steps exceeded B200LP_MAX_STEPS
  int32_t pad;
  long long n_poses;  // sum num_steps over [t_begin, t_end)
};

// Result block of a single-robot cycle in mapped pinned host memory: the last CTA of plan_kernel stores it straight into
// host memory and raises `seq`, the host spins on `seq` instead of enqueueing two copies and synchronising the stream.
struct DirectOut {
  b200lp_result r;
  RobotMeta m;
  uint32_t cycle_ns;     // device time of the cycle: first CTA of prep_kernel -> last CTA of plan_kernel (globaltimer)
  uint32_t pad;
  uint32_t peer_ns[16];  // sample-sharded cycles: the same figure of every rank (what the shard cuts are moved with)
  unsigned long long seq;
};

struct alignas(16) BlockBest {    // 32 bytes: read back with two 16-byte loads
  unsigned long long cost_bits;  // ~0 = none
  int32_t id;
  int32_t n_collided;
  int32_t n_scored;              // trajectories this CTA actually scored (checked against prep_kernel's count)
  int32_t pad;
  long long poses_scored;        // sum of their num_steps
};
static_assert(sizeof(BlockBest) == 32, "BlockBest layout");

// ---------------------------------------------------------------------------------------------
// double-precision pose algebra (mirrors oracle: quat_to_matrix, affine_mul, affine_inverse, matrix_to_quat)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void quat_to_matrix(double x, double y, double z, double w, double* R /*9 row-major*/) {
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
  R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}

__device__ __forceinline__ void matrix_to_quat(const double* m, double* q /*x,y,z,w*/) {
  double t = (m[0] + m[4]) + m[8];
  if (t > 0.0) {
    t = lpm::dsqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[7] - m[5]) * t;
    q[1] = (m[2] - m[6]) * t;
    q[2] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[i * 4]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = lpm::dsqrt(((m[i * 4] - m[j * 4]) - m[k * 4]) + 1.0);
    double qq[4];
    qq[i] = 0.5 * t;
    t = 0.5 / t;
    qq[3] = (m[k * 3 + j] - m[j * 3 + k]) * t;
    qq[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
    qq[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
    q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2]; q[3] = qq[3];
  }
}

__device__ __forceinline__ double cof3(const double* m, int i, int j) {
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
}

// PurePursuitModel::scoreTrajectory body (pure_pursuit_model.cpp:86-113) for a trajectory end pose given
// as affine (L,t) and the plan end pose as affine (gL,gt).
// Writes the translation distance and the wrapped yaw; the critic's value is tw * distance + ow * yaw.
__device__ __noinline__ void pure_pursuit_terms(const double* L, const double* t, const double* gL, const double* gt,
                                                double* distance_out, double* yaw_out) {
  // PoseStamped orientation = Quaterniond(L); the critic turns it back into a matrix
  double q[4];
  matrix_to_quat(L, q);
  double A[9];
  quat_to_matrix(q[0], q[1], q[2], q[3], A);
  // inverse (Affine mode)
  const double c00 = cof3(A, 0, 0), c10 = cof3(A, 1, 0), c20 = cof3(A, 2, 0);
  const double det = (c00 * A[0] + c10 * A[3]) + c20 * A[6];
  const double invdet = 1.0 / det;
  double I[9];
  I[0] = c00 * invdet; I[1] = c10 * invdet; I[2] = c20 * invdet;
  I[3] = cof3(A, 0, 1) * invdet; I[4] = cof3(A, 1, 1) * invdet; I[5] = cof3(A, 2, 1) * invdet;
  I[6] = cof3(A, 0, 2) * invdet; I[7] = cof3(A, 1, 2) * invdet; I[8] = cof3(A, 2, 2) * invdet;
  double it[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) it[i] = -((I[i * 3] * t[0] + I[i * 3 + 1] * t[1]) + I[i * 3 + 2] * t[2]);
  // diff = inv * goal
  double D[9], dt[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) D[i * 3 + j] = (I[i * 3] * gL[j] + I[i * 3 + 1] * gL[3 + j]) + I[i * 3 + 2] * gL[6 + j];
    dt[i] = ((I[i * 3] * gt[0] + I[i * 3 + 1] * gt[1]) + I[i * 3 + 2] * gt[2]) + it[i];
  }
  matrix_to_quat(D, q);
  const double d = ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3];
  const double s = 2.0 / d;
  const double ys = q[1] * s, zs = q[2] * s;
  const double wy = q[3] * ys, wz = q[3] * zs;
  const double xy = q[0] * ys, xz = q[0] * zs;
  const double yy = q[1] * ys, zz = q[2] * zs;
  const double m00 = 1.0 - (yy + zz), m10 = xy + wz, m20 = xz - wy;
  double yaw;
  if (lpm::dabs(m20) >= 1.0) {
    yaw = 0.0;
  } else {
    const double pitch = -lpm::asin(m20);
    const double cp = lpm::cos(pitch);
    yaw = lpm::atan2(m10 / cp, m00 / cp);
  }
  {
    const double v = yaw + 3.1416;
    const double r = lpm::fmod_pos(lpm::dabs(v), 3.1416);
    yaw = v < 0.0 ? -r : r;
  }
  *distance_out = lpm::dsqrt((dt[0] * dt[0] + dt[1] * dt[1]) + dt[2] * dt[2]);
  *yaw_out = yaw;
}

// ---------------------------------------------------------------------------------------------
// grid helpers
// ---------------------------------------------------------------------------------------------
// Cell coordinate of a world coordinate. The SAME function bins points at build time and maps query
// boxes to cell ranges; float subtraction, multiplication and floor are monotone, so a point with
// v >= lo can never land in a cell below cell(lo).
__device__ __forceinline__ float cell_f(float v, float org, float inv) { return floorf((v - org) * inv); }
__device__ __forceinline__ int cell_clamped(float v, float org, float inv, int n) {
  const float f = cell_f(v, org, inv);
  return (int)fminf(fmaxf(f, 0.0f), (float)(n - 1));
}

// slack of the conservative pre-test and of the candidate box, both scaled by the map extent
__device__ __forceinline__ float pretest_delta(const GridDev& g) { return 4e-6f + 1e-6f * g.cmax; }
__device__ __forceinline__ float box_margin(const GridDev& g) { return 1e-3f + 1e-5f * g.cmax; }

// ---------------------------------------------------------------------------------------------
// rollout: one warp, 32 consecutive steps. State (x,y,th) is carried in every lane.
// DD simple / rotate: dd_simple…cpp:457-464; omni: omni_simple…cpp:498-505.
// On return lane k holds pose base+k (after step base+k) in (px,py,pth).
// ---------------------------------------------------------------------------------------------
struct Carry {
  float x, y, th;
};

__device__ __forceinline__ void rollout32(Carry& c, int lane, int theory, float vx, float vy, float w, double dt,
                                          float& px, float& py, float& pth) {
  // heading chain: th' = (float)(th + w*dt); w*dt is loop invariant
  const double wdt = (double)w * dt;
  float th_old_mine = 0.f, th_new_mine = 0.f;
  float th = c.th;
#pragma unroll 8
  for (int k = 0; k < 32; ++k) {
    const float tn = (float)((double)th + wdt);
    if (k == lane) {
      th_old_mine = th;
      th_new_mine = tn;
    }
    th = tn;
  }
  c.th = th;
  // per-step increments, one step per lane
  double ex, ey;
  if (theory == B200LP_THEORY_OMNI_SIMPLE) {
    const double a = 1.57079632679489661923 + (double)th_old_mine;  // M_PI_2 + pos[2]
    ex = ((double)(vx * lpm::cosf(th_old_mine)) + (double)vy * lpm::cos(a)) * dt;
    ey = ((double)(vx * lpm::sinf(th_old_mine)) + (double)vy * lpm::sin(a)) * dt;
  } else {
    ex = (double)(vx * lpm::cosf(th_old_mine)) * dt;
    ey = (double)(vx * lpm::sinf(th_old_mine)) * dt;
  }
  // position chains
  float x = c.x, y = c.y;
  float xm = 0.f, ym = 0.f;
#pragma unroll 8
  for (int k = 0; k < 32; ++k) {
    const double exk = __shfl_sync(kFull, ex, k);
    const double eyk = __shfl_sync(kFull, ey, k);
    x = (float)((double)x + exk);
    y = (float)((double)y + eyk);
    if (k == lane) {
      xm = x;
      ym = y;
    }
  }
  c.x = x;
  c.y = y;
  px = xm;
  py = ym;
  pth = th_new_mine;
}

// World transform of one pose: G = pos_af3 * [Rz(th), (x,y,0)] (dd_simple…cpp:416-433).
// Terms that multiply an exact zero of the planar transform are dropped: a + (±0) == a.
__device__ __forceinline__ void pose_affine(const double* R0, const double* t0, float x, float y, float th,
                                            double* L, double* t) {
  const double ang = (double)th;
  double s, co;
  lpm::sincos(ang, &s, &co);
  const double r22 = (1.0 - co) + co;
  const double xd = (double)x, yd = (double)y;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double a = R0[i * 3], b = R0[i * 3 + 1], cc = R0[i * 3 + 2];
    L[i * 3 + 0] = a * co + b * s;
    L[i * 3 + 1] = a * (-s) + b * co;
    L[i * 3 + 2] = cc * r22;
    t[i] = (a * xd + b * yd) + t0[i];
  }
}

// Cell range of the box [lo, hi] clamped to the grid, and whether any cloud point lies in those cells (inclusion-exclusion
// on the summed-volume table, exact in modular u32 arithmetic): most poses drive through free space and are culled right
// here. *cb is the range when the answer is yes, the empty box (x0 > x1) otherwise.
__device__ __forceinline__ bool cells_with_points(const GridDev& g, const float* lo, const float* hi, CellBox* cb) {
  const float fx0 = cell_f(lo[0], g.org[0], g.inv_xy), fx1 = cell_f(hi[0], g.org[0], g.inv_xy);
  const float fy0 = cell_f(lo[1], g.org[1], g.inv_xy), fy1 = cell_f(hi[1], g.org[1], g.inv_xy);
  const float fz0 = cell_f(lo[2], g.org[2], g.inv_z), fz1 = cell_f(hi[2], g.org[2], g.inv_z);
  const bool empty = !(lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2]) || fx1 < 0.f || fy1 < 0.f || fz1 < 0.f ||
                     fx0 > (float)(g.nx - 1) || fy0 > (float)(g.ny - 1) || fz0 > (float)(g.nz - 1) || g.n_kept == 0;
  cb->x0 = cb->y0 = cb->z0 = 0x7fffffff;
  cb->x1 = cb->y1 = cb->z1 = -1;
  if (empty) return false;
  const int x0 = (int)fmaxf(fx0, 0.f), x1 = (int)fminf(fx1, (float)(g.nx - 1));
  const int y0 = (int)fmaxf(fy0, 0.f), y1 = (int)fminf(fy1, (float)(g.ny - 1));
  const int z0 = (int)fmaxf(fz0, 0.f), z1 = (int)fminf(fz1, (float)(g.nz - 1));
  const size_t sx = (size_t)(g.nx + 1), sy = (size_t)(g.ny + 1) * sx;
  const uint32_t* s0 = g.sat + (size_t)z0 * sy;
  const uint32_t* s1 = g.sat + (size_t)(z1 + 1) * sy;
  const size_t a0 = (size_t)y0 * sx, a1 = (size_t)(y1 + 1) * sx;
  const uint32_t up = (__ldg(s1 + a1 + x1 + 1) - __ldg(s1 + a1 + x0)) - (__ldg(s1 + a0 + x1 + 1) - __ldg(s1 + a0 + x0));
  const uint32_t dn = (__ldg(s0 + a1 + x1 + 1) - __ldg(s0 + a1 + x0)) - (__ldg(s0 + a0 + x1 + 1) - __ldg(s0 + a0 + x0));
  if (up - dn == 0u) return false;
  cb->x0 = x0; cb->x1 = x1;
  cb->y0 = y0; cb->y1 = y1;
  cb->z0 = z0; cb->z1 = z1;
  return true;
}

// Cheap, conservative stand-in for pose_geometry's candidate box, in float: a box that CONTAINS the exact candidate box
// of the pose (x, y, th), so a pose whose loose box holds no cloud point cannot collide and never needs the
// double-precision cuboid. R0f / t0f are the robot's world transform rounded to float. The cuboid is replaced by its
// robot-frame bounding box (C.box_c +- C.box_h), whose world extent along axis a is sum_j |L[a][j]| h[j]; sin/cos come
// from the fast intrinsics after a two-constant range reduction. Every rounding on the way — the intrinsics (< 5e-7 on
// [-pi, pi]), the float transform, the float-rounded vertices of the exact path — is far below C.box_slack +
// 2e-6 * cmax, which widens the box on every side.
__device__ __forceinline__ void loose_box(const Consts& C, const GridDev& g, const float* R0f, const float* t0f, float x,
                                          float y, float th, float* lo, float* hi) {
  const float k = rintf(th * 0.15915494309189535f);
  float r = __fmaf_rn(-k, 6.2831855f, th);
  r = __fmaf_rn(-k, -1.7484555e-07f, r);  // 2 pi = 6.2831855 - 1.7484555e-07
  float sn, cs;
  __sincosf(r, &sn, &cs);
  const float m = box_margin(g);
  const float slack = C.box_slack + 2e-6f * g.cmax + m;
  const float rad = 1.0f + slack;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float ra = R0f[a * 3], rb = R0f[a * 3 + 1], rc = R0f[a * 3 + 2];
    const float l0 = __fmaf_rn(ra, cs, rb * sn), l1 = __fmaf_rn(rb, cs, -(ra * sn));
    const float t = __fmaf_rn(ra, x, __fmaf_rn(rb, y, t0f[a]));
    const float c = __fmaf_rn(l0, C.box_c[0], __fmaf_rn(l1, C.box_c[1], __fmaf_rn(rc, C.box_c[2], t)));
    const float e = __fmaf_rn(fabsf(l0), C.box_h[0], __fmaf_rn(fabsf(l1), C.box_h[1], fabsf(rc) * C.box_h[2])) + slack;
    lo[a] = fmaxf(c - e, t - rad);
    hi[a] = fminf(c + e, t + rad);
  }
}

// Everything CollisionModel derives per pose from the transformed cuboid, written to the warp stash:
//  * stash (SoA [field][lane]): the reference's own quantities, read by the exact tests and the path critics;
//  * pre (AoS, kPreStride float4 per lane, may be nullptr): coefficients of the conservative pre-test of
//    sweep_points — the box axes with -k, k = fl(centre . axis) (x and y axis interleaved component by component, then the
//    z axis), and the half extents rounded up by delta;
//  * *cb (may be nullptr): the cell range of the pose's candidate box.
__device__ __forceinline__ void pose_geometry(const Consts& C, const GridDev& g, const double* L, const double* t,
                                              float* stash /* [F_COUNT][32] */, float4* pre, CellBox* cb, int lane,
                                              bool live, float* verts_out /* optional 24 floats, may be nullptr */) {
  float v[8][3];
  float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
  float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double x = (double)C.cuboid[k][0], y = (double)C.cuboid[k][1], z = (double)C.cuboid[k][2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      v[k][a] = (float)(((L[a * 3] * x + L[a * 3 + 1] * y) + L[a * 3 + 2] * z) + t[a]);
      mn[a] = fminf(mn[a], v[k][a]);
      mx[a] = fmaxf(mx[a], v[k][a]);
    }
  }
  if (verts_out) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int a = 0; a < 3; ++a) verts_out[k * 3 + a] = v[k][a];
  }
  float c[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float s = v[0][a];
#pragma unroll
    for (int k = 1; k < 8; ++k) s = s + v[k][a];
    c[a] = s / 8.0f;
  }
  const int other[3] = {3, 1, 2};
  float ax[3][3], half[3];
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    float d[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) d[a] = v[other[e]][a] - v[0][a];
    const float len = lpm::fsqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
    half[e] = len * 0.5f;  // == sqrtf(..)/2. (exact)
    // (float)((double)d / (2.*half)) == correctly rounded float quotient d/len (double rounding is
    // innocuous for division when the wide format has >= 2p+2 bits)
#pragma unroll
    for (int a = 0; a < 3; ++a) ax[e][a] = __fdiv_rn(d[a], len);
  }
  const float p[3] = {(float)t[0], (float)t[1], (float)t[2]};
  const float m = box_margin(g);
  if (!live) {  // lanes past num_steps: an empty candidate box and an unsatisfiable half extent
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      half[a] = -1.0f;
      mn[a] = 3.402823466e+38f;
      mx[a] = -3.402823466e+38f;
    }
  }
  stash[F_CX * 32 + lane] = c[0]; stash[F_CY * 32 + lane] = c[1]; stash[F_CZ * 32 + lane] = c[2];
  stash[F_AXX * 32 + lane] = ax[0][0]; stash[F_AXY * 32 + lane] = ax[0][1]; stash[F_AXZ * 32 + lane] = ax[0][2];
  stash[F_AYX * 32 + lane] = ax[1][0]; stash[F_AYY * 32 + lane] = ax[1][1]; stash[F_AYZ * 32 + lane] = ax[1][2];
  stash[F_AZX * 32 + lane] = ax[2][0]; stash[F_AZY * 32 + lane] = ax[2][1]; stash[F_AZZ * 32 + lane] = ax[2][2];
  stash[F_HX * 32 + lane] = half[0]; stash[F_HY * 32 + lane] = half[1]; stash[F_HZ * 32 + lane] = half[2];
  stash[F_PX * 32 + lane] = p[0]; stash[F_PY * 32 + lane] = p[1]; stash[F_PZ * 32 + lane] = p[2];
  stash[F_MNX * 32 + lane] = mn[0]; stash[F_MNY * 32 + lane] = mn[1]; stash[F_MNZ * 32 + lane] = mn[2];
  stash[F_MXX * 32 + lane] = mx[0]; stash[F_MXY * 32 + lane] = mx[1]; stash[F_MXZ * 32 + lane] = mx[2];
  if (pre) {
    const float delta = pretest_delta(g);
    float hb[3], k[3];
#pragma unroll
    for (int e = 0; e < 3; ++e) {
      k[e] = (float)(((double)c[0] * ax[e][0] + (double)c[1] * ax[e][1]) + (double)c[2] * ax[e][2]);
      hb[e] = (half[e] < 0.f) ? -1.0f : __fadd_ru(half[e], delta);
    }
    // the x and y axes interleaved: the sweep evaluates both dot products of a pose with packed (2 x fp32) FMAs
    pre[lane * kPreStride + 0] = make_float4(ax[0][0], ax[1][0], ax[0][1], ax[1][1]);
    pre[lane * kPreStride + 1] = make_float4(ax[0][2], ax[1][2], -k[0], -k[1]);
    pre[lane * kPreStride + 2] = make_float4(ax[2][0], ax[2][1], ax[2][2], -k[2]);
    pre[lane * kPreStride + 3] = make_float4(hb[0], hb[1], hb[2], 0.f);
  }
  if (cb) {
    // candidate box: every point that can pass BOTH exact tests lies inside it
    const float r = 1.0f + m;
    float lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      lo[a] = fmaxf(mn[a] - m, p[a] - r);
      hi[a] = fminf(mx[a] + m, p[a] + r);
    }
    cells_with_points(g, lo, hi, cb);
  }
}

// union of the cell boxes of each aligned group of kGroup lanes (empty boxes are neutral); every lane of a group
// ends up with the group's box
__device__ __forceinline__ void group_union(CellBox& b) {
#pragma unroll
  for (int o = 1; o < kGroup; o <<= 1) {
    b.x0 = min(b.x0, __shfl_xor_sync(kFull, b.x0, o)); b.x1 = max(b.x1, __shfl_xor_sync(kFull, b.x1, o));
    b.y0 = min(b.y0, __shfl_xor_sync(kFull, b.y0, o)); b.y1 = max(b.y1, __shfl_xor_sync(kFull, b.y1, o));
    b.z0 = min(b.z0, __shfl_xor_sync(kFull, b.z0, o)); b.z1 = max(b.z1, __shfl_xor_sync(kFull, b.z1, o));
  }
}

// exact tests, reference arithmetic ---------------------------------------------------------------
// FLANN L2_Simple<float>: ((dx*dx) + dy*dy) + dz*dz, strict < 1.0f
__device__ __forceinline__ float l2_simple(float qx, float qy, float qz, float px, float py, float pz) {
  float d = qx - px;
  float r = d * d;
  d = qy - py;
  r = r + d * d;
  d = qz - pz;
  r = r + d * d;
  return r;
}

// CollisionModel point-in-cuboid (collision_model.cpp:124-139) for stash column `col`
__device__ __forceinline__ bool exact_in_box(const float* stash, int col, float px, float py, float pz) {
  const float dx = px - stash[F_CX * 32 + col], dy = py - stash[F_CY * 32 + col], dz = pz - stash[F_CZ * 32 + col];
  const float xv = fabsf((dx * stash[F_AXX * 32 + col] + dy * stash[F_AXY * 32 + col]) + dz * stash[F_AXZ * 32 + col]);
  const float yv = fabsf((dx * stash[F_AYX * 32 + col] + dy * stash[F_AYY * 32 + col]) + dz * stash[F_AYZ * 32 + col]);
  const float zv = fabsf((dx * stash[F_AZX * 32 + col] + dy * stash[F_AZY * 32 + col]) + dz * stash[F_AZZ * 32 + col]);
  return xv <= stash[F_HX * 32 + col] && yv <= stash[F_HY * 32 + col] && zv <= stash[F_HZ * 32 + col];
}

// CollisionMinMaxModel containment (collision_min_max_model.cpp:74-77)
__device__ __forceinline__ bool exact_in_aabb(const float* stash, int col, float px, float py, float pz) {
  return px >= stash[F_MNX * 32 + col] && px <= stash[F_MXX * 32 + col] && py >= stash[F_MNY * 32 + col] &&
         py <= stash[F_MXY * 32 + col] && pz >= stash[F_MNZ * 32 + col] && pz <= stash[F_MXZ * 32 + col];
}

__device__ __forceinline__ bool exact_in_radius(const float* stash, int col, float px, float py, float pz) {
  return l2_simple(stash[F_PX * 32 + col], stash[F_PY * 32 + col], stash[F_PZ * 32 + col], px, py, pz) < 1.0f;
}

// ---------------------------------------------------------------------------------------------
// The obstacle query for kGroup consecutive poses (stash columns col0 .. col0+3) whose united cell box is `ub`
// (warp-uniform). Returns a 4-bit mask (warp-uniform) whose LOWEST set bit g is exact: pose col0+g is the first
// pose of the group that collides (higher bits may be missing once a lower pose is known to collide).
//   minmax == false: CollisionModel; minmax == true: CollisionMinMaxModel.
// Candidate set: the cell rows overlapping the box. A row is the run of x-adjacent cells [x0, x1] at fixed
// (iy, iz); cell-sorted storage makes it ONE contiguous range of float4, which the warp streams with coalesced
// 512-byte loads.
// ---------------------------------------------------------------------------------------------
#ifndef B200LP_CHECKS
#define B200LP_CHECKS 0  // 1: the checking build (libb200lp_checks.so): device-side bounds assertions at every indexed store and
                         // list access of the cycle and the grid build; a failed one prints its source line and traps, which
                         // the host sees as a CUDA error (compute-sanitizer is not available on the GPU pool). Never timed.
#endif
#if B200LP_CHECKS
#define LP_CHECK(cond)                                                                      \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      printf("b200lp check failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__);               \
      __trap();                                                                             \
    }                                                                                       \
  } while (0)
#else
#define LP_CHECK(cond) do { } while (0)
#endif

#ifndef B200LP_COUNT
#define B200LP_COUNT 0  // 1: the counting build (libb200lp_count.so): work counters for the roofline accounting, never timed
#endif
#if B200LP_COUNT
__device__ unsigned long long g_counters[4];  // candidates pre-tested, 32-candidate rounds, exact re-tests, groups swept
#endif

// What the sweep needs of the grid, by value: the function is NOT inlined (its inner loop wants the whole predicate file and
// a register allocation of its own; inlined into plan_kernel's nested loops its twelve compares per candidate serialised
// through the few predicates left over, +18 % on the kernel), and a reference to a kernel parameter would cost a local copy.
struct SweepGrid {
  const float4* pts;
  const uint32_t* cell_start;
  int nx, ny;
  float cmax;
};
#ifndef B200LP_SWEEP_INLINE
#define B200LP_SWEEP_INLINE 1
#endif
#if B200LP_SWEEP_INLINE
#define B200LP_SWEEP_ATTR __forceinline__
#else
#define B200LP_SWEEP_ATTR __noinline__
#endif
// How the sweep gets its candidate points (A/B builds; DESIGN.md §4.1 holds the measurements):
//   0  every lane loads its point of the current 32-candidate round with __ldg, then tests it;
//   1  the same, software-pipelined: the loads of the NEXT round are issued before the current one is tested;
//   2  bulk-asynchronous staging: one lane issues cp.async.bulk (the TMA unit's 1-D copy: a candidate run is ONE contiguous
//      range of float4) of the next <= 64 candidates into a two-deep per-warp ring in shared memory, completion signalled
//      through an mbarrier (complete_tx); the warp tests the current chunk out of shared memory meanwhile.
#ifndef B200LP_STREAM
#define B200LP_STREAM 0
#endif
constexpr int kRingPoints = 64;  // candidates per bulk copy (1 KB)
struct alignas(16) SweepRing {   // per warp
  float4 buf[2][kRingPoints];
  unsigned long long bar[2];
};
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ring_init(SweepRing* r, int lane) {
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&r->bar[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&r->bar[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
}
// one lane: announce `bytes` to the slot's barrier and start the bulk copy global -> shared
__device__ __forceinline__ void ring_issue(SweepRing* r, int slot, const float4* src, uint32_t n_points) {
  const uint32_t bytes = n_points * 16u, bar = smem_addr(&r->bar[slot]);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(&r->buf[slot][0])),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void ring_wait(SweepRing* r, int slot, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "RING_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra RING_DONE;\n"
      "bra RING_WAIT;\n"
      "RING_DONE:\n"
      "}\n" ::"r"(smem_addr(&r->bar[slot])),
      "r"(parity)
      : "memory");
}

// packed 2 x fp32 arithmetic (FFMA2): only ever used by the conservative pre-test, whose rounding does not matter
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

template <bool kMinMax>
__device__ B200LP_SWEEP_ATTR unsigned sweep_points(const SweepGrid g, const float* stash, const float4* pre, int col0,
                                                   int lane, const CellBox ub, SweepRing* ring, unsigned& ring_state) {
  if (ub.x0 > ub.x1) return 0u;
  const int nyr = ub.y1 - ub.y0 + 1;
  const int nrows = nyr * (ub.z1 - ub.z0 + 1);
  const float inv_nyr = 1.0f / (float)nyr;

  // Pre-test coefficients of the kGroup poses, in registers. CollisionModel: per pose the x and y axes of the box with
  // -k = -fl(centre . axis) folded in, interleaved so that both dot products of a pose are three packed FMAs (8 floats);
  // the half extents of x and y shared by the group (their maximum: the poses carry the same cuboid, the extents differ by
  // roundings only); and ONE slab for the z axis of all poses: the box's z axis is the robot's, which the rollout never
  // tilts, so the poses' slabs a_q . p - k_q in [-h_q, h_q] are intervals of (nearly) the same linear form. With A the
  // axis of the group's first pose, a_q . p = A . p - (A - a_q) . p and |(A - a_q) . p| <= |A - a_q|_1 cmax for every cloud
  // point, so A . p in [k_q - h_q - D_q, k_q + h_q + D_q]; the slab tested is the hull of those intervals, widened by the
  // rounding of both FMA chains and of this arithmetic (<= 2e-6 cmax in all). A superset of the per-pose test, like
  // everything in the pre-test: survivors are re-decided exactly. Per candidate: 12 packed + 3 scalar FMAs and 9 compares
  // chained through predicates; which pose a survivor belongs to is only worked out when a lane has one.
  // A pose that is no longer worth testing (past the end, or above the lowest pose known to collide) gets -k = +inf.
  // CollisionMinMaxModel: the poses' AABBs (exact compares need no slack).
  f32x2 c0[kGroup], c1[kGroup], c2[kGroup], c3[kGroup];  // (ax.x, ay.x), (ax.y, ay.y), (ax.z, ay.z), (-kx, -ky)
  float4 ca[kGroup], cb[kGroup];                         // min-max: AABB corners
  float hx = -1.0f, hy = -1.0f, zc = 0.f, zr = -1.0f;
  float ax_z = 0.f, ay_z = 0.f, az_z = 0.f;
  unsigned alive = 0u;  // poses still worth testing: live ones below the lowest pose known to collide (warp-uniform)
  const f32x2 never = pack2(__int_as_float(0x7f800000), __int_as_float(0x7f800000));
  {
    float lo = 3.402823466e+38f, hi = -3.402823466e+38f;
#pragma unroll
    for (int q = 0; q < kGroup; ++q) {
      const int col = col0 + q;
      if (kMinMax) {
        ca[q] = make_float4(stash[F_MNX * 32 + col], stash[F_MNY * 32 + col], stash[F_MNZ * 32 + col], 0.f);
        cb[q] = make_float4(stash[F_MXX * 32 + col], stash[F_MXY * 32 + col], stash[F_MXZ * 32 + col], 0.f);
        alive |= 1u << q;  // (lanes past the end carry an inverted box: never inside)
      } else {
        const float4 p0 = pre[col * kPreStride + 0];
        const float4 p1 = pre[col * kPreStride + 1];
        const float4 cz = pre[col * kPreStride + 2];
        const float4 h = pre[col * kPreStride + 3];
        c0[q] = pack2(p0.x, p0.y);
        c1[q] = pack2(p0.z, p0.w);
        c2[q] = pack2(p1.x, p1.y);
        c3[q] = pack2(p1.z, p1.w);
        if (q == 0) { ax_z = cz.x; ay_z = cz.y; az_z = cz.z; }
        if (h.x >= 0.f) {  // a live pose (lanes past the end carry half extents of -1)
          alive |= 1u << q;
          hx = fmaxf(hx, h.x);
          hy = fmaxf(hy, h.y);
          const float d = ((fabsf(ax_z - cz.x) + fabsf(ay_z - cz.y)) + fabsf(az_z - cz.z)) * g.cmax;
          const float s_q = h.z + d;
          lo = fminf(lo, -cz.w - s_q);
          hi = fmaxf(hi, -cz.w + s_q);
        } else {
          c3[q] = never;
        }
      }
    }
    if (!kMinMax && alive) {
      zc = 0.5f * (lo + hi);
      zr = 0.5f * (hi - lo) + (4e-6f + 4e-6f * g.cmax);
    }
  }

  unsigned hit = 0u;     // per-lane, exact hits
#if B200LP_COUNT
  unsigned long long c_cand = 0ull, c_rounds = 0ull, c_exact = 0ull;
#define B200LP_COUNT_FLUSH() do { if (lane == 0) { atomicAdd(&g_counters[0], c_cand); atomicAdd(&g_counters[1], c_rounds); \
    atomicAdd(&g_counters[2], c_exact); atomicAdd(&g_counters[3], 1ull); } } while (0)
#else
#define B200LP_COUNT_FLUSH() do { } while (0)
#endif
  // One 32-candidate round: pre-test, and — rarely — the exact decision. Returns true when the group's lowest pose
  // collides (nothing can precede it: the sweep is over).
  auto test_round = [&](const float4 p, uint32_t n_valid) -> bool {
#if B200LP_COUNT
    c_cand += (unsigned long long)n_valid * (unsigned long long)__popc(alive);  // (candidate, pose) pre-tests
    c_rounds += 1ull;
#else
    (void)n_valid;
#endif
    bool any = false;
    if (kMinMax) {
#pragma unroll
      for (int q = 0; q < kGroup; ++q)
        any |= (p.x >= ca[q].x) & (p.x <= cb[q].x) & (p.y >= ca[q].y) & (p.y <= cb[q].y) & (p.z >= ca[q].z) & (p.z <= cb[q].z);
    } else {
      // conservative superset of the exact test: |v' - v| <= delta
      const float tz = __fmaf_rn(p.x, ax_z, __fmaf_rn(p.y, ay_z, __fmaf_rn(p.z, az_z, -zc)));
      const f32x2 px = pack2(p.x, p.x), py = pack2(p.y, p.y), pz = pack2(p.z, p.z);
#pragma unroll
      for (int q = 0; q < kGroup; ++q) {
        float vx, vy;
        unpack2(fma2(px, c0[q], fma2(py, c1[q], fma2(pz, c2[q], c3[q]))), vx, vy);
        any |= (fabsf(vx) <= hx) & (fabsf(vy) <= hy);
      }
      any &= fabsf(tz) <= zr;
    }
    if (!__any_sync(kFull, any)) return false;
    // rare: find the poses concerned and decide with the reference's own arithmetic
#pragma unroll 1
    for (int q = 0; q < kGroup; ++q) {
      if (!((alive >> q) & 1u)) continue;  // (warp-uniform)
      const int col = col0 + q;
      bool in;
      if (kMinMax) {
        in = exact_in_aabb(stash, col, p.x, p.y, p.z);
      } else {
        const float4 p0 = pre[col * kPreStride + 0];
        const float4 p1 = pre[col * kPreStride + 1];
        const float vx = __fmaf_rn(p.x, p0.x, __fmaf_rn(p.y, p0.z, __fmaf_rn(p.z, p1.x, p1.z)));
        const float vy = __fmaf_rn(p.x, p0.y, __fmaf_rn(p.y, p0.w, __fmaf_rn(p.z, p1.y, p1.w)));
        in = any & (fabsf(vx) <= hx) & (fabsf(vy) <= hy);
        if (__any_sync(kFull, in)) in = in && exact_in_box(stash, col, p.x, p.y, p.z);
      }
      if (in && exact_in_radius(stash, col, p.x, p.y, p.z)) hit |= 1u << q;
    }
#if B200LP_COUNT
    c_exact += (unsigned long long)__popc(__ballot_sync(kFull, any));
#endif
    const unsigned wh = __reduce_or_sync(kFull, hit);
    if (wh & 1u) return true;
    if (wh) {
      alive &= (wh & (0u - wh)) - 1u;
      if (!kMinMax) {
#pragma unroll
        for (int q = 1; q < kGroup; ++q)
          if (!((alive >> q) & 1u)) c3[q] = never;
      } else {
#pragma unroll
        for (int q = 1; q < kGroup; ++q)
          if (!((alive >> q) & 1u)) ca[q].x = 3.402823466e+38f;
      }
    }
    return false;
  };

  constexpr uint32_t kChunk = B200LP_STREAM == 2 ? (uint32_t)kRingPoints : 32u;  // candidates fetched at a time
#if B200LP_STREAM == 2
  int slot = (int)(ring_state >> 2) & 1;  // ring_state: bit 0 / 1 = phase parity of barrier 0 / 1, bit 2 = slot to fill next
#else
  (void)ring; (void)ring_state;
#endif
  for (int r0 = 0; r0 < nrows; r0 += 32) {
    const int r = r0 + lane;
    uint32_t beg = 0, end = 0;
    if (r < nrows) {
      // r / nyr and r % nyr without the integer-division sequence: float estimate, then one correction step
      int qz = (int)((float)r * inv_nyr);
      int ry = r - qz * nyr;
      if (ry < 0) { --qz; ry += nyr; }
      else if (ry >= nyr) { ++qz; ry -= nyr; }
      const int iy = ub.y0 + ry, iz = ub.z0 + qz;
      const size_t base = ((size_t)iz * g.ny + iy) * (size_t)g.nx;
      beg = __ldg(g.cell_start + base + ub.x0);
      end = __ldg(g.cell_start + base + ub.x1 + 1);
    }
    unsigned rows = __ballot_sync(kFull, beg < end);
    // cursor over the chunks of these rows, one chunk ahead of the tests (all warp-uniform)
    uint32_t nj = 0u, ne = 0u;
    auto advance = [&]() -> bool {
      if (nj + kChunk < ne) { nj += kChunk; return true; }
      if (!rows) return false;
      const int src = __ffs(rows) - 1;
      rows &= rows - 1;
      nj = __shfl_sync(kFull, beg, src);
      ne = __shfl_sync(kFull, end, src);
      return true;
    };
    bool more = advance();
#if B200LP_STREAM == 1
    float4 pn = make_float4(0.f, 0.f, 0.f, 0.f);
    if (more) pn = __ldg(g.pts + min(nj + lane, ne - 1u));
#elif B200LP_STREAM == 2
    if (more && lane == 0) ring_issue(ring, slot, g.pts + nj, min(kChunk, ne - nj));
#endif
    while (more) {
      const uint32_t j0 = nj, e = ne;
      const uint32_t cnt = min(kChunk, e - j0);
      more = advance();
#if B200LP_STREAM == 0
      // lanes past the end of the row look at its last point again: no divergence, and a duplicate cannot change an any-hit
      LP_CHECK(e >= 1u && j0 < e);
      if (test_round(__ldg(g.pts + min(j0 + lane, e - 1u)), cnt)) { B200LP_COUNT_FLUSH(); return __reduce_or_sync(kFull, hit); }
#elif B200LP_STREAM == 1
      const float4 p = pn;
      if (more) pn = __ldg(g.pts + min(nj + lane, ne - 1u));  // in flight while this round is tested
      if (test_round(p, cnt)) { B200LP_COUNT_FLUSH(); return __reduce_or_sync(kFull, hit); }
#else
      if (more && lane == 0) ring_issue(ring, slot ^ 1, g.pts + nj, min(kChunk, ne - nj));  // in flight while this chunk is tested
      ring_wait(ring, slot, (ring_state >> slot) & 1u);
      ring_state ^= 1u << slot;
      bool over = false;
      for (uint32_t o = 0; o < cnt && !over; o += 32) over = test_round(ring->buf[slot][min(o + lane, cnt - 1u)], min(32u, cnt - o));
      __syncwarp();  // every lane has read the slot before it is filled again
      slot ^= 1;
      if (over) {
        if (more) {  // the copy already under way must land before the ring is used again
          ring_wait(ring, slot, (ring_state >> slot) & 1u);
          ring_state ^= 1u << slot;
          slot ^= 1;
        }
        ring_state = (ring_state & 3u) | ((unsigned)slot << 2);
        B200LP_COUNT_FLUSH();
        return __reduce_or_sync(kFull, hit);
      }
#endif
    }
  }
#if B200LP_STREAM == 2
  ring_state = (ring_state & 3u) | ((unsigned)slot << 2);
#endif
  B200LP_COUNT_FLUSH();
  return __reduce_or_sync(kFull, hit);
}

// nearestKSearch(K=1) against the prune-plan cloud: min float squared distance
// (stick_path_model.cpp:61-68, toward_global_plan_model.cpp:62-71)
__device__ __forceinline__ float plan_nn_d2(const float4* plan, int n, float qx, float qy, float qz) {
  float best = 3.402823466e+38f;
#pragma unroll 4
  for (int i = 0; i < n; ++i) {
    const float4 p = plan[i];
    const float d = l2_simple(qx, qy, qz, p.x, p.y, p.z);
    best = fminf(best, d);
  }
  return best;
}

// The same scan over the warp's shared-memory copy of the plan, two points per trip: pair k holds points 2k and 2k + 1 as
// xy[k] = (ax, ay, bx, by), z[k] = (az, bz) (an odd last point is doubled: a duplicate cannot change a minimum). Per point
// the arithmetic is l2_simple's, operation for operation — dx = qx - px, r = dx * dx, dy = qy - py, r = r + dy * dy, ... —
// each an individually rounded IEEE operation; only the INDEPENDENT subtractions and squares of a pair share an
// instruction (FADD2 / FMUL2: two fp32 lanes, each rounded like the scalar operation). The additions stay scalar:
// ptxas contracts a packed multiply feeding a packed add into FFMA2 even under --fmad=false (checked on CUDA 12.9), which
// would change the rounding; it does not split a packed multiply to feed a scalar add (SASS checked: no FFMA in this loop).
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float plan_nn_d2_pairs(const float4* xy, const float2* z, int n_pairs, float qx, float qy, float qz) {
  float best = 3.402823466e+38f;
  const f32x2 qxy = pack2(qx, qy), qzz = pack2(qz, qz);
#pragma unroll 2
  for (int k = 0; k < n_pairs; ++k) {
    const float4 p = xy[k];
    const float2 pz = z[k];
    const f32x2 da = sub2(qxy, pack2(p.x, p.y)), db = sub2(qxy, pack2(p.z, p.w)), dz = sub2(qzz, pack2(pz.x, pz.y));
    float ax2, ay2, bx2, by2, az2, bz2;
    unpack2(mul2(da, da), ax2, ay2);
    unpack2(mul2(db, db), bx2, by2);
    unpack2(mul2(dz, dz), az2, bz2);
    const float ra = __fadd_rn(__fadd_rn(ax2, ay2), az2);
    const float rb = __fadd_rn(__fadd_rn(bx2, by2), bz2);
    best = fminf(best, fminf(ra, rb));
  }
  return best;
}

}  // namespace lp
