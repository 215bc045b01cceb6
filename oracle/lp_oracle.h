/* lp_oracle.h — C ABI of the CPU oracle (TEST INFRASTRUCTURE, not product code).
 *
 * The oracle restates, on the CPU and in the reference's own arithmetic, the local-planner
 * rollout-and-score path of dddmr_navigation. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY STATUS: the reference ships no tests or golden vectors for this path and its own build
 * (colcon + ROS 2 Humble, PCL 1.15, FLANN, Eigen, tf2) cannot run in this image. What the oracle IS
 * pinned to: the reference's OWN sources for this path — three theories, StackedGenerator,
 * base_trajectory::Trajectory, seven critics, StackedScoringModel — compiled where they lie under
 * /root/reference into oracle/_ref/liblpref.so against stand-ins for the third-party headers they
 * include (oracle/ref_shims/) with the reference's vendored nanoflann as kd-tree;
 * tests/test_reference_sources.py requires IDENTICAL bits from that code and from this restatement
 * (libm mode) for every trajectory, pose, cuboid, AABB, critic value, cost and the selected id, on
 * the playground / C1 / C2 scenes, all theories, edge cases and 48 randomised scenes. What stays
 * UNPINNED: the arithmetic inside Eigen / PCL / FLANN / tf2 themselves (which
 * both sides follow), plus glibc sinf/cosf which lp_math.h matches exhaustively. See DESIGN.md §2.
 *
 * It reuses the product's public POD structs (include/b200lp.h) so tests feed both sides the
 * same bytes.
 */
#ifndef LP_ORACLE_H_
#define LP_ORACLE_H_
#include "../include/b200lp.h"

#ifdef __cplusplus
extern "C" {
#endif

#define LPORACLE_MATH_SHARED 0 /* dddmr_navigation_b200/csrc/lp_math.h — what the GPU evaluates */
#define LPORACLE_MATH_LIBM 1   /* glibc libm — what the reference binary calls */

#define LPORACLE_INDEX_BRUTE 0     /* scan the whole cloud per query */
#define LPORACLE_INDEX_GRID 1      /* uniform 1 m bucket grid (exact) */
#define LPORACLE_INDEX_NANOFLANN 2 /* reference's vendored nanoflann 1.5.1 kd-tree; only in oracle/_ref builds */

typedef struct lporacle_ctx lporacle_ctx;

int lporacle_has_nanoflann(void);
int lporacle_create(lporacle_ctx** out, const b200lp_limits* limits, const b200lp_params* params,
                    const float* cuboid_xyz, const b200lp_critic* critics, int n_critics, int math_mode,
                    int index_mode);
void lporacle_destroy(lporacle_ctx* ctx);
const char* lporacle_last_error(const lporacle_ctx* ctx);
int lporacle_set_cloud(lporacle_ctx* ctx, const void* pts, size_t n, size_t stride_bytes);
int lporacle_set_plan(lporacle_ctx* ctx, const double* xyz_qxyzw, size_t n);
/* Score only samples whose index % stride == phase (bounded-sample CPU baseline). Default 1, 0. */
int lporacle_set_sample_stride(lporacle_ctx* ctx, int stride, int phase);
/* The reference rebuilds its kd-tree every cycle (model_shared_data.h:78-81) and so does lporacle_plan; keep != 0 reuses an
 * index that is still valid for the current cloud (same contents, same answers) — for tests that plan many robots on one map. */
int lporacle_set_keep_index(lporacle_ctx* ctx, int keep);
/* Process-wide: right != 0 evaluates every 3-term inner product of the restated Eigen products (Affine3d * Affine3d,
 * Affine3d::inverse) as a0*b0 + (a1*b1 + a2*b2) instead of (a0*b0 + a1*b1) + a2*b2. Exists to MEASURE how much of the
 * result depends on an association that cannot be pinned offline (Eigen is not vendored under /root/reference). */
int lporacle_set_eigen_association(int right);
/* One cycle: (re)build the spatial index like ModelSharedData::updateData, roll out, score, argmin.
 * n_threads > 1 splits the trajectories over std::threads (a courtesy upper bound; the reference
 * is single-threaded). seconds_index / seconds_rollout / seconds_score may be NULL. */
int lporacle_plan(lporacle_ctx* ctx, const b200lp_query* q, int n_threads, b200lp_result* out,
                  double* seconds_index, double* seconds_rollout, double* seconds_score);
int lporacle_read_trajectories(lporacle_ctx* ctx, const b200lp_traj_view* view);
int lporacle_read_poses(lporacle_ctx* ctx, int32_t traj_id, const b200lp_pose_view* view);
int lporacle_count_radius(lporacle_ctx* ctx, int64_t* sum_n_r1, int64_t* n_poses);

/* Local_Planner::prunePlan and PathBlockedStrategy::selfMark restated (see lp_oracle.cpp). */
int lporacle_prune_plan(const double* global_plan7, size_t n, const double robot_xyz[3], double forward_distance,
                        double backward_distance, double* out_poses7, float* out_pcl_xyzi, size_t capacity,
                        b200lp_prune_info* info);
int lporacle_path_blocked(lporacle_ctx* ctx, const float* pcl_xyzi, size_t n, double check_radius, b200lp_blocked* out);

/* MultiLayerSpinningLidar::cbSensor's transform -> pass-through -> voxel filter -> transform,
 * PCL's filters restated from their published algorithms (UNPINNED: PCL is not vendored under /root/reference).
 * order_mode 0 = points of a voxel added in scan order (the device's order), 1 = in std::sort's (unstable) order. */
int lporacle_sensor_observation(const void* scan, size_t n, size_t stride_bytes, const double base_from_sensor[7],
                                const double global_from_base[7], const b200lp_sensor_params* params, int order_mode,
                                float* out_xyz1, size_t capacity, b200lp_observation_info* info);

/* Velocity samples exactly as initialise() leaves them in sample_params_ (xv,yv,thetav floats).
 * Returns the count; fills up to cap samples. */
int lporacle_samples(lporacle_ctx* ctx, const b200lp_query* q, float* out_xyz, int cap);

/* scalar math probes for tests/test_math.py (mode = LPORACLE_MATH_*) */
float lporacle_sinf(int mode, float x);
float lporacle_cosf(int mode, float x);
double lporacle_sin(int mode, double x);
double lporacle_cos(int mode, double x);
double lporacle_asin(int mode, double x);
double lporacle_atan2(int mode, double y, double x);
double lporacle_fmod(int mode, double x, double y);
/* vectorised mismatch counter: compares shared vs libm sinf and cosf over the float bit patterns
 * [lo, hi) with the given step (sign bit cleared and set); returns the number of mismatches. */
int64_t lporacle_sincosf_mismatches(uint32_t lo, uint32_t hi, uint32_t step);

#ifdef __cplusplus
}
#endif
#endif
