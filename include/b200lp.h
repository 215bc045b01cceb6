/* b200lp.h — C ABI of the B200-native local-planner rollout-and-score path.
 *
 * This is the drop-in boundary for dddmr_navigation's local planner hot path. Every entry
 * point replaces a piece of the reference's in-process C++ interface (paths relative to the
 * reference tree, src/dddmr_local_planner/ = LP/):
 *
 *   b200lp_create            <- plugin parameter loading: TrajectoryGeneratorTheory::onInitialize
 *                               (LP/trajectory_generators/theories/dd_simple_trajectory_generator_theory.cpp:43-234,
 *                                omni_simple…cpp:43-256, dd_rotate_inplace_theory.cpp:43-227) and
 *                               ScoringModel::onInitialize + the per-generator critic list
 *                               (LP/mpc_critics/src/mpc_critics_ros.cpp:60-82).
 *   b200lp_set_cloud         <- ModelSharedData::pcl_perception_ assignment + updateData() kd-tree build
 *                               (LP/local_planner/src/local_planner.cpp:582,586;
 *                                LP/mpc_critics/include/mpc_critics/model_shared_data.h:74-81).
 *   b200lp_set_plan          <- prune_plan_ hand-over to generator and critic shared data
 *                               (local_planner.cpp:530,583; model_shared_data.h:83-91).
 *   b200lp_plan              <- initializeTheories_wi_Shared_data() + the rollout loop + getBestTrajectory()
 *                               (local_planner.cpp:535-587; LP/trajectory_generators/src/stacked_generator.cpp:67-111;
 *                                LP/mpc_critics/src/stacked_scoring_model.cpp:75-93; local_planner.cpp:447-480).
 *   b200lp_plan_shard / b200lp_plan_shard_exchange / b200lp_peer_* <- the same cycle with the velocity samples split over
 *                               several GPUs (no reference analogue), argmin exchanged by the caller's collective or through
 *                               peer device memory over NVLink.
 *   b200lp_plan_batch        <- the same cycle for many independent robots (fleet sharding; no reference analogue,
 *                               one reference process serves one robot).
 *   b200lp_traj_count /
 *   b200lp_read_trajectories <- what the reference keeps in std::vector<base_trajectory::Trajectory>
 *                               (LP/base_trajectory/include/base_trajectory/trajectory.h:47-126) — read back for
 *                               RViz publishing (local_planner.cpp:554,569) and for parity tests.
 *   b200lp_sensor_observation /
 *   b200lp_aggregate_observations <- the producer of that cloud: MultiLayerSpinningLidar::cbSensor's transform -> pass-through ->
 *                               0.1 m voxel filter -> transform (dddmr_perception_3d/plugins/multilayer_spinning_lidar.cpp:232-269)
 *                               and StackedPerception::aggregateObservations (dddmr_perception_3d/src/stacked_perception.cpp:128-140);
 *                               the observation stays on the device and becomes the critics' cloud without a host round trip.
 *   b200lp_count_radius      <- diagnostic: |radiusSearch(pose, 1.0)| per pose (LP/mpc_critics/models/collision_model.cpp:122),
 *                               the n_r1 figure the roofline accounting is defined on.
 *
 * Conventions: extern "C"; plain pointers and sizes; every function returns 0 on success and a
 * negative B200LP_E_* code on failure, with text from b200lp_last_error(); the caller owns every
 * input buffer and may free it on return; a ctx owns its device memory and one CUDA stream, is
 * bound to one device and is NOT re-entrant (the reference serialises the same calls under its
 * perception and critics mutexes, local_planner.cpp:498,577). There is no CPU fallback: without a
 * usable CUDA device b200lp_create fails with B200LP_E_CUDA. Calls return when their RESULT is on the
 * host, not necessarily when the stream is idle: b200lp_set_cloud returns once the caller's buffer is
 * consumed (grid kernels still in flight), b200lp_plan once the kernel has written the result into
 * pinned host memory; everything later on the ctx is ordered behind that work on the same stream.
 */
#ifndef B200LP_H_
#define B200LP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200LP_ABI_VERSION 6

/* status codes */
#define B200LP_OK 0
#define B200LP_E_INVALID (-1) /* bad argument / parameter set outside the supported envelope */
#define B200LP_E_CUDA (-2)    /* CUDA runtime error or no device */
#define B200LP_E_STATE (-3)   /* call order violated (e.g. read before plan) */
#define B200LP_E_NOMEM (-4)

/* trajectory generator theories (LP/trajectory_generators/trajectory_generators.xml) */
#define B200LP_THEORY_DD_SIMPLE 0         /* trajectory_generators::DDSimpleTrajectoryGeneratorTheory */
#define B200LP_THEORY_OMNI_SIMPLE 1       /* trajectory_generators::OmniSimpleTrajectoryGeneratorTheory */
#define B200LP_THEORY_DD_ROTATE_INPLACE 2 /* trajectory_generators::DDRotateInplaceTheory */

/* critics (LP/mpc_critics/mpc_critics.xml) */
#define B200LP_CRITIC_COLLISION 0          /* mpc_critics::CollisionModel */
#define B200LP_CRITIC_COLLISION_MIN_MAX 1  /* mpc_critics::CollisionMinMaxModel */
#define B200LP_CRITIC_STICK_PATH 2         /* mpc_critics::StickPathModel */
#define B200LP_CRITIC_PURE_PURSUIT 3       /* mpc_critics::PurePursuitModel */
#define B200LP_CRITIC_TOWARD_GLOBAL_PLAN 4 /* mpc_critics::TowardGlobalPlanModel */
#define B200LP_CRITIC_SHORTEST_ANGLE 5     /* mpc_critics::ShortestAngleModel */
#define B200LP_CRITIC_TWIRLING 6           /* mpc_critics::TwirlingModel */
#define B200LP_MAX_CRITICS 8

/* Upper bound on poses per trajectory (num_steps). The reference has none; the shipped parameter
 * sets produce <= 252 (rotate in place). b200lp_create rejects parameter sets that could exceed it. */
#define B200LP_MAX_STEPS 512
/* Upper bound on prune-plan poses per robot. */
#define B200LP_MAX_PLAN 1024

typedef struct b200lp_ctx b200lp_ctx;

/* DDTrajectoryGeneratorLimits / OmniTrajectoryGeneratorLimits
 * (LP/trajectory_generators/include/trajectory_generators/dd_simple_trajectory_generator_limits.h:40-109,
 *  omni_simple_trajectory_generator_limits.h:40-125); rotation_speed from dd_rotate_inplace_theory.cpp:127. */
typedef struct b200lp_limits {
  double max_vel_x, min_vel_x;
  double max_vel_y, min_vel_y;         /* omni only */
  double max_vel_trans, min_vel_trans; /* omni only */
  double max_vel_theta, min_vel_theta;
  double acc_lim_x, acc_lim_y, acc_lim_theta;
  double deceleration_ratio;
  double max_motor_shaft_rpm, wheel_diameter, gear_ratio, robot_radius;
  double rotation_speed; /* rotate-in-place only */
  int32_t use_motor_constraint;
  int32_t reserved_;
} b200lp_limits;

/* DDTrajectoryGeneratorParams / OmniTrajectoryGeneratorParams
 * (…/dd_simple_trajectory_generator_params.h:42-76, omni_simple_trajectory_generator_params.h:42-80).
 * The *_sample fields are doubles truncated to int exactly as the reference passes them to
 * VelocityIterator (velocity_iterator.h:44). */
typedef struct b200lp_params {
  int32_t theory; /* B200LP_THEORY_* */
  int32_t reserved_;
  double controller_frequency;
  double sim_time;
  double linear_x_sample, linear_y_sample, angular_z_sample;
  double sim_granularity, angular_sim_granularity;
} b200lp_params;

/* One entry of the generator's ordered critic list (YAML plugins: order, mpc_critics_ros.cpp:63-82). */
typedef struct b200lp_critic {
  int32_t kind; /* B200LP_CRITIC_* */
  int32_t reserved_;
  double weight;             /* ScoringModel::weight_ */
  double translation_weight; /* PurePursuitModel only */
  double orientation_weight; /* PurePursuitModel only */
} b200lp_critic;

/* Voxel-grid configuration of the lethal cloud (no reference analogue: replaces KdTreeFLANN).
 * Zero-initialise for defaults. */
typedef struct b200lp_grid_config {
  float cell_xy;      /* cell edge in x and y, metres (default 0.2) */
  float cell_z;       /* cell height, metres (default 0.4) */
  uint32_t max_cells; /* cap on the dense cell count; cells are coarsened to fit (default 1<<26) */
  uint32_t reserved_;
} b200lp_grid_config;

/* Per-cycle inputs the reference copies into the shared-data structs (local_planner.cpp:528-533,579-583). */
typedef struct b200lp_query {
  double pose[7];            /* robot_pose_: translation x,y,z then rotation x,y,z,w (map -> base_link) */
  double twist[3];           /* robot_state_.twist.twist: linear.x, linear.y, angular.z */
  double max_speed_override; /* current_allowed_max_linear_speed_ (<= 0: ignored) */
  double heading_deviation;  /* ModelSharedData::heading_deviation_ (ShortestAngleModel) */
} b200lp_query;

/* What Local_Planner::getBestTrajectory leaves in best_traj plus cycle counters. */
typedef struct b200lp_result {
  int32_t best_id;   /* index into the generated-trajectory list; -1 = all rejected (ALL_TRAJECTORIES_FAIL) */
  int32_t n_samples; /* |sample_params_| after initialise() */
  int32_t n_traj;    /* trajectories generateTrajectory() accepted (= |trajectories_|) */
  int32_t n_collided;/* trajectories whose cost_ is a collision critic's -1 */
  int64_t n_poses;   /* sum of num_steps over generated trajectories */
  double best_cost;  /* best_traj.cost_ (-1 when best_id < 0) */
  double xv, yv, thetav; /* best_traj.{xv_,yv_,thetav_} */
} b200lp_result;

/* Optional per-trajectory read-back. Any pointer may be NULL. Arrays are indexed by trajectory id
 * (0..n_traj-1) of robot `robot` of the last plan / plan_batch call. */
typedef struct b200lp_traj_view {
  int32_t* sample_index;  /* [n_traj] index of the velocity sample that produced the trajectory */
  float* vel;             /* [n_traj*3] xv, yv, thetav as the float sample */
  int32_t* num_steps;     /* [n_traj] */
  double* time_delta;     /* [n_traj] Trajectory::time_delta_ */
  double* cost;           /* [n_traj] Trajectory::cost_ after StackedScoringModel::scoreTrajectory */
  double* critic_scores;  /* [n_traj*n_critics] value each critic returned; NaN = not evaluated (after an early-out) */
  int32_t* first_hit_pose;/* [n_traj] first pose index a collision critic rejected at, -1 = none / not evaluated */
} b200lp_traj_view;

/* Optional per-pose read-back of ONE trajectory (recomputed on the device on demand). */
typedef struct b200lp_pose_view {
  double* pose;     /* [num_steps*7] PoseStamped position xyz + orientation xyzw */
  float* pcl_pose;  /* [num_steps*3] Trajectory::getPCLPoint */
  float* cuboid;    /* [num_steps*8*3] Trajectory::getCuboid vertices, order blb,brb,blt,flb,brt,frt,flt,frb */
  float* aabb;      /* [num_steps*6] Trajectory::getCuboidMinMax min xyz, max xyz */
  uint8_t* collide; /* [num_steps] 1 if the pose collides under the first collision critic of the stack
                       (evaluated for every pose, no early exit) */
  int32_t* n_r1;    /* [num_steps] |radiusSearch(pose, 1.0)| */
} b200lp_pose_view;

const char* b200lp_last_error(const b200lp_ctx* ctx); /* ctx may be NULL: last create() error of this thread */
int b200lp_abi_version(void);

int b200lp_create(b200lp_ctx** out, int device, const b200lp_limits* limits, const b200lp_params* params,
                  const float* cuboid_xyz /* 8*3, order blb,brb,blt,flb,brt,frt,flt,frb */,
                  const b200lp_critic* critics, int n_critics, const b200lp_grid_config* grid /* may be NULL */);
void b200lp_destroy(b200lp_ctx* ctx);

/* Upload the aggregated observation cloud and (re)build the voxel grid. `pts` is host memory,
 * n points of `stride_bytes` each (32 = pcl::PointXYZI, 16 = pcl::PointXYZ), x,y,z = first three floats. */
/* Host clouds of 2 MB or more whose points carry padding (stride >= 16) are packed to 12 bytes per point by a few host
 * threads of the ctx into a pinned staging buffer while the chunks already packed are copied: 24 MB instead of 64 MB
 * cross PCIe for 2 M PointXYZI points, and `pts` may be ordinary pageable memory. B200LP_PACK_THREADS in the
 * environment sets the thread count (default: 3/4 of the CPUs the process may run on divided by the visible GPUs, at most 12,
 * none below 4; 0 = copy the caller's buffer as is). */
int b200lp_set_cloud(b200lp_ctx* ctx, const void* pts, size_t n, size_t stride_bytes);
/* How the last b200lp_set_cloud moved the cloud: bytes copied host -> device and the host threads that packed them
 * (0 = the caller's buffer was copied as is). */
int b200lp_last_upload(const b200lp_ctx* ctx, size_t* h2d_bytes, int32_t* pack_threads);
/* Same, `pts` already resident on ctx's device (e.g. produced by a device-side perception stage). */
int b200lp_set_cloud_device(b200lp_ctx* ctx, const void* dev_pts, size_t n, size_t stride_bytes);

/* Prune plan of the single-robot path: n poses of 7 doubles (position xyz, orientation xyzw). */
int b200lp_set_plan(b200lp_ctx* ctx, const double* xyz_qxyzw, size_t n);

/* One local-plan cycle for one robot. */
int b200lp_plan(b200lp_ctx* ctx, const b200lp_query* q, b200lp_result* out);

/* Sample-sharded cycle: this ctx scores only shard `rank` of `count` — a contiguous range of the velocity-sample grid.
 * The cuts are placed at equal shares of the estimated pose count (the grid is ordered by rising linear speed, and fast
 * trajectories have more poses), computed on the device from the query alone, so every rank derives the same cuts;
 * b200lp_traj_count reports the trajectory-id range that was scored. ids and counters in `out` stay global / local
 * resp.: best_id is the GLOBAL trajectory id, n_traj/n_poses/n_collided count the local shard. The caller
 * reduces (best_cost, best_id) across shards: min cost, ties -> largest id (local_planner.cpp:460). */
int b200lp_plan_shard(b200lp_ctx* ctx, const b200lp_query* q, int rank, int count, b200lp_result* out);

/* The same exchange through peer device memory instead of a host-launched collective (one process per GPU on one
 * NVLink / NVSwitch box). Set-up, once: every rank calls b200lp_peer_export, the 64-byte handles are gathered by whatever
 * transport the application has (torch.distributed, MPI, a socket), every rank calls b200lp_peer_attach with all of them.
 * Per cycle: b200lp_plan_shard_exchange scores shard `rank` of `world`, stores its local best into every peer's slot over
 * NVLink from a kernel, waits — in that kernel — for the peers' stores, applies the reference rule (min cost, ties ->
 * largest id) and returns the GLOBAL best_id / best_cost / xv / yv / thetav on every rank (n_samples, n_traj, n_collided,
 * n_poses stay the local shard's). All ranks must call it the same number of times; a peer that does not deliver within
 * about two seconds fails the call with B200LP_E_STATE. */
#define B200LP_PEER_HANDLE_BYTES 64
#define B200LP_MAX_PEERS 16
int b200lp_peer_export(b200lp_ctx* ctx, uint8_t handle[B200LP_PEER_HANDLE_BYTES]);
/* (A ctx attached for the first time may start exchanging at once. Attaching a ctx AGAIN — a new group, or the same one
 * after a failure — clears its slots: the application then puts a barrier between the attach and the first exchange.) */
int b200lp_peer_attach(b200lp_ctx* ctx, int rank, int world, const uint8_t* handles /* world * B200LP_PEER_HANDLE_BYTES */);
int b200lp_plan_shard_exchange(b200lp_ctx* ctx, const b200lp_query* q, b200lp_result* out);
/* One map for all ranks of a peer group (the replicated-map configurations: fleets and sample shards share ONE cloud,
 * model_shared_data.h:74-81 gives every planner the same pcl_perception_). Collective: every rank of the group calls it
 * once per new cloud. The root passes the host cloud (what b200lp_set_cloud takes); it is packed to 12-byte rows, uploaded
 * once, and every upload piece is pushed into every peer's row buffer over NVLink by a kernel as soon as it has landed. The
 * other ranks pass pts = NULL: they wait for the root's header (size, bounds), count the pieces into their histogram as the
 * pieces' flags come up and build their own grid. The host link is crossed once per cloud, not once per rank.
 * b200lp_peer_reserve_cloud(max_points) must be called on every rank BEFORE b200lp_peer_export: it sizes the row buffer
 * that travels with the exported handle. A root that does not deliver within about two seconds fails the peers' call with
 * B200LP_E_STATE (then: b200lp_peer_resync). */
int b200lp_peer_reserve_cloud(b200lp_ctx* ctx, size_t max_points);
int b200lp_set_cloud_shared(b200lp_ctx* ctx, int root, const void* pts /* root only */, size_t n, size_t stride_bytes);
/* After a failed b200lp_plan_shard_exchange (a peer timed out, or the ranks disagreed on the cuts) EVERY rank calls this —
 * with a barrier of the application before and after — and the exchange starts over: sequence numbers, slots and shard
 * cuts return to their initial state. */
int b200lp_peer_resync(b200lp_ctx* ctx);
/* Where sample-sharded cycles cut the sample grid. The estimated work of the grid (expected pose count per linear-speed
 * row plus a fixed term per trajectory) is cut at the shares shares[0] = 0 <= shares[1] <= ... <= shares[count] = 1; rank r
 * scores the samples between cuts r and r + 1. Default (count = 0): equal shares. Every rank of a cycle must use the same
 * cuts: b200lp_plan_shard_exchange checks that and, unless b200lp_set_adaptive_cuts(ctx, 0) was called, moves the cuts
 * of the next cycle with the device times all ranks needed for this one (they travel with the exchanged results, so every
 * rank derives the same new cuts). With b200lp_plan_shard the cuts are the caller's business. */
int b200lp_set_shard_cuts(b200lp_ctx* ctx, const float* shares /* count + 1 */, int count /* 0: equal shares */);
int b200lp_get_shard_cuts(const b200lp_ctx* ctx, float* shares /* B200LP_MAX_PEERS + 1 */, int* count);
int b200lp_set_adaptive_cuts(b200lp_ctx* ctx, int on);

/* Fleet cycle: n_robots independent queries on the shared cloud. Robot i's prune plan is
 * plans[plan_offsets[i] .. plan_offsets[i+1]) (7 doubles per pose). A plan table in page-locked host memory
 * (cudaHostAlloc / cudaHostRegister) is uploaded from where it is; any other buffer is staged first (one extra host copy). */
int b200lp_plan_batch(b200lp_ctx* ctx, const b200lp_query* qs, size_t n_robots, const double* plans,
                      const int64_t* plan_offsets /* n_robots+1 */, b200lp_result* outs);

/* Size of robot's trajectory-id space in the last cycle (= n_traj of an unsharded run) and the id range
 * [t_begin, t_end) the last call actually scored (the whole space unless b200lp_plan_shard was used). */
int b200lp_traj_count(const b200lp_ctx* ctx, size_t robot, int32_t* n_traj_global, int32_t* t_begin, int32_t* t_end);
int b200lp_read_trajectories(b200lp_ctx* ctx, size_t robot, const b200lp_traj_view* view);
int b200lp_read_poses(b200lp_ctx* ctx, size_t robot, int32_t traj_id, const b200lp_pose_view* view);
/* The same for trajectories [t_begin, t_end) in ONE launch — what the generator adapters use to materialise
 * base_trajectory::Trajectory objects for the reference's PoseArray publishers (local_planner.cpp:554,569).
 * Trajectory t's poses occupy rows pose_offsets[t - t_begin] .. pose_offsets[t - t_begin + 1) of every view array;
 * pose_offsets (t_end - t_begin + 1 entries, may be NULL) is written by the call. Fails with B200LP_E_INVALID when
 * the range holds more than capacity_poses poses (pose_offsets is still filled, so the caller can size and retry). */
int b200lp_read_pose_batch(b200lp_ctx* ctx, size_t robot, int32_t t_begin, int32_t t_end, int64_t* pose_offsets,
                           const b200lp_pose_view* view, size_t capacity_poses);

/* ---- the steps either side of the path -------------------------------------------------------- */
/* What Local_Planner::prunePlan decided (LP/local_planner/src/local_planner.cpp:374-445). */
typedef struct b200lp_prune_info {
  int32_t status;        /* 0 = pruned; 1 = global plan has < 3 poses: the reference returns before touching the prune
                            plan, the previous one stays in effect; 2 = robot farther than 1 m from the plan: the reference
                            has already cleared the prune plan (:379-380) when it returns, so it is now EMPTY */
  int32_t nearest_index; /* global-plan index nearest to the robot (-1 when status == 1) */
  int32_t n_prune;       /* poses of the prune plan now in effect (the nearest pose appears twice, as upstream) */
  int32_t n_backward;    /* how many of them are tagged backward (pcl_prune_plan_ intensity -1) */
} b200lp_prune_info;
/* What perception_3d::PathBlockedStrategy::selfMark decided (dddmr_perception_3d/plugins/path_blocked_strategy.cpp:56-100). */
typedef struct b200lp_blocked {
  int32_t n_blocked; /* forward prune-plan points with a cloud point inside check_radius (strict float d^2 < r^2) */
  int32_t n_checked; /* points with intensity >= 0 */
  int32_t n_total;   /* |pcl_prune_plan_| */
  int32_t opinion;   /* 0 = perception_3d::PASS, 1 = PATH_BLOCKED_WAIT (ratio > 0) */
  double ratio;      /* prune_plan_blocked_ratio_ = (float)n_blocked / (float)n_total * 100.0 */
} b200lp_blocked;

/* Local_Planner::setPlan (local_planner.cpp:322-343): keep the global plan on the device. n >= 3 as upstream
 * (smaller plans are rejected with B200LP_E_INVALID and the previous plan stays). */
int b200lp_set_global_plan(b200lp_ctx* ctx, const double* xyz_qxyzw, size_t n);
/* Local_Planner::prunePlan on the device: nearest global-plan pose to the robot (float L2, lowest index on exact
 * ties), walk backward / forward until the distances are used up. On status 0 and 2 the result BECOMES the prune plan
 * of the following b200lp_plan / b200lp_plan_shard calls without leaving the device (it replaces b200lp_set_plan). */
int b200lp_prune_plan(b200lp_ctx* ctx, const double robot_xyz[3], double forward_distance, double backward_distance,
                      b200lp_prune_info* out);
/* Read the device-side prune plan back: poses7 = prune_plan_.poses (backward part reversed, then forward part),
 * pcl_xyzi = pcl_prune_plan_ in ITS order (backward part as walked, then forward part; w = intensity tag).
 * Either pointer may be NULL; capacity in poses. */
int b200lp_read_prune_plan(b200lp_ctx* ctx, double* poses7, float* pcl_xyzi, size_t capacity);
/* PathBlockedStrategy::selfMark of the device-side prune plan against the current cloud (the same voxel grid the
 * critics query). Requires a successful b200lp_prune_plan. */
int b200lp_path_blocked(b200lp_ctx* ctx, double check_radius, b200lp_blocked* out);

/* ---- the observation producer in front of the path --------------------------------------- */
#define B200LP_MAX_SENSORS 8
/* Parameters MultiLayerSpinningLidar reads for cbSensor (multilayer_spinning_lidar.cpp:64-131). */
typedef struct b200lp_sensor_params {
  double perception_window_size; /* pass-through limits of x and y in base_link: [-w, w] (:242, :246) */
  double marking_height;         /* pass-through limits of z in base_link: [0, marking_height] (:249) */
  float leaf_size;               /* pcl::VoxelGrid leaf, 0.1f upstream (:254); 0 selects 0.1f */
  int32_t is_local_planner;      /* non-zero: the observation is moved to the global frame (:264-268) */
} b200lp_sensor_params;
typedef struct b200lp_observation_info {
  int64_t n_scan;   /* points handed in */
  int64_t n_window; /* points left by the three pass-through filters (non-finite points are removed there too) */
  int64_t n_points; /* voxels = points of the observation */
  float ms_device;  /* CUDA-event time of the device work of this call (upload included) */
  int32_t n_launches;
  float ms_upload;  /* the part of ms_device the host -> device copy of the scan took */
  int32_t reserved_;
} b200lp_observation_info;
/* MultiLayerSpinningLidar::cbSensor on one scan (after pcl::fromROSMsg / stitching): `scan` is host memory, n points of
 * stride_bytes each (16 = pcl::PointXYZ), x,y,z = first three floats, in the sensor frame. base_from_sensor =
 * trans_b2s_, global_from_base = trans_gbl2b_ (translation xyz, rotation xyzw: geometry_msgs Transform order).
 * The result is sensor `sensor`'s current observation (Sensor::sensor_current_observation_), kept on the device.
 * Voxels leave in ascending voxel-index order like pcl::VoxelGrid; inside a voxel the points are added in scan order
 * (PCL adds them in the order an unstable sort of the voxel indices leaves them).
 * B200LP_E_INVALID when the window holds more than 2^27 voxels of that leaf or the scan more than 2^26 points. */
int b200lp_sensor_observation(b200lp_ctx* ctx, int sensor, const void* scan, size_t n, size_t stride_bytes,
                              const double base_from_sensor[7], const double global_from_base[7],
                              const b200lp_sensor_params* params, b200lp_observation_info* info /* may be NULL */);
/* Read sensor's current observation back (for the reference's current_observation publisher, :273-278, and for parity
 * tests): stride_bytes 16 = pcl::PointXYZ (x, y, z, 1), 32 = pcl::PointXYZI (+ intensity 0). Returns B200LP_E_INVALID
 * when it holds more than capacity_points. */
int b200lp_read_observation(b200lp_ctx* ctx, int sensor, void* out, size_t capacity_points, size_t stride_bytes,
                            size_t* n_points);
/* StackedPerception::aggregateObservations: concatenate the current observations of `sensors` (plugin order) on the
 * device and make the result the critics' cloud (what b200lp_set_cloud does with a host cloud). */
int b200lp_aggregate_observations(b200lp_ctx* ctx, const int32_t* sensors, int n_sensors, size_t* n_total /* may be NULL */);

/* Roofline accounting helper: sum over all scored poses of the last plan call of
 * |{cloud points with float d^2 < 1.0 to the pose}| (the reference's radiusSearch candidate set). */
int b200lp_count_radius(b200lp_ctx* ctx, int64_t* sum_n_r1, int64_t* n_poses);

/* Work counters of the sweep (roofline accounting): out[0] = (candidate point, pose) pre-tests, out[1] = 32-candidate
 * rounds, out[2] = candidates that went on to the exact test, out[3] = pose groups swept, summed over every plan call since
 * the last reset. Only the counting build (libb200lp_count.so, compiled with -DB200LP_COUNT=1, never timed) keeps them;
 * the product library answers B200LP_E_STATE. */
int b200lp_work_counters(b200lp_ctx* ctx, uint64_t out[4], int reset);

/* Device-timeline instrumentation, CUDA events on ctx's stream: ms_upload / ms_grid_build of the last
 * b200lp_set_cloud* (which returns once the caller's buffer is consumed, with the grid kernels still in
 * flight — asking here waits for them), ms_plan_kernels / ms_readback of the last plan call. Any pointer may be NULL. */
int b200lp_last_timing(const b200lp_ctx* ctx, float* ms_upload, float* ms_grid_build, float* ms_plan_kernels,
                       float* ms_readback);
/* Per-kernel split of ms_plan_kernels for the last plan call: prep_kernel (velocity sampling, trajectory list,
 * forward simulation), plan_kernel (pose geometry + obstacle query + critics; the one the roofline is reported
 * for) and argmin_kernel (best trajectory per robot). */
int b200lp_last_kernel_ms(const b200lp_ctx* ctx, float* ms_prep_kernel, float* ms_plan_kernel, float* ms_argmin_kernel);
/* The same with every kernel of the cycle: ms[0] = prep_kernel, ms[1] = cull_kernel (float pre-cull of every pose + the work
 * lists plan_kernel drains), ms[2] = plan_kernel, ms[3] = argmin_kernel (fleets). Writes min(n, 4) values. */
int b200lp_last_kernel_times(const b200lp_ctx* ctx, float* ms, int n);
/* Device time of the last single-robot cycle, nanoseconds of the GPU's global timer from the first CTA of prep_kernel to
 * the last CTA of plan_kernel (the one that writes the result into host memory); after b200lp_plan_shard_exchange
 * peer_ns[r] (B200LP_MAX_PEERS entries, may be NULL) holds the same figure of every rank r < world. */
int b200lp_last_cycle_ns(const b200lp_ctx* ctx, uint32_t* cycle_ns, uint32_t* peer_ns);
/* Number of kernels this library launched on ctx's stream since creation. */
int64_t b200lp_launch_count(const b200lp_ctx* ctx);
/* Grid geometry of the current cloud (for tests / docs). dims = nx,ny,nz; origin xyz; cell xy,z. */
int b200lp_grid_info(const b200lp_ctx* ctx, int32_t dims[3], float origin[3], float cell[2], int64_t* n_points_kept);
/* The CUDA stream (cudaStream_t as void*) all work of ctx is enqueued on. */
void* b200lp_stream(const b200lp_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* B200LP_H_ */
