// lpref_driver.cpp — runs the REFERENCE's own rollout-and-score sources (compiled where they lie under
// /root/reference against the third-party stand-ins of oracle/ref_shims/, see oracle/Makefile) through one
// local-plan cycle. TEST INFRASTRUCTURE ONLY: it exists to pin oracle/lp_oracle.cpp — the restatement — to the
// reference's real control flow, types and call order. Nothing under dddmr_navigation_b200/ may load it.
//
// The cycle is Local_Planner::computeVelocityCommand (local_planner.cpp:528-587) + getBestTrajectory (:447-480),
// driven through the real StackedGenerator / StackedScoringModel; local_planner.cpp itself (tf, publishers,
// perception_3d_ros) is not compiled, those ~40 lines are restated here.
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

// Only the plugin BASE headers are included: several of the reference's plugin headers have no include guards, so the
// concrete classes are created through the PLUGINLIB_EXPORT_CLASS registrations of their own .cpp files (ref_shims/pluginlib).
#include <mpc_critics/stacked_scoring_model.h>
#include <trajectory_generators/stacked_generator.h>

#include "../include/b200lp.h"

struct lpref_ctx {
  std::string err;
  std::shared_ptr<rclcpp::Node> tg_node, mc_node;
  std::shared_ptr<trajectory_generators::StackedGenerator> gen;
  std::shared_ptr<mpc_critics::StackedScoringModel> stack;
  std::vector<std::shared_ptr<mpc_critics::ScoringModel>> models;  // the same objects the stack holds, in order
  std::vector<int> kinds;
  pcl::PointCloud<pcl::PointXYZI>::Ptr cloud;
  nav_msgs::msg::Path plan;
  int theory = 0;
  std::vector<base_trajectory::Trajectory> trajs;  // Local_Planner::trajectories_ after scoring
  std::vector<std::vector<double>> scores;         // per trajectory: what each critic returned (NaN = not asked)
};

static const char* kGen = "gen";

extern "C" {

const char* lpref_last_error(const lpref_ctx* c) { return c ? c->err.c_str() : ""; }

int lpref_create(lpref_ctx** out, const b200lp_limits* L, const b200lp_params* P, const float* cuboid,
                 const b200lp_critic* critics, int n_critics) {
  if (!out || !L || !P || !cuboid) return B200LP_E_INVALID;
  auto c = std::make_unique<lpref_ctx>();
  try {
    c->theory = P->theory;
    c->tg_node = std::make_shared<rclcpp::Node>("trajectory_generators");
    c->mc_node = std::make_shared<rclcpp::Node>("mpc_critics");
    auto set = [&](const char* key, double v) { c->tg_node->set_parameter_override(std::string(kGen) + "." + key, v); };
    set("min_vel_x", L->min_vel_x); set("max_vel_x", L->max_vel_x);
    set("min_vel_y", L->min_vel_y); set("max_vel_y", L->max_vel_y);
    set("min_vel_trans", L->min_vel_trans); set("max_vel_trans", L->max_vel_trans);
    set("min_vel_theta", L->min_vel_theta); set("max_vel_theta", L->max_vel_theta);
    set("acc_lim_x", L->acc_lim_x); set("acc_lim_y", L->acc_lim_y); set("acc_lim_theta", L->acc_lim_theta);
    set("deceleration_ratio", L->deceleration_ratio);
    set("max_motor_shaft_rpm", L->max_motor_shaft_rpm); set("wheel_diameter", L->wheel_diameter);
    set("gear_ratio", L->gear_ratio); set("robot_radius", L->robot_radius); set("rotation_speed", L->rotation_speed);
    set("controller_frequency", P->controller_frequency); set("sim_time", P->sim_time);
    set("linear_x_sample", P->linear_x_sample); set("linear_y_sample", P->linear_y_sample);
    set("angular_z_sample", P->angular_z_sample);
    set("sim_granularity", P->sim_granularity); set("angular_sim_granularity", P->angular_sim_granularity);
    c->tg_node->set_parameter_override(std::string(kGen) + ".use_motor_constraint", L->use_motor_constraint != 0);
    static const char* kOrder[8] = {"blb", "brb", "blt", "flb", "brt", "frt", "flt", "frb"};
    for (int k = 0; k < 8; ++k)
      c->tg_node->set_parameter_override(std::string(kGen) + ".cuboid." + kOrder[k],
                                         std::vector<double>{(double)cuboid[k * 3], (double)cuboid[k * 3 + 1], (double)cuboid[k * 3 + 2]});
    auto logger = std::make_shared<rclcpp::node_interfaces::NodeLoggingInterface>();
    auto tfbuf = std::make_shared<tf2_ros::Buffer>();
    c->gen = std::make_shared<trajectory_generators::StackedGenerator>(logger, tfbuf);
    static const char* kTheory[3] = {"trajectory_generators::DDSimpleTrajectoryGeneratorTheory",
                                     "trajectory_generators::OmniSimpleTrajectoryGeneratorTheory",
                                     "trajectory_generators::DDRotateInplaceTheory"};
    if (P->theory < 0 || P->theory > 2) return B200LP_E_INVALID;
    auto th = lpref::create<trajectory_generators::TrajectoryGeneratorTheory>(kTheory[P->theory]);
    c->gen->addPlugin(kGen, th);                 // trajectory_generators_ros.cpp:78-80
    th->initialize(kGen, c->tg_node);

    c->stack = std::make_shared<mpc_critics::StackedScoringModel>(logger, tfbuf);
    for (int k = 0; k < n_critics; ++k) {
      const std::string name = "critic" + std::to_string(k);
      c->mc_node->set_parameter_override(name + ".weight", critics[k].weight);
      c->mc_node->set_parameter_override(name + ".translation_weight", critics[k].translation_weight);
      c->mc_node->set_parameter_override(name + ".orientation_weight", critics[k].orientation_weight);
      static const char* kCritic[7] = {"mpc_critics::CollisionModel", "mpc_critics::CollisionMinMaxModel",
                                       "mpc_critics::StickPathModel", "mpc_critics::PurePursuitModel",
                                       "mpc_critics::TowardGlobalPlanModel", "mpc_critics::ShortestAngleModel",
                                       "mpc_critics::TwirlingModel"};
      if (critics[k].kind < 0 || critics[k].kind > 6) return B200LP_E_INVALID;
      auto m = lpref::create<mpc_critics::ScoringModel>(kCritic[critics[k].kind]);
      c->kinds.push_back(critics[k].kind);
      c->stack->addPluginByTraj(kGen, m);        // mpc_critics_ros.cpp:75-79
      m->initialize(name, c->mc_node);
      c->models.push_back(m);
    }
    c->cloud.reset(new pcl::PointCloud<pcl::PointXYZI>);
  } catch (const std::exception& e) {
    return B200LP_E_INVALID;
  }
  *out = c.release();
  return B200LP_OK;
}

void lpref_destroy(lpref_ctx* c) { delete c; }

int lpref_set_cloud(lpref_ctx* c, const void* pts, size_t n, size_t stride) {
  if (!c || (n && !pts) || stride < 12) return B200LP_E_INVALID;
  // aggregateObservations builds a fresh cloud object every cycle (stacked_perception.cpp:128-140)
  c->cloud.reset(new pcl::PointCloud<pcl::PointXYZI>);
  c->cloud->points.resize(n);
  for (size_t i = 0; i < n; ++i) {
    const float* p = (const float*)((const char*)pts + i * stride);
    pcl::PointXYZI q;
    q.x = p[0]; q.y = p[1]; q.z = p[2];
    c->cloud->points[i] = q;
  }
  return B200LP_OK;
}

int lpref_set_plan(lpref_ctx* c, const double* p, size_t n) {
  if (!c || (n && !p)) return B200LP_E_INVALID;
  c->plan.poses.clear();
  for (size_t i = 0; i < n; ++i) {
    geometry_msgs::msg::PoseStamped ps;
    ps.pose.position.x = p[i * 7]; ps.pose.position.y = p[i * 7 + 1]; ps.pose.position.z = p[i * 7 + 2];
    ps.pose.orientation.x = p[i * 7 + 3]; ps.pose.orientation.y = p[i * 7 + 4]; ps.pose.orientation.z = p[i * 7 + 5];
    ps.pose.orientation.w = p[i * 7 + 6];
    c->plan.poses.push_back(ps);
  }
  return B200LP_OK;
}

int lpref_plan(lpref_ctx* c, const b200lp_query* q, b200lp_result* out) {
  if (!c || !q || !out) return B200LP_E_INVALID;
  try {
    geometry_msgs::msg::TransformStamped pose;
    pose.header.frame_id = "map";
    pose.child_frame_id = "base_link";
    pose.transform.translation.x = q->pose[0]; pose.transform.translation.y = q->pose[1]; pose.transform.translation.z = q->pose[2];
    pose.transform.rotation.x = q->pose[3]; pose.transform.rotation.y = q->pose[4]; pose.transform.rotation.z = q->pose[5];
    pose.transform.rotation.w = q->pose[6];
    nav_msgs::msg::Odometry odom;
    odom.twist.twist.linear.x = q->twist[0]; odom.twist.twist.linear.y = q->twist[1]; odom.twist.twist.angular.z = q->twist[2];

    // local_planner.cpp:528-535
    auto tg = c->gen->getSharedDataPtr();
    tg->robot_pose_ = pose;
    tg->robot_state_ = odom;
    tg->prune_plan_ = c->plan;
    tg->current_allowed_max_linear_speed_ = q->max_speed_override;
    c->gen->initializeTheories_wi_Shared_data();
    // :549-557
    c->trajs.clear();
    while (c->gen->hasMoreTrajectories(kGen)) {
      base_trajectory::Trajectory a_traj;
      if (c->gen->nextTrajectory(kGen, a_traj)) c->trajs.push_back(a_traj);
    }
    // :577-587
    auto mc = c->stack->getSharedDataPtr();
    mc->robot_pose_ = pose;
    mc->robot_state_ = odom;
    mc->pcl_perception_ = c->cloud;
    mc->prune_plan_ = c->plan;
    mc->heading_deviation_ = q->heading_deviation;
    mc->updateData();
    // getBestTrajectory :447-480
    const double nan = std::numeric_limits<double>::quiet_NaN();
    c->scores.assign(c->trajs.size(), std::vector<double>(c->models.size(), nan));
    double minimum_cost = 9999999;
    int best = -1, n_collided = 0;
    long long n_poses = 0;
    base_trajectory::Trajectory best_traj;
    best_traj.cost_ = -1;
    for (size_t i = 0; i < c->trajs.size(); ++i) {
      base_trajectory::Trajectory& traj = c->trajs[i];
      n_poses += traj.getPointsSize();
      // per-critic values for the parity report: the same calls StackedScoringModel::scoreTrajectory makes, in its order
      for (size_t k = 0; k < c->models.size(); ++k) {
        const double v = c->models[k]->scoreTrajectory(traj);
        c->scores[i][k] = v;
        if (v < 0) {
          if (c->kinds[k] == B200LP_CRITIC_COLLISION || c->kinds[k] == B200LP_CRITIC_COLLISION_MIN_MAX) ++n_collided;
          break;
        }
      }
      c->stack->scoreTrajectory(kGen, traj);  // the real accumulation into traj.cost_
      if (traj.cost_ >= 0 && traj.cost_ <= minimum_cost) {
        best_traj = traj;
        minimum_cost = traj.cost_;
        best = (int)i;
      }
    }
    out->best_id = best;
    out->n_samples = -1;  // sample_params_ is private to the theory
    out->n_traj = (int32_t)c->trajs.size();
    out->n_collided = n_collided;
    out->n_poses = n_poses;
    out->best_cost = best_traj.cost_;
    out->xv = best >= 0 ? best_traj.xv_ : 0.0;
    out->yv = best >= 0 ? best_traj.yv_ : 0.0;
    out->thetav = best >= 0 ? best_traj.thetav_ : 0.0;
  } catch (const std::exception& e) {
    c->err = e.what();
    return B200LP_E_INVALID;
  }
  return B200LP_OK;
}

int lpref_read_trajectories(lpref_ctx* c, const b200lp_traj_view* v) {
  if (!c || !v) return B200LP_E_INVALID;
  const size_t nc = c->models.size();
  for (size_t i = 0; i < c->trajs.size(); ++i) {
    const base_trajectory::Trajectory& t = c->trajs[i];
    if (v->vel) { v->vel[i * 3] = (float)t.xv_; v->vel[i * 3 + 1] = (float)t.yv_; v->vel[i * 3 + 2] = (float)t.thetav_; }
    if (v->num_steps) v->num_steps[i] = (int32_t)t.getPointsSize();
    if (v->time_delta) v->time_delta[i] = t.time_delta_;
    if (v->cost) v->cost[i] = t.cost_;
    if (v->critic_scores)
      for (size_t k = 0; k < nc; ++k) v->critic_scores[i * nc + k] = c->scores[i][k];
    if (v->sample_index) v->sample_index[i] = -1;
    if (v->first_hit_pose) v->first_hit_pose[i] = -1;
  }
  return B200LP_OK;
}

int lpref_read_poses(lpref_ctx* c, int32_t id, const b200lp_pose_view* v) {
  if (!c || !v || id < 0 || (size_t)id >= c->trajs.size()) return B200LP_E_INVALID;
  const base_trajectory::Trajectory& t = c->trajs[(size_t)id];
  for (unsigned int i = 0; i < t.getPointsSize(); ++i) {
    if (v->pose) {
      const auto ps = t.getPoint(i);
      const double row[7] = {ps.pose.position.x, ps.pose.position.y, ps.pose.position.z, ps.pose.orientation.x,
                             ps.pose.orientation.y, ps.pose.orientation.z, ps.pose.orientation.w};
      memcpy(v->pose + i * 7, row, sizeof(row));
    }
    if (v->pcl_pose) { const auto p = t.getPCLPoint(i); v->pcl_pose[i * 3] = p.x; v->pcl_pose[i * 3 + 1] = p.y; v->pcl_pose[i * 3 + 2] = p.z; }
    if (v->cuboid) {
      const auto cu = t.getCuboid(i);
      for (size_t k = 0; k < 8 && k < cu.points.size(); ++k) {
        v->cuboid[i * 24 + k * 3] = cu.points[k].x; v->cuboid[i * 24 + k * 3 + 1] = cu.points[k].y; v->cuboid[i * 24 + k * 3 + 2] = cu.points[k].z;
      }
    }
    if (v->aabb) {
      const auto mm = t.getCuboidMinMax(i);
      const float row[6] = {mm.first.x, mm.first.y, mm.first.z, mm.second.x, mm.second.y, mm.second.z};
      memcpy(v->aabb + i * 6, row, sizeof(row));
    }
  }
  return B200LP_OK;
}

}  // extern "C"
