// plugin_cycle.cpp — drives ONE local-plan cycle through the reference's plugin interfaces (host layer mirrors) the way
// Local_Planner::computeVelocityCommand does (local_planner.cpp:482-621), and dumps what the caller sees so pytest can
// compare it with the CPU oracle. Usage:
//   plugin_cycle <params.yaml> <scenario.bin> <out_prefix> <generator_name> <early|late> [prune <fwd> <bwd> <check_radius>]
//   plugin_cycle <params.yaml> <scenario.bin> <out_prefix> <generator_name> <early|late> scan <scans.bin> <window> <marking_height>
// With `scan`, the scenario's cloud is ignored: the lidar scans of scans.bin (int64 n_sensors; per sensor int64 n,
// float64 base_from_sensor[7], float32 points[n][4] in the sensor frame) go through the MultiLayerSpinningLidar::cbSensor
// mirrors and StackedPerception::aggregateObservations before every cycle; .obs.f32 (n x 8) is the
// host copy of the aggregate.
// With `prune`, the scenario's plan is the GLOBAL plan: Local_Planner::setPlan + prunePlan run on the device and a
// PathBlockedStrategy gives its opinion after scoring; .prune.f64 / .prunepcl.f32 are dumped too.
// With `rotate <max_passes> <tolerance>`, the OTHER caller of the path runs instead: recovery_behaviors::RotateInPlaceBehavior's
// control loop (rotate_inplace_behavior.cpp:137-305) with a robot that turns by cmd.angular.z / frequency per pass;
// .rotate.f64 holds one row per pass: yaw, cmd.angular.z, cmd.linear.x, finished, result, best_id, best_cost, got_180,
// dist_left, size of the critics' cloud after the pass (the behaviour resets it, :254-256).
// scenario.bin: int64 n_points, int64 n_plan, float32 points[n][8] (PointXYZI), float64 plan[m][7], float64 pose[7],
//               float64 twist[3], float64 max_speed, float64 heading_deviation
// outputs:      <out_prefix>.summary.txt (key=value), .traj.f64 (n x 6: cost,xv,yv,thetav,time_delta,n_points),
//               .pose.f64 (P x 7), .pcl.f32 (P x 3), .cuboid.f32 (P x 24), .aabb.f32 (P x 6)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "b200lp/session.hpp"
#include "local_planner/local_planner.h"
#include "recovery_behaviors/rotate_inplace_behavior.h"

static std::string slurp(const char* path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) {
    std::fprintf(stderr, "cannot open %s\n", path);
    std::exit(2);
  }
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}
template <class T>
static void dump(const std::string& path, const std::vector<T>& v) {
  std::ofstream f(path, std::ios::binary);
  f.write((const char*)v.data(), (std::streamsize)(v.size() * sizeof(T)));
}

int main(int argc, char** argv) {
  if (argc < 6) {
    std::fprintf(stderr, "usage: %s params.yaml scenario.bin out_prefix generator early|late\n", argv[0]);
    return 2;
  }
  const std::string yaml = slurp(argv[1]), sc = slurp(argv[2]), out = argv[3], gen = argv[4];
  const bool early = std::string(argv[5]) == "early";
  try {
    // ---- node + plugin bring-up, as p2p_move_base_node.cpp / local_planner_node.cpp do ----
    auto tg = std::make_shared<trajectory_generators::Trajectory_Generators_ROS>("trajectory_generators");
    auto mc = std::make_shared<mpc_critics::MPC_Critics_ROS>("mpc_critics");
    tg->load_parameters_yaml(yaml);
    mc->load_parameters_yaml(yaml);
    tg->initial();
    mc->initial();
    auto perception = std::make_shared<perception_3d::SharedData>();
    local_planner::Local_Planner lp("local_planner");
    lp.initial(perception, mc, tg);
    lp.setEarlyObservationHandOver(early);

    // ---- scenario ----
    const char* p = sc.data();
    int64_t n_points, n_plan;
    std::memcpy(&n_points, p, 8); p += 8;
    std::memcpy(&n_plan, p, 8); p += 8;
    perception->aggregate_observation_.reset(new pcl::PointCloud<pcl::PointXYZI>);
    perception->aggregate_observation_->points.resize((size_t)n_points);
    std::memcpy((void*)perception->aggregate_observation_->points.data(), p, (size_t)n_points * 32); p += n_points * 32;
    nav_msgs::msg::Path plan;
    for (int64_t i = 0; i < n_plan; ++i) {
      double r[7];
      std::memcpy(r, p, 56); p += 56;
      geometry_msgs::msg::PoseStamped ps;
      ps.pose.position.x = r[0]; ps.pose.position.y = r[1]; ps.pose.position.z = r[2];
      ps.pose.orientation.x = r[3]; ps.pose.orientation.y = r[4]; ps.pose.orientation.z = r[5]; ps.pose.orientation.w = r[6];
      plan.poses.push_back(ps);
    }
    double tail[12];
    std::memcpy(tail, p, sizeof(tail));
    geometry_msgs::msg::TransformStamped pose;
    pose.header.frame_id = "map";
    pose.child_frame_id = "base_link";
    pose.transform.translation.x = tail[0]; pose.transform.translation.y = tail[1]; pose.transform.translation.z = tail[2];
    pose.transform.rotation.x = tail[3]; pose.transform.rotation.y = tail[4]; pose.transform.rotation.z = tail[5]; pose.transform.rotation.w = tail[6];
    nav_msgs::msg::Odometry odom;
    odom.twist.twist.linear.x = tail[7]; odom.twist.twist.linear.y = tail[8]; odom.twist.twist.angular.z = tail[9];
    perception->current_allowed_max_linear_speed_ = tail[10];
    mc->getSharedDataPtr()->heading_deviation_ = tail[11];
    lp.setGlobalPose(pose);
    lp.cbOdom(odom);
    const bool scan_mode = argc >= 10 && std::string(argv[6]) == "scan";
    std::vector<std::shared_ptr<perception_3d::MultiLayerSpinningLidar>> lidars;
    std::vector<pcl::PointCloud<pcl::PointXYZ>> scans;
    std::vector<geometry_msgs::msg::TransformStamped> mounts;
    perception_3d::StackedPerception stacked(perception);
    auto observe = [&] {  // what the sensor callbacks + the perception loop do between two planner cycles
      for (size_t k = 0; k < lidars.size(); ++k) lidars[k]->cbSensor(scans[k], mounts[k], pose);
      stacked.aggregateObservations();
    };
    if (scan_mode) {
      const std::string sb = slurp(argv[7]);
      const char* q = sb.data();
      int64_t n_sensors;
      std::memcpy(&n_sensors, q, 8); q += 8;
      for (int64_t k = 0; k < n_sensors; ++k) {
        int64_t n;
        double m[7];
        std::memcpy(&n, q, 8); q += 8;
        std::memcpy(m, q, 56); q += 56;
        pcl::PointCloud<pcl::PointXYZ> sc_k;
        sc_k.points.resize((size_t)n);
        std::memcpy((void*)sc_k.points.data(), q, (size_t)n * 16); q += n * 16;
        sc_k.is_dense = false;
        scans.push_back(sc_k);
        geometry_msgs::msg::TransformStamped t;
        t.transform.translation.x = m[0]; t.transform.translation.y = m[1]; t.transform.translation.z = m[2];
        t.transform.rotation.x = m[3]; t.transform.rotation.y = m[4]; t.transform.rotation.z = m[5]; t.transform.rotation.w = m[6];
        mounts.push_back(t);
        lidars.push_back(std::make_shared<perception_3d::MultiLayerSpinningLidar>("lidar" + std::to_string(k), gen, (int)k,
                                                                                 std::atof(argv[8]), std::atof(argv[9]), true));
        stacked.addPluginToVector(lidars.back());
      }
      observe();
    }
    if (argc >= 9 && std::string(argv[6]) == "rotate") {
      recovery_behaviors::RotateInPlaceBehavior beh("rotate_inplace");
      beh.initial(perception, mc, tg, gen, std::atof(argv[8]), 10.0);
      const double yaw0 = recovery_behaviors::yaw_of(tail[3], tail[4], tail[5], tail[6]);
      double yaw = yaw0, now = 0.0;
      beh.begin(pose, now);
      std::vector<double> rows;
      for (int it = 0; it < std::atoi(argv[7]); ++it) {
        geometry_msgs::msg::TransformStamped t = pose;  // the robot turns on the spot, about z
        t.transform.rotation.x = 0.0; t.transform.rotation.y = 0.0;
        t.transform.rotation.z = std::sin(yaw / 2); t.transform.rotation.w = std::cos(yaw / 2);
        // aggregateObservations() builds a fresh cloud object every pass (stacked_perception.cpp:128-140)
        perception->aggregate_observation_.reset(new pcl::PointCloud<pcl::PointXYZI>(*perception->aggregate_observation_));
        const auto st = beh.step(t, odom, now);
        rows.insert(rows.end(), {yaw, st.cmd_angular_z, st.cmd_linear_x, st.finished ? 1.0 : 0.0, (double)st.result, (double)st.best_id,
                                 st.best_cost, st.got_180 ? 1.0 : 0.0, st.dist_left,
                                 (double)mc->getSharedDataPtr()->pcl_perception_->points.size()});
        if (st.finished) break;
        yaw += st.cmd_angular_z / beh.frequency();
        now += 1.0 / beh.frequency();
      }
      dump(out + ".rotate.f64", rows);
      b200lp::Session::resetAll();
      return 0;
    }
    const bool prune = argc >= 10 && std::string(argv[6]) == "prune";
    std::shared_ptr<perception_3d::PathBlockedStrategy> blocked;
    if (prune) {
      lp.setPlan(plan.poses, gen);
      lp.prunePlan(std::atof(argv[7]), std::atof(argv[8]), gen);
      blocked = std::make_shared<perception_3d::PathBlockedStrategy>(std::atof(argv[9]));
      lp.setPathBlockedStrategy(blocked);
    } else {
      lp.setPrunePlan(plan);
    }

    // ---- two cycles: the second one must reproduce the first (cached state is per cycle) ----
    base_trajectory::Trajectory best;
    dddmr_sys_core::PlannerState state = lp.computeVelocityCommand(gen, best);
    auto session = b200lp::Session::forGenerator(gen);
    const int launches_first = session->launchesThisCycle();
    // aggregateObservations() builds a fresh cloud object every cycle (stacked_perception.cpp:128-140)
    if (scan_mode) observe();
    else perception->aggregate_observation_.reset(new pcl::PointCloud<pcl::PointXYZI>(*perception->aggregate_observation_));
    base_trajectory::Trajectory best2;
    const dddmr_sys_core::PlannerState state2 = lp.computeVelocityCommand(gen, best2);
    const int launches_second = session->launchesThisCycle();

    std::vector<double> traj, pose7;
    std::vector<float> pcl3, cub, aabb;
    for (const auto& t : *lp.trajectories_) {
      traj.insert(traj.end(), {t.cost_, t.xv_, t.yv_, t.thetav_, t.time_delta_, (double)t.getPointsSize()});
      for (unsigned int i = 0; i < t.getPointsSize(); ++i) {
        const auto ps = t.getPoint(i);
        pose7.insert(pose7.end(), {ps.pose.position.x, ps.pose.position.y, ps.pose.position.z, ps.pose.orientation.x,
                                   ps.pose.orientation.y, ps.pose.orientation.z, ps.pose.orientation.w});
        const auto pp = t.getPCLPoint(i);
        pcl3.insert(pcl3.end(), {pp.x, pp.y, pp.z});
        const auto c = t.getCuboid(i);
        for (size_t k = 0; k < c.size(); ++k) cub.insert(cub.end(), {c[k].x, c[k].y, c[k].z});
        const auto mm = t.getCuboidMinMax(i);
        aabb.insert(aabb.end(), {mm.first.x, mm.first.y, mm.first.z, mm.second.x, mm.second.y, mm.second.z});
      }
    }
    dump(out + ".traj.f64", traj);
    dump(out + ".pose.f64", pose7);
    dump(out + ".pcl.f32", pcl3);
    dump(out + ".cuboid.f32", cub);
    dump(out + ".aabb.f32", aabb);
    if (scan_mode) {
      std::vector<float> ob;
      for (const auto& pt : perception->aggregate_observation_->points)
        ob.insert(ob.end(), {pt.x, pt.y, pt.z, pt.pad_, pt.intensity, pt.pad2_[0], pt.pad2_[1], pt.pad2_[2]});
      dump(out + ".obs.f32", ob);
    }
    if (prune) {
      std::vector<double> pp;
      std::vector<float> pc;
      for (const auto& ps : lp.getPrunePlan().poses)
        pp.insert(pp.end(), {ps.pose.position.x, ps.pose.position.y, ps.pose.position.z, ps.pose.orientation.x,
                             ps.pose.orientation.y, ps.pose.orientation.z, ps.pose.orientation.w});
      for (const auto& pt : lp.getPCLPrunePlan().points) pc.insert(pc.end(), {pt.x, pt.y, pt.z, pt.intensity});
      dump(out + ".prune.f64", pp);
      dump(out + ".prunepcl.f32", pc);
    }
    const b200lp_result& r = session->result();
    std::ofstream s(out + ".summary.txt");
    s.precision(17);
    s << "state=" << (int)state << "\nstate2=" << (int)state2 << "\nbest_id=" << best.id_ << "\nbest_id2=" << best2.id_
      << "\nbest_cost=" << best.cost_ << "\nbest_xv=" << best.xv_ << "\nbest_thetav=" << best.thetav_
      << "\nn_traj=" << lp.trajectories_->size() << "\nlaunches_first=" << launches_first
      << "\nlaunches_second=" << launches_second << "\ndevice_best_id=" << r.best_id << "\ndevice_best_cost=" << r.best_cost
      << "\ndevice_n_samples=" << r.n_samples << "\ndevice_n_poses=" << r.n_poses << "\n";
    if (scan_mode) {
      s << "n_observation=" << perception->aggregate_observation_->points.size() << "\n";
      for (size_t k = 0; k < lidars.size(); ++k)
        s << "sensor" << k << "_n_points=" << lidars[k]->lastInfo().n_points << "\nsensor" << k << "_n_window=" << lidars[k]->lastInfo().n_window << "\n";
    }
    if (prune) s << "blocked_ratio=" << blocked->getBlockedRatio() << "\nblocked_opinion=" << (int)blocked->getOpinion() << "\n";
    b200lp::Session::resetAll();
  } catch (const b200lp::Error& e) {
    std::fprintf(stderr, "b200lp::Error(%d): %s\n", e.code(), e.what());
    return 10;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 11;
  }
  return 0;
}
