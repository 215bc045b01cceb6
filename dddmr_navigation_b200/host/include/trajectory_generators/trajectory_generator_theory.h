// trajectory_generators::TrajectoryGeneratorTheory — the generator plugin base, same surface as the reference
// (src/dddmr_local_planner/trajectory_generators/include/trajectory_generators/trajectory_generator_theory.h:49-73).
#ifndef B200LP_TRAJECTORY_GENERATOR_THEORY_H_
#define B200LP_TRAJECTORY_GENERATOR_THEORY_H_
#include <memory>
#include <string>

#include "base_trajectory/trajectory.h"
#include "trajectory_generators/trajectory_shared_data.h"

namespace trajectory_generators {
class TrajectoryGeneratorTheory {
 public:
  TrajectoryGeneratorTheory() {}
  virtual ~TrajectoryGeneratorTheory() {}
  void initialize(const std::string name, const rclcpp::Node::WeakPtr& weak_node) {
    name_ = name;
    node_ = weak_node.lock();
    onInitialize();
  }
  void setSharedData(std::shared_ptr<trajectory_generators::TrajectoryGeneratorSharedData> shared_data) {
    shared_data_ = shared_data;
  }
  virtual bool hasMoreTrajectories() = 0;
  virtual bool nextTrajectory(base_trajectory::Trajectory& _traj) = 0;
  virtual void initialise() = 0;  // per-cycle reset, called through StackedGenerator

 protected:
  rclcpp::Node::SharedPtr node_;
  virtual void onInitialize() = 0;  // read the plugin's parameters
  std::shared_ptr<trajectory_generators::TrajectoryGeneratorSharedData> shared_data_;
  std::string name_;
};
}  // namespace trajectory_generators
#endif
