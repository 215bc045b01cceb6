// mpc_critics::ScoringModel — the critic plugin base, same surface as the reference
// (src/dddmr_local_planner/mpc_critics/include/mpc_critics/scoring_model.h:44-73).
#ifndef B200LP_SCORING_MODEL_H_
#define B200LP_SCORING_MODEL_H_
#include <memory>
#include <string>

#include "base_trajectory/trajectory.h"
#include "mpc_critics/model_shared_data.h"

namespace mpc_critics {
class ScoringModel {
 public:
  ScoringModel() : weight_(1.0) {}
  virtual ~ScoringModel() {}
  void initialize(const std::string name, const rclcpp::Node::WeakPtr& weak_node) {
    name_ = name;
    node_ = weak_node.lock();
    onInitialize();
  }
  /** score for trajectory traj; negative rejects it */
  virtual double scoreTrajectory(base_trajectory::Trajectory& traj) = 0;
  void setSharedData(std::shared_ptr<mpc_critics::ModelSharedData> shared_data) { shared_data_ = shared_data; }
  std::string getModelName() { return name_; }

 protected:
  rclcpp::Node::SharedPtr node_;
  virtual void onInitialize() = 0;
  std::string name_;
  std::shared_ptr<mpc_critics::ModelSharedData> shared_data_;
  double weight_;
};
}  // namespace mpc_critics
#endif
