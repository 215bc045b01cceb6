// mpc_critics::StackedScoringModel — generator name -> ordered critic list, summed with the negative early-out
// (reference: include/mpc_critics/stacked_scoring_model.h:42-75, src/stacked_scoring_model.cpp:57-93).
#ifndef B200LP_STACKED_SCORING_MODEL_H_
#define B200LP_STACKED_SCORING_MODEL_H_
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "mpc_critics/scoring_model.h"

namespace mpc_critics {
class StackedScoringModel {
 public:
  StackedScoringModel() : shared_data_(std::make_shared<ModelSharedData>()) {}
  void addPluginByTraj(std::string traj_name, std::shared_ptr<ScoringModel> model) {
    model->setSharedData(shared_data_);
    models_map_[traj_name].push_back(model);
  }
  std::shared_ptr<ModelSharedData> getSharedDataPtr() { return shared_data_; }
  void scoreTrajectory(std::string traj_gen_name, base_trajectory::Trajectory& one_traj) {
    for (auto& model : models_map_[traj_gen_name]) {
      const double return_cost = model->scoreTrajectory(one_traj);
      if (return_cost < 0) {
        one_traj.cost_ = return_cost;
        break;
      }
      one_traj.cost_ += return_cost;
    }
  }
  typedef std::recursive_mutex model_mutex_t;
  model_mutex_t* getMutex() { return &access_; }

 private:
  std::map<std::string, std::vector<std::shared_ptr<ScoringModel>>> models_map_;
  std::shared_ptr<ModelSharedData> shared_data_;
  model_mutex_t access_;
};
}  // namespace mpc_critics
#endif
