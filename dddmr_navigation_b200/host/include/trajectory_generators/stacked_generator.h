// trajectory_generators::StackedGenerator — name -> theory map with the shared-data hand-over
// (reference: trajectory_generators/include/trajectory_generators/stacked_generator.h:42-75, src/stacked_generator.cpp:61-111).
#ifndef B200LP_STACKED_GENERATOR_H_
#define B200LP_STACKED_GENERATOR_H_
#include <map>
#include <memory>
#include <mutex>
#include <string>

#include "trajectory_generators/trajectory_generator_theory.h"

namespace trajectory_generators {
class StackedGenerator {
 public:
  StackedGenerator() : shared_data_(std::make_shared<TrajectoryGeneratorSharedData>()) {}
  void addPlugin(std::string pname, std::shared_ptr<TrajectoryGeneratorTheory> theory) {
    theories_[pname] = theory;
    theory->setSharedData(shared_data_);
  }
  std::shared_ptr<TrajectoryGeneratorSharedData> getSharedDataPtr() { return shared_data_; }
  void initializeTheories_wi_Shared_data() {
    for (auto& kv : theories_) kv.second->initialise();
  }
  bool hasMoreTrajectories(std::string pname) {
    auto it = theories_.find(pname);
    return it != theories_.end() && it->second->hasMoreTrajectories();  // unknown name: false, like the reference
  }
  bool nextTrajectory(std::string pname, base_trajectory::Trajectory& comp_traj) {
    auto it = theories_.find(pname);
    if (it == theories_.end() || !it->second->hasMoreTrajectories()) return false;
    return it->second->nextTrajectory(comp_traj);
  }
  typedef std::recursive_mutex theory_mutex_t;
  theory_mutex_t* getMutex() { return &access_; }

 private:
  std::map<std::string, std::shared_ptr<TrajectoryGeneratorTheory>> theories_;
  std::shared_ptr<TrajectoryGeneratorSharedData> shared_data_;
  theory_mutex_t access_;
};
}  // namespace trajectory_generators
#endif
