// The seven critic plugins of the reference (mpc_critics.xml:1-33) as adapters over a b200lp::Session. Each keeps its
// reference class name and parameters and returns, for the trajectory it is handed, the value the fused kernel
// computed for THAT critic — so the reference's StackedScoringModel (sum in order, negative early-out) reproduces
// Trajectory::cost_ exactly:
//   CollisionModel (models/collision_model.cpp:51-148)          CollisionMinMaxModel (collision_min_max_model.cpp:51-88)
//   StickPathModel (stick_path_model.cpp:51-77)                 PurePursuitModel (pure_pursuit_model.cpp:60-114)
//   TowardGlobalPlanModel (toward_global_plan_model.cpp:52-78)  ShortestAngleModel (shortest_angle_model.cpp:51-69)
//   TwirlingModel (twirling_model.cpp:51-55)
#ifndef B200LP_B200_MODELS_H_
#define B200LP_B200_MODELS_H_
#include "b200lp/session.hpp"
#include "mpc_critics/scoring_model.h"

namespace mpc_critics {

class B200ModelBase : public ScoringModel {
 public:
  double scoreTrajectory(base_trajectory::Trajectory& traj) override;

 protected:
  // reads `<name>.weight` (+ the pure-pursuit weights), finds the generator the critic is bound to
  // (`<name>.trajectory_generator`) and takes the next slot of that generator's stack
  void bind(int kind);
  std::shared_ptr<b200lp::Session> session_;
  int index_ = -1;
};

#define B200LP_DECLARE_CRITIC(Class, KIND)        \
  class Class : public B200ModelBase {            \
   protected:                                     \
    void onInitialize() override { bind(KIND); }  \
  }
B200LP_DECLARE_CRITIC(CollisionModel, B200LP_CRITIC_COLLISION);
B200LP_DECLARE_CRITIC(CollisionMinMaxModel, B200LP_CRITIC_COLLISION_MIN_MAX);
B200LP_DECLARE_CRITIC(StickPathModel, B200LP_CRITIC_STICK_PATH);
B200LP_DECLARE_CRITIC(PurePursuitModel, B200LP_CRITIC_PURE_PURSUIT);
B200LP_DECLARE_CRITIC(TowardGlobalPlanModel, B200LP_CRITIC_TOWARD_GLOBAL_PLAN);
B200LP_DECLARE_CRITIC(ShortestAngleModel, B200LP_CRITIC_SHORTEST_ANGLE);
B200LP_DECLARE_CRITIC(TwirlingModel, B200LP_CRITIC_TWIRLING);
#undef B200LP_DECLARE_CRITIC

}  // namespace mpc_critics
#endif
