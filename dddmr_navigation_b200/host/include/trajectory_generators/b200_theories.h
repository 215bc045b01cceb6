// The three generator plugins of the reference (trajectory_generators.xml:1-17), re-implemented as adapters over a
// b200lp::Session. Class names and parameter names are the reference's, so an unchanged YAML selects them:
//   trajectory_generators::DDSimpleTrajectoryGeneratorTheory    (theories/dd_simple_trajectory_generator_theory.cpp)
//   trajectory_generators::OmniSimpleTrajectoryGeneratorTheory  (theories/omni_simple_trajectory_generator_theory.cpp)
//   trajectory_generators::DDRotateInplaceTheory                (theories/dd_rotate_inplace_theory.cpp)
// initialise() opens a device cycle with the data in the shared struct; the sampling, the rollout and — fused into the
// same launch — the critics all run on the GPU; nextTrajectory() hands out the generated list in the reference's order.
#ifndef B200LP_B200_THEORIES_H_
#define B200LP_B200_THEORIES_H_
#include "b200lp/session.hpp"
#include "trajectory_generators/trajectory_generator_theory.h"

namespace trajectory_generators {

class B200TheoryBase : public TrajectoryGeneratorTheory {
 public:
  bool hasMoreTrajectories() override;
  bool nextTrajectory(base_trajectory::Trajectory& _traj) override;
  void initialise() override;
  std::shared_ptr<b200lp::Session> session() { return session_; }

 protected:
  // declares + reads the parameter set every theory shares, the theory-specific ones selected by `theory`
  void readParameters(int theory);
  std::shared_ptr<b200lp::Session> session_;
  int next_ = 0;
  bool materialize_points_ = true;  // `<name>.b200_materialize_points`: fill poses/cuboids into every Trajectory
};

class DDSimpleTrajectoryGeneratorTheory : public B200TheoryBase {
 protected:
  void onInitialize() override { readParameters(B200LP_THEORY_DD_SIMPLE); }
};
class OmniSimpleTrajectoryGeneratorTheory : public B200TheoryBase {
 protected:
  void onInitialize() override { readParameters(B200LP_THEORY_OMNI_SIMPLE); }
};
class DDRotateInplaceTheory : public B200TheoryBase {
 protected:
  void onInitialize() override { readParameters(B200LP_THEORY_DD_ROTATE_INPLACE); }
};

}  // namespace trajectory_generators
#endif
