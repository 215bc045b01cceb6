// dddmr_sys_core::PlannerState — the return codes of Local_Planner::computeVelocityCommand
// (src/dddmr_sys_core/include/dddmr_sys_core/dddmr_enum_states.h:46-54), same enumerators in the same order.
#ifndef B200LP_DDDMR_ENUM_STATES_H_
#define B200LP_DDDMR_ENUM_STATES_H_
namespace dddmr_sys_core {
enum PlannerState {
  TF_FAIL,
  PRUNE_PLAN_FAIL,
  ALL_TRAJECTORIES_FAIL,
  PERCEPTION_MALFUNCTION,
  TRAJECTORY_FOUND,
  PATH_BLOCKED_WAIT,
  PATH_BLOCKED_REPLANNING
};
// the return codes of a recovery behaviour (dddmr_enum_states.h:56-62), same enumerators in the same order
enum RecoveryState {
  RECOVERY_BEHAVIOR_NOT_FOUND,
  INTERRUPT_BY_CANCEL,
  INTERRUPT_BY_NEW_GOAL,
  RECOVERY_DONE,
  RECOVERY_FAIL
};
}
#endif
