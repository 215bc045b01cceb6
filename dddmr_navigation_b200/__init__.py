"""dddmr_navigation_b200 — B200-native (sm_100a) local-planner rollout-and-score path of dddmr_navigation.

Scope: the hot path only: velocity sampling -> rollout -> voxel-grid obstacle query ->
critics -> argmin, behind the C ABI in include/b200lp.h. The CUDA library must be built first
(`python -c "import __graft_entry__ as g; g.build()"`); there is no CPU fallback.
"""
from . import abi  # noqa: F401
from .config import PlannerConfig, make_query  # noqa: F401
from .planner import LocalPlanner, Local_Planner, PlannerState, Trajectory  # noqa: F401

__all__ = ["abi", "PlannerConfig", "make_query", "LocalPlanner", "Local_Planner", "PlannerState", "Trajectory"]
