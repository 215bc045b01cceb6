"""ctypes mirror of include/b200lp.h (the C ABI) and the loader of libb200lp.so.

The structs here are the single Python definition of the ABI's POD types; the CPU oracle's wrapper
(test infrastructure, under oracle/) reuses them so tests hand both sides the same bytes.

There is no CPU fallback: :func:`load_library` raises if the CUDA library has not been built.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
COUNT_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libb200lp_count.so")
CHECKS_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libb200lp_checks.so")
LIB_PATH = os.path.join(HERE, "csrc", "libb200lp.so")

ABI_VERSION = 6

# status codes
OK, E_INVALID, E_CUDA, E_STATE, E_NOMEM = 0, -1, -2, -3, -4

# theories (reference: trajectory_generators/trajectory_generators.xml)
THEORY_DD_SIMPLE, THEORY_OMNI_SIMPLE, THEORY_DD_ROTATE_INPLACE = 0, 1, 2
THEORY_BY_PLUGIN = {
    "trajectory_generators::DDSimpleTrajectoryGeneratorTheory": THEORY_DD_SIMPLE,
    "trajectory_generators::OmniSimpleTrajectoryGeneratorTheory": THEORY_OMNI_SIMPLE,
    "trajectory_generators::DDRotateInplaceTheory": THEORY_DD_ROTATE_INPLACE,
}

# critics (reference: mpc_critics/mpc_critics.xml)
(CRITIC_COLLISION, CRITIC_COLLISION_MIN_MAX, CRITIC_STICK_PATH, CRITIC_PURE_PURSUIT,
 CRITIC_TOWARD_GLOBAL_PLAN, CRITIC_SHORTEST_ANGLE, CRITIC_TWIRLING) = range(7)
CRITIC_BY_PLUGIN = {
    "mpc_critics::CollisionModel": CRITIC_COLLISION,
    "mpc_critics::CollisionMinMaxModel": CRITIC_COLLISION_MIN_MAX,
    "mpc_critics::StickPathModel": CRITIC_STICK_PATH,
    "mpc_critics::PurePursuitModel": CRITIC_PURE_PURSUIT,
    "mpc_critics::TowardGlobalPlanModel": CRITIC_TOWARD_GLOBAL_PLAN,
    "mpc_critics::ShortestAngleModel": CRITIC_SHORTEST_ANGLE,
    "mpc_critics::TwirlingModel": CRITIC_TWIRLING,
}
MAX_CRITICS = 8
MAX_STEPS = 512
MAX_PLAN = 1024
MAX_SENSORS = 8
PEER_HANDLE_BYTES = 64
MAX_PEERS = 16


class Limits(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "max_vel_x", "min_vel_x", "max_vel_y", "min_vel_y", "max_vel_trans", "min_vel_trans",
        "max_vel_theta", "min_vel_theta", "acc_lim_x", "acc_lim_y", "acc_lim_theta", "deceleration_ratio",
        "max_motor_shaft_rpm", "wheel_diameter", "gear_ratio", "robot_radius", "rotation_speed")] + [
        ("use_motor_constraint", C.c_int32), ("reserved_", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("theory", C.c_int32), ("reserved_", C.c_int32)] + [(n, C.c_double) for n in (
        "controller_frequency", "sim_time", "linear_x_sample", "linear_y_sample", "angular_z_sample",
        "sim_granularity", "angular_sim_granularity")]


class Critic(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved_", C.c_int32), ("weight", C.c_double),
                ("translation_weight", C.c_double), ("orientation_weight", C.c_double)]


class GridConfig(C.Structure):
    _fields_ = [("cell_xy", C.c_float), ("cell_z", C.c_float), ("max_cells", C.c_uint32), ("reserved_", C.c_uint32)]


class Query(C.Structure):
    _fields_ = [("pose", C.c_double * 7), ("twist", C.c_double * 3), ("max_speed_override", C.c_double),
                ("heading_deviation", C.c_double)]


class Result(C.Structure):
    _fields_ = [("best_id", C.c_int32), ("n_samples", C.c_int32), ("n_traj", C.c_int32), ("n_collided", C.c_int32),
                ("n_poses", C.c_int64), ("best_cost", C.c_double), ("xv", C.c_double), ("yv", C.c_double),
                ("thetav", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class TrajView(C.Structure):
    _fields_ = [("sample_index", C.POINTER(C.c_int32)), ("vel", C.POINTER(C.c_float)),
                ("num_steps", C.POINTER(C.c_int32)), ("time_delta", C.POINTER(C.c_double)),
                ("cost", C.POINTER(C.c_double)), ("critic_scores", C.POINTER(C.c_double)),
                ("first_hit_pose", C.POINTER(C.c_int32))]


class PoseView(C.Structure):
    _fields_ = [("pose", C.POINTER(C.c_double)), ("pcl_pose", C.POINTER(C.c_float)), ("cuboid", C.POINTER(C.c_float)),
                ("aabb", C.POINTER(C.c_float)), ("collide", C.POINTER(C.c_uint8)), ("n_r1", C.POINTER(C.c_int32))]


class PruneInfo(C.Structure):
    _fields_ = [("status", C.c_int32), ("nearest_index", C.c_int32), ("n_prune", C.c_int32), ("n_backward", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Blocked(C.Structure):
    _fields_ = [("n_blocked", C.c_int32), ("n_checked", C.c_int32), ("n_total", C.c_int32), ("opinion", C.c_int32),
                ("ratio", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SensorParams(C.Structure):
    """What MultiLayerSpinningLidar reads for cbSensor (multilayer_spinning_lidar.cpp:64-131)."""
    _fields_ = [("perception_window_size", C.c_double), ("marking_height", C.c_double), ("leaf_size", C.c_float),
                ("is_local_planner", C.c_int32)]


class ObservationInfo(C.Structure):
    _fields_ = [("n_scan", C.c_int64), ("n_window", C.c_int64), ("n_points", C.c_int64), ("ms_device", C.c_float),
                ("n_launches", C.c_int32), ("ms_upload", C.c_float), ("reserved_", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/b200lp.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "b200lp_last_error": (C.c_char_p, [_P]),
    "b200lp_abi_version": (C.c_int, []),
    "b200lp_create": (C.c_int, [C.POINTER(_P), C.c_int, C.POINTER(Limits), C.POINTER(Params), C.POINTER(C.c_float),
                                C.POINTER(Critic), C.c_int, C.POINTER(GridConfig)]),
    "b200lp_destroy": (None, [_P]),
    "b200lp_set_cloud": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t]),
    "b200lp_last_upload": (C.c_int, [_P, C.POINTER(C.c_size_t), C.POINTER(C.c_int32)]),
    "b200lp_set_cloud_device": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t]),
    "b200lp_set_plan": (C.c_int, [_P, C.POINTER(C.c_double), C.c_size_t]),
    "b200lp_plan": (C.c_int, [_P, C.POINTER(Query), C.POINTER(Result)]),
    "b200lp_plan_shard": (C.c_int, [_P, C.POINTER(Query), C.c_int, C.c_int, C.POINTER(Result)]),
    "b200lp_peer_export": (C.c_int, [_P, C.POINTER(C.c_uint8)]),
    "b200lp_peer_attach": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_uint8)]),
    "b200lp_plan_shard_exchange": (C.c_int, [_P, C.POINTER(Query), C.POINTER(Result)]),
    "b200lp_peer_resync": (C.c_int, [_P]),
    "b200lp_peer_reserve_cloud": (C.c_int, [_P, C.c_size_t]),
    "b200lp_set_cloud_shared": (C.c_int, [_P, C.c_int, _P, C.c_size_t, C.c_size_t]),
    "b200lp_set_shard_cuts": (C.c_int, [_P, C.POINTER(C.c_float), C.c_int]),
    "b200lp_get_shard_cuts": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "b200lp_set_adaptive_cuts": (C.c_int, [_P, C.c_int]),
    "b200lp_last_cycle_ns": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "b200lp_plan_batch": (C.c_int, [_P, C.POINTER(Query), C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                    C.POINTER(Result)]),
    "b200lp_traj_count": (C.c_int, [_P, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "b200lp_read_trajectories": (C.c_int, [_P, C.c_size_t, C.POINTER(TrajView)]),
    "b200lp_read_poses": (C.c_int, [_P, C.c_size_t, C.c_int32, C.POINTER(PoseView)]),
    "b200lp_read_pose_batch": (C.c_int, [_P, C.c_size_t, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(PoseView),
                                         C.c_size_t]),
    "b200lp_set_global_plan": (C.c_int, [_P, C.POINTER(C.c_double), C.c_size_t]),
    "b200lp_prune_plan": (C.c_int, [_P, C.POINTER(C.c_double), C.c_double, C.c_double, C.POINTER(PruneInfo)]),
    "b200lp_read_prune_plan": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_size_t]),
    "b200lp_path_blocked": (C.c_int, [_P, C.c_double, C.POINTER(Blocked)]),
    "b200lp_sensor_observation": (C.c_int, [_P, C.c_int, _P, C.c_size_t, C.c_size_t, C.POINTER(C.c_double),
                                            C.POINTER(C.c_double), C.POINTER(SensorParams), C.POINTER(ObservationInfo)]),
    "b200lp_read_observation": (C.c_int, [_P, C.c_int, _P, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t)]),
    "b200lp_aggregate_observations": (C.c_int, [_P, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_size_t)]),
    "b200lp_count_radius": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "b200lp_work_counters": (C.c_int, [_P, C.POINTER(C.c_uint64), C.c_int]),
    "b200lp_last_timing": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                     C.POINTER(C.c_float)]),
    "b200lp_last_kernel_ms": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "b200lp_last_kernel_times": (C.c_int, [_P, C.POINTER(C.c_float), C.c_int]),
    "b200lp_launch_count": (C.c_int64, [_P]),
    "b200lp_grid_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                   C.POINTER(C.c_int64)]),
    "b200lp_stream": (_P, [_P]),
}

_lib = None


class B200LPError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"b200lp error {code}: {text}")
        self.code = code


def load_library(path: str | None = None):
    """dlopen libb200lp.so and type every entry point. Raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    # B200LP_LIB: run everything against another build of the SAME sources (A/B builds of tools/variants); still CUDA only
    p = path or os.environ.get("B200LP_LIB") or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(
            f"{p} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "dddmr_navigation_b200 has no CPU fallback.")
    lib = C.CDLL(p)
    lib.b200lp_abi_version.restype = C.c_int
    older = path is not None and lib.b200lp_abi_version() < ABI_VERSION  # an A/B build of an earlier round (tools/time_variants.py)
    for name, (res, args) in SYMBOLS.items():
        if path is not None and not hasattr(lib, name):  # explicitly named A/B builds may predate an entry point
            continue
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.b200lp_abi_version() != ABI_VERSION and not older:
        raise ImportError(f"{p}: ABI version {lib.b200lp_abi_version()} != {ABI_VERSION}")
    if path is None:
        _lib = lib
    return lib
