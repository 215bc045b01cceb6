"""Parameter sets in the reference's own YAML vocabulary, and their conversion to the C ABI structs.

The reference configures the path through ROS 2 parameters declared per plugin
(trajectory_generators/theories/dd_simple_trajectory_generator_theory.cpp:43-234,
mpc_critics/src/mpc_critics_ros.cpp:60-82). :class:`PlannerConfig` accepts the same keys (a dict shaped like
the ``trajectory_generators`` / ``mpc_critics`` ``ros__parameters`` blocks of
dddmr_p2p_move_base/config/p2p_move_base_localization.yaml:150-247) so a user can paste their YAML.
"""
from __future__ import annotations

import copy
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import abi

# Cuboid vertex order the collision critic relies on (dd_simple…cpp:211-218): blb, brb, blt, flb, brt, frt, flt, frb
CUBOID_ORDER = ("blb", "brb", "blt", "flb", "brt", "frt", "flt", "frb")

# dddmr_p2p_move_base/config/p2p_move_base_localization.yaml:170-214
DEFAULT_CUBOID = {
    "flb": [0.42, 0.36, 0.0], "frb": [0.42, -0.36, 0.0], "flt": [0.42, 0.36, 0.6], "frt": [0.42, -0.36, 0.6],
    "blb": [-0.35, 0.36, 0.0], "brb": [-0.35, -0.36, 0.0], "blt": [-0.35, 0.36, 0.6], "brt": [-0.35, -0.36, 0.6],
}

DD_SIMPLE_DEFAULT = {  # p2p_move_base_localization.yaml:185-214
    "plugin": "trajectory_generators::DDSimpleTrajectoryGeneratorTheory",
    "max_vel_x": 1.0, "min_vel_x": 0.1, "max_vel_theta": 0.6, "min_vel_theta": 0.15,
    "acc_lim_x": 1.0, "acc_lim_theta": 3.0, "deceleration_ratio": 2.0,
    "max_motor_shaft_rpm": 3000.0, "wheel_diameter": 0.16, "gear_ratio": 1.0, "robot_radius": 0.25,
    "controller_frequency": 10.0, "sim_time": 2.0, "linear_x_sample": 5.0, "angular_z_sample": 10.0,
    "sim_granularity": 0.05, "angular_sim_granularity": 0.025,
    "cuboid": DEFAULT_CUBOID,
}

DD_ROTATE_INPLACE_DEFAULT = {  # p2p_move_base_localization.yaml:172-183; every other key takes its code default
    "plugin": "trajectory_generators::DDRotateInplaceTheory",
    "controller_frequency": 10.0, "rotation_speed": 0.5,
    "cuboid": DEFAULT_CUBOID,
}

OMNI_SIMPLE_DEFAULT = {  # dddmr_p2p_move_base/config/p2p_wo_mcl.yaml:86-118
    "plugin": "trajectory_generators::OmniSimpleTrajectoryGeneratorTheory",
    "max_vel_x": 1.0, "min_vel_x": -1.0, "max_vel_y": 1.0, "min_vel_y": -1.0,
    "max_vel_theta": 0.6, "min_vel_theta": 0.15, "min_vel_trans": 0.1, "max_vel_trans": 1.0,
    "acc_lim_x": 2.0, "acc_lim_y": 2.0, "acc_lim_theta": 3.0, "deceleration_ratio": 2.0,
    "use_motor_constraint": False,
    "controller_frequency": 10.0, "sim_time": 2.0, "linear_x_sample": 5.0, "linear_y_sample": 5.0,
    "angular_z_sample": 10.0, "sim_granularity": 0.05, "angular_sim_granularity": 0.025,
    "cuboid": DEFAULT_CUBOID,
}

OMNI_SIMPLE_CRITICS = [  # p2p_wo_mcl.yaml:120-143
    {"name": "collision", "plugin": "mpc_critics::CollisionModel", "weight": 1.0},
    {"name": "stick_path", "plugin": "mpc_critics::StickPathModel", "weight": 0.1},
    {"name": "pure_pursuit", "plugin": "mpc_critics::PurePursuitModel", "translation_weight": 1.0,
     "orientation_weight": 0.01},
    {"name": "toward_global_plan", "plugin": "mpc_critics::TowardGlobalPlanModel", "weight": 1.0},
    {"name": "twirling", "plugin": "mpc_critics::TwirlingModel", "weight": 1.0},
]

# p2p_move_base_localization.yaml:216-235 — order matters (stacked_scoring_model.cpp:75-93)
DD_SIMPLE_CRITICS = [
    {"name": "collision", "plugin": "mpc_critics::CollisionModel", "weight": 1.0},
    {"name": "stick_path", "plugin": "mpc_critics::StickPathModel", "weight": 0.1},
    {"name": "pure_pursuit", "plugin": "mpc_critics::PurePursuitModel", "translation_weight": 1.0,
     "orientation_weight": 0.01},
    {"name": "toward_global_plan", "plugin": "mpc_critics::TowardGlobalPlanModel", "weight": 1.0},
]
ROTATE_CRITICS = [  # collision_rotate_shortest + prefer_rotate_shortest (p2p_move_base_localization.yaml:240-247)
    {"name": "collision_rotate_shortest", "plugin": "mpc_critics::CollisionModel", "weight": 1.0},
    {"name": "prefer_rotate_shortest", "plugin": "mpc_critics::ShortestAngleModel", "weight": 1.0},
]

# code defaults of declare_parameter(...) when a key is absent from the YAML
_CODE_DEFAULTS = {
    "min_vel_x": 0.01, "max_vel_x": 0.1, "min_vel_y": 0.01, "max_vel_y": 0.1, "min_vel_trans": 0.01,
    "max_vel_trans": 0.1, "min_vel_theta": 0.1, "max_vel_theta": 0.1, "acc_lim_x": 0.3, "acc_lim_y": 0.3,
    "acc_lim_theta": 0.5, "deceleration_ratio": 2.0, "use_motor_constraint": False, "max_motor_shaft_rpm": 3000.0,
    "wheel_diameter": 0.15, "gear_ratio": 30.0, "robot_radius": 0.25, "rotation_speed": 0.4,
    "controller_frequency": 10.0, "sim_time": 2.0, "linear_x_sample": 10.0, "linear_y_sample": 10.0,
    "angular_z_sample": 10.0, "sim_granularity": 0.1, "angular_sim_granularity": 0.05,
}


@dataclass
class PlannerConfig:
    """One trajectory generator + its ordered critic list + the voxel-grid settings."""
    generator: dict = field(default_factory=lambda: copy.deepcopy(DD_SIMPLE_DEFAULT))
    critics: list = field(default_factory=lambda: copy.deepcopy(DD_SIMPLE_CRITICS))
    cell_xy: float = 0.0  # 0 = library default
    cell_z: float = 0.0
    max_cells: int = 0

    @classmethod
    def from_ros_yaml(cls, params: dict, generator_name: str) -> "PlannerConfig":
        """params: {'trajectory_generators': {'ros__parameters': {...}}, 'mpc_critics': {'ros__parameters': {...}}}."""
        tg = params["trajectory_generators"]["ros__parameters"]
        mc = params["mpc_critics"]["ros__parameters"]
        gen = copy.deepcopy(tg[generator_name])
        critics = []
        for name in mc["plugins"]:  # YAML order is evaluation order (mpc_critics_ros.cpp:63-82)
            c = mc[name]
            if c.get("trajectory_generator") == generator_name:
                d = dict(c)
                d["name"] = name
                critics.append(d)
        return cls(generator=gen, critics=critics)

    def to_ros_yaml(self, generator_name: str = "differential_drive_simple") -> str:
        """The same configuration as the text of a ROS 2 params file in the reference's layout (two nodes,
        `trajectory_generators` and `mpc_critics`) — what the C++ host layer's plugin loaders read."""
        def scalar(v):
            if isinstance(v, bool):
                return "true" if v else "false"
            if isinstance(v, str):
                return f'"{v}"'
            return repr(float(v))
        out = ["trajectory_generators:", "  ros__parameters:", f'    plugins: ["{generator_name}"]', f"    {generator_name}:"]
        for k, v in self.generator.items():
            if k == "cuboid":
                out.append("      cuboid:")
                for name, xyz in v.items():
                    out.append(f"        {name}: [{', '.join(repr(float(c)) for c in xyz)}]")
            else:
                out.append(f"      {k}: {scalar(v)}")
        names = [c.get("name", f"critic{i}") for i, c in enumerate(self.critics)]
        out += ["", "mpc_critics:", "  ros__parameters:", "    plugins: [" + ", ".join(f'"{n}"' for n in names) + "]"]
        for n, c in zip(names, self.critics):
            out.append(f"    {n}:")
            out.append(f'      plugin: "{c["plugin"]}"')
            out.append(f"      trajectory_generator: {generator_name}")
            for k, v in c.items():
                if k not in ("name", "plugin", "trajectory_generator"):
                    out.append(f"      {k}: {scalar(v)}")
        return "\n".join(out) + "\n"

    # ---- conversion to ABI structs -----------------------------------------------------------
    def _g(self, key):
        return self.generator.get(key, _CODE_DEFAULTS[key])

    def limits(self) -> abi.Limits:
        L = abi.Limits()
        for n in ("max_vel_x", "min_vel_x", "max_vel_y", "min_vel_y", "max_vel_trans", "min_vel_trans",
                  "max_vel_theta", "min_vel_theta", "acc_lim_x", "acc_lim_y", "acc_lim_theta",
                  "deceleration_ratio", "max_motor_shaft_rpm", "wheel_diameter", "gear_ratio", "robot_radius",
                  "rotation_speed"):
            setattr(L, n, float(self._g(n)))
        L.use_motor_constraint = 1 if self._g("use_motor_constraint") else 0
        return L

    def params(self) -> abi.Params:
        P = abi.Params()
        P.theory = abi.THEORY_BY_PLUGIN[self.generator["plugin"]]
        for n in ("controller_frequency", "sim_time", "linear_x_sample", "linear_y_sample", "angular_z_sample",
                  "sim_granularity", "angular_sim_granularity"):
            setattr(P, n, float(self._g(n)))
        return P

    def cuboid(self) -> np.ndarray:
        cub = self.generator["cuboid"]
        # the reference stores vertices in pcl::PointXYZ, i.e. rounded to float
        return np.ascontiguousarray([cub[k] for k in CUBOID_ORDER], dtype=np.float32)

    def critic_array(self):
        arr = (abi.Critic * max(1, len(self.critics)))()
        for i, c in enumerate(self.critics):
            arr[i].kind = abi.CRITIC_BY_PLUGIN[c["plugin"]]
            arr[i].weight = float(c.get("weight", 1.0))
            arr[i].translation_weight = float(c.get("translation_weight", 0.5))
            arr[i].orientation_weight = float(c.get("orientation_weight", 0.5))
        return arr, len(self.critics)

    def grid_config(self) -> abi.GridConfig:
        g = abi.GridConfig()
        g.cell_xy = float(self.cell_xy)
        g.cell_z = float(self.cell_z)
        g.max_cells = int(self.max_cells)
        return g


def make_query(pose7, twist3, max_speed_override=-1.0, heading_deviation=0.0) -> abi.Query:
    q = abi.Query()
    q.pose[:] = [float(v) for v in pose7]
    q.twist[:] = [float(v) for v in twist3]
    q.max_speed_override = float(max_speed_override)
    q.heading_deviation = float(heading_deviation)
    return q


def as_c(arr: np.ndarray, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))
