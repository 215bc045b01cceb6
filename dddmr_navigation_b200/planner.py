"""Host-side mirror of the reference's local-planner cycle on top of the C ABI (include/b200lp.h).

:class:`LocalPlanner` is the thin, literal wrapper (one method per C entry point).
:class:`Local_Planner` keeps the reference's operator names and return codes for the hot path —
``setPlan`` / ``computeVelocityCommand`` returning a ``PlannerState`` and a ``Trajectory``
(src/dddmr_local_planner/local_planner/include/local_planner/local_planner.h:72-82,
src/dddmr_sys_core/include/dddmr_sys_core/dddmr_enum_states.h:46-54) — so tests and callers read like
the reference. Everything numeric happens in libb200lp.so on the GPU; nothing here computes.
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass

import numpy as np

from . import abi
from .config import PlannerConfig, make_query

_P = C.c_void_p


class PlannerState(enum.IntEnum):
    """dddmr_sys_core::PlannerState (dddmr_enum_states.h:46-54) — the values the hot path can produce."""
    TF_FAIL = 0
    PRUNE_PLAN_FAIL = 1
    ALL_TRAJECTORIES_FAIL = 2
    PERCEPTION_MALFUNCTION = 3
    TRAJECTORY_FOUND = 4
    PATH_BLOCKED_WAIT = 5
    PATH_BLOCKED_REPLANNING = 6


@dataclass
class Trajectory:
    """The fields of base_trajectory::Trajectory the caller consumes (trajectory.h:62-66)."""
    xv_: float = 0.0
    yv_: float = 0.0
    thetav_: float = 0.0
    cost_: float = -1.0
    time_delta_: float = 0.0
    id: int = -1


def _cloud_bytes(pts: np.ndarray):
    """Accept (N,3) xyz, (N,4) PointXYZ layout or (N,8) PointXYZI layout float32; returns (array, stride)."""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    if pts.ndim != 2 or pts.shape[1] not in (3, 4, 8):
        raise ValueError("cloud must be (N,3), (N,4) [pcl::PointXYZ] or (N,8) [pcl::PointXYZI] float32")
    return pts, pts.shape[1] * 4


class LocalPlanner:
    """One generator + its critic stack on one GPU (a b200lp_ctx)."""

    def __init__(self, config: PlannerConfig | None = None, device: int = 0, lib_path: str | None = None):
        self.lib = abi.load_library(lib_path)
        self._batch_cache = None
        self.config = config or PlannerConfig()
        self._L, self._Pm = self.config.limits(), self.config.params()
        self._cub = self.config.cuboid()
        self._crit, self.n_critics = self.config.critic_array()
        self._grid = self.config.grid_config()
        h = _P()
        rc = self.lib.b200lp_create(C.byref(h), device, C.byref(self._L), C.byref(self._Pm),
                                    self._cub.ctypes.data_as(C.POINTER(C.c_float)), self._crit, self.n_critics,
                                    C.byref(self._grid))
        if rc != 0:
            raise abi.B200LPError(rc, (self.lib.b200lp_last_error(None) or b"").decode())
        self.h = h
        self.last = None
        self._n_robots = 0
        self._results = None

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.b200lp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise abi.B200LPError(rc, (self.lib.b200lp_last_error(self.h) or b"").decode())

    # -- inputs ---------------------------------------------------------------------------------
    def set_cloud(self, pts: np.ndarray):
        pts, stride = _cloud_bytes(pts)
        self._ck(self.lib.b200lp_set_cloud(self.h, pts.ctypes.data_as(_P), pts.shape[0], stride))

    def set_cloud_device(self, dev_ptr: int, n: int, stride: int):
        self._ck(self.lib.b200lp_set_cloud_device(self.h, _P(dev_ptr), n, stride))

    def set_plan(self, plan: np.ndarray):
        plan = np.ascontiguousarray(plan, dtype=np.float64).reshape(-1, 7)
        self._ck(self.lib.b200lp_set_plan(self.h, plan.ctypes.data_as(C.POINTER(C.c_double)), plan.shape[0]))

    # -- the steps either side of the cycle ------------------------------------------
    def set_global_plan(self, plan: np.ndarray):
        """Local_Planner::setPlan (local_planner.cpp:322-343): (n,7) position + orientation xyzw, n >= 3."""
        plan = np.ascontiguousarray(plan, dtype=np.float64).reshape(-1, 7)
        self._ck(self.lib.b200lp_set_global_plan(self.h, plan.ctypes.data_as(C.POINTER(C.c_double)), plan.shape[0]))

    def prune_plan(self, robot_xyz, forward_distance: float, backward_distance: float) -> abi.PruneInfo:
        """Local_Planner::prunePlan (local_planner.cpp:374-445) on the device; the result becomes the cycle's prune plan."""
        info = abi.PruneInfo()
        xyz = (C.c_double * 3)(*[float(v) for v in robot_xyz])
        self._ck(self.lib.b200lp_prune_plan(self.h, xyz, float(forward_distance), float(backward_distance), C.byref(info)))
        return info

    def read_prune_plan(self, n: int):
        """-> (prune_plan_.poses (n,7) float64, pcl_prune_plan_ (n,4) float32 x,y,z,intensity in its own order)."""
        poses = np.zeros((n, 7), np.float64)
        pcl = np.zeros((n, 4), np.float32)
        self._ck(self.lib.b200lp_read_prune_plan(self.h, poses.ctypes.data_as(C.POINTER(C.c_double)),
                                                 pcl.ctypes.data_as(C.POINTER(C.c_float)), n))
        return poses, pcl

    def path_blocked(self, check_radius: float) -> abi.Blocked:
        """perception_3d::PathBlockedStrategy::selfMark (path_blocked_strategy.cpp:56-100) of the device-side prune plan."""
        b = abi.Blocked()
        self._ck(self.lib.b200lp_path_blocked(self.h, float(check_radius), C.byref(b)))
        return b

    def sensor_observation(self, sensor: int, scan: np.ndarray, base_from_sensor, global_from_base, window: float,
                           marking_height: float, leaf: float = 0.1, is_local_planner: bool = True) -> abi.ObservationInfo:
        """perception_3d::MultiLayerSpinningLidar::cbSensor (multilayer_spinning_lidar.cpp:232-269) on one scan:
        transform -> pass-through -> voxel filter -> transform; the observation stays on the device.
        scan: (n,3|4|8) float32 in the sensor frame; the transforms are (x,y,z, qx,qy,qz,qw)."""
        pts, stride = _cloud_bytes(scan)
        sp = abi.SensorParams(float(window), float(marking_height), float(leaf), int(bool(is_local_planner)))
        b2s = (C.c_double * 7)(*[float(v) for v in base_from_sensor])
        g2b = (C.c_double * 7)(*[float(v) for v in global_from_base])
        info = abi.ObservationInfo()
        self._ck(self.lib.b200lp_sensor_observation(self.h, int(sensor), pts.ctypes.data_as(_P), pts.shape[0], stride, b2s, g2b,
                                                    C.byref(sp), C.byref(info)))
        return info

    def read_observation(self, sensor: int, n: int, stride: int = 16) -> np.ndarray:
        """Sensor::sensor_current_observation_ of `sensor`: (n,4) x,y,z,1 for stride 16, (n,8) PointXYZI rows for 32."""
        out = np.zeros((max(n, 1), stride // 4), np.float32)
        got = C.c_size_t()
        self._ck(self.lib.b200lp_read_observation(self.h, int(sensor), out.ctypes.data_as(_P), n, stride, C.byref(got)))
        return out[:got.value]

    def aggregate_observations(self, sensors) -> int:
        """StackedPerception::aggregateObservations (stacked_perception.cpp:128-140) on the device: the concatenation
        becomes the critics' cloud."""
        ids = (C.c_int32 * len(sensors))(*[int(v) for v in sensors])
        total = C.c_size_t()
        self._ck(self.lib.b200lp_aggregate_observations(self.h, ids, len(sensors), C.byref(total)))
        return total.value

    # -- cycle ----------------------------------------------------------------------------------
    def plan(self, q: abi.Query) -> abi.Result:
        r = abi.Result()
        self._ck(self.lib.b200lp_plan(self.h, C.byref(q), C.byref(r)))
        self.last, self._n_robots = r, 1
        self._results = [r]
        return r

    def plan_shard(self, q: abi.Query, rank: int, count: int) -> abi.Result:
        r = abi.Result()
        self._ck(self.lib.b200lp_plan_shard(self.h, C.byref(q), rank, count, C.byref(r)))
        self.last, self._n_robots = r, 1
        self._results = [r]
        return r

    # -- sample sharding with the argmin exchanged through peer device memory ---------
    def peer_export(self) -> bytes:
        h = (C.c_uint8 * abi.PEER_HANDLE_BYTES)()
        self._ck(self.lib.b200lp_peer_export(self.h, h))
        return bytes(h)

    def peer_attach(self, rank: int, handles) -> None:
        """handles: the peer_export() bytes of every rank, in rank order (this rank's own included)."""
        blob = b"".join(handles)
        arr = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self._ck(self.lib.b200lp_peer_attach(self.h, int(rank), len(handles), arr))

    def plan_shard_exchange(self, q: abi.Query) -> abi.Result:
        """plan_shard(rank, world) of the attached group + the in-kernel exchange: the GLOBAL best on every rank."""
        r = abi.Result()
        self._ck(self.lib.b200lp_plan_shard_exchange(self.h, C.byref(q), C.byref(r)))
        self.last, self._n_robots = r, 1
        self._results = [r]
        return r

    def peer_reserve_cloud(self, max_points: int) -> None:
        """Size the row buffer of set_cloud_shared; before peer_export."""
        self._ck(self.lib.b200lp_peer_reserve_cloud(self.h, int(max_points)))

    def set_cloud_shared(self, root: int, cloud=None, host_ptr=None, n=0, stride=0) -> None:
        """Collective over the attached peer group: the root passes the cloud (array, or host_ptr/n/stride), the others nothing."""
        if cloud is not None:
            cloud = np.ascontiguousarray(cloud, dtype=np.float32)
            self._keep_cloud = cloud
            host_ptr, n, stride = cloud.ctypes.data, cloud.shape[0], cloud.shape[1] * 4
        self._ck(self.lib.b200lp_set_cloud_shared(self.h, int(root), _P(host_ptr) if host_ptr else None, int(n), int(stride)))

    def peer_resync(self) -> None:
        """After a failed plan_shard_exchange: every rank calls this between two barriers; the exchange starts over."""
        self._ck(self.lib.b200lp_peer_resync(self.h))

    def set_shard_cuts(self, shares=None) -> None:
        """shares: rising floats from 0.0 to 1.0, one more than there are shards; None = equal shares."""
        if shares is None:
            self._ck(self.lib.b200lp_set_shard_cuts(self.h, None, 0))
            return
        arr = (C.c_float * len(shares))(*[float(v) for v in shares])
        self._ck(self.lib.b200lp_set_shard_cuts(self.h, arr, len(shares) - 1))

    def shard_cuts(self):
        arr, n = (C.c_float * (abi.MAX_PEERS + 1))(), C.c_int()
        self._ck(self.lib.b200lp_get_shard_cuts(self.h, arr, C.byref(n)))
        return [arr[k] for k in range(n.value + 1)] if n.value else None

    def set_adaptive_cuts(self, on: bool) -> None:
        self._ck(self.lib.b200lp_set_adaptive_cuts(self.h, 1 if on else 0))

    def last_cycle_ns(self) -> dict:
        """Device time of the last single-robot cycle (GPU global timer) and, after plan_shard_exchange, of every rank."""
        a, peers = C.c_uint32(), (C.c_uint32 * abi.MAX_PEERS)()
        self._ck(self.lib.b200lp_last_cycle_ns(self.h, C.byref(a), peers))
        return {"cycle_ns": a.value, "peer_ns": list(peers)}

    def plan_batch(self, queries, plans, plan_offsets):
        """queries: ctypes array of abi.Query; plans: (sum,7) float64; plan_offsets: (n+1,) int64.
        The ctypes views of the two arrays and the result block are kept between calls with the same array objects (a fleet
        re-plans with the same buffers every cycle; building them costs ~25 us per call, 3 % of a 512-robot step)."""
        n = len(queries)
        c = self._batch_cache
        if c is None or c[0] is not plans or c[1] is not plan_offsets or c[2] != n:
            p = np.ascontiguousarray(plans, dtype=np.float64).reshape(-1, 7)
            o = np.ascontiguousarray(plan_offsets, dtype=np.int64)
            assert o.shape == (n + 1,)
            c = (plans, plan_offsets, n, p, o, p.ctypes.data_as(C.POINTER(C.c_double)), o.ctypes.data_as(C.POINTER(C.c_int64)))
            # (kept only when the views ARE the caller's buffers: a converted copy would go stale when the caller refills its array)
            direct = all(isinstance(a, np.ndarray) and a.flags.c_contiguous and a.dtype == d
                         for a, d in ((plans, np.float64), (plan_offsets, np.int64)))
            self._batch_cache = c if direct else None
        res = (abi.Result * n)()
        self._ck(self.lib.b200lp_plan_batch(self.h, queries, n, c[5], c[6], res))
        self._n_robots = n
        self._results = res
        self.last = res[0]
        return res

    # -- read-back ------------------------------------------------------------------------------
    def read_trajectories(self, robot: int = 0) -> dict:
        n = self.traj_count(robot)[0]  # arrays span the robot's whole id space, also after a sharded run
        nc = max(1, self.n_critics)
        d = {
            "sample_index": np.zeros(n, np.int32), "vel": np.zeros((n, 3), np.float32),
            "num_steps": np.zeros(n, np.int32), "time_delta": np.zeros(n, np.float64),
            "cost": np.zeros(n, np.float64), "critic_scores": np.zeros((n, nc), np.float64),
            "first_hit_pose": np.zeros(n, np.int32),
        }
        if self.n_critics == 0:
            d["critic_scores"] = np.zeros((n, 0), np.float64)
        v = abi.TrajView(*[d[k].ctypes.data_as(t) if d[k].size else None for k, t in abi.TrajView._fields_])
        self._ck(self.lib.b200lp_read_trajectories(self.h, robot, C.byref(v)))
        return d

    def traj_count(self, robot: int = 0):
        """-> (n_traj_global, t_begin, t_end) of the last cycle."""
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self.lib.b200lp_traj_count(self.h, robot, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def read_poses(self, traj_id: int, num_steps: int, robot: int = 0) -> dict:
        n = num_steps
        d = {
            "pose": np.zeros((n, 7), np.float64), "pcl_pose": np.zeros((n, 3), np.float32),
            "cuboid": np.zeros((n, 8, 3), np.float32), "aabb": np.zeros((n, 6), np.float32),
            "collide": np.zeros(n, np.uint8), "n_r1": np.zeros(n, np.int32),
        }
        v = abi.PoseView(*[d[k].ctypes.data_as(t) for k, t in abi.PoseView._fields_])
        self._ck(self.lib.b200lp_read_poses(self.h, robot, traj_id, C.byref(v)))
        return d

    def read_pose_batch(self, t_begin: int = 0, t_end: int | None = None, robot: int = 0, fields=None) -> dict:
        """Per-pose quantities of trajectories [t_begin, t_end) in one launch; "offsets" maps a trajectory to its rows."""
        if t_end is None:
            t_end = self.traj_count(robot)[0]
        nt = t_end - t_begin
        offs = np.zeros(nt + 1, np.int64)
        empty = abi.PoseView()
        rc = self.lib.b200lp_read_pose_batch(self.h, robot, t_begin, t_end, offs.ctypes.data_as(C.POINTER(C.c_int64)),
                                             C.byref(empty), 0)
        n = int(offs[-1])
        if rc != 0 and n == 0:
            self._ck(rc)
        shapes = {"pose": ((n, 7), np.float64), "pcl_pose": ((n, 3), np.float32), "cuboid": ((n, 8, 3), np.float32),
                  "aabb": ((n, 6), np.float32), "collide": ((n,), np.uint8), "n_r1": ((n,), np.int32)}
        want = list(fields) if fields is not None else list(shapes)
        d = {k: np.zeros(*shapes[k]) for k in want}
        v = abi.PoseView(*[d[k].ctypes.data_as(t) if k in d and d[k].size else None for k, t in abi.PoseView._fields_])
        if n:
            self._ck(self.lib.b200lp_read_pose_batch(self.h, robot, t_begin, t_end, offs.ctypes.data_as(C.POINTER(C.c_int64)),
                                                     C.byref(v), n))
        d["offsets"] = offs
        return d

    def count_radius(self):
        s, n = C.c_int64(), C.c_int64()
        self._ck(self.lib.b200lp_count_radius(self.h, C.byref(s), C.byref(n)))
        return s.value, n.value

    def last_timing(self) -> dict:
        a, b, c, d = C.c_float(), C.c_float(), C.c_float(), C.c_float()
        self.lib.b200lp_last_timing(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return {"ms_upload": a.value, "ms_grid_build": b.value, "ms_plan_kernels": c.value, "ms_readback": d.value}

    def work_counters(self, reset: bool = False) -> dict:
        """Counting build only (LocalPlanner(..., lib_path=abi.COUNT_LIB_PATH))."""
        v = (C.c_uint64 * 4)()
        self._ck(self.lib.b200lp_work_counters(self.h, v, 1 if reset else 0))
        return {"pretests": v[0], "rounds": v[1], "exact_tests": v[2], "groups": v[3]}

    def last_kernel_ms(self) -> dict:
        if hasattr(self.lib, "b200lp_last_kernel_times"):
            v = (C.c_float * 4)()
            self.lib.b200lp_last_kernel_times(self.h, v, 4)
            return {"prep_kernel": v[0], "cull_kernel": v[1], "plan_kernel": v[2], "argmin_kernel": v[3]}
        a, b, c = C.c_float(), C.c_float(), C.c_float()  # (A/B builds of earlier rounds)
        self.lib.b200lp_last_kernel_ms(self.h, C.byref(a), C.byref(b), C.byref(c))
        return {"prep_kernel": a.value, "cull_kernel": 0.0, "plan_kernel": b.value, "argmin_kernel": c.value}

    def set_cloud_ptr(self, host_ptr: int, n: int, stride: int):
        """set_cloud from a raw host address (e.g. a pinned buffer)."""
        self._ck(self.lib.b200lp_set_cloud(self.h, _P(host_ptr), n, stride))

    def last_upload(self) -> dict:
        """Bytes the last set_cloud copied host -> device and the host threads that packed them (0: copied as is)."""
        b, t = C.c_size_t(), C.c_int32()
        self._ck(self.lib.b200lp_last_upload(self.h, C.byref(b), C.byref(t)))
        return {"h2d_bytes": b.value, "pack_threads": t.value}

    def launch_count(self) -> int:
        return int(self.lib.b200lp_launch_count(self.h))

    def grid_info(self) -> dict:
        dims = (C.c_int32 * 3)()
        org = (C.c_float * 3)()
        cell = (C.c_float * 2)()
        kept = C.c_int64()
        self._ck(self.lib.b200lp_grid_info(self.h, dims, org, cell, C.byref(kept)))
        return {"dims": tuple(dims), "origin": tuple(org), "cell": tuple(cell), "n_points_kept": kept.value}

    def stream(self) -> int:
        return int(self.lib.b200lp_stream(self.h) or 0)


class Local_Planner:
    """The reference's Local_Planner surface for the hot path (local_planner.h:72-82).

    The caller supplies what the reference pulls from ROS at the top of computeVelocityCommand: the robot pose
    (tf map->base_link), the odometry twist, the aggregated observation cloud and the prune plan.
    """

    def __init__(self, config: PlannerConfig | None = None, device: int = 0):
        self.planner = LocalPlanner(config, device)
        self._have_plan = False
        self.robot_pose = None
        self.robot_twist = None
        self.current_allowed_max_linear_speed_ = -1.0
        self.heading_deviation_ = 0.0

    def setPlan(self, prune_plan: np.ndarray):
        """prune plan, (n,7) position + orientation xyzw — prunePlan()'s output (local_planner.cpp:374-445)."""
        self.planner.set_plan(prune_plan)
        self._have_plan = True

    def setObservation(self, cloud: np.ndarray):
        """aggregate_observation_ (perception_3d/src/stacked_perception.cpp:128-140)."""
        self.planner.set_cloud(cloud)

    def setRobotState(self, pose7, twist3):
        self.robot_pose = [float(v) for v in pose7]
        self.robot_twist = [float(v) for v in twist3]

    def computeVelocityCommand(self, traj_gen_name: str = "differential_drive_simple"):
        """-> (PlannerState, Trajectory). Mirrors local_planner.cpp:482-621 for the states the path decides."""
        if self.robot_pose is None:
            return PlannerState.TF_FAIL, Trajectory()
        q = make_query(self.robot_pose, self.robot_twist, self.current_allowed_max_linear_speed_, self.heading_deviation_)
        r = self.planner.plan(q)
        best = Trajectory(r.xv, r.yv, r.thetav, r.best_cost, 0.0, r.best_id)
        if r.best_id < 0:
            return PlannerState.ALL_TRAJECTORIES_FAIL, best
        return PlannerState.TRAJECTORY_FOUND, best
