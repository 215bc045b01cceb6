"""ctypes wrapper of the CPU oracle (oracle/lp_oracle.cpp). TEST INFRASTRUCTURE — see oracle/lp_oracle.h.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The class mirrors dddmr_navigation_b200.planner.LocalPlanner method for method so parity tests read
symmetrically.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from dddmr_navigation_b200 import abi
from dddmr_navigation_b200.config import PlannerConfig

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liblporacle.so")
LIB_REF = os.path.join(HERE, "_ref", "liblporacle_ref.so")

MATH_SHARED, MATH_LIBM = 0, 1
INDEX_BRUTE, INDEX_GRID, INDEX_NANOFLANN = 0, 1, 2

_P = C.c_void_p
_SYMS = {
    "lporacle_has_nanoflann": (C.c_int, []),
    "lporacle_create": (C.c_int, [C.POINTER(_P), C.POINTER(abi.Limits), C.POINTER(abi.Params), C.POINTER(C.c_float),
                                  C.POINTER(abi.Critic), C.c_int, C.c_int, C.c_int]),
    "lporacle_destroy": (None, [_P]),
    "lporacle_last_error": (C.c_char_p, [_P]),
    "lporacle_set_cloud": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t]),
    "lporacle_set_plan": (C.c_int, [_P, C.POINTER(C.c_double), C.c_size_t]),
    "lporacle_set_sample_stride": (C.c_int, [_P, C.c_int, C.c_int]),
    "lporacle_set_keep_index": (C.c_int, [_P, C.c_int]),
    "lporacle_set_eigen_association": (C.c_int, [C.c_int]),
    "lporacle_plan": (C.c_int, [_P, C.POINTER(abi.Query), C.c_int, C.POINTER(abi.Result), C.POINTER(C.c_double),
                                C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "lporacle_read_trajectories": (C.c_int, [_P, C.POINTER(abi.TrajView)]),
    "lporacle_read_poses": (C.c_int, [_P, C.c_int32, C.POINTER(abi.PoseView)]),
    "lporacle_count_radius": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "lporacle_prune_plan": (C.c_int, [C.POINTER(C.c_double), C.c_size_t, C.POINTER(C.c_double), C.c_double, C.c_double,
                                      C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_size_t, C.POINTER(abi.PruneInfo)]),
    "lporacle_path_blocked": (C.c_int, [_P, C.POINTER(C.c_float), C.c_size_t, C.c_double, C.POINTER(abi.Blocked)]),
    "lporacle_sensor_observation": (C.c_int, [_P, C.c_size_t, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                              C.POINTER(abi.SensorParams), C.c_int, C.POINTER(C.c_float), C.c_size_t,
                                              C.POINTER(abi.ObservationInfo)]),
    "lporacle_samples": (C.c_int, [_P, C.POINTER(abi.Query), C.POINTER(C.c_float), C.c_int]),
    "lporacle_sinf": (C.c_float, [C.c_int, C.c_float]),
    "lporacle_cosf": (C.c_float, [C.c_int, C.c_float]),
    "lporacle_sin": (C.c_double, [C.c_int, C.c_double]),
    "lporacle_cos": (C.c_double, [C.c_int, C.c_double]),
    "lporacle_asin": (C.c_double, [C.c_int, C.c_double]),
    "lporacle_atan2": (C.c_double, [C.c_int, C.c_double, C.c_double]),
    "lporacle_fmod": (C.c_double, [C.c_int, C.c_double, C.c_double]),
    "lporacle_sincosf_mismatches": (C.c_int64, [C.c_uint32, C.c_uint32, C.c_uint32]),
}
_libs = {}


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present) with oracle/Makefile."""
    if force or not os.path.exists(LIB):
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s"], check=True, capture_output=True)
    else:
        subprocess.run(["make", "-C", HERE, "-s"], check=True, capture_output=True)


def load(ref: bool = False):
    key = "ref" if ref else "own"
    if key in _libs:
        return _libs[key]
    path = LIB_REF if ref else LIB
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    for name, (res, args) in _SYMS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _libs[key] = lib
    return lib


def have_ref() -> bool:
    return os.path.exists(LIB_REF)


class OraclePlanner:
    """CPU restatement of one generator + its critic stack (same surface as LocalPlanner)."""

    def __init__(self, config: PlannerConfig, math_mode: int = MATH_SHARED, index_mode: int = INDEX_GRID):
        self.lib = load(ref=(index_mode == INDEX_NANOFLANN))
        self.config = config
        self._L, self._Pm = config.limits(), config.params()
        self._cub = config.cuboid()
        self._crit, self.n_critics = config.critic_array()
        h = _P()
        rc = self.lib.lporacle_create(C.byref(h), C.byref(self._L), C.byref(self._Pm),
                                      self._cub.ctypes.data_as(C.POINTER(C.c_float)), self._crit, self.n_critics,
                                      math_mode, index_mode)
        if rc != 0:
            raise RuntimeError(f"lporacle_create failed: {rc}")
        self.h = h
        self.last = None
        self.timing = (0.0, 0.0, 0.0)

    def close(self):
        if self.h:
            self.lib.lporacle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_cloud(self, pts: np.ndarray):
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        assert pts.ndim == 2 and pts.shape[1] in (3, 4, 8)
        stride = pts.shape[1] * 4
        if pts.shape[1] == 3:  # repack to 16-byte stride
            p4 = np.zeros((pts.shape[0], 4), np.float32)
            p4[:, :3] = pts
            pts, stride = p4, 16
        rc = self.lib.lporacle_set_cloud(self.h, pts.ctypes.data_as(_P), pts.shape[0], stride)
        assert rc == 0

    def set_plan(self, plan: np.ndarray):
        plan = np.ascontiguousarray(plan, dtype=np.float64).reshape(-1, 7)
        rc = self.lib.lporacle_set_plan(self.h, plan.ctypes.data_as(C.POINTER(C.c_double)), plan.shape[0])
        assert rc == 0

    def set_sample_stride(self, stride: int, phase: int = 0):
        assert self.lib.lporacle_set_sample_stride(self.h, stride, phase) == 0

    def set_keep_index(self, keep: bool = True):
        """Do not rebuild the radius-search index on every plan() while the cloud stays the same (results are unaffected)."""
        assert self.lib.lporacle_set_keep_index(self.h, 1 if keep else 0) == 0

    def plan(self, q: abi.Query, n_threads: int = 1) -> abi.Result:
        r = abi.Result()
        t = [C.c_double(), C.c_double(), C.c_double()]
        rc = self.lib.lporacle_plan(self.h, C.byref(q), n_threads, C.byref(r), C.byref(t[0]), C.byref(t[1]),
                                    C.byref(t[2]))
        assert rc == 0
        self.last = r
        self.timing = tuple(x.value for x in t)
        return r

    def samples(self, q: abi.Query) -> np.ndarray:
        n = self.lib.lporacle_samples(self.h, C.byref(q), None, 0)
        out = np.zeros((max(n, 1), 3), np.float32)
        self.lib.lporacle_samples(self.h, C.byref(q), out.ctypes.data_as(C.POINTER(C.c_float)), n)
        return out[:n]

    def read_trajectories(self) -> dict:
        n = self.last.n_traj
        nc = max(1, self.n_critics)
        d = {
            "sample_index": np.zeros(n, np.int32), "vel": np.zeros((n, 3), np.float32),
            "num_steps": np.zeros(n, np.int32), "time_delta": np.zeros(n, np.float64),
            "cost": np.zeros(n, np.float64), "critic_scores": np.zeros((n, nc), np.float64),
            "first_hit_pose": np.zeros(n, np.int32),
        }
        v = abi.TrajView(*[d[k].ctypes.data_as(t) for k, t in abi.TrajView._fields_])
        assert self.lib.lporacle_read_trajectories(self.h, C.byref(v)) == 0
        d["critic_scores"] = d["critic_scores"][:, :self.n_critics]
        return d

    def read_poses(self, traj_id: int, num_steps: int) -> dict:
        n = num_steps
        d = {
            "pose": np.zeros((n, 7), np.float64), "pcl_pose": np.zeros((n, 3), np.float32),
            "cuboid": np.zeros((n, 8, 3), np.float32), "aabb": np.zeros((n, 6), np.float32),
            "collide": np.zeros(n, np.uint8), "n_r1": np.zeros(n, np.int32),
        }
        v = abi.PoseView(*[d[k].ctypes.data_as(t) for k, t in abi.PoseView._fields_])
        assert self.lib.lporacle_read_poses(self.h, traj_id, C.byref(v)) == 0
        return d

    def count_radius(self):
        s, n = C.c_int64(), C.c_int64()
        assert self.lib.lporacle_count_radius(self.h, C.byref(s), C.byref(n)) == 0
        return s.value, n.value

    def path_blocked(self, pcl_xyzi, check_radius: float) -> abi.Blocked:
        """PathBlockedStrategy::selfMark restated, against the cloud given to set_cloud."""
        pcl = np.ascontiguousarray(pcl_xyzi, np.float32).reshape(-1, 4)
        b = abi.Blocked()
        assert self.lib.lporacle_path_blocked(self.h, pcl.ctypes.data_as(C.POINTER(C.c_float)), pcl.shape[0],
                                              float(check_radius), C.byref(b)) == 0
        return b


def set_eigen_association(right: bool) -> None:
    """Process-wide switch of the restated Eigen products' 3-term association (see lp_oracle.h); default left."""
    assert load().lporacle_set_eigen_association(1 if right else 0) == 0
    if have_ref():
        assert load(ref=True).lporacle_set_eigen_association(1 if right else 0) == 0


def prune_plan(global_plan, robot_xyz, forward_distance, backward_distance, capacity=abi.MAX_PLAN):
    """Local_Planner::prunePlan restated (lp_oracle.cpp). -> (PruneInfo, poses (n,7), pcl (n,4))."""
    lib = load()
    g = np.ascontiguousarray(global_plan, np.float64).reshape(-1, 7)
    poses = np.zeros((capacity, 7), np.float64)
    pcl = np.zeros((capacity, 4), np.float32)
    info = abi.PruneInfo()
    xyz = (C.c_double * 3)(*[float(v) for v in robot_xyz])
    rc = lib.lporacle_prune_plan(g.ctypes.data_as(C.POINTER(C.c_double)), g.shape[0], xyz, float(forward_distance),
                                 float(backward_distance), poses.ctypes.data_as(C.POINTER(C.c_double)),
                                 pcl.ctypes.data_as(C.POINTER(C.c_float)), capacity, C.byref(info))
    if rc != 0:
        raise RuntimeError(f"lporacle_prune_plan: {rc}")
    n = info.n_prune if info.status == 0 else 0
    return info, poses[:n], pcl[:n]


def sensor_observation(scan, base_from_sensor, global_from_base, window, marking_height, leaf=0.1, is_local_planner=True,
                       order_mode=0):
    """MultiLayerSpinningLidar::cbSensor's filter chain restated (lp_oracle.cpp). -> (ObservationInfo, (n,4) float32)."""
    lib = load()
    pts = np.ascontiguousarray(scan, np.float32)
    if pts.ndim != 2 or pts.shape[1] not in (3, 4, 8):
        raise ValueError("scan must be (n,3|4|8) float32")
    sp = abi.SensorParams(float(window), float(marking_height), float(leaf), int(bool(is_local_planner)))
    b2s = (C.c_double * 7)(*[float(v) for v in base_from_sensor])
    g2b = (C.c_double * 7)(*[float(v) for v in global_from_base])
    out = np.zeros((max(pts.shape[0], 1), 4), np.float32)
    info = abi.ObservationInfo()
    rc = lib.lporacle_sensor_observation(pts.ctypes.data_as(_P), pts.shape[0], pts.shape[1] * 4, b2s, g2b, C.byref(sp),
                                         int(order_mode), out.ctypes.data_as(C.POINTER(C.c_float)), out.shape[0], C.byref(info))
    if rc != 0:
        raise RuntimeError(f"lporacle_sensor_observation: {rc}")
    return info, out[:info.n_points]


# ---------------------------------------------------------------------------------------------------------------------
# oracle/_ref/liblpref.so: the REFERENCE's own theory / critic / stacked-model sources compiled where they lie under
# /root/reference against the third-party stand-ins of oracle/ref_shims/ (oracle/Makefile target `lpref`). Used by
# tests/test_reference_sources.py to pin the restatement above to the reference's real control flow.
# ---------------------------------------------------------------------------------------------------------------------
LIB_LPREF = os.path.join(HERE, "_ref", "liblpref.so")
_LPREF_SYMS = {
    "lpref_last_error": (C.c_char_p, [_P]),
    "lpref_create": (C.c_int, [C.POINTER(_P), C.POINTER(abi.Limits), C.POINTER(abi.Params), C.POINTER(C.c_float),
                               C.POINTER(abi.Critic), C.c_int]),
    "lpref_destroy": (None, [_P]),
    "lpref_set_cloud": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t]),
    "lpref_set_plan": (C.c_int, [_P, C.POINTER(C.c_double), C.c_size_t]),
    "lpref_plan": (C.c_int, [_P, C.POINTER(abi.Query), C.POINTER(abi.Result)]),
    "lpref_read_trajectories": (C.c_int, [_P, C.POINTER(abi.TrajView)]),
    "lpref_read_poses": (C.c_int, [_P, C.c_int32, C.POINTER(abi.PoseView)]),
    "lpref_path_blocked": (C.c_int, [_P, C.c_size_t, C.c_size_t, C.POINTER(C.c_float), C.c_size_t, C.c_double, C.POINTER(C.c_int32)]),
}
_lpref = None


def have_reference_sources() -> bool:
    return os.path.exists(LIB_LPREF)


def load_lpref():
    global _lpref
    if _lpref is None:
        lib = C.CDLL(LIB_LPREF)
        for name, (res, args) in _LPREF_SYMS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lpref = lib
    return _lpref


class ReferencePlanner:
    """The reference's real C++ (DD/omni/rotate theories, seven critics, StackedGenerator, StackedScoringModel,
    base_trajectory::Trajectory) behind the same surface as OraclePlanner."""

    def __init__(self, config: PlannerConfig):
        self.lib = load_lpref()
        self.config = config
        self._L, self._Pm = config.limits(), config.params()
        self._cub = config.cuboid()
        self._crit, self.n_critics = config.critic_array()
        h = _P()
        rc = self.lib.lpref_create(C.byref(h), C.byref(self._L), C.byref(self._Pm), self._cub.ctypes.data_as(C.POINTER(C.c_float)),
                                   self._crit, self.n_critics)
        assert rc == 0, rc
        self.h = h
        self.last = None

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.lpref_destroy(self.h)
            self.h = None

    def set_cloud(self, pts):
        pts = np.ascontiguousarray(pts, np.float32)
        assert self.lib.lpref_set_cloud(self.h, pts.ctypes.data_as(_P), pts.shape[0], pts.shape[1] * 4) == 0

    def set_plan(self, plan):
        plan = np.ascontiguousarray(plan, np.float64).reshape(-1, 7)
        assert self.lib.lpref_set_plan(self.h, plan.ctypes.data_as(C.POINTER(C.c_double)), plan.shape[0]) == 0

    def plan(self, q: abi.Query) -> abi.Result:
        r = abi.Result()
        rc = self.lib.lpref_plan(self.h, C.byref(q), C.byref(r))
        assert rc == 0, (rc, self.lib.lpref_last_error(self.h))
        self.last = r
        return r

    def read_trajectories(self) -> dict:
        n, nc = self.last.n_traj, max(1, self.n_critics)
        d = {"sample_index": np.zeros(n, np.int32), "vel": np.zeros((n, 3), np.float32), "num_steps": np.zeros(n, np.int32),
             "time_delta": np.zeros(n, np.float64), "cost": np.zeros(n, np.float64),
             "critic_scores": np.zeros((n, nc if self.n_critics else 0), np.float64), "first_hit_pose": np.zeros(n, np.int32)}
        v = abi.TrajView(*[d[k].ctypes.data_as(t) if d[k].size else None for k, t in abi.TrajView._fields_])
        assert self.lib.lpref_read_trajectories(self.h, C.byref(v)) == 0
        return d

    def read_poses(self, traj_id: int, num_steps: int) -> dict:
        n = num_steps
        d = {"pose": np.zeros((n, 7), np.float64), "pcl_pose": np.zeros((n, 3), np.float32),
             "cuboid": np.zeros((n, 8, 3), np.float32), "aabb": np.zeros((n, 6), np.float32)}
        v = abi.PoseView(d["pose"].ctypes.data_as(C.POINTER(C.c_double)), d["pcl_pose"].ctypes.data_as(C.POINTER(C.c_float)),
                         d["cuboid"].ctypes.data_as(C.POINTER(C.c_float)), d["aabb"].ctypes.data_as(C.POINTER(C.c_float)), None, None)
        assert self.lib.lpref_read_poses(self.h, traj_id, C.byref(v)) == 0
        return d


def reference_path_blocked_opinion(cloud, pcl_xyzi, check_radius: float) -> int:
    """perception_3d::PathBlockedStrategy::selfMark — the reference's own code — on (cloud, pcl_prune_plan_): its opinion
    (0 PASS, 1 PATH_BLOCKED_WAIT)."""
    lib = load_lpref()
    cloud = np.ascontiguousarray(cloud, np.float32)
    pcl = np.ascontiguousarray(pcl_xyzi, np.float32).reshape(-1, 4)
    out = C.c_int32(-1)
    rc = lib.lpref_path_blocked(cloud.ctypes.data_as(_P), cloud.shape[0], cloud.shape[1] * 4, pcl.ctypes.data_as(C.POINTER(C.c_float)),
                                pcl.shape[0], float(check_radius), C.byref(out))
    assert rc == 0, rc
    return out.value
