"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/b200lp.h declares, the ctypes
mirror matches the C struct layout, argument validation works, and — no CPU fallback — creation fails loudly
without a CUDA device."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from dddmr_navigation_b200 import PlannerConfig, abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200lp.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200lp_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    return abi.load_library()


def test_every_declared_symbol_is_exported_and_typed(lib):
    declared = _declared_symbols()
    assert len(declared) >= 18
    assert sorted(abi.SYMBOLS) == declared, (sorted(set(declared) ^ set(abi.SYMBOLS)))
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.b200lp_abi_version() == abi.ABI_VERSION


def test_ctypes_structs_match_the_c_layout(tmp_path):
    names = {"b200lp_limits": abi.Limits, "b200lp_params": abi.Params, "b200lp_critic": abi.Critic,
             "b200lp_grid_config": abi.GridConfig, "b200lp_query": abi.Query, "b200lp_result": abi.Result,
             "b200lp_traj_view": abi.TrajView, "b200lp_pose_view": abi.PoseView,
             "b200lp_prune_info": abi.PruneInfo, "b200lp_blocked": abi.Blocked,
             "b200lp_sensor_params": abi.SensorParams, "b200lp_observation_info": abi.ObservationInfo}
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, cls in names.items():
        prog.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            prog.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    prog.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run([os.environ.get("CC", "gcc"), "-std=c99", "-o", str(exe), str(src)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in names.items():
        assert int(out[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def _create(lib, cfg, device=0):
    L, P, cub = cfg.limits(), cfg.params(), cfg.cuboid()
    crit, n = cfg.critic_array()
    g = cfg.grid_config()
    h = C.c_void_p()
    rc = lib.b200lp_create(C.byref(h), device, C.byref(L), C.byref(P), cub.ctypes.data_as(C.POINTER(C.c_float)), crit, n, C.byref(g))
    return rc, h


def test_create_validates_parameters_before_touching_the_device(lib):
    cfg = PlannerConfig()
    cfg.generator["sim_time"] = 100.0     # 1.0 m/s * 100 s / 0.05 m = 2000 poses > B200LP_MAX_STEPS
    rc, h = _create(lib, cfg)
    assert rc == abi.E_INVALID and b"B200LP_MAX_STEPS" in lib.b200lp_last_error(None)
    cfg = PlannerConfig()
    cfg.generator["sim_granularity"] = 0.0
    assert _create(lib, cfg)[0] == abi.E_INVALID
    cfg = PlannerConfig()
    cfg.generator["linear_x_sample"] = 1e6
    assert _create(lib, cfg)[0] == abi.E_INVALID
    cfg = PlannerConfig()
    cfg.critics = cfg.critics * 3          # 12 critics > B200LP_MAX_CRITICS
    crit, n = cfg.critic_array()
    assert n == 12 and _create(lib, cfg)[0] == abi.E_INVALID
    assert lib.b200lp_create(None, 0, None, None, None, None, 0, None) == abi.E_INVALID


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    rc, h = _create(lib, PlannerConfig())
    assert rc == abi.E_CUDA and not h.value
    assert b"no CPU fallback" in lib.b200lp_last_error(None)
    from dddmr_navigation_b200 import LocalPlanner
    with pytest.raises(abi.B200LPError):
        LocalPlanner()


def test_null_ctx_calls_return_errors_not_crashes(lib):
    assert lib.b200lp_set_cloud(None, None, 0, 32) == abi.E_INVALID
    assert lib.b200lp_plan(None, None, None) == abi.E_INVALID
    assert lib.b200lp_launch_count(None) == 0
    assert lib.b200lp_stream(None) is None
    lib.b200lp_destroy(None)


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under dddmr_navigation_b200/ may reference it."""
    pkg = os.path.join(ROOT, "dddmr_navigation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "lporacle" not in text and "lp_oracle" not in text and "import oracle" not in text, os.path.join(dirpath, f)
    code = "import sys; import dddmr_navigation_b200; assert not any(m.startswith('oracle') for m in sys.modules), 'oracle imported'"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_config_from_reference_yaml_shape():
    """PlannerConfig.from_ros_yaml takes the reference's own YAML structure (p2p_move_base_localization.yaml:150-247)."""
    from dddmr_navigation_b200.config import DD_SIMPLE_DEFAULT
    params = {
        "trajectory_generators": {"ros__parameters": {"plugins": ["differential_drive_simple"],
                                                      "differential_drive_simple": dict(DD_SIMPLE_DEFAULT)}},
        "mpc_critics": {"ros__parameters": {
            "plugins": ["collision", "stick_path", "pure_pursuit", "toward_global_plan", "collision_rotate"],
            "collision": {"plugin": "mpc_critics::CollisionModel", "trajectory_generator": "differential_drive_simple", "weight": 1.0},
            "stick_path": {"plugin": "mpc_critics::StickPathModel", "trajectory_generator": "differential_drive_simple", "weight": 0.1},
            "pure_pursuit": {"plugin": "mpc_critics::PurePursuitModel", "trajectory_generator": "differential_drive_simple",
                             "translation_weight": 1.0, "orientation_weight": 0.01},
            "toward_global_plan": {"plugin": "mpc_critics::TowardGlobalPlanModel", "trajectory_generator": "differential_drive_simple", "weight": 1.0},
            "collision_rotate": {"plugin": "mpc_critics::CollisionModel", "trajectory_generator": "differential_drive_rotate_inplace", "weight": 1.0},
        }},
    }
    cfg = PlannerConfig.from_ros_yaml(params, "differential_drive_simple")
    arr, n = cfg.critic_array()
    assert n == 4 and [arr[i].kind for i in range(n)] == [abi.CRITIC_COLLISION, abi.CRITIC_STICK_PATH, abi.CRITIC_PURE_PURSUIT,
                                                          abi.CRITIC_TOWARD_GLOBAL_PLAN]
    assert arr[2].translation_weight == 1.0 and arr[2].orientation_weight == 0.01
    assert cfg.params().theory == abi.THEORY_DD_SIMPLE and cfg.limits().max_vel_x == 1.0
    assert np.allclose(cfg.cuboid()[0], [-0.35, 0.36, 0.0])  # blb first (dd_simple…cpp:211)


def test_host_packing_pool_against_a_scalar_reference(tmp_path):
    """The host side of the packing upload (csrc/lp_hostpack.h) needs no GPU: tests/cpp/pack_check.cpp runs the thread pool
    over strides 16-48, sizes 0-300 001, 1-8 threads and clouds with NaN / inf points (and NaN padding) and compares the packed
    rows, the chunk hand-over and the bounds with a scalar loop."""
    import shutil
    cxx = shutil.which(os.environ.get("CXX", "g++"))
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not cxx or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    exe = tmp_path / "pack_check"
    subprocess.run([cxx, "-O2", "-std=c++17", "-pthread", "-I" + os.path.join(cuda, "include"), "-o", str(exe),
                    os.path.join(ROOT, "tests", "cpp", "pack_check.cpp"), "-L" + os.path.join(cuda, "lib64"), "-lcudart",
                    "-Wl,-rpath," + os.path.join(cuda, "lib64")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "0 failed" in out.stdout
