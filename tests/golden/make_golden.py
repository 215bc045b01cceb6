#!/usr/bin/env python
"""Regenerates the golden fixtures in this directory from the CPU oracle (shared math, brute-force index).

    python tests/golden/make_golden.py

The reference ships no golden vectors for this path and cannot be built here, so these fixtures pin
the ORACLE (and, through the GPU tests, the CUDA path) against regressions; they are not reference outputs.
Inputs are regenerated from seeds by dddmr_navigation_b200.synth; only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from dddmr_navigation_b200 import make_query, synth  # noqa: E402
from oracle import lporacle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    yield "playground", synth.playground()
    yield "c1_ramp_20k", synth.c1_ramp(n_points=20_000)


def run_case(sc):
    o = O.OraclePlanner(sc.config, O.MATH_SHARED, O.INDEX_BRUTE)
    o.set_cloud(sc.cloud)
    o.set_plan(sc.plan)
    r = o.plan(make_query(sc.pose, sc.twist))
    t = o.read_trajectories()
    out = {"result_" + k: np.asarray(v) for k, v in r.as_dict().items()}
    out.update({"traj_" + k: v for k, v in t.items()})
    ids = sorted(set(np.linspace(0, r.n_traj - 1, 5).astype(int).tolist()))
    out["pose_ids"] = np.asarray(ids)
    for i in ids:
        p = o.read_poses(i, int(t["num_steps"][i]))
        for k, v in p.items():
            out[f"pose{i}_{k}"] = v
    return out


OBSERVATION_CASES = {  # name -> (lidar_scan kwargs, window, marking_height, leaf, is_local_planner)
    "observation_16x512": (dict(n_beams=16, n_azimuth=512, seed=synth.SEED0 + 60), 10.0, 2.0, 0.1, True),
    "observation_24x700_base_frame": (dict(n_beams=24, n_azimuth=700, seed=synth.SEED0 + 61, room=(30.0, 20.0, 3.0), n_pillars=30),
                                      6.0, 1.2, 0.25, False),
}


def run_observation(name):
    """the restated cbSensor filter chain on a seeded synthetic scan (only the output is stored)."""
    kw, window, height, leaf, local = OBSERVATION_CASES[name]
    scan, b2s, g2b = synth.lidar_scan(**kw)
    info, obs = O.sensor_observation(scan, b2s, g2b, window, height, leaf, local)
    return {"n_scan": np.asarray(info.n_scan), "n_window": np.asarray(info.n_window), "n_points": np.asarray(info.n_points),
            "observation": obs}


if __name__ == "__main__":
    O.build()
    for name in OBSERVATION_CASES:
        out = run_observation(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, int(out["n_scan"]), int(out["n_window"]), int(out["n_points"]))
    for name, sc in cases():
        out = run_case(sc)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: out[k].tolist() for k in out if k.startswith("result_")})
