"""CPU check of the closed-form velocity axes a single-robot launch hands prep_kernel (csrc/lp_kernels.cuh: AxisPlan,
axis_value): tests/cpp/axis_check.cu compares them entry by entry, bit for bit, with the VelocityIterator chain
(trajectory_generators/velocity_iterator.h:44-69) over 60 000 windows x 15 sample counts, rebuilds every axis from closed
form + exception list the way the kernel does, and reports how often exceptions are needed. Host code, built with nvcc
because the header is CUDA."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_closed_form_axes_plus_exceptions_equal_the_chains(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("needs nvcc")
    exe = tmp_path / "axis_check"
    subprocess.run([nvcc, "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false", "-Xcompiler",
                    "-ffp-contract=off,-pthread", "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "axis_check.cu")], check=True,
                   capture_output=True, timeout=600)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "\n0 failed" in out.stdout
    # the statistics DESIGN.md quotes: exceptions are common, the fall-back to the kernel's own chains is rare
    line = out.stdout.strip().splitlines()[-2]
    words = line.split()
    windows, differing_windows, beyond = int(words[0]), int(words[7]), int(words[words.index("window:") + 2])
    assert differing_windows > windows // 50 and beyond < windows // 1000, line
