"""The C++ drop-in (lidar_odometry_b200/shim/b2lo_dropin.h): compiles against stand-ins of the reference's util/ types and
links libb2lo.so (CPU); on the GPU box the Estimator-style call sequence runs through it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "lidar_odometry_b200", "shim")
EXE = os.path.join(SHIM, "test", "dropin_smoke")


def _build():
    from lidar_odometry_b200 import capi
    capi.lib()
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + SHIM, os.path.join(SHIM, "test", "dropin_smoke.cpp"),
           "-o", EXE, "-L" + os.path.join(ROOT, "lidar_odometry_b200"), "-lb2lo", "-Wl,-rpath," + os.path.join(ROOT, "lidar_odometry_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_dropin_compiles_and_links():
    exe = _build()
    assert os.path.exists(exe)
    # without a GPU the first call fails loudly (no CPU fallback) instead of computing anything
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_dropin_runs_estimator_call_sequence():
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "DROPIN PASS" in r.stdout, r.stdout + r.stderr
