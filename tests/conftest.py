import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests skip (instead of failing) on a box without a CUDA device, whatever fixtures they use."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure).  Built on demand with g++."""
    from oracle import orc as _orc
    _orc.build()
    return _orc


@pytest.fixture(scope="session")
def b2():
    """The product: host mirror over libb2lo.so.  Skips when no CUDA device is present."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lidar_odometry_b200 import api
    return api


@pytest.fixture(scope="session")
def small_kitti():
    """6 KITTI-shaped scans at reduced angular resolution (64 rings x 600 azimuth steps)."""
    from lidar_odometry_b200 import synth
    scans, poses = synth.kitti_sequence(n_scans=6, seed=7, n_rings=64, n_az=600)
    return scans, poses


@pytest.fixture(scope="session")
def small_mid360():
    from lidar_odometry_b200 import synth
    scans, poses = synth.mid360_sequence(n_scans=5, seed=11, n_pts=6000)
    return scans, poses
