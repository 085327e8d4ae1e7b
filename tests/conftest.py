import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run through gpurun); everything else runs on CPU")


@pytest.fixture(scope="session")
def scene():
    from rgc_slam_b200 import synth
    return synth.Scene.make(synth.BASE_SEED)


@pytest.fixture(scope="session")
def traj():
    from rgc_slam_b200 import synth
    return synth.trajectory(64, seed=1)


@pytest.fixture(scope="session")
def scan_pair(scene, traj):
    """Config C1: two consecutive full VLP-16 sweeps (xyz1 float32) and the true relative pose."""
    from rgc_slam_b200 import synth
    tgt = synth.to_xyz1(synth.lidar_scan(scene, traj[30], seed=synth.BASE_SEED + 1000 + 30))
    src = synth.to_xyz1(synth.lidar_scan(scene, traj[31], seed=synth.BASE_SEED + 1000 + 31))
    return src, tgt, synth.relative_pose(traj[31], traj[30])


@pytest.fixture(scope="session")
def small_pair(scene, traj):
    """Quarter-resolution sweeps (450 azimuth steps) for the quick cases."""
    from rgc_slam_b200 import synth
    tgt = synth.to_xyz1(synth.lidar_scan(scene, traj[10], n_azimuth=450, seed=5))
    src = synth.to_xyz1(synth.lidar_scan(scene, traj[11], n_azimuth=450, seed=6))
    return src, tgt, synth.relative_pose(traj[11], traj[10])


def rot_angle(Ra, Rb):
    """Angle of Ra^T Rb, from the skew part (well conditioned near zero, unlike acos(trace))."""
    R = Ra.astype(np.float64).T @ Rb.astype(np.float64)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
    return float(np.arctan2(np.linalg.norm(v), (np.trace(R) - 1) / 2))
