"""CPU: properties of the feature oracle (restating scanRegistration.cpp:110-663) and its golden vectors."""
import os

import numpy as np

from oracle import oracle as orc
from rgc_slam_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_feature_oracle_invariants(scene, traj):
    scan = synth.lidar_scan(scene, traj[12], seed=31)
    f = orc.extract_features(scan, n_scans=16)
    m = f["cloud_size"]
    assert 0 < m <= len(scan)
    ring = np.floor(f["cloud"][:, 3]).astype(int)
    assert (np.diff(ring) >= 0).all() and ring.min() >= 0 and ring.max() <= 15       # ring-ordered
    for r in range(16):
        idx = f["src_index"][ring == r]
        assert (np.diff(idx) > 0).all()                                            # firing order kept inside a ring
    assert np.array_equal(f["cloud"][:, :3], scan[f["src_index"], :3])
    assert set(np.unique(f["label"])) <= {-1, 0, 1, 2}
    # per (ring, sextant) caps (:493-513, :546-556)
    for r in range(16):
        s, e = f["scan_start"][r], f["scan_end"][r]
        if e - s < 10:
            continue
        for j in range(6):
            sp, ep = s + (e - s) * j // 6, s + (e - s) * (j + 1) // 6 - 1
            lab = f["label"][sp:ep + 1]
            assert (lab == 2).sum() <= 20 and (lab == 1).sum() <= 1 and (lab == -1).sum() <= 40
    assert not (f["ground_marked"][f["corner_sharp"]] == 1).any()                  # ground is never an edge (:490)
    assert (f["curvature"][f["corner_sharp"]] > 0.1).all() and (f["curvature"][f["surf_flat"]] < 0.3).all()
    # flat scene: ground normal ~ -z (flipped toward the centroid, :374-377), distance ~ sensor height
    gp = f["groundparam"]
    assert abs(abs(gp[2]) - 1) < 1e-3 and abs(gp[9] - synth.SENSOR_HEIGHT) < 0.02


def test_feature_golden_vectors():
    z = np.load(os.path.join(GOLD, "features_small.npz"))
    for beams in (16, 32):
        f = orc.extract_features(z[f"scan{beams}"], n_scans=beams)
        for k in ("label", "inten_label", "neighbor_picked", "inten_neighbor_picked", "ground_marked", "curvature", "inten_curvature",
                  "curvature2", "corner_sharp", "surf_flat", "inten_sharp", "corner_less_sharp", "src_index"):
            assert np.array_equal(f[k], z[f"{k}{beams}"]), (beams, k)
        assert np.allclose(f["groundparam"], z[f"groundparam{beams}"], atol=1e-12)
