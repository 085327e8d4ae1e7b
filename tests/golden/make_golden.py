"""Generates tests/golden/*.npz from the CPU oracle (the reference ships no golden vectors and is
not importable — SURVEY.md §8c).  Run once after the oracle's cross-checks pass:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402
from rgc_slam_b200 import synth  # noqa: E402


def gicp_small():
    scene = synth.Scene.make(synth.BASE_SEED + 1)
    traj = synth.trajectory(12, seed=9)
    tgt = synth.to_xyz1(synth.lidar_scan(scene, traj[5], n_azimuth=300, seed=101))
    src = synth.to_xyz1(synth.lidar_scan(scene, traj[6], n_azimuth=300, seed=102))
    idx, d2 = orc.knn(tgt, tgt, 20)
    o = orc.FastGICP()
    o.setInputTarget(tgt)
    o.setInputSource(src)
    T_lin = np.eye(4)
    T_lin[:3, 3] = [0.08, 0.01, -0.01]
    e, H, b = o.linearize(T_lin)
    corr, _ = o.correspondences()
    T = o.align()
    np.savez_compressed(os.path.join(HERE, "gicp_small.npz"), src=src, tgt=tgt, knn_idx=idx, knn_d2=d2, T_lin=T_lin, lin_err=e, lin_H=H,
                        lin_b=b, corr=corr, T_final=T, iterations=o.last["iterations"], covs_tgt=o.getTargetCovariances().astype(np.float64))


def features_small():
    scene = synth.Scene.make(synth.BASE_SEED + 2)
    traj = synth.trajectory(12, seed=10)
    out = {}
    for beams, az in ((16, 600), (32, 300)):
        scan = synth.lidar_scan(scene, traj[4], n_beams=beams, n_azimuth=az, seed=200 + beams)
        f = orc.extract_features(scan, n_scans=beams)
        out[f"scan{beams}"] = scan
        for k in ("label", "inten_label", "neighbor_picked", "inten_neighbor_picked", "ground_marked", "curvature", "inten_curvature",
                  "curvature2", "groundparam", "corner_sharp", "surf_flat", "inten_sharp", "corner_less_sharp", "src_index"):
            out[f"{k}{beams}"] = f[k]
    np.savez_compressed(os.path.join(HERE, "features_small.npz"), **out)


if __name__ == "__main__":
    gicp_small()
    features_small()
    print("golden vectors written to", HERE)
