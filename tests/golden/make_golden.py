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


def frontend_small():
    """pre-step (de-skew + voxel filter) and mapping association (N4)"""
    sys.path.insert(0, os.path.join(os.path.dirname(HERE)))
    from test_oracle_mapping import _pose, _scene
    scene = synth.Scene.make(synth.BASE_SEED + 3)
    traj = synth.trajectory(12, seed=11)
    rng = np.random.default_rng(77)
    scan = synth.lidar_scan(scene, traj[4], n_azimuth=400, seed=301)
    scan[:, 3] = rng.integers(0, 16, len(scan)) + np.float32(0.1) * rng.uniform(0, 1, len(scan)).astype(np.float32)
    q = np.array([0.99995, 0.001, -0.002, 0.009])
    q /= np.linalg.norm(q)
    t = np.array([0.12, 0.01, -0.004])
    desk = orc.deskew(scan, q, t)
    corner, surf, r2 = _scene(21)
    rot, qm, tm = _pose(r2)
    fe = np.zeros((300, 4), np.float32)
    fe[:, :3] = rot.inv().apply(corner[r2.choice(len(corner), 300), :3] + r2.normal(0, 0.05, (300, 3)) - tm)
    fp = np.zeros((600, 4), np.float32)
    fp[:, :3] = rot.inv().apply(surf[r2.choice(len(surf), 600), :3] + r2.normal(0, 0.03, (600, 3)) - tm)
    ev, ea, eb = orc.assoc_edges(corner, fe, qm, tm)
    pv, pn, pd = orc.assoc_planes(surf, fp, qm, tm)
    np.savez_compressed(os.path.join(HERE, "frontend_small.npz"), scan=scan, q=q, t=t, deskewed=desk, vg02=orc.voxel_grid(desk, 0.2), vg03=orc.voxel_grid(scan, 0.3),
                        corner=corner, surf=surf, qm=qm, tm=tm, edge_feats=fe, plane_feats=fp, edge_valid=ev, edge_a=ea, edge_b=eb,
                        plane_valid=pv, plane_norm=pn, plane_d=pd)


def vgicp_small():
    """FastVGICP on the clouds of gicp_small.npz: voxel map, one linearize, one align"""
    z = np.load(os.path.join(HERE, "gicp_small.npz"))
    src, tgt = z["src"], z["tgt"]
    out = {}
    for name, search in (("d1", orc.DIRECT1), ("d7", orc.DIRECT7)):
        v = orc.FastVGICP(resolution=1.0, search_method=search)
        v.setInputTarget(tgt)
        v.setInputSource(src)
        e, H, b = v.linearize(z["T_lin"])
        ncorr = v.num_correspondences()
        coords, num, mean, cov = v.voxels()
        order = np.lexsort((coords[:, 2], coords[:, 1], coords[:, 0]))
        T = v.align()
        out.update({f"{name}_err": e, f"{name}_H": H, f"{name}_b": b, f"{name}_ncorr": ncorr, f"{name}_T": T,
                    f"{name}_iterations": v.last["iterations"]})
        if name == "d1":
            out.update(vox_coords=coords[order], vox_num=num[order], vox_mean=mean[order], vox_cov=cov[order])
    np.savez_compressed(os.path.join(HERE, "vgicp_small.npz"), **out)


if __name__ == "__main__":
    gicp_small()
    features_small()
    frontend_small()
    vgicp_small()
    print("golden vectors written to", HERE)
