"""GPU parity of the A-LOAM feature path against the CPU oracle: labels, masks and feature lists
bit-exact (north_star), per-point floats bit-exact, ground plane to 1e-9."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EXACT = ["src_index", "intensity_num", "range_vec", "scan_angle", "curvature", "inten_curvature", "curvature2", "distance_source", "other_source",
         "label", "inten_label", "neighbor_picked", "inten_neighbor_picked", "ground_marked",
         "corner_sharp", "corner_less_sharp", "surf_flat", "inten_sharp", "inten_less_sharp", "corner_sharp_w", "surf_flat_w", "inten_sharp_w",
         # surfPointsLessFlatScan (:586-592) and GroundPoints (:338: push order, duplicates included)
         "surf_less_flat", "ground_points"]


def _compare(g, o, tag):
    assert g["cloud_size"] == o["cloud_size"], tag
    assert np.array_equal(g["scan_start"], o["scan_start"]) and np.array_equal(g["scan_end"], o["scan_end"]), tag
    assert np.array_equal(g["cloud"][:, :3], o["cloud"][:, :3]), tag
    # intensity = scanID + 0.1 * relTime: atan2 comes from libm on the CPU and from CUDA on the GPU
    # (at most one float ulp: 1.9e-6 on rings 16-31, 3.8e-6 on rings 32-63)
    assert (np.abs(g["cloud"][:, 3] - o["cloud"][:, 3]) <= np.spacing(np.maximum(np.abs(o["cloud"][:, 3]), np.float32(1.0)))).all(), tag
    assert np.array_equal(np.floor(g["cloud"][:, 3]), np.floor(o["cloud"][:, 3])), tag
    for k in EXACT:
        assert np.array_equal(g[k], o[k]), f"{tag}: {k} differs ({(np.asarray(g[k]) != np.asarray(o[k])).sum()} entries)"
    assert g["ground_size"] == o["ground_size"] and g["inten_merged"] == o["inten_merged"], tag
    gp, op = g["groundparam"], o["groundparam"]
    assert np.abs(gp[:3] - op[:3]).max() < 1e-9 and abs(gp[9] - op[9]) < 1e-9 and abs(gp[10] - op[10]) < 1e-9, tag
    for a in (3, 6):  # in-plane eigenvectors: sign is implementation defined in Eigen
        assert min(np.abs(gp[a:a + 3] - op[a:a + 3]).max(), np.abs(gp[a:a + 3] + op[a:a + 3]).max()) < 1e-6, tag


@pytest.mark.parametrize("beams,az", [(16, 1800), (32, 900)])
def test_features_match_oracle_on_a_batch(scene, traj, beams, az):
    from oracle import oracle as orc
    from rgc_slam_b200 import synth
    from rgc_slam_b200.features import extract_features
    scans = [synth.lidar_scan(scene, traj[f], n_beams=beams, n_azimuth=az, seed=700 + f) for f in (3, 9, 17, 25)]
    # a near obstacle so the r < 2 m incidence-angle / intensity-smoothing branch runs
    near = scans[0].copy()
    sel = (near[:, 0] > 0) & (np.abs(near[:, 1]) < 0.17 * near[:, 0])
    rng = np.random.default_rng(5)
    near[sel, :3] *= ((1.2 + 0.5 * rng.random(sel.sum())) / np.maximum(np.linalg.norm(near[sel, :3], axis=1), 1e-3))[:, None]
    scans.append(near)
    scans.append(scans[1][:137])  # ragged: a scan with almost nothing in it
    res, ms = extract_features(scans, n_rings=beams)
    assert ms > 0
    picked = 0
    for b, s in enumerate(scans):
        o = orc.extract_features(s, n_scans=beams)
        _compare(res[b], o, f"scan {b}")
        picked += len(o["corner_sharp"]) + len(o["surf_flat"])
    assert picked > 1000  # the comparison is not vacuous
    o0 = orc.extract_features(scans[0], n_scans=beams)
    assert len(o0["surf_less_flat"]) > 10000
    if beams == 16:  # (the 32-beam elevations put no ring at the ground-ring ranges: its GroundPoints cloud is empty)
        assert len(o0["ground_points"]) > len(np.unique(o0["ground_points"])) > 100  # duplicates exist
    # a ground list shorter than the cloud it describes is truncated, never overrun
    res2, _ = extract_features(scans[:2], n_rings=beams, want_arrays=False, ground_cap=100)
    for b in range(2):
        assert res2[b]["ground_size"] == res[b]["ground_size"] and np.array_equal(res2[b]["ground_points"], res[b]["ground_points"][:100])
    near_o = orc.extract_features(near, n_scans=beams)
    assert (near_o["scan_angle"] > 0).sum() > 10


def test_features_edge_cases():
    from oracle import oracle as orc
    from rgc_slam_b200.features import extract_features
    rng = np.random.default_rng(0)
    # everything outside the range gate / NaNs -> empty ordered cloud
    far = np.zeros((50, 4), np.float32)
    far[:, 0] = 500.0
    nan = np.full((20, 4), np.nan, np.float32)
    tiny = np.zeros((12, 4), np.float32)
    tiny[:, 0] = np.linspace(3, 4, 12)
    tiny[:, 2] = -0.5
    res, _ = extract_features([far, nan, tiny], n_rings=16)
    assert res[0]["cloud_size"] == 0 and res[1]["cloud_size"] == 0
    _compare(res[2], orc.extract_features(tiny, n_scans=16), "tiny")
    from rgc_slam_b200 import RgcError
    with pytest.raises(RgcError):
        extract_features([tiny], n_rings=20)
