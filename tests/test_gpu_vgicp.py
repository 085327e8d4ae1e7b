"""GPU parity of the voxelised path (fast_gicp::FastVGICP) against the CPU oracle."""
import numpy as np
import pytest

from conftest import rot_angle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 2])
def test_voxel_map_matches_oracle(small_pair, mode):
    import rgc_slam_b200 as rgc
    from oracle import oracle as orc
    src, tgt, _ = small_pair
    g = rgc.FastVGICP()
    g.setVoxelAccumulationMode(mode)
    g.setInputTarget(tgt)
    g.setInputSource(src)
    coords, num, mean, cov = g.voxels()
    o = orc.FastVGICP(voxel_mode=mode)
    o.setInputTarget(tgt)
    o.setInputSource(src)
    oc, on, om, ov = o.voxels()
    assert np.array_equal(coords, oc) and np.array_equal(num, on)      # voxel keys and populations: bit-exact
    assert np.abs(mean - om).max() <= 1e-9 * max(1.0, np.abs(om).max())
    assert np.abs(cov - ov).max() <= 1e-8 * np.abs(ov).max()
    if mode == 0:  # additive sums: same terms, different (fixed) order
        assert np.abs(mean - om).max() <= 1e-12 * max(1.0, np.abs(om).max())


@pytest.mark.parametrize("search", [2, 1, 0])
@pytest.mark.parametrize("mode", [0, 2])
def test_vgicp_linearize_and_compute_error(scan_pair, search, mode):
    import rgc_slam_b200 as rgc
    from oracle import oracle as orc
    src, tgt, _ = scan_pair
    g = rgc.FastVGICP()
    g.setNeighborSearchMethod(search)
    g.setVoxelAccumulationMode(mode)
    g.setInputTarget(tgt)
    g.setInputSource(src)
    o = orc.FastVGICP(search_method=search, voxel_mode=mode)
    o.setInputTarget(tgt)
    o.setInputSource(src)
    T = np.eye(4)
    T[:3, 3] = [0.1, -0.05, 0.02]
    c, s = np.cos(0.01), np.sin(0.01)
    T[:3, :3] = [[c, -s, 0], [s, c, 0], [0, 0, 1]]
    e, H, b = g.linearize(T)
    oe, oH, ob = o.linearize(T)
    assert g.last_inliers() == o.num_correspondences()
    assert abs(e - oe) <= 1e-9 * abs(oe)
    assert np.abs(H - oH).max() <= 1e-9 * np.abs(oH).max()
    assert np.abs(b - ob).max() <= 1e-9 * np.abs(ob).max()
    T2 = T.copy()
    T2[:3, 3] += [0.02, 0.01, -0.01]
    assert abs(g.compute_error(T2) - o.compute_error(T2)) <= 1e-9 * abs(oe)
    e2, H2, b2 = g.linearize(T)
    assert e2 == e and np.array_equal(H2, H) and np.array_equal(b2, b)   # deterministic


@pytest.mark.parametrize("search", [2, 1, 0])
def test_vgicp_align_call_site(scan_pair, search):
    """RGC_odometer.cpp:998-1011: resolution 1, 25 iterations, trans eps 1e-6."""
    import rgc_slam_b200 as rgc
    from oracle import oracle as orc
    src, tgt, Ttrue = scan_pair
    g = rgc.FastVGICP()
    g.setResolution(1.0)
    g.setNeighborSearchMethod(search)
    g.setMaximumIterations(25)
    g.setMaxCorrespondenceDistance(2.0)
    g.setTransformationEpsilon(1e-6)
    g.setInputTarget(tgt)
    g.setInputSource(src)
    T = g.align()
    o = orc.FastVGICP(resolution=1.0, search_method=search, max_iterations=25, transformation_epsilon=1e-6)
    o.setInputTarget(tgt)
    o.setInputSource(src)
    To = o.align()
    assert np.abs(T[:3, 3].astype(np.float64) - To[:3, 3]).max() < 1e-4 and rot_angle(T[:3, :3], To[:3, :3]) < 1e-5
    assert g.last_result["iterations"] == o.last["iterations"] and g.hasConverged() == o.last["converged"]
    assert np.abs(T[:3, 3] - Ttrue[:3, 3]).max() < 0.06
    with pytest.raises(rgc.RgcError):
        g.correspondences()


def test_vgicp_against_golden_fixture():
    """the committed fixture tests/golden/vgicp_small.npz (oracle outputs frozen by make_golden.py)"""
    import os
    import rgc_slam_b200 as rgc
    gold = os.path.join(os.path.dirname(__file__), "golden")
    z, v = np.load(os.path.join(gold, "gicp_small.npz")), np.load(os.path.join(gold, "vgicp_small.npz"))
    for name, search in (("d1", rgc.DIRECT1), ("d7", rgc.DIRECT7)):
        g = rgc.FastVGICP()
        g.setNeighborSearchMethod(search)
        g.setInputTarget(z["tgt"])
        g.setInputSource(z["src"])
        e, H, b = g.linearize(z["T_lin"])
        assert g.last_inliers() == int(v[f"{name}_ncorr"])
        assert abs(e - v[f"{name}_err"]) <= 1e-9 * abs(v[f"{name}_err"])
        assert np.abs(H - v[f"{name}_H"]).max() <= 1e-9 * np.abs(v[f"{name}_H"]).max()
        assert np.abs(b - v[f"{name}_b"]).max() <= 1e-9 * np.abs(v[f"{name}_b"]).max()
        if name == "d1":
            coords, num, mean, cov = g.voxels()
            assert np.array_equal(coords, v["vox_coords"]) and np.array_equal(num, v["vox_num"])
            assert np.abs(mean - v["vox_mean"]).max() <= 1e-9 and np.abs(cov - v["vox_cov"]).max() <= 1e-8 * np.abs(v["vox_cov"]).max()
        T = g.align()
        To = v[f"{name}_T"]
        assert np.abs(T[:3, 3].astype(np.float64) - To[:3, 3]).max() < 1e-4 and rot_angle(T[:3, :3], To[:3, :3]) < 1e-5
        assert g.last_result["iterations"] == int(v[f"{name}_iterations"])
