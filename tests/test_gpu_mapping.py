"""GPU parity of the mapping-node association (include/rgc_mapping.h) against the oracle:
validity flags identical, line end points / plane parameters within 1e-9 (both sides are fp64; the
device contracts a*b+c into FMAs and uses its own eigen-solver iteration order)."""
import numpy as np
import pytest

from test_oracle_mapping import _pose, _scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rgc():
    import rgc_slam_b200
    return rgc_slam_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _features(map_pts, rot, t, rng, n, noise, far):
    f = np.zeros((n, 12), np.float32)                    # pcl::PointXYZINormal records: 48 bytes
    f[:, :3] = rot.inv().apply(map_pts[rng.choice(len(map_pts), n), :3] + rng.normal(0, noise, (n, 3)) - t)
    f[n // 2:, :3] += rng.normal(0, far, (n - n // 2, 3))
    f[:, 4] = rng.uniform(0.5, 1.5, n)                   # normal_x = weight, untouched by the library
    return f


def test_edges_match_oracle(rgc, orc):
    corner, _, rng = _scene(11)
    rot, q, t = _pose(rng)
    feats = _features(corner, rot, t, rng, 3000, 0.05, 3.0)
    m = rgc.FeatureMap(corner)
    valid, pa, pb = m.associate_edges(feats, q, t)
    ov, oa, ob = orc.assoc_edges(corner, np.ascontiguousarray(feats[:, :4]), q, t)
    # a flag may only differ when the eigenvalue ratio sits on the 3x threshold to rounding
    assert (valid != ov).sum() <= 1 and 300 < valid.sum() < 3000
    both = valid & ov
    same = np.abs(pa[both] - oa[both]).max(1) < 1e-9
    swap = np.abs(pa[both] - ob[both]).max(1) < 1e-9     # eigenvector sign: a and b swap, the residual is the same
    assert (same | swap).all()
    mid_g, mid_o = 0.5 * (pa[both] + pb[both]), 0.5 * (oa[both] + ob[both])
    assert np.abs(mid_g - mid_o).max() < 1e-11
    assert np.abs(np.linalg.norm(pa[both] - pb[both], axis=1) - 0.2).max() < 1e-12
    # same map, another pose (the "last frame" loop of the reference)
    rot2, q2, t2 = _pose(rng)
    v2, _, _ = m.associate_edges(feats, q2, t2)
    o2, _, _ = orc.assoc_edges(corner, np.ascontiguousarray(feats[:, :4]), q2, t2)
    assert (v2 != o2).sum() <= 1
    m.close()


def test_planes_match_oracle(rgc, orc):
    _, surf, rng = _scene(12)
    rot, q, t = _pose(rng)
    feats = _features(surf, rot, t, rng, 5000, 0.03, 4.0)
    m = rgc.FeatureMap(surf)
    valid, norm, dist = m.associate_planes(feats, q, t)
    ov, on, od = orc.assoc_planes(surf, np.ascontiguousarray(feats[:, :4]), q, t)
    assert (valid != ov).sum() <= 1 and 1000 < valid.sum() < 5000
    both = valid & ov
    assert np.abs(norm[both] - on[both]).max() < 1e-9
    assert (np.abs(dist[both] - od[both]) <= 1e-9 * np.maximum(1.0, od[both])).all()
    assert np.abs(np.linalg.norm(norm[both], axis=1) - 1).max() < 1e-12
    m.close()


def test_small_maps_and_errors(rgc, orc):
    rng = np.random.default_rng(0)
    tiny = np.ones((4, 4), np.float32)
    tiny[:, :3] = rng.normal(0, 0.1, (4, 3))
    m = rgc.FeatureMap(tiny)                             # fewer than 5 map points: nothing can be valid
    f = np.zeros((10, 4), np.float32)
    v, _, _ = m.associate_edges(f, [1, 0, 0, 0], [0, 0, 0])
    assert not v.any()
    v, _, _ = m.associate_planes(f, [1, 0, 0, 0], [0, 0, 0])
    assert not v.any()
    m.close()
    with pytest.raises(rgc.RgcError):
        rgc.FeatureMap(np.zeros((0, 4), np.float32))
