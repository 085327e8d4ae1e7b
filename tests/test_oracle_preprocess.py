"""CPU tests of the pre-step oracle (oracle/orc_preprocess.hpp) against independent numpy / scipy
restatements: pcl::VoxelGrid centroids and RGC_odometer::adjustDistortion."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation as R

from oracle import oracle as orc


def _cloud(n, seed, spread=6.0):
    rng = np.random.default_rng(seed)
    P = np.zeros((n, 4), np.float32)
    P[:, :3] = rng.normal(0, spread, (n, 3))
    P[:, 3] = rng.integers(0, 16, n) + np.float32(0.1) * rng.uniform(0, 1, n).astype(np.float32)
    return P


def numpy_voxel_grid(P, leaf):
    """independent restatement: same float index arithmetic, sequential float32 sums in input order"""
    inv = np.float32(1.0) / np.float32(leaf)
    mn, mx = P[:, :3].min(0), P[:, :3].max(0)
    min_b = np.floor(mn * inv).astype(np.int64)
    div_b = np.floor(mx * inv).astype(np.int64) - min_b + 1
    ijk = (np.floor(P[:, :3] * inv) - min_b.astype(np.float32)).astype(np.int64)
    idx = ijk[:, 0] + ijk[:, 1] * div_b[0] + ijk[:, 2] * div_b[0] * div_b[1]
    order = np.argsort(idx, kind="stable")
    out = []
    s = 0
    while s < len(order):
        e = s
        acc = np.zeros(4, np.float32)
        while e < len(order) and idx[order[e]] == idx[order[s]]:
            acc = (acc + P[order[e]]).astype(np.float32)
            e += 1
        out.append(acc / np.float32(e - s))
        s = e
    return np.array(out, np.float32)


@pytest.mark.parametrize("leaf", [0.2, 0.3, 1.0])
def test_voxel_grid_matches_numpy(leaf):
    P = _cloud(6000, 1)
    V = orc.voxel_grid(P, leaf)
    W = numpy_voxel_grid(P, leaf)
    assert len(V) == len(W) and len(V) < len(P)
    assert np.array_equal(V, W)


def test_voxel_grid_properties():
    P = _cloud(4000, 2, spread=3.0)
    V = orc.voxel_grid(P, 0.5)
    # every centroid lies inside the bounding box, total mass is preserved per voxel count
    assert (V[:, :3] >= P[:, :3].min(0) - 1e-5).all() and (V[:, :3] <= P[:, :3].max(0) + 1e-5).all()
    # idempotent on a cloud that already has one point per voxel when the grid origin is unchanged
    G = np.zeros((27, 4), np.float32)
    G[:, :3] = np.stack(np.meshgrid(*[np.arange(3)] * 3, indexing="ij"), -1).reshape(-1, 3) + 0.5
    assert len(orc.voxel_grid(G, 1.0)) == 27
    assert np.array_equal(np.sort(orc.voxel_grid(G, 1.0), axis=0), np.sort(G, axis=0))
    # a single point and coincident points
    one = np.array([[1.0, 2.0, 3.0, 7.0]], np.float32)
    assert np.array_equal(orc.voxel_grid(one, 0.2), one)
    assert np.array_equal(orc.voxel_grid(np.repeat(one, 5, 0), 0.2), one)
    # leaf so small that the int32 index would overflow: PCL returns the input unchanged
    far = np.array([[0, 0, 0, 0], [3000, 3000, 3000, 1]], np.float32)
    assert np.array_equal(orc.voxel_grid(far, 0.001), far)


def test_deskew_matches_scipy():
    P = _cloud(3000, 3, spread=20.0)
    rot = R.from_rotvec([0.01, -0.02, 0.05])
    x, y, z, w = rot.as_quat()
    t = np.array([0.3, -0.05, 0.01])
    out = orc.deskew(P, [w, x, y, z], t)
    frac = P[:, 3] - np.trunc(P[:, 3]).astype(np.float32)
    s = (np.float32(1) - frac / np.float32(0.1)).astype(np.float64)
    rv = rot.inv().as_rotvec()
    exp = np.stack([R.from_rotvec(si * rv).apply(P[i, :3].astype(np.float64) - si * t) for i, si in enumerate(s)])
    assert np.abs(out[:, :3] - exp).max() < 1e-5          # float output at |p| ~ 60 m
    assert np.array_equal(out[:, 3], P[:, 3])
    # identity motion leaves the cloud untouched (the slerp takes its absD >= 1 - eps branch)
    same = orc.deskew(P, [1, 0, 0, 0], [0, 0, 0])
    assert np.array_equal(same, P)
    # a point stamped at the end of the sweep (s = 0) does not move, one at the start moves by the full inverse motion
    Q = np.array([[5, 1, 0.5, 3 + 0.1], [5, 1, 0.5, 3.0]], np.float32)
    o = orc.deskew(Q, [w, x, y, z], t)
    s0 = float(np.float32(1) - (Q[0, 3] - np.float32(3)) / np.float32(0.1))
    assert abs(s0) < 1e-5 and np.abs(o[0, :3] - Q[0, :3]).max() < 1e-4
    assert np.abs(o[1, :3] - rot.inv().apply(Q[1, :3].astype(np.float64) - t)).max() < 1e-5


def test_frontend_golden_vectors():
    """Frozen oracle outputs (tests/golden/make_golden.py: frontend_small): protects the oracle from drift."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "frontend_small.npz"))
    assert np.array_equal(orc.deskew(g["scan"], g["q"], g["t"]), g["deskewed"])
    assert np.array_equal(orc.voxel_grid(g["deskewed"], 0.2), g["vg02"])
    assert np.array_equal(orc.voxel_grid(g["scan"], 0.3), g["vg03"])
    ev, ea, eb = orc.assoc_edges(g["corner"], g["edge_feats"], g["qm"], g["tm"])
    assert np.array_equal(ev, g["edge_valid"]) and np.array_equal(ea[ev], g["edge_a"][ev]) and np.array_equal(eb[ev], g["edge_b"][ev])
    pv, pn, pd = orc.assoc_planes(g["surf"], g["plane_feats"], g["qm"], g["tm"])
    assert np.array_equal(pv, g["plane_valid"]) and np.array_equal(pn[pv], g["plane_norm"][pv]) and np.array_equal(pd[pv], g["plane_d"][pv])
    assert ev.sum() > 20 and pv.sum() > 100
