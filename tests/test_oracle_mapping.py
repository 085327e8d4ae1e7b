"""CPU tests of the mapping-association oracle (oracle/orc_mapping.hpp) against numpy / scipy."""
import numpy as np
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation as R

from oracle import oracle as orc


def _scene(seed):
    """corner map = points along vertical poles, surface map = points on a few planes (+ noise)"""
    rng = np.random.default_rng(seed)
    poles = rng.uniform(-20, 20, (40, 2))
    corner = np.concatenate([np.c_[np.repeat(p[None], 60, 0) + rng.normal(0, 0.01, (60, 2)), rng.uniform(0, 3, 60)] for p in poles]).astype(np.float32)
    surf = []
    for _ in range(6):
        n = rng.normal(0, 1, 3)
        n /= np.linalg.norm(n)
        u = np.cross(n, [0.3, 0.5, 0.8])
        u /= np.linalg.norm(u)
        v = np.cross(n, u)
        o = rng.uniform(-10, 10, 3)
        ab = rng.uniform(-6, 6, (3000, 2))
        surf.append(o + ab[:, :1] * u + ab[:, 1:] * v + rng.normal(0, 0.01, (3000, 1)) * n)
    surf = np.concatenate(surf).astype(np.float32)
    corner = np.c_[corner, np.ones(len(corner), np.float32)]   # xyz1
    surf = np.c_[surf, np.ones(len(surf), np.float32)]
    return corner, surf, rng


def _pose(rng):
    rot = R.from_rotvec(rng.normal(0, 0.02, 3))
    x, y, z, w = rot.as_quat()
    return rot, np.array([w, x, y, z]), rng.normal(0, 0.05, 3)


def test_colpiv_qr_matches_lstsq():
    rng = np.random.default_rng(0)
    for _ in range(300):
        A = rng.normal(0, 1, (5, 3)) + rng.normal(0, 30, 3)
        x = orc.colpiv_qr_solve_5x3(A, -np.ones(5))
        xr = np.linalg.lstsq(A, -np.ones(5), rcond=None)[0]
        assert np.abs(x - xr).max() <= 1e-9 * max(1.0, np.abs(xr).max())


def test_edges_against_numpy():
    corner, _, rng = _scene(1)
    rot, q, t = _pose(rng)
    feats = np.zeros((400, 4), np.float32)
    feats[:, :3] = rot.inv().apply(corner[rng.choice(len(corner), 400), :3] + rng.normal(0, 0.05, (400, 3)) - t)
    feats[200:, :3] += rng.normal(0, 3, (200, 3))        # some far from any pole: d2[4] >= 1 or not a line
    valid, pa, pb = orc.assoc_edges(corner, feats, q, t)
    assert 50 < valid.sum() < 400
    sel = (rot.apply(feats[:, :3].astype(np.float64)) + t).astype(np.float32)
    tree = cKDTree(corner[:, :3].astype(np.float64))
    for i in range(len(feats)):
        d, idx = tree.query(sel[i].astype(np.float64), 5)
        P = corner[idx, :3].astype(np.float64)
        c = P.mean(0)
        w, V = np.linalg.eigh((P - c).T @ (P - c))
        exp_valid = (np.float32(d[4] ** 2) < 1.0) and (w[2] > 3 * w[1])
        if abs(d[4] ** 2 - 1.0) < 1e-4 or abs(w[2] - 3 * w[1]) < 1e-9 * w[2]:
            continue                                      # too close to a threshold to call
        assert bool(valid[i]) == bool(exp_valid), i
        if valid[i]:
            a, b = c + 0.1 * V[:, 2], c - 0.1 * V[:, 2]
            same = np.abs(pa[i] - a).max() < 1e-8 and np.abs(pb[i] - b).max() < 1e-8
            swapped = np.abs(pa[i] - b).max() < 1e-8 and np.abs(pb[i] - a).max() < 1e-8   # eigenvector sign is free
            assert same or swapped


def test_planes_against_numpy():
    _, surf, rng = _scene(2)
    rot, q, t = _pose(rng)
    feats = np.zeros((500, 4), np.float32)
    feats[:, :3] = rot.inv().apply(surf[rng.choice(len(surf), 500), :3] + rng.normal(0, 0.03, (500, 3)) - t)
    feats[300:, :3] += rng.normal(0, 4, (200, 3))
    valid, norm, dist = orc.assoc_planes(surf, feats, q, t)
    assert 100 < valid.sum() < 500
    sel = (rot.apply(feats[:, :3].astype(np.float64)) + t).astype(np.float32)
    tree = cKDTree(surf[:, :3].astype(np.float64))
    for i in range(len(feats)):
        d, idx = tree.query(sel[i].astype(np.float64), 5)
        if abs(d[4] ** 2 - 2.0) < 1e-4:
            continue
        P = surf[idx, :3].astype(np.float64)
        x = np.linalg.lstsq(P, -np.ones(5), rcond=None)[0]
        nn = np.linalg.norm(x)
        res = np.abs(P @ (x / nn) + 1 / nn)
        if np.abs(res - 0.2).min() < 1e-6:
            continue
        exp_valid = d[4] ** 2 < 2.0 and (res <= 0.2).all()
        assert bool(valid[i]) == bool(exp_valid), i
        if valid[i]:
            assert np.abs(norm[i] - x / nn).max() < 1e-7 and abs(dist[i] - 1 / nn) < 1e-7 * max(1.0, 1 / nn)
            assert abs(np.linalg.norm(norm[i]) - 1) < 1e-12
