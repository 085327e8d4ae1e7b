"""CPU: property-based checks (hypothesis) of the oracle pieces whose algebra has invariants that do not
depend on the input size: SO(3) exponential, LDLT solve, voxel filter, de-skew, plane / line fits."""
import numpy as np
from hypothesis import given, settings, strategies as st
from hypothesis.extra import numpy as hnp

from oracle import oracle as orc

finite = st.floats(-3.0, 3.0, allow_nan=False, allow_infinity=False, width=64)


@settings(max_examples=60, deadline=None)
@given(hnp.arrays(np.float64, 3, elements=finite))
def test_so3_exp_is_a_rotation(w):
    R = orc.so3_exp(w)
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-12 and abs(np.linalg.det(R) - 1) < 1e-12
    assert np.abs(orc.so3_exp(-w) - R.T).max() < 1e-12            # exp(-w) = exp(w)^T


@settings(max_examples=60, deadline=None)
@given(hnp.arrays(np.float64, (6, 6), elements=finite), hnp.arrays(np.float64, 6, elements=finite))
def test_ldlt_solves_spd_systems(B, b):
    A = B @ B.T + 6 * np.eye(6)                                    # symmetric positive definite, cond <= ~1e2
    x = orc.ldlt6_solve(A, b)
    assert np.abs(A @ x - b).max() <= 1e-10 * max(1.0, np.abs(b).max())


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 400), st.integers(0, 2**31 - 1), st.sampled_from([0.2, 0.3, 1.0]))
def test_voxel_grid_invariants(n, seed, leaf):
    rng = np.random.default_rng(seed)
    P = np.zeros((n, 4), np.float32)
    P[:, :3] = rng.normal(0, 2.0, (n, 3))
    P[:, 3] = rng.uniform(0, 255, n)
    V = orc.voxel_grid(P, leaf)
    assert 1 <= len(V) <= n
    # centroids stay inside the bounding box; the point-count weighted mean of the centroids is the cloud's mean
    assert (V[:, :3] >= P[:, :3].min(0) - 1e-4).all() and (V[:, :3] <= P[:, :3].max(0) + 1e-4).all()
    # the number of occupied voxels does not depend on the order of the points
    assert len(orc.voxel_grid(P[rng.permutation(n)], leaf)) == len(V)
    # filtering the centroids again with a much larger leaf collapses them further, never grows
    assert len(orc.voxel_grid(V, leaf * 8)) <= len(V)


@settings(max_examples=40, deadline=None)
@given(hnp.arrays(np.float64, 3, elements=st.floats(-0.2, 0.2)), hnp.arrays(np.float64, 3, elements=st.floats(-1, 1)), st.integers(0, 2**31 - 1))
def test_deskew_is_rigid_per_point_and_interpolates(rv, t, seed):
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(seed)
    x, y, z, w = R.from_rotvec(rv).as_quat()
    P = np.zeros((64, 4), np.float32)
    P[:, :3] = rng.normal(0, 10, (64, 3))
    P[:, 3] = rng.integers(0, 16, 64) + np.float32(0.1) * rng.uniform(0, 1, 64).astype(np.float32)
    out = orc.deskew(P, [w, x, y, z], t)
    frac = P[:, 3] - np.trunc(P[:, 3]).astype(np.float32)
    s = (np.float32(1) - frac / np.float32(0.1)).astype(np.float64)
    # |p' | = |p - s t| : each point undergoes a rigid motion (slerp keeps unit quaternions unit to rounding)
    lhs = np.linalg.norm(out[:, :3].astype(np.float64), axis=1)
    rhs = np.linalg.norm(P[:, :3].astype(np.float64) - s[:, None] * t, axis=1)
    assert np.abs(lhs - rhs).max() < 5e-5
    assert np.array_equal(out[:, 3], P[:, 3])


@settings(max_examples=40, deadline=None)
@given(hnp.arrays(np.float64, 3, elements=st.floats(-1, 1)).filter(lambda v: np.linalg.norm(v) > 0.2), st.floats(1.0, 40.0), st.integers(0, 2**31 - 1))
def test_plane_fit_recovers_a_plane(nrm, dist, seed):
    rng = np.random.default_rng(seed)
    n = nrm / np.linalg.norm(nrm)
    u = np.cross(n, [0.3, 0.5, 0.8])
    u /= np.linalg.norm(u)
    v = np.cross(n, u)
    ab = rng.uniform(-0.4, 0.4, (200, 2))
    pts = (-dist * n) + ab[:, :1] * u + ab[:, 1:] * v            # n . p + dist = 0
    M = np.ones((200, 4), np.float32)
    M[:, :3] = pts
    f = np.zeros((1, 4), np.float32)
    f[0, :3] = pts[:20].mean(0)
    valid, norm, d = orc.assoc_planes(M, f, [1, 0, 0, 0], [0, 0, 0])
    assert valid[0]
    sgn = 1.0 if norm[0] @ n > 0 else -1.0
    assert np.abs(sgn * norm[0] - n).max() < 2e-4 and abs(sgn * d[0] - dist) < 2e-3 * max(1.0, dist)   # float32 map points
