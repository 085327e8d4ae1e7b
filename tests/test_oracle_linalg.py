"""CPU: the oracle's small linear algebra (restating Eigen routines the reference depends on)
against numpy / scipy — SURVEY.md §8c "what pins results instead" item (ii)."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from oracle import oracle as orc


def _rand_spd(rng, scale=1.0, rank=3):
    A = rng.normal(size=(3, rank)) * scale
    return A @ A.T


def test_jacobi_svd3_reconstructs_and_orders():
    rng = np.random.default_rng(0)
    for i in range(200):
        A = _rand_spd(rng, 10 ** rng.uniform(-3, 2)) if i % 2 else rng.normal(size=(3, 3))
        U, s, V = orc.jacobi_svd3(A)
        assert np.allclose(U @ np.diag(s) @ V.T, A, atol=1e-12 * max(1, np.abs(A).max()))
        assert np.all(np.diff(s) <= 1e-15) and np.all(s >= 0)
        assert np.allclose(U.T @ U, np.eye(3), atol=1e-12) and np.allclose(V.T @ V, np.eye(3), atol=1e-12)
        assert np.allclose(s, np.linalg.svd(A, compute_uv=False), rtol=1e-10, atol=1e-13)


def test_plane_regularisation_equals_normal_form():
    """PLANE: U diag(1,1,1e-3) V^T == I - 0.999 n n^T for a well-conditioned PSD covariance."""
    rng = np.random.default_rng(1)
    for _ in range(100):
        C = _rand_spd(rng)
        U, s, V = orc.jacobi_svd3(C)
        R = U @ np.diag([1, 1, 1e-3]) @ V.T
        w, E = np.linalg.eigh(C)
        n = E[:, 0]
        assert np.allclose(R, np.eye(3) - 0.999 * np.outer(n, n), atol=1e-9)


def test_eigh3():
    rng = np.random.default_rng(2)
    for _ in range(200):
        A = _rand_spd(rng, 10 ** rng.uniform(-3, 2))
        w, V = orc.eigh3(A)
        wn = np.linalg.eigvalsh(A)
        assert np.allclose(w, wn, rtol=1e-10, atol=1e-12 * np.abs(wn).max())
        assert np.allclose(A @ V, V @ np.diag(w), atol=1e-10 * max(1, np.abs(A).max()))
        assert np.allclose(np.linalg.norm(V, axis=0), 1.0)


def test_inverse4_and_ldlt6():
    rng = np.random.default_rng(3)
    for _ in range(100):
        A = rng.normal(size=(4, 4)) + 3 * np.eye(4)
        assert np.allclose(orc.inverse4(A), np.linalg.inv(A), rtol=1e-9, atol=1e-11)
        B = rng.normal(size=(6, 8))
        H = B @ B.T + 1e-3 * np.eye(6)
        b = rng.normal(size=6)
        assert np.allclose(orc.ldlt6_solve(H, b), np.linalg.solve(H, b), rtol=1e-8, atol=1e-10)
    # indefinite but non-singular (LM can see this for rho < 0 steps)
    H = np.diag([1.0, -2.0, 3.0, 4.0, 5.0, 6.0])
    H[0, 1] = H[1, 0] = 0.5
    b = np.arange(6.0)
    assert np.allclose(orc.ldlt6_solve(H, b), np.linalg.solve(H, b))


def test_so3_exp_matches_scipy_including_small_angle_branch():
    rng = np.random.default_rng(4)
    for scale in (1e-8, 1e-6, 1e-4, 1e-2, 1.0, 3.0):
        for _ in range(20):
            w = rng.normal(size=3) * scale
            assert np.allclose(orc.so3_exp(w), Rotation.from_rotvec(w).as_matrix(), atol=1e-12)
