"""Seeded synthetic sparse-channel LiDAR scans.

The reference's bag datasets are unavailable offline (README.md:39-53), so every test and
benchmark in this repo runs on procedural scenes with known ground truth: a ground plane
0.56 m below the sensor (``laderH``, rgc_slam/src/scanRegistration.cpp:39), axis-aligned boxes
(walls / buildings) and thin vertical cylinders (poles), ray-cast by an N-beam spinning sensor
in firing order (azimuth-major, beam-minor, clockwise like a Velodyne so that the reference's
``startOri/endOri/halfPassed`` logic, scanRegistration.cpp:117-204, behaves) with Gaussian range
noise, which makes exact distance / curvature ties measure-zero.

Pure numpy; deterministic for a given seed (numpy PCG64).  Used by tests/, bench.py and
__graft_entry__.smoke().  Nothing here touches the GPU.
"""
from __future__ import annotations

import dataclasses
import numpy as np

SENSOR_HEIGHT = 0.56  # scanRegistration.cpp:39 laderH
BASE_SEED = 20240913


def beam_elevations_deg(n_beams: int) -> np.ndarray:
    """Elevation angle of each ring, matched to the reference's ring-id formulas."""
    if n_beams == 16:  # scanID = int((v + 15) / 2 + 0.5), scanRegistration.cpp:147
        return -15.0 + 2.0 * np.arange(16)
    if n_beams == 32:  # scanID = int((v + 92/3) * 3/4), scanRegistration.cpp:156 -> bin centres
        return -92.0 / 3.0 + (np.arange(32) + 0.5) * 4.0 / 3.0
    if n_beams == 64:  # HDL-64 split fan, scanRegistration.cpp:160-170: 1/3 deg steps down to -8.83 deg, then 1/2 deg steps
        upper = 2.0 - np.arange(33) / 3.0                  # scanID = int((2 - v) * 3 + 0.5) = 0..32
        lower = -8.83 - np.arange(1, 32) / 2.0             # scanID = 32 + int((-8.83 - v) * 2 + 0.5) = 33..63 (> 50 is dropped there)
        return np.concatenate([upper, lower])
    # generic: uniform fan (used for the 128-beam GICP-only config C5)
    return np.linspace(-25.0, 25.0, n_beams)


@dataclasses.dataclass
class Scene:
    boxes: np.ndarray      # (B, 6): xmin, ymin, zmin, xmax, ymax, zmax  (world, ground z = 0)
    cylinders: np.ndarray  # (C, 4): cx, cy, radius, height
    extent: float

    @staticmethod
    def make(seed: int, extent: float = 200.0, n_boxes: int | None = None, n_cyl: int = 30) -> "Scene":
        rng = np.random.Generator(np.random.PCG64(seed))
        if n_boxes is None:
            n_boxes = int(rng.integers(40, 81))
        half = extent / 2
        boxes = []
        while len(boxes) < n_boxes:
            cx, cy = rng.uniform(-half, half, 2)
            w, d = rng.uniform(2.0, 20.0, 2)
            h = rng.uniform(2.0, 12.0)
            # keep a corridor around the trajectory (|y| < 4 m near the x axis) free
            if abs(cy) - d / 2 < 4.0:
                continue
            boxes.append([cx - w / 2, cy - d / 2, 0.0, cx + w / 2, cy + d / 2, h])
        cyl = []
        while len(cyl) < n_cyl:
            cx, cy = rng.uniform(-half, half, 2)
            if abs(cy) < 2.5:
                continue
            cyl.append([cx, cy, 0.15, rng.uniform(3.0, 6.0)])
        return Scene(np.asarray(boxes, np.float64), np.asarray(cyl, np.float64).reshape(-1, 4), extent)


def trajectory(n: int, step: float = 0.3, max_yaw_deg: float = 2.0, seed: int = 0) -> np.ndarray:
    """Smooth planar curve: (n, 3) of x, y, yaw; ~`step` metres and <= max_yaw per frame."""
    rng = np.random.Generator(np.random.PCG64(seed + 7))
    yaw_rate = np.deg2rad(max_yaw_deg) * 0.5 * np.sin(np.arange(n) * 0.11 + rng.uniform(0, 6.28))
    yaw = np.cumsum(yaw_rate) * 0.3
    yaw = np.clip(yaw, -0.25, 0.25)
    x = np.cumsum(step * np.cos(yaw)) - step * n / 2
    y = np.cumsum(step * np.sin(yaw))
    y = np.clip(y, -1.5, 1.5)
    return np.stack([x, y, yaw], 1)


def pose_matrix(p: np.ndarray) -> np.ndarray:
    """World-from-sensor 4x4 for a planar pose (x, y, yaw), sensor SENSOR_HEIGHT above ground."""
    c, s = np.cos(p[2]), np.sin(p[2])
    T = np.eye(4)
    T[:3, :3] = [[c, -s, 0], [s, c, 0], [0, 0, 1]]
    T[:3, 3] = [p[0], p[1], SENSOR_HEIGHT]
    return T


def _raycast(scene: Scene, origin: np.ndarray, dirs: np.ndarray):
    """First hit of rays origin + t*dirs (world frame).  Returns t (inf = miss) and surface id."""
    n = dirs.shape[0]
    t_best = np.full(n, np.inf)
    surf = np.full(n, -1, np.int32)
    # ground z = 0
    dz = dirs[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = np.where(dz < -1e-9, -origin[2] / dz, np.inf)
    hit = tg < t_best
    t_best = np.where(hit, tg, t_best)
    surf = np.where(hit, 0, surf)
    # boxes (slab test), chunked over boxes to bound memory
    inv = 1.0 / np.where(np.abs(dirs) < 1e-12, 1e-12, dirs)
    B = scene.boxes
    for b0 in range(0, B.shape[0], 16):
        bb = B[b0:b0 + 16]
        t1 = (bb[None, :, 0:3] - origin[None, None, :]) * inv[:, None, :]
        t2 = (bb[None, :, 3:6] - origin[None, None, :]) * inv[:, None, :]
        tmin = np.minimum(t1, t2).max(axis=2)
        tmax = np.maximum(t1, t2).min(axis=2)
        ok = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 1e-6)
        tb = np.where(ok, tmin, np.inf)
        j = tb.argmin(axis=1)
        tbm = tb[np.arange(n), j]
        hit = tbm < t_best
        t_best = np.where(hit, tbm, t_best)
        surf = np.where(hit, 1 + b0 + j, surf)
    # vertical cylinders
    C = scene.cylinders
    if C.shape[0]:
        ox = origin[0] - C[None, :, 0]
        oy = origin[1] - C[None, :, 1]
        dx, dy = dirs[:, 0:1], dirs[:, 1:2]
        a = dx * dx + dy * dy
        bq = 2 * (ox * dx + oy * dy)
        c = ox * ox + oy * oy - C[None, :, 2] ** 2
        disc = bq * bq - 4 * a * c
        with np.errstate(divide="ignore", invalid="ignore"):
            tc = (-bq - np.sqrt(np.maximum(disc, 0))) / (2 * a)
        zc = origin[2] + tc * dirs[:, 2:3]
        ok = (disc > 0) & (tc > 1e-6) & (zc >= 0) & (zc <= C[None, :, 3])
        tc = np.where(ok, tc, np.inf)
        j = tc.argmin(axis=1)
        tcm = tc[np.arange(n), j]
        hit = tcm < t_best
        t_best = np.where(hit, tcm, t_best)
        surf = np.where(hit, 1000 + j, surf)
    return t_best, surf


def lidar_scan(scene: Scene, pose: np.ndarray, n_beams: int = 16, n_azimuth: int = 1800, seed: int = 0,
               sigma: float = 0.01, max_range: float = 80.0, min_range: float = 0.5,
               extra_T: np.ndarray | None = None) -> np.ndarray:
    """One sweep in the SENSOR frame: float32 (n, 4) = x, y, z, intensity (integer-valued 0..255),
    firing order.  `extra_T` (4x4) perturbs the sensor pose (full SE(3)) after the planar pose."""
    rng = np.random.Generator(np.random.PCG64(seed))
    elev = np.deg2rad(beam_elevations_deg(n_beams))
    az = -np.arange(n_azimuth) * (2 * np.pi / n_azimuth)  # clockwise
    A, E = np.meshgrid(az, elev, indexing="ij")           # azimuth-major, beam-minor
    A = A.ravel()
    E = E.ravel()
    d_s = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], 1)
    T = pose_matrix(pose)
    if extra_T is not None:
        T = T @ extra_T
    d_w = d_s @ T[:3, :3].T
    t, surf = _raycast(scene, T[:3, 3], d_w)
    t = t + rng.normal(0.0, sigma, t.shape)
    ok = np.isfinite(t) & (t < max_range) & (t > min_range)
    pts_s = d_s[ok] * t[ok, None]
    # integer intensity from a world-space pattern so intensity edges exist (lane-marking stripes
    # on the ground, banding on walls); deterministic per surface
    pw = pts_s @ T[:3, :3].T + T[:3, 3]
    s = surf[ok]
    inten = np.where(s == 0,
                     np.where((np.floor(pw[:, 1] / 0.75).astype(np.int64) % 4) == 0, 220, 40),
                     np.where(s >= 1000, 150,
                              60 + 15 * (s % 7) + np.where((np.floor(pw[:, 2] / 1.0).astype(np.int64) % 2) == 0, 100, 0)))
    inten = np.clip(inten + rng.integers(-3, 4, inten.shape), 0, 255)
    out = np.empty((pts_s.shape[0], 4), np.float32)
    out[:, :3] = pts_s.astype(np.float32)
    out[:, 3] = inten.astype(np.float32)
    return out


def to_xyz1(scan_xyzi: np.ndarray) -> np.ndarray:
    """PCL-style homogeneous float4 (data[3] = 1; fast_gicp_impl.hpp:131 relies on it)."""
    out = np.ascontiguousarray(scan_xyzi[:, :4], dtype=np.float32).copy()
    out[:, 3] = 1.0
    return out


def relative_pose(pose_from: np.ndarray, pose_to: np.ndarray) -> np.ndarray:
    """T such that points in `pose_from`'s sensor frame map into `pose_to`'s frame."""
    return np.linalg.inv(pose_matrix(pose_to)) @ pose_matrix(pose_from)


def make_submap(scene: Scene, traj: np.ndarray, frame: int, n_points: int, n_beams: int = 16, window: int | None = None,
                seed: int = 0, n_azimuth: int = 1800) -> np.ndarray:
    """Target for config C2: accumulation of the scans preceding `frame`, expressed in the sensor
    frame of pose `frame - 1`, subsampled (without replacement) to exactly `n_points`."""
    rng = np.random.Generator(np.random.PCG64(seed + 99))
    ref = traj[frame - 1]
    chunks = []
    total = 0
    f = frame - 1
    while total < n_points * 1.02 and f >= 0 and (window is None or frame - 1 - f < window):
        sc = lidar_scan(scene, traj[f], n_beams, n_azimuth, seed=seed * 100003 + f)
        Tr = relative_pose(traj[f], ref)
        p = sc[:, :3].astype(np.float64) @ Tr[:3, :3].T + Tr[:3, 3]
        chunks.append(p.astype(np.float32))
        total += p.shape[0]
        f -= 1
    pts = np.concatenate(chunks, 0)
    if pts.shape[0] < n_points:
        raise ValueError(f"trajectory too short: only {pts.shape[0]} points for a {n_points}-point submap")
    sel = rng.permutation(pts.shape[0])[:n_points]
    sel.sort()
    out = np.ones((n_points, 4), np.float32)
    out[:, :3] = pts[sel]
    return out


def small_perturbation(rng: np.random.Generator, trans: float = 0.5, rot_deg: float = 5.0) -> np.ndarray:
    """U(+-trans m, +-rot_deg) perturbation about identity (config C4 initial error)."""
    w = np.deg2rad(rng.uniform(-rot_deg, rot_deg, 3))
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    R = np.eye(3) if th < 1e-12 else np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * K @ K
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = rng.uniform(-trans, trans, 3)
    return T
