"""Seeded synthetic workloads of BASELINE.json's configs, shared by bench.py, tools/ and
tests/ so that every number and every parity case is quoted on the same inputs.

  C2  build_c2_pairs        VLP-16 sweep vs an n_submap-point accumulation of the preceding sweeps
  C4  make_c4_pairs         sweep vs 100 000-point submap, initial error U(+-0.5 m, +-5 deg) about the truth
  C5  make_c5_map / make_c5_case   sweep vs a very large map (tiled replication of the C2 submap)
"""
from __future__ import annotations

import numpy as np

from . import synth

N_SUBMAP_C2 = 500_000
N_SUBMAP_C4 = 100_000
CALL_SITE = dict(max_iterations=25, corr_dist=2.0, transformation_epsilon=1e-6)  # RGC_odometer.cpp:1000-1006


def build_c2_pairs(rank: int, n_submap: int = N_SUBMAP_C2, n_pairs: int = 2):
    """Seeded synthetic stream: sweeps along a trajectory, submap = accumulation of the preceding
    sweeps in the previous frame's coordinates, subsampled to exactly n_submap points."""
    scene = synth.Scene.make(synth.BASE_SEED + 2000)
    need = int(np.ceil(n_submap * 1.03 / 20000.0)) + 2
    if n_pairs * 3 + rank > 64:
        raise ValueError("at most 64 frames of the stream are used (n_pairs * 3 + rank <= 64)")
    traj = synth.trajectory(need + 72, seed=2)  # fixed length: the same stream whatever n_pairs / rank
    scans = {}

    def scan(f):
        if f not in scans:
            scans[f] = synth.lidar_scan(scene, traj[f], seed=synth.BASE_SEED + 2000 + f)
        return scans[f]

    pairs = []
    for p in range(n_pairs):
        frame = need + p * 3 + rank
        ref = traj[frame - 1]
        chunks, total, f = [], 0, frame - 1
        while total < n_submap * 1.02 and f >= 0:
            sc = scan(f)
            Tr = synth.relative_pose(traj[f], ref)
            chunks.append((sc[:, :3].astype(np.float64) @ Tr[:3, :3].T + Tr[:3, 3]).astype(np.float32))
            total += len(sc)
            f -= 1
        pts = np.concatenate(chunks, 0)
        if len(pts) < n_submap:
            raise RuntimeError("not enough points for the submap")
        rng = np.random.Generator(np.random.PCG64(1234 + p + 100 * rank))
        sel = np.sort(rng.permutation(len(pts))[:n_submap])
        tgt = np.ones((n_submap, 4), np.float32)
        tgt[:, :3] = pts[sel]
        src = synth.to_xyz1(scan(frame))
        # guess = previous frame-to-frame motion
        guess = synth.relative_pose(traj[frame - 1], traj[frame - 2]).astype(np.float32)
        truth = synth.relative_pose(traj[frame], traj[frame - 1])
        pairs.append(dict(src=src, tgt=tgt, guess=guess, truth=truth))
    return pairs


def c4_perturbation(rng) -> np.ndarray:
    """U(+-0.5 m, +-5 deg) about the identity"""
    t = rng.uniform(-0.5, 0.5, 3)
    w = np.deg2rad(rng.uniform(-5.0, 5.0, 3))
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    R = np.eye(3) + (np.sin(th) / th) * K + ((1 - np.cos(th)) / th**2) * K @ K if th > 1e-12 else np.eye(3)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def make_c4_pairs(first: int, n_pairs: int, n_base: int = 4, n_submap: int = N_SUBMAP_C4):
    """Pairs [first, first + n_pairs) of the C4 list: pair i = base pair (i mod n_base) with the i-th draw
    of the initial error.  The list is the same whatever the sharding: rank r of N takes a contiguous
    slice (sharded.shard_range), so results are comparable across N."""
    base = build_c2_pairs(0, n_submap, n_base)
    out = []
    for i in range(first, first + n_pairs):
        rng = np.random.Generator(np.random.PCG64([synth.BASE_SEED + 4000, i]))
        b = base[i % n_base]
        guess = (c4_perturbation(rng) @ b["truth"]).astype(np.float32)
        out.append(dict(src=b["src"], tgt=b["tgt"], guess=guess, truth=b["truth"], base=i % n_base, index=i))
    return out


def make_c5_case(n_tiles: int, n_beams: int = 128, n_submap: int = N_SUBMAP_C2, tile_pitch: float = 400.0):
    """A large map = the C2 submap replicated on a grid of `tile_pitch` offsets (n_tiles x n_submap points),
    one n_beams sweep placed in a central tile, and the guess conjugated into that tile."""
    pair = build_c2_pairs(0, n_submap, 1)[0]
    tgt0 = pair["tgt"]
    side = int(np.ceil(np.sqrt(n_tiles)))
    offs = [(tile_pitch * (i % side), tile_pitch * (i // side)) for i in range(n_tiles)]
    tgt = np.concatenate([tgt0 + np.array([ox, oy, 0, 0], np.float32) for ox, oy in offs], 0)
    ci = (side // 2) * side + side // 2
    centre = offs[ci] if ci < n_tiles else offs[0]
    scene = synth.Scene.make(synth.BASE_SEED + 2000)
    need = int(np.ceil(n_submap * 1.03 / 20000.0)) + 2
    traj = synth.trajectory(need + 72, seed=2)
    src = synth.to_xyz1(synth.lidar_scan(scene, traj[need], n_beams=n_beams, seed=4242))
    src[:, 0] += centre[0]
    src[:, 1] += centre[1]
    Toff = np.eye(4)
    Toff[:2, 3] = centre
    guess = (Toff @ pair["guess"].astype(np.float64) @ np.linalg.inv(Toff)).astype(np.float32)
    return dict(src=src, tgt=tgt, guess=guess, centre=centre)
