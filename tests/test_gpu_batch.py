"""GPU: the batched registration path (include/rgc_batch.h, config C4) against (a) the single-registration
API pair by pair — bit-identical, the batch kernels replay the same reduction order — and (b) the oracle."""
import numpy as np
import pytest

from conftest import rot_angle

pytestmark = pytest.mark.gpu


def _single(rgc, p, prm):
    g = rgc.FastGICP()
    g.setMaximumIterations(prm["max_iterations"])
    g.setMaxCorrespondenceDistance(prm["corr"])
    if "optimizer" in prm:
        g.setOptimizer(prm["optimizer"])
    g.setInputTarget(p["tgt"])
    g.setInputSource(p["src"])
    T = g.align(p["guess"])
    return g, T


def _params(prm):
    from rgc_slam_b200 import batch
    q = batch.default_params()
    q.max_iterations = prm["max_iterations"]
    q.max_correspondence_distance = prm["corr"]
    if "optimizer" in prm:
        q.optimizer = prm["optimizer"]
    return q


@pytest.fixture(scope="module")
def mixed_pairs(scene, traj):
    """ragged batch: sweeps of different resolutions against single sweeps and accumulated submaps, near and far guesses"""
    from rgc_slam_b200 import synth
    rng = np.random.default_rng(11)
    pairs = []
    for i, (f, az_s, az_t, n_acc) in enumerate([(10, 450, 450, 1), (14, 900, 450, 3), (18, 225, 900, 2), (22, 450, 1800, 1), (26, 1800, 900, 4), (30, 450, 450, 1), (34, 300, 600, 2)]):
        chunks = []
        for j in range(n_acc):
            sc = synth.lidar_scan(scene, traj[f - j], n_azimuth=az_t, seed=900 + 10 * i + j)
            Tr = synth.relative_pose(traj[f - j], traj[f])
            chunks.append((sc[:, :3].astype(np.float64) @ Tr[:3, :3].T + Tr[:3, 3]).astype(np.float32))
        tgt = np.ones((sum(len(c) for c in chunks), 4), np.float32)
        tgt[:, :3] = np.concatenate(chunks, 0)
        src = synth.to_xyz1(synth.lidar_scan(scene, traj[f + 1], n_azimuth=az_s, seed=950 + i))
        guess = np.eye(4, dtype=np.float32)
        guess[:3, 3] = rng.uniform(-0.4, 0.4, 3)
        a = np.deg2rad(rng.uniform(-4, 4))
        guess[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
        pairs.append(dict(src=src, tgt=tgt, guess=guess))
    pairs.append(dict(src=pairs[0]["src"][:40].copy(), tgt=pairs[0]["tgt"][:15].copy(), guess=np.eye(4, dtype=np.float32)))  # fewer target points than k
    return pairs


@pytest.mark.parametrize("chunk", [0, 3])
@pytest.mark.parametrize("prm", [dict(max_iterations=64, corr=2.0), dict(max_iterations=3, corr=0.5), dict(max_iterations=20, corr=2.0, optimizer=0)])
def test_batch_equals_single_registrations_bit_for_bit(mixed_pairs, prm, chunk):
    import rgc_slam_b200 as rgc
    from rgc_slam_b200 import batch
    res = batch.align_batch(mixed_pairs, params=_params(prm), want_fitness=True, max_chunk_pairs=chunk)
    assert len(res) == len(mixed_pairs)
    for p, r in zip(mixed_pairs, res):
        g, T = _single(rgc, p, prm)
        assert np.array_equal(r["T"], T)
        lr = g.last_result
        assert (r["converged"], r["iterations"], r["n_linearize"], r["n_compute_error"], r["n_inliers"]) == \
               (lr["converged"], lr["iterations"], lr["n_linearize"], lr["n_compute_error"], lr["n_inliers"])
        assert r["final_error"] == lr["final_error"]
        Hs = g.getFinalHessian()
        assert np.array_equal(r["final_hessian"], Hs), f"final Hessian differs: max rel {np.abs(r['final_hessian'] - Hs).max() / np.abs(Hs).max():.3e}, entries {np.argwhere(r['final_hessian'] != Hs).tolist()}"
        assert r["fitness"] == g.getFitnessScore()
    assert batch.last_stage_ms()["rounds"] >= 2


def test_batch_rejected_steps_and_lm_failure(scene, traj, capfd):
    """the far-guess cases of test_gpu_round2 (rejected LM trials, 'lm not converged') through the batch path"""
    import rgc_slam_b200 as rgc
    from rgc_slam_b200 import batch, synth
    from test_gpu_round2 import far_guess
    tgt = synth.to_xyz1(synth.lidar_scan(scene, traj[10], n_azimuth=225, seed=5))
    src = synth.to_xyz1(synth.lidar_scan(scene, traj[11], n_azimuth=225, seed=6))
    pairs = [dict(src=src, tgt=tgt, guess=far_guess(t)) for t in (3, 19, 29, 0, 14)]
    for lm_max in (10, 1):
        q = batch.default_params()
        q.max_iterations, q.max_correspondence_distance, q.lm_max_iterations = 30, 1.0, lm_max
        res = batch.align_batch(pairs, params=q)
        rejected = 0
        for p, r in zip(pairs, res):
            g = rgc.FastGICP()
            g.setMaximumIterations(30)
            g.setMaxCorrespondenceDistance(1.0)
            g.setLMMaxIterations(lm_max)
            g.setInputTarget(p["tgt"])
            g.setInputSource(p["src"])
            T = g.align(p["guess"])
            lr = g.last_result
            assert np.array_equal(r["T"], T)
            assert (r["converged"], r["iterations"], r["n_linearize"], r["n_compute_error"]) == (lr["converged"], lr["iterations"], lr["n_linearize"], lr["n_compute_error"])
            assert np.array_equal(r["final_hessian"], g.getFinalHessian())
            rejected += int(r["n_compute_error"] > r["n_linearize"]) if lm_max == 10 else int(not r["converged"] and r["iterations"] < 29)
        assert rejected >= 3
    assert "lm not converged" in capfd.readouterr().err


def test_batch_c4_pairs_match_oracle():
    """config C4 inputs through the batch entry point vs the ORACLE (1e-4 m / 1e-5 rad) and the caller's gate"""
    from oracle import oracle as orc
    from rgc_slam_b200 import batch, workloads
    pairs = workloads.make_c4_pairs(32, 16, n_base=4)
    q = batch.default_params()
    q.max_iterations, q.max_correspondence_distance = 64, 2.0
    res = batch.align_batch(pairs, params=q, want_fitness=True)
    oracles = {}
    for p, r in zip(pairs, res):
        if p["base"] not in oracles:
            o = orc.FastGICP(max_iterations=64, corr_dist=2.0)
            o.setInputTarget(p["tgt"])
            o.setInputSource(p["src"])
            oracles[p["base"]] = o
        o = oracles[p["base"]]
        To = o.align(p["guess"])
        assert np.abs(r["T"][:3, 3] - To[:3, 3]).max() < 1e-4 and rot_angle(r["T"][:3, :3], To[:3, :3]) < 1e-5
        assert (r["converged"], r["iterations"]) == (o.last["converged"], o.last["iterations"])
        fo = o.getFitnessScore()
        assert abs(r["fitness"] - fo) <= 1e-6 * fo
        assert (r["converged"] and r["fitness"] <= 0.1) == (o.last["converged"] and fo <= 0.1)  # RGC_mapping.cpp:2070-2071


def test_batch_rejects_bad_input():
    import rgc_slam_b200 as rgc
    from rgc_slam_b200 import batch
    ok = np.ones((100, 4), np.float32)
    ok[:, :3] = np.random.default_rng(0).normal(0, 3, (100, 3))
    bad = ok.copy()
    bad[3, 2] = np.nan
    with pytest.raises(rgc.RgcError, match="non-finite"):
        batch.align_batch([dict(src=ok, tgt=ok, guess=None), dict(src=bad, tgt=ok, guess=None)])
    with pytest.raises(rgc.RgcError):
        batch.align_batch([dict(src=ok[:0], tgt=ok, guess=None)])
    res = batch.align_batch([dict(src=ok, tgt=ok, guess=None)])  # the context is still usable
    assert res[0]["converged"] and np.abs(res[0]["T"] - np.eye(4)).max() < 1e-5
