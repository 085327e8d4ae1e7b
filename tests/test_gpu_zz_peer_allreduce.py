"""GPU (one device): the all-reduce kernel of the sharded registration (k_peer_allreduce, config C5) with the ranks played
by streams of one device — every "rank" has its own mailbox, stores its partial sums into all of them, polls its own and
adds in rank order.  On several GPUs the mailboxes are mapped through CUDA IPC (tests/multigpu/run_sharded.py,
bench.py --gpus N); the protocol — sequence tags packed into the data words, two parities, back-to-back calls with ranks
running ahead of each other — is the same and is what this checks.  (Named to run last: its launches wait for each other.)"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,n,n_a", [(2, 30, 1), (4, 29, 29), (8, 30, 1), (8, 64, 0), (1, 2, 2)])
def test_peer_allreduce_sums_in_rank_order_on_every_rank(world, n, n_a):
    import rgc_slam_b200 as rgc
    from rgc_slam_b200 import api
    L = rgc.lib()
    L.rgc_debug_peer_allreduce_selftest.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    ctx = api.default_context(0)
    reps = 64
    rng = np.random.default_rng(world * 100 + n)
    x = rng.normal(size=(reps, world, n)) * 10.0 ** rng.integers(-8, 8, size=(reps, world, n))   # sums whose order matters
    x[3, :, 0] = [np.nan if r == world - 1 else 1.0 for r in range(world)]                      # NaN / inf travel as bit patterns
    x[4, :, 1] = np.inf
    out, pub = np.zeros_like(x), np.zeros_like(x)
    ctx.check(L.rgc_debug_peer_allreduce_selftest(ctx._h, world, n, n_a, reps, x.ctypes.data, out.ctypes.data, pub.ctypes.data))
    want = np.zeros((reps, n))
    for r in range(world):                                                                      # 0.0 + x[0] + x[1] + ... in rank order
        want = want + x[:, r, :]
    for r in range(world):
        assert np.array_equal(out[:, r, :], want, equal_nan=True), f"rank {r}"
        assert np.array_equal(pub[:, r, :], want, equal_nan=True), f"rank {r} (published copy)"
    assert np.isnan(want[3, 0]) and np.isinf(want[4, 1])
