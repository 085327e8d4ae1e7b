"""CPU (gloo, world_size 2): host-side logic of the two multi-GPU modes — the C4 partitioner +
result gather, and the C5 slab ownership rule + all-reduce of partial sums."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rgc_slam_b200 import sharded


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 4096):
        for world in (1, 2, 3, 8):
            blocks = [sharded.shard_range(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1
    w = np.array([1, 1, 1, 1, 10, 1, 1, 1, 1, 10], float)
    blocks = [sharded.shard_range(len(w), 2, r, w) for r in range(2)]
    assert blocks[0][1] == blocks[1][0] and abs(w[blocks[0][0]:blocks[0][1]].sum() - w[blocks[1][0]:blocks[1][1]].sum()) <= 10


def test_slab_ownership_is_exclusive_and_total():
    rng = np.random.default_rng(0)
    pts = rng.normal(0, 50, (20000, 3)).astype(np.float32)
    for world in (1, 2, 4, 8):
        b = sharded.slab_boundaries(pts[:, 0], world)
        assert len(b) == world + 1 and np.isneginf(b[0]) and np.isposinf(b[-1]) and (np.diff(b) > 0).all()
        q = rng.normal(0, 80, 5000).astype(np.float32)
        owners = sum(((q >= b[r]) & (q < b[r + 1])).astype(int) for r in range(world))
        assert (owners == 1).all()
        counts = [len(sharded.slab_select(pts, 0, b[r], b[r + 1], 0.0)) for r in range(world)]
        assert sum(counts) == len(pts) and max(counts) - min(counts) < 0.02 * len(pts) + 2
        with_halo = [len(sharded.slab_select(pts, 0, b[r], b[r + 1], 5.0)) for r in range(world)]
        assert all(h >= c for h, c in zip(with_halo, counts))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # ---- C4: every rank registers its block of pairs, results gathered in pair order
    n_pairs = 11
    lo, hi = sharded.shard_range(n_pairs, world, rank)
    local = [(p, float(p) ** 2) for p in range(lo, hi)]      # stand-in for (pair id, result)
    allres = sharded.gather_results(local, world)
    ok_c4 = [r[0] for r in allres] == list(range(n_pairs))
    # ---- C5: partial sums over owned queries + one all-reduce == the global sum
    rng = np.random.default_rng(42)                          # same data on every rank
    tgt = rng.normal(0, 30, (5000, 3)).astype(np.float32)
    qry = rng.normal(0, 30, (800, 3)).astype(np.float32)
    b = sharded.slab_boundaries(tgt[:, 0], world)
    mine = (qry[:, 0] >= b[rank]) & (qry[:, 0] < b[rank + 1])
    corr = 4.0
    keep = sharded.slab_select(tgt, 0, b[rank], b[rank + 1], corr)
    local_t = tgt[keep]
    part = np.zeros(29)
    for x in qry[mine]:
        d2 = ((local_t - x) ** 2).sum(1)
        j = int(d2.argmin())
        if d2[j] < corr * corr:                              # the halo guarantees the true NN is local
            part[0] += d2[j]
            part[28] += 1
    t = torch.from_numpy(part)
    dist.all_reduce(t)
    full = np.zeros(29)
    for x in qry:
        d2 = ((tgt - x) ** 2).sum(1)
        j = int(d2.argmin())
        if d2[j] < corr * corr:
            full[0] += d2[j]
            full[28] += 1
    ok_c5 = abs(t[0].item() - full[0]) <= 1e-9 * full[0] and t[28].item() == full[28]
    q.put((rank, ok_c4, ok_c5))
    dist.destroy_process_group()


def test_two_rank_gloo_partition_and_allreduce():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res
