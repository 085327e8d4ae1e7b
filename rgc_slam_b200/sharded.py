"""Multi-GPU modes of the scan-matching path.  One process per GPU; plumbing is
torch.distributed (NCCL on GPUs, gloo in the CPU tests).

* `shard_range` / `shard_pairs` — config C4: independent registrations (batched loop-closure
  verification) split across ranks in contiguous, work-balanced blocks.  No data-path collective;
  results are gathered at the end.
* `ShardedFastGICP` — config C5: ONE registration against a very large target that is split into
  spatial slabs, one per rank.  Every rank keeps its slab plus a halo, handles exactly the source
  points whose transformed position falls into its slab, and the 29 partial doubles of every
  linearize (1 for compute_error, 2 for the fitness score) are summed by a single all-reduce, so
  all ranks run the same host LM loop on identical numbers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api


# --------------------------------------------------------------------------------------------- C4
def shard_range(n_items: int, world: int, rank: int, weights=None) -> tuple[int, int]:
    """Contiguous block [lo, hi) of `n_items` for `rank`; with `weights` (e.g. N_src + N_tgt per
    pair) the blocks are balanced by total weight instead of by count."""
    if weights is None:
        return (n_items * rank) // world, (n_items * (rank + 1)) // world
    w = np.asarray(weights, np.float64)
    assert len(w) == n_items
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [int(np.searchsorted(cum, cum[-1] * r / world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, n_items
    for r in range(1, world + 1):
        cuts[r] = max(cuts[r], cuts[r - 1])
    return cuts[rank], cuts[rank + 1]


def gather_results(local, world: int, group=None):
    """All ranks end up with the concatenation (rank order) of every rank's list of results."""
    if world == 1:
        return list(local)
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, list(local), group=group)
    return [x for part in out for x in part]


# --------------------------------------------------------------------------------------------- C5
def slab_boundaries(coords_1d: np.ndarray, world: int) -> np.ndarray:
    """world+1 slab edges along one axis with (approximately) equal point counts; the outer edges
    are -inf / +inf so every query position is owned by exactly one rank."""
    q = np.quantile(np.asarray(coords_1d, np.float64), np.linspace(0, 1, world + 1)[1:-1]) if world > 1 else np.array([])
    return np.concatenate([[-np.inf], q.astype(np.float32), [np.inf]]).astype(np.float32)


def slab_select(points: np.ndarray, axis: int, lo: float, hi: float, halo: float) -> np.ndarray:
    """Indices of the target points a rank stores: its slab widened by `halo` on both sides."""
    x = points[:, axis]
    return np.nonzero((x >= lo - halo) & (x < hi + halo))[0]


_REDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int)


class ShardedFastGICP(api.FastGICP):
    """fast_gicp::FastGICP against a voxel/slab-sharded target.  Exact w.r.t. the unsharded result
    (up to the fp64 summation order) provided (a) `max_correspondence_distance` is finite — the
    halo must cover it — and (b) every target point that can become a correspondence has its k
    covariance neighbours inside the stored region, i.e. its kNN radius <= `cov_halo`."""

    def __init__(self, ctx=None, device: int = 0, group=None, cov_halo: float = 3.0):
        super().__init__(ctx, device)
        import torch
        import torch.distributed as dist
        self._torch, self._dist, self._group = torch, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.cov_halo = float(cov_halo)
        self._buf = torch.zeros(32, dtype=torch.float64, device=f"cuda:{self.ctx.device}")
        self._ext = torch.cuda.ExternalStream(self.ctx.stream, device=self.ctx.device)
        self._cb = _REDUCE_FN(self._reduce)  # keep a reference: ctypes callbacks must outlive their use
        self.n_allreduce = 0
        L = api.lib()
        L.rgc_reg_set_owner_slab.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.rgc_reg_set_allreduce.argtypes = [C.c_void_p, _REDUCE_FN, C.c_void_p, C.c_void_p]
        if self.world > 1:
            self.ctx.check(L.rgc_reg_set_allreduce(self._h, self._cb, None, self._buf.data_ptr()))

    def _reduce(self, user, d_buf, n):
        try:
            with self._torch.cuda.stream(self._ext):
                self._dist.all_reduce(self._buf[:n], op=self._dist.ReduceOp.SUM, group=self._group)
            self.n_allreduce += 1
            return 0
        except Exception as e:  # noqa: BLE001 — must not propagate through the C frame
            print(f"[rgc sharded] all_reduce failed: {e}")
            return 1

    def setInputTarget(self, cloud, axis: int | None = None, boundaries=None):
        """`cloud`: the FULL target on the host (every rank passes the same array) — each rank keeps
        only its slab + halo.  `boundaries`: world+1 edges along `axis` (default: equal-count
        quantiles along the longest axis)."""
        pts = np.asarray(cloud, np.float32)
        if not np.isfinite(self._p.max_correspondence_distance) or self._p.max_correspondence_distance > 1e6:
            raise api.RgcError("a sharded target needs a finite max correspondence distance (the halo must cover it)")
        if axis is None:
            axis = int(np.argmax(pts[:, :3].max(0) - pts[:, :3].min(0)))
        if boundaries is None:
            boundaries = slab_boundaries(pts[:, axis], self.world)
        lo, hi = float(boundaries[self.rank]), float(boundaries[self.rank + 1])
        halo = float(self._p.max_correspondence_distance) + self.cov_halo
        self.local_index = slab_select(pts, axis, lo, hi, halo)
        self.local_target = np.ascontiguousarray(pts[self.local_index])
        super().setInputTarget(self.local_target)
        self.slab = (axis, lo, hi)
        big = float(np.finfo(np.float32).max)
        self.ctx.check(api.lib().rgc_reg_set_owner_slab(self._h, axis, max(lo, -big), min(hi, big)))
