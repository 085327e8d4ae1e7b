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
def slab_boundaries(coords_1d: np.ndarray, world: int, subsample: int = 1) -> np.ndarray:
    """world+1 slab edges along one axis with (approximately) equal point counts; the outer edges
    are -inf / +inf so every query position is owned by exactly one rank.  `subsample` > 1 estimates
    the quantiles from every subsample-th point (a 50 M-point map does not need all of them)."""
    c = np.asarray(coords_1d)[::max(1, int(subsample))]
    q = np.quantile(c.astype(np.float64), np.linspace(0, 1, world + 1)[1:-1]) if world > 1 else np.array([])
    return np.concatenate([[-np.inf], q.astype(np.float32), [np.inf]]).astype(np.float32)


def slab_select(points: np.ndarray, axis: int, lo: float, hi: float, halo: float) -> np.ndarray:
    """Indices of the target points a rank stores: its slab widened by `halo` on both sides (host-side
    version of rgc_reg_set_target_slab; used by the fake-shard tests)."""
    x = points[:, axis]
    return np.nonzero((x >= lo - halo) & (x < hi + halo))[0]


_REDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int)


class ShardedFastGICP(api.FastGICP):
    """fast_gicp::FastGICP against a voxel/slab-sharded target.  Exact w.r.t. the unsharded result
    (up to the fp64 summation order) provided (a) `max_correspondence_distance` is finite — the
    halo must cover it — and (b) every target point that can become a correspondence has its k
    covariance neighbours inside the stored region, i.e. its kNN radius <= `cov_halo`.

    The cross-rank sum is ONE ncclAllReduce per LM step, issued by the library on its own stream through a
    communicator it owns (include/rgc_gicp.h: rgc_comm_*); torch.distributed only carries the 128-byte NCCL
    id at construction.  `use_nccl=False` keeps the transport-agnostic callback (torch.distributed, any backend)."""

    def __init__(self, ctx=None, device: int = 0, group=None, cov_halo: float = 3.0, use_nccl: bool = True):
        super().__init__(ctx, device)
        import torch
        import torch.distributed as dist
        self._torch, self._dist, self._group = torch, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.cov_halo = float(cov_halo)
        self._comm = C.c_void_p()
        self._cb = None
        self._cb_calls = 0
        self.n_local = 0
        self.local_index = None
        self.allreduce_kind = "none (single rank)"
        L = api.lib()
        L.rgc_reg_set_owner_slab.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.rgc_reg_set_allreduce.argtypes = [C.c_void_p, _REDUCE_FN, C.c_void_p, C.c_void_p]
        L.rgc_reg_set_target_slab.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_float, C.c_float, C.c_float, C.c_uint64,
                                              C.POINTER(C.c_size_t), C.c_void_p]
        L.rgc_comm_unique_id.argtypes = [C.c_char_p]
        L.rgc_comm_create.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.rgc_comm_destroy.argtypes = [C.c_void_p]
        L.rgc_comm_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        L.rgc_reg_set_comm.argtypes = [C.c_void_p, C.c_void_p]
        L.rgc_comm_allreduce_us.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
        if self.world > 1 and use_nccl:
            buf = C.create_string_buffer(128)
            if self.rank == 0:
                self.ctx.check(L.rgc_comm_unique_id(buf))
            box = [buf.raw]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            self.ctx.check(L.rgc_comm_create(self.ctx._h, box[0], self.rank, self.world, C.byref(self._comm)))
            self.ctx.check(L.rgc_reg_set_comm(self._h, self._comm))
            ver = C.c_int(0)
            L.rgc_comm_info(self._comm, None, None, None, C.byref(ver))
            v = ver.value
            L.rgc_comm_transport.argtypes = [C.c_void_p]
            if L.rgc_comm_transport(self._comm) == 1:
                self.allreduce_kind = ("k_peer_allreduce: one 1-block launch per LM step over NVLink peer memory (CUDA-IPC mailboxes on every rank; "
                                       "store partials + sequence tag to all ranks, poll, sum in rank order, publish to the host) — compute_error + "
                                       "look-ahead linearize summed together")
            else:
                self.allreduce_kind = (f"ncclAllReduce (NCCL {v // 10000}.{(v // 100) % 100}.{v % 100}) issued by the library on its stream, "
                                       "one per LM step (compute_error + look-ahead linearize summed together)")
        elif self.world > 1:
            self._buf = torch.zeros(32, dtype=torch.float64, device=f"cuda:{self.ctx.device}")
            self._ext = torch.cuda.ExternalStream(self.ctx.stream, device=self.ctx.device)
            self._cb = _REDUCE_FN(self._reduce)  # keep a reference: ctypes callbacks must outlive their use
            self.ctx.check(L.rgc_reg_set_allreduce(self._h, self._cb, None, self._buf.data_ptr()))
            self.allreduce_kind = "torch.distributed.all_reduce through a host callback (two per LM step)"

    @property
    def n_allreduce(self) -> int:
        if self._comm:
            n = C.c_uint64(0)
            api.lib().rgc_comm_info(self._comm, None, None, C.byref(n), None)
            return int(n.value)
        return self._cb_calls

    def allreduce_us(self, n_doubles: int = 30, reps: int = 50):
        """mean latency of the LM loop's all-reduce (collective: every rank calls it); None without a communicator"""
        if not self._comm:
            return None
        us = C.c_float(0.0)
        self.ctx.check(api.lib().rgc_comm_allreduce_us(self._comm, n_doubles, reps, C.byref(us)))
        return float(us.value)

    def _reduce(self, user, d_buf, n):
        try:
            with self._torch.cuda.stream(self._ext):
                self._dist.all_reduce(self._buf[:n], op=self._dist.ReduceOp.SUM, group=self._group)
            self._cb_calls += 1
            return 0
        except Exception as e:  # noqa: BLE001 — must not propagate through the C frame
            print(f"[rgc sharded] all_reduce failed: {e}")
            return 1

    def _halo(self) -> float:
        if not np.isfinite(self._p.max_correspondence_distance) or self._p.max_correspondence_distance > 1e6:
            raise api.RgcError("a sharded target needs a finite max correspondence distance (the halo must cover it)")
        return float(self._p.max_correspondence_distance) + self.cov_halo

    def setInputTarget(self, cloud, axis: int | None = None, boundaries=None, want_index: bool = False, force: bool = False):
        """`cloud`: the FULL target on the host (every rank passes the same array, numpy or a pinned torch
        tensor) — each rank uploads it once and keeps only its slab + halo, selected on the device.
        `boundaries`: world+1 edges along `axis` (default: equal-count quantiles of a 1/64 subsample along
        the longest axis)."""
        ptr, n, stride, on_dev, keep = api._as_cloud(cloud)
        if on_dev:
            raise ValueError("the full target is a host cloud")
        pts = cloud if hasattr(cloud, "numpy") else np.asarray(cloud)
        halo = self._halo()
        if axis is None or boundaries is None:
            xyz = pts.numpy() if hasattr(pts, "numpy") else pts
            sub = xyz[::64, :3]
            if axis is None:
                axis = int(np.argmax(sub.max(0) - sub.min(0)))
            if boundaries is None:
                boundaries = slab_boundaries(sub[:, axis], self.world)
        lo, hi = float(boundaries[self.rank]), float(boundaries[self.rank + 1])
        m = C.c_size_t(0)
        idx = np.empty(n, np.int32) if want_index else None
        self.ctx.check(api.lib().rgc_reg_set_target_slab(self._h, ptr, n, stride, int(axis), lo, hi, halo, 0 if force else id(cloud), C.byref(m),
                                                         None if idx is None else idx.ctypes.data))
        self.n_local = int(m.value)
        self.local_index = None if idx is None else idx[:self.n_local].copy()
        self._tgt, self._tgt_id, self._n_tgt = keep, (id(cloud), cloud), self.n_local
        self.slab = (int(axis), lo, hi)

    def setInputTargetLocal(self, local_points, axis: int, lo: float, hi: float):
        """this rank already holds its share (slab + halo) of the target: no selection, just the ownership slab"""
        super().setInputTarget(local_points)
        self.n_local = self._n_tgt
        big = float(np.finfo(np.float32).max)
        self.ctx.check(api.lib().rgc_reg_set_owner_slab(self._h, int(axis), max(float(lo), -big), min(float(hi), big)))
        self.slab = (int(axis), float(lo), float(hi))

    def close(self):
        if self._comm:
            if self._h:
                api.lib().rgc_reg_set_comm(self._h, None)
            api.lib().rgc_comm_destroy(self._comm)
            self._comm = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
        super().__del__()
