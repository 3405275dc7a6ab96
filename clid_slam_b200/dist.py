"""Multi-GPU plumbing for the mapping step: one process per GPU, torch.distributed (NCCL over
NVLink 5 / NVSwitch on the box, gloo on CPU for the host-logic tests).

The reference is single-process, so this is new functionality whose contract
is "N ranks on one global batch == 1 rank on the same batch" (up to fp32 summation order).
Every sample's forward / loss / backward only reads map state, so the batch is sharded across
ranks; the reductions the reference does over the whole batch are completed with all-reduces.

Two sharding modes (FusedTrainer.iteration):

* replicated (`sync=True`, no shards): any sample may sit on any rank.  Per iteration one flat
  all-reduce of [decoder grads | loss] plus a dense all-reduce of the replicated feature gradient
  (35 MB at 1 M points) -- correct but bandwidth-bound; kept as the simple reference mode.

* spatial (`shards=SpatialShards(...)`): space is cut into slabs along one axis, a sample belongs
  to the rank that owns the slab of its voxel.  A sample only touches neural points within
  `reach` voxels, so feature rows are private to their slab's rank except for a thin band around
  every slab boundary (`shared_rows`).  Per iteration a 3 kB all-reduce carries [decoder grads | loss]
  and every rank completes the gradients of its two bands with its slab neighbours (`NeighbourExchange`,
  grouped send/recv); when bands overlap (slabs narrower than two bands) ONE flat all-reduce carries
  [decoder grads | loss | gradients of all shared rows] instead.  Ranks then apply the identical Adam
  step to the rows they share and their own step to their private rows.  Other ranks' private rows go stale
  locally but are never read; `SpatialShards.gather_features` re-replicates the table once per
  mapping() call.  Certainty / ts_update side effects are likewise reduced once per call
  (`reduce_side_effects`).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of an n-sample batch owned by `rank`; sizes differ by at
    most one and the slices tile [0, n) in rank order."""
    base, extra = divmod(n, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def decimated_count(begin: int, end: int, decimation: int) -> int:
    """How many global indices i with i % decimation == 0 fall into [begin, end)."""
    first = -(-begin // decimation) * decimation
    return 0 if first >= end else (end - 1 - first) // decimation + 1


class FlatAllReduce:
    """Sum-all-reduce several tensors through one flat buffer (one collective launch)."""

    def __init__(self, tensors: Sequence[Optional[torch.Tensor]], group=None):
        self.tensors = [t for t in tensors if t is not None]
        self.group = group
        total = sum(t.numel() for t in self.tensors)
        ref = self.tensors[0]
        self.flat = torch.empty(total, dtype=ref.dtype, device=ref.device)

    def __call__(self) -> None:
        if world()[1] == 1:
            return
        off = 0
        for t in self.tensors:
            self.flat[off:off + t.numel()].copy_(t.reshape(-1))
            off += t.numel()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        off = 0
        for t in self.tensors:
            t.copy_(self.flat[off:off + t.numel()].view_as(t))
            off += t.numel()


def all_reduce_sum(t: Optional[torch.Tensor], group=None) -> None:
    if t is not None and world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def all_reduce_max(t: Optional[torch.Tensor], group=None) -> None:
    if t is not None and world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)


def reduce_side_effects(certainty: torch.Tensor, certainty_before: torch.Tensor, ts_update: torch.Tensor,
                        group=None) -> None:
    """After a sharded mapping() call: certainty increments are summed over ranks, ts_update takes
    the maximum -- the result every rank would hold had it processed the whole batch."""
    if world()[1] == 1:
        return
    delta = certainty - certainty_before
    dist.all_reduce(delta, op=dist.ReduceOp.SUM, group=group)
    certainty.copy_(certainty_before + delta)
    dist.all_reduce(ts_update, op=dist.ReduceOp.MAX, group=group)


def slab_boundaries(cell: torch.Tensor, world_size: int) -> torch.Tensor:
    """[world-1] strictly increasing int64 cell coordinates that cut `cell` (the slab-axis voxel coordinate of
    every point) into world_size slabs of equal point counts."""
    dev = cell.device
    if world_size <= 1 or cell.numel() == 0:
        return torch.empty(0, dtype=torch.int64, device=dev)
    q = torch.arange(1, world_size, device=dev, dtype=torch.float32) / world_size
    srt = torch.sort(cell).values
    pick = (q * (srt.numel() - 1)).long()
    # a clustered map can repeat a quantile: keep exactly world-1 strictly increasing boundaries by bumping
    # repeats (the slabs in between are then empty, which is valid)
    bnd = srt[pick].tolist()
    for i in range(1, len(bnd)):
        if bnd[i] <= bnd[i - 1]:
            bnd[i] = bnd[i - 1] + 1
    return torch.tensor(bnd, dtype=torch.int64, device=dev)


def axis_cells(points: torch.Tensor, resolution: float, axis: int) -> torch.Tensor:
    from .utils.tools import ieee_div

    return torch.floor(ieee_div(points[:, axis], float(resolution))).to(torch.int64)


def partition_mask(points: torch.Tensor, resolution: float, axis: int, boundaries: torch.Tensor, rank: int,
                   band: int) -> torch.Tensor:
    """Points a rank of a PARTITIONED map holds: those of its own slab [b_rank, b_rank+1) plus the neighbours'
    halves of its two boundary bands (`band` = reach + margin cells): every neural point a sample of the slab can
    reach.  The selection is by voxel cell, so a voxel is held completely or not at all and every rank that holds
    it derives the same neural point from it."""
    cell = axis_cells(points, resolution, axis)
    b = boundaries.tolist()
    world_size = len(b) + 1
    keep = torch.ones(cell.shape, dtype=torch.bool, device=cell.device)
    if rank > 0:
        keep &= cell >= b[rank - 1] - band
    if rank < world_size - 1:
        keep &= cell <= b[rank] + band - 1
    return keep


def hash_owner_mask(slots: torch.Tensor, buffer_size: int) -> torch.Tensor:
    """Which of the points inserted in this order end up OWNING their voxel-hash slot (the last writer of a slot
    wins, NeuralPoints._store_slots): the reference reaches a neural point only through its slot, so a point that
    lost its slot to a colliding voxel is dead weight.  A partitioned map is cut from the owners only -- with the
    losers dropped on every rank, each rank's private hash table answers every probe exactly like the one global
    table of a single process would (a collision whose two voxels sit on different ranks would otherwise keep
    both alive)."""
    slots = torch.remainder(slots, int(buffer_size))
    srt, perm = torch.sort(slots, stable=True)
    last = torch.ones(srt.shape, dtype=torch.bool, device=slots.device)
    last[:-1] = srt[1:] != srt[:-1]
    keep = torch.zeros(slots.shape, dtype=torch.bool, device=slots.device)
    keep[perm[last]] = True
    return keep


def voxel_keys(points: torch.Tensor, resolution: float) -> torch.Tensor:
    """One int64 per point that identifies its voxel (21 bits per axis): the rank-independent name of a neural
    point of a partitioned map."""
    from .utils.tools import ieee_div

    c = torch.floor(ieee_div(points, float(resolution))).to(torch.int64) + (1 << 20)
    return (c[:, 0] << 42) | (c[:, 1] << 21) | c[:, 2]


class SpatialShards:
    """Slab partition of the local map along one axis, in voxel units.

    boundaries  [world-1] int64 cell coordinates b_1 < ... ; rank r owns cells in [b_r, b_{r+1})
    shared_rows rows of the local feature table whose voxel lies within `reach + margin` cells of
                a boundary: the only rows that can receive gradient from two ranks
    row_owner   [rows] rank that owns each local row (shared rows included: they have one owner for
                the final gather, although every rank keeps them up to date)
    The margin (default 1 voxel) covers the numerical-gradient probes, which are shifted by a
    fraction of a voxel and may fall into the next cell.
    """

    def __init__(self, points: torch.Tensor, resolution: float, reach: int, world_size: int, axis: Optional[int] = None,
                 margin: int = 1, boundaries: Optional[torch.Tensor] = None, pad_rows: int = 1):
        self.resolution = float(resolution)
        self.world_size = int(world_size)
        dev = points.device
        if axis is None:  # cut across the longest extent of the map
            ext = points.amax(0) - points.amin(0)
            axis = int(torch.argmax(ext).item())
        self.axis = axis
        from .utils.tools import ieee_div

        cell = torch.floor(ieee_div(points[:, axis], self.resolution)).to(torch.int64)
        if boundaries is None:
            boundaries = slab_boundaries(cell, world_size)
        self.boundaries = boundaries.to(device=dev, dtype=torch.int64).contiguous()
        band = reach + margin
        self.band = int(band)
        shared = torch.zeros(cell.shape[0], dtype=torch.bool, device=dev)
        for b in self.boundaries.tolist():
            shared |= (cell >= b - band) & (cell <= b + band - 1)
        owner = torch.searchsorted(self.boundaries, cell, right=True)
        pad = torch.zeros(pad_rows, dtype=torch.bool, device=dev)  # the feature table's padding row
        self.shared_mask = torch.cat((shared, pad))
        self.shared_rows = torch.nonzero(self.shared_mask).flatten()
        # rows in the band of each boundary (boundary j separates ranks j and j + 1).  When no row lies in
        # two bands (slabs wider than two bands) a shared row has exactly two contributors and its
        # gradient can be completed by a neighbour exchange instead of an all-reduce over every band.
        self.band_rows = []
        hits = torch.zeros(cell.shape[0], dtype=torch.int32, device=dev)
        for b in self.boundaries.tolist():
            in_band = (cell >= b - band) & (cell <= b + band - 1)
            hits += in_band.to(torch.int32)
            self.band_rows.append(torch.nonzero(in_band).flatten())
        self.pairwise = bool(int(hits.max().item()) <= 1) if cell.numel() else True
        self.row_owner = torch.cat((owner, torch.zeros(pad_rows, dtype=owner.dtype, device=dev)))

    def neighbour_rows(self, rank: int):
        """(rows shared with rank - 1 | None, rows shared with rank + 1 | None)."""
        left = self.band_rows[rank - 1] if rank > 0 else None
        right = self.band_rows[rank] if rank < len(self.band_rows) else None
        return left, right

    def owner_of(self, x: torch.Tensor) -> torch.Tensor:
        """Rank that processes each sample of x [n,3]."""
        from .utils.tools import ieee_div

        cell = torch.floor(ieee_div(x[:, self.axis], self.resolution)).to(torch.int64)
        return torch.searchsorted(self.boundaries, cell, right=True)

    def gather_features(self, features: torch.Tensor, rank: int, group=None) -> None:
        """Re-replicate the feature table after training: every row is taken from its owner."""
        if world()[1] == 1:
            return
        mine = (self.row_owner == rank).unsqueeze(1)
        contrib = torch.where(mine, features, torch.zeros_like(features))
        dist.all_reduce(contrib, op=dist.ReduceOp.SUM, group=group)
        features.copy_(contrib)


def peer_row_tables(shards: SpatialShards, rank: int, points: torch.Tensor, rows_total: int, group=None):
    """Row translation of a PARTITIONED map (ClidTrainFusedArgs.peer_row): (lower, upper) int32 [rows_total] tables
    giving, for every local row in the band shared with the lower / upper slab neighbour, that neural point's row
    in the neighbour's table (-1 elsewhere); None on the sides without a neighbour.  `points` is this rank's
    local neural-point table.  Both sides of a boundary hold exactly the voxels of its band, so sorting the band
    rows by voxel key gives both ranks the same order; the row lists travel once through torch.distributed."""
    _, world_size = world()
    left, right = shards.neighbour_rows(rank)
    mine = []
    for rows in (left, right):
        if rows is None:
            mine.append(None)
            continue
        key = voxel_keys(points[rows], shards.resolution)
        order = torch.argsort(key)
        mine.append((key[order].cpu(), rows[order].to(torch.int32).cpu()))
    table = [None] * world_size
    if world_size > 1:
        dist.all_gather_object(table, mine, group=group)
    else:
        table[0] = mine
    out = []
    for side, peer in ((0, rank - 1), (1, rank + 1)):
        if mine[side] is None or peer < 0 or peer >= world_size:
            out.append(None)
            continue
        theirs = table[peer][1 - side]  # my lower band is the neighbour's upper band
        if theirs is None or not torch.equal(theirs[0], mine[side][0]):
            raise RuntimeError(f"rank {rank} and rank {peer} hold different voxels in the band they share: the "
                               "partitions were not cut from the same map")
        tab = torch.full((rows_total,), -1, dtype=torch.int32, device=points.device)
        tab[mine[side][1].to(points.device).long()] = theirs[1].to(points.device)
        out.append(tab)
    return tuple(out)


def exchange_band_values(shards: SpatialShards, rank: int, points: torch.Tensor, values: torch.Tensor, op: str,
                         group=None) -> None:
    """Once per mapping() call on a PARTITIONED map: complete the per-row side effects of the band rows with the
    slab neighbours' (op 'sum': certainty increments, 'max': ts_update).  `values` [rows] is updated in place."""
    _, world_size = world()
    if world_size == 1:
        return
    left, right = shards.neighbour_rows(rank)
    mine = []
    for rows in (left, right):
        if rows is None:
            mine.append(None)
            continue
        order = torch.argsort(voxel_keys(points[rows], shards.resolution))
        mine.append((rows[order], values[rows[order]].cpu()))
    table = [None] * world_size
    dist.all_gather_object(table, [None if m is None else m[1] for m in mine], group=group)
    for side, peer in ((0, rank - 1), (1, rank + 1)):
        if mine[side] is None or peer < 0 or peer >= world_size or table[peer][1 - side] is None:
            continue
        rows, theirs = mine[side][0], table[peer][1 - side].to(values.device)
        if op == "sum":
            values[rows] += theirs
        else:
            values[rows] = torch.maximum(values[rows], theirs)


class NeighbourExchange:
    """Completes the gradients of a rank's shared band rows with its two slab neighbours: every
    boundary band has exactly two contributors (SpatialShards.pairwise), so each rank sends its
    partial rows to the neighbour across the boundary, receives the neighbour's and adds them --
    two grouped NCCL send/recv pairs of one band each (~0.4 MB at 1 M points), independent of
    the number of ranks, instead of an all-reduce over the bands of all N - 1 boundaries.

        ex.pack(grad)      gather the band rows into the send buffers      (graph-capturable)
        ex.exchange()      grouped isend / irecv with rank - 1 and rank + 1 (eager)
        ex.unpack(grad)    grad[band rows] += received partials             (graph-capturable)
    """

    def __init__(self, shards: SpatialShards, rank: int, like: torch.Tensor, group=None):
        if not shards.pairwise:
            raise ValueError("slabs are narrower than two bands: use the flat all-reduce")
        self.rank, self.group = int(rank), group
        self._ops = None
        self.sides = []  # (peer, rows, send, recv)
        left, right = shards.neighbour_rows(rank)
        for peer, rows in ((rank - 1, left), (rank + 1, right)):
            if rows is not None and rows.numel() > 0:
                send = torch.zeros(rows.numel(), like.shape[1], dtype=like.dtype, device=like.device)
                self.sides.append((peer, rows, send, torch.zeros_like(send)))

    def rows(self) -> torch.Tensor:
        return torch.cat([r for _, r, _, _ in self.sides]) if self.sides else torch.empty(0, dtype=torch.int64)

    def pack(self, grad: torch.Tensor) -> None:
        for _, rows, send, _ in self.sides:
            torch.index_select(grad, 0, rows, out=send)

    def exchange(self) -> None:
        if world()[1] == 1 or not self.sides:
            return
        if self._ops is None:  # the buffers are static: build the op list once
            self._ops = []
            for peer, _, send, recv in self.sides:
                self._ops.append(dist.P2POp(dist.isend, send, peer, self.group))
                self._ops.append(dist.P2POp(dist.irecv, recv, peer, self.group))
        for req in dist.batch_isend_irecv(self._ops):
            req.wait()

    def unpack(self, grad: torch.Tensor) -> None:
        for _, rows, _, recv in self.sides:
            grad.index_add_(0, rows, recv)


class PeerLink:
    """Peer-mapped buffers of all ranks of one box (CUDA IPC over NVLink / NVSwitch) for the fused gradient
    exchange of the spatially sharded step -- no NCCL call on the step's data path:

      grad[2]   [rows, F]  feature-gradient tables, ping-pong by step parity.  clid_train_fused adds the
                contributions to a boundary-band row into the slab neighbour's table as well
                (ClidTrainFusedArgs.peer_grad, red.global.add over NVLink); a rank can run at most one step ahead
                of its neighbours (the flag wait below), so two tables make a second barrier unnecessary.
      slots     [world, stride] + flags [world]: one-shot all-reduce of [decoder gradients | loss]
                (clid_peer_publish / clid_peer_reduce), whose flag wait is also the barrier that orders the
                neighbours' remote adds before this rank's optimiser step.

    Everything lives in ONE torch allocation per rank; its cudaIpcMemHandle (torch's `_share_cuda_`) travels through
    torch.distributed once, and every other rank opens it with ITS device current (clid_ipc_open), so the mapping
    belongs to the importing device with lazy peer access to the exporter -- memory imported under another
    device's context is not reachable from this device's kernels."""

    def __init__(self, rows: int, feat_dim: int, n_small: int, device: torch.device, group=None,
                 capacity_rows: Optional[int] = None):
        import ctypes as C

        from . import _lib

        self.rank, self.world = world()
        if self.world > 8:
            raise ValueError("PeerLink covers one box (<= 8 GPUs)")
        self.device = torch.device(device)
        self.stride = (int(n_small) + 31) // 32 * 32
        # every rank lays its allocation out alike (a rank addresses its peers' buffers by offset): with a PARTITIONED
        # map the ranks' tables differ in length, so the layout follows the longest one
        # capacity_rows > rows: room for view_rows() to follow a table that changes from call to call (Mapper)
        layout_rows = max(int(rows), int(capacity_rows or 0))
        if self.world > 1:
            all_rows = [None] * self.world
            dist.all_gather_object(all_rows, layout_rows, group=group)
            layout_rows = max(all_rows)
        self.capacity_rows, self.feat_dim, self.group = layout_rows, int(feat_dim), group
        grad_bytes = (layout_rows * feat_dim * 4 + 255) // 256 * 256
        slots_bytes = (max(self.world, 1) * self.stride * 4 + 255) // 256 * 256
        self._off = {"grad0": 0, "grad1": grad_bytes, "slots": 2 * grad_bytes, "flags": 2 * grad_bytes + slots_bytes}
        total = 2 * grad_bytes + slots_bytes + 256
        lib = _lib.load()
        # one zero-filled cudaMalloc allocation owned by the library (clid_peer_alloc): its base pointer and IPC handle
        # are exact (a block of torch's caching allocator has neither), torch sees it through the CUDA array interface
        base_ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        with torch.cuda.device(self.device):
            _lib.check(lib.clid_peer_alloc(total, C.byref(base_ptr), handle), "clid_peer_alloc")
        self._base = int(base_ptr.value)

        class _Raw:  # zero-copy torch view of the allocation
            __cuda_array_interface__ = {"shape": (total,), "typestr": "|u1", "data": (self._base, False), "version": 3}

        self.buf = torch.as_tensor(_Raw(), device=self.device)
        assert self.buf.data_ptr() == self._base

        def view(name, nbytes, dtype, shape):
            o = self._off[name]
            return self.buf[o:o + nbytes].view(dtype).view(shape)

        self.grad = [view("grad0", rows * feat_dim * 4, torch.float32, (rows, feat_dim)),
                     view("grad1", rows * feat_dim * 4, torch.float32, (rows, feat_dim))]
        self.slots = view("slots", max(self.world, 1) * self.stride * 4, torch.float32, (max(self.world, 1), self.stride))
        self.flags = view("flags", 64, torch.int32, (16,))
        self.epoch = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.error = torch.zeros(1, dtype=torch.int32, device=self.device)
        torch.cuda.synchronize(self.device)
        self.base_of = {self.rank: self._base}
        self._opened = []
        if self.world > 1:
            table = [None] * self.world
            dist.all_gather_object(table, bytes(handle.raw), group=group)
            with torch.cuda.device(self.device):
                for r, h in enumerate(table):
                    if r == self.rank:
                        continue
                    ptr = C.c_void_p()
                    _lib.check(lib.clid_ipc_open(h, C.byref(ptr)), f"clid_ipc_open (rank {r})")
                    self._opened.append(ptr)
                    self.base_of[r] = int(ptr.value)
            torch.cuda.synchronize(self.device)
            dist.barrier(group=group)  # nobody publishes before everybody has opened everything

    def view_rows(self, rows: int) -> None:
        """Re-shape the two gradient tables for a feature table of `rows` rows (<= capacity_rows) and clear them.
        Collective: every rank calls it between steps (the barrier keeps a neighbour's late remote adds of the
        previous call out of the cleared tables)."""
        if rows > self.capacity_rows:
            raise ValueError(f"{rows} rows exceed the link's capacity of {self.capacity_rows}")
        if self.world > 1:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
        nbytes = rows * self.feat_dim * 4
        self.grad = [self.buf[self._off[k]:self._off[k] + nbytes].view(torch.float32).view(rows, self.feat_dim)
                     for k in ("grad0", "grad1")]
        for g in self.grad:
            g.zero_()
        if self.world > 1:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)

    def close(self) -> None:
        """Unmap the peers' allocations and free this rank's (collective: nobody may still be stepping)."""
        from . import _lib

        lib = _lib.load()
        if self.world > 1:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            for ptr in self._opened:
                lib.clid_ipc_close(ptr)
            self._opened = []
            if self._base:
                self.grad, self.slots, self.flags, self.buf = [], None, None, None
                lib.clid_peer_free(self._base)
                self._base = 0

    def ptr(self, rank: int, name: str) -> int:
        """Device address of buffer `name` ('grad0' | 'grad1' | 'slots' | 'flags') of `rank`, valid on this device."""
        return self.base_of[rank] + self._off[name]

    def grad_ptr(self, rank: int, parity: int) -> int:
        import os

        if os.environ.get("CLID_PEER_DEBUG_SELF") == "1":  # developer knob: exercise the kernel path without NVLink
            rank = self.rank
        return self.ptr(rank, "grad1" if parity else "grad0")

    def args(self, n0: int, n1: int):
        from . import _lib

        a = _lib.ClidPeerArgs()
        for r in range(self.world):
            a.slots_of[r] = self.ptr(r, "slots")
            a.flags_of[r] = self.ptr(r, "flags")
        a.epoch, a.error = self.epoch.data_ptr(), self.error.data_ptr()
        a.rank, a.world, a.n0, a.n1, a.stride, a.timeout_ms = self.rank, self.world, int(n0), int(n1), self.stride, 2000
        return a

    def check(self) -> None:
        if int(self.error.item()) != 0:
            raise RuntimeError("clid_peer_reduce timed out waiting for a peer rank (a rank died or the ranks ran a "
                               "different number of steps)")
