"""Brick index: a compact per-frame voxel index that answers the hash probes of
NeuralPoints.radius_neighborhood_search (model/neural_points.py:971-1030) without touching the
400 MB `buffer_pt_index` table.

Why it is exact.  A probe of cell C returns v = table[hash(C)].  By construction of the table
(update / recreate_hash) hash(cell(v)) == hash(C).  If cell(v) != C the point sits in an aliasing
cell C' = C + d with  d . primes == 0 (mod buffer_size),  d != 0.  A candidate is only kept when
|p_v - x|^2 <= max_valid_dist2 = 3 ((n+1) res)^2, which bounds |C' - Q|_inf <= floor(sqrt(3)(n+1)+1)
for the query cell Q, hence |d|_inf <= floor(sqrt(3)(n+1)+1) + n.  `hash_is_alias_free` checks by
enumeration that no such d exists for the (primes, buffer_size) in use; then every surviving
candidate of cell C is the point that (a) lies in C and (b) owns C's slot -- exactly what the index
stores, with the per-point predicates (travel-distance window, global->local remap) folded in at
build time.  When the check fails (tiny tables) callers stay on the hashed kernel.

The index is rebuilt lazily whenever the map changes (per frame), with torch ops on the device.
"""
from __future__ import annotations

import math
import os
from functools import lru_cache
from typing import Optional

import numpy as np
import torch

from .. import _lib

_PRIMES = (73856093, 19349669, 83492791)
MAX_BRICKS = 1 << 26  # dense header array cap (1 GiB); larger boxes use the hashed kernel


@lru_cache(maxsize=64)
def hash_is_alias_free(buffer_size: int, reach: int) -> bool:
    bound = int(math.floor(math.sqrt(3.0) * (reach + 1) + 1)) + reach
    r = np.arange(-bound, bound + 1, dtype=np.int64)
    d = np.stack(np.meshgrid(r, r, r, indexing="ij"), -1).reshape(-1, 3)
    h = (d * np.array(_PRIMES, dtype=np.int64)).sum(-1) % int(buffer_size)
    return int((h == 0).sum()) == 1  # only d == 0


_STENCILS = {}

# why a query could not use the brick index and ran on the (20x slower) probe of the reference's hash table:
# reason -> number of index builds that gave up.  Each reason warns once per process.
FALLBACKS: dict = {}


def _fallback(reason: str):
    import warnings

    if reason not in FALLBACKS:
        warnings.warn(f"clid_slam_b200: brick index unavailable ({reason}); queries fall back to the hash-table probe, "
                      "which is ~20x slower (DESIGN.md section 3)", RuntimeWarning, stacklevel=3)
    FALLBACKS[reason] = FALLBACKS.get(reason, 0) + 1
    return None


def _stencil_table(offsets_cpu: torch.Tensor, reach: int, span: int) -> torch.Tensor:
    key = (offsets_cpu.numpy().tobytes(), reach, span)
    if key not in _STENCILS:
        _STENCILS[key] = _make_stencil_table(offsets_cpu, reach, span)
    return _STENCILS[key]


_STENCILS_DEV = {}


def _stencil_on(dev, offsets_cpu: torch.Tensor, reach: int, span: int) -> torch.Tensor:
    """Device copy of the stencil table, uploaded once per (neighbourhood, device)."""
    key = (offsets_cpu.numpy().tobytes(), reach, span, str(dev))
    if key not in _STENCILS_DEV:
        _STENCILS_DEV[key] = _stencil_table(offsets_cpu, reach, span).to(dev)
    return _STENCILS_DEV[key]


def _make_stencil_table(offsets_cpu: torch.Tensor, reach: int, span: int) -> torch.Tensor:
    inside = {tuple(int(v) for v in row) for row in offsets_cpu.tolist()}
    table = np.zeros((64, span**3), dtype=np.uint64)
    width = 2 * reach
    for lz in range(4):
        for ly in range(4):
            for lx in range(4):
                pos = (lz * 4 + ly) * 4 + lx
                for dz in range(span):
                    for dy in range(span):
                        for dx in range(span):
                            bits = 0
                            for k in range(4):
                                tz = 4 * dz + k - lz
                                if tz < 0 or tz > width:
                                    continue
                                for j in range(4):
                                    ty = 4 * dy + j - ly
                                    if ty < 0 or ty > width:
                                        continue
                                    for i in range(4):
                                        tx = 4 * dx + i - lx
                                        if tx < 0 or tx > width:
                                            continue
                                        if (tx - reach, ty - reach, tz - reach) in inside:
                                            bits |= 1 << (i + 4 * j + 16 * k)
                            table[pos, (dz * span + dy) * span + dx] = bits
    return torch.from_numpy(table.view(np.int64).reshape(-1).copy())


class BrickIndex:
    """Device tensors of one index plus the ClidBricks struct that points at them."""

    def __init__(self, headers, records, stencil, origin, dims, span, reach, apron=0, hood=None):
        self.headers, self.records, self.stencil, self.hood = headers, records, stencil, hood
        s = _lib.ClidBricks()
        s.headers = headers.data_ptr()
        s.hood = None if hood is None else hood.data_ptr()
        s.records = records.data_ptr()
        s.stencil = stencil.data_ptr()
        for i in range(3):
            s.origin[i] = int(origin[i])
            s.dims[i] = int(dims[i])
        s.span, s.reach, s.n_records = int(span), int(reach), int(records.shape[0])
        s.apron = int(apron)
        self.struct = s
        self.n_bricks = int(dims[0]) * int(dims[1]) * int(dims[2])

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.headers, self.records, self.stencil, self.hood)
                   if t is not None)


def build(npm, query_locally: bool) -> Optional[BrickIndex]:
    """Index of the points a query of `npm` can return; None when the fast path is not exact or
    not applicable (then the hashed kernel is used)."""
    offsets_cpu = getattr(npm, "_neighbor_dx_cpu", None)
    if offsets_cpu is None or tuple(offsets_cpu.shape) != tuple(npm.neighbor_dx.shape):
        offsets_cpu = npm.neighbor_dx.detach().cpu()  # a map whose table was set from outside (tests, old pickles)
    reach = int(offsets_cpu.abs().max().item()) if offsets_cpu.numel() else 0
    span = (2 * reach + 7) // 4
    if npm.count() == 0:
        return None
    if span != 2:  # the kernels walk exactly 2x2x2 bricks (num_nei_cells 1 or 2)
        return _fallback(f"num_nei_cells = {reach}: a neighbourhood must span 2 bricks per axis") if reach > 2 else None
    if not hash_is_alias_free(int(npm.buffer_size), reach):
        return _fallback(f"hash table of {int(npm.buffer_size)} slots aliases cells inside a neighbourhood")
    dev = npm.neural_points.device
    if dev.type == "cuda" and NATIVE_BUILD:
        return _build_native(npm, query_locally, offsets_cpu, reach, span)
    return _build_torch(npm, query_locally, offsets_cpu, reach, span)


NATIVE_BUILD = os.environ.get("CLID_NATIVE_BRICK_BUILD", "1") != "0"


def _build_native(npm, query_locally: bool, offsets_cpu, reach: int, span: int) -> Optional[BrickIndex]:
    """The index through clid_brick_keep / clid_brick_keys / clid_brick_fill (csrc/feeder.cuh) around one torch.sort:
    four launches and ONE 28-byte read-back (the bounding box sizes the dense header array) per rebuild."""
    import ctypes as C

    from . import query as _q

    dev = npm.neural_points.device
    lib = _lib.load()
    if query_locally:
        pts = npm.local_neural_points.contiguous()
        gids = getattr(npm, "_local_gids", None)
        if gids is None or gids.shape[0] != pts.shape[0] or gids.device != dev:
            gids = torch.nonzero(npm.local_mask[:-1]).flatten()
        gids = gids.contiguous()
        gid_ptr = gids.data_ptr()
    else:
        pts = npm.neural_points.contiguous()
        gids, gid_ptr = None, None
    n = pts.shape[0]
    if n == 0:
        return None
    m, _ = _q.map_struct(npm, query_locally)
    time_filter = bool(query_locally and npm.temporal_local_map_on)
    ts_ptr = _lib.ptr(npm.point_ts_create, torch.int32, "point_ts_create") if time_filter else None
    cells = torch.empty(n, 3, dtype=torch.int32, device=dev)
    keep = torch.empty(n, dtype=torch.uint8, device=dev)
    i32 = torch.iinfo(torch.int32)
    bbox = torch.tensor([i32.max] * 3 + [i32.min] * 3 + [0], dtype=torch.int32, device=dev)
    stream = _lib.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.clid_brick_keep(C.byref(m), pts.data_ptr(), gid_ptr, n, ts_ptr, cells.data_ptr(), keep.data_ptr(),
                                       bbox.data_ptr(), stream), "clid_brick_keep")
    bb = bbox.cpu().tolist()  # the one synchronisation of a rebuild
    n_kept = bb[6]
    if n_kept == 0:
        return None
    lo = [bb[a] - 4 for a in range(3)]  # one empty brick all around (ClidBricks.apron)
    dims = [(bb[3 + a] + 4 - lo[a]) // 4 + 1 for a in range(3)]
    n_bricks = dims[0] * dims[1] * dims[2]
    if n_bricks > MAX_BRICKS:
        return _fallback(f"bounding box of {dims} bricks exceeds the dense header cap of {MAX_BRICKS}")
    lo_c, dims_c = (C.c_int32 * 3)(*lo), (C.c_int32 * 3)(*dims)
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.clid_brick_keys(cells.data_ptr(), keep.data_ptr(), n, lo_c, dims_c, keys.data_ptr(), stream), "clid_brick_keys")
    sorted_keys, order = torch.sort(keys)  # kept points first, in (brick, cell) order
    records = torch.empty(n_kept, 4, dtype=torch.float32, device=dev)
    headers = torch.empty(n_bricks, 4, dtype=torch.int32, device=dev)
    hood = torch.empty(n_bricks, 32, dtype=torch.int32, device=dev) if USE_HOOD and n_bricks <= MAX_HOOD_BRICKS else None
    with torch.cuda.device(dev):
        _lib.check(lib.clid_brick_fill(sorted_keys.data_ptr(), order.data_ptr(), n_kept, pts.data_ptr(), dims_c,
                                       records.data_ptr(), headers.data_ptr(), None if hood is None else hood.data_ptr(),
                                       stream), "clid_brick_fill")
    return BrickIndex(headers, records, _stencil_on(dev, offsets_cpu, reach, span), lo, dims, span, reach, apron=1, hood=hood)


def _build_torch(npm, query_locally: bool, offsets_cpu, reach: int, span: int) -> Optional[BrickIndex]:
    """The same index with torch ops (host logic for CPU maps and the cross-check of the native build)."""
    dev = npm.neural_points.device
    res = float(npm.resolution)
    primes = npm.primes

    if query_locally:
        gids = torch.nonzero(npm.local_mask[:-1]).flatten()
        pts = npm.local_neural_points
        rows = torch.arange(pts.shape[0], device=dev, dtype=torch.int32)
    else:
        gids = torch.arange(npm.count(), device=dev)
        pts = npm.neural_points
        rows = gids.to(torch.int32)
    if pts.shape[0] == 0:
        return None
    from ..utils.tools import ieee_div

    cells = ieee_div(pts, res).floor().to(torch.int64)
    slots = torch.fmod((cells * primes).sum(-1), int(npm.buffer_size))
    keep = npm.buffer_pt_index[slots] == gids  # the point owns its voxel's slot
    if query_locally and npm.temporal_local_map_on:
        td = npm.travel_dist.to(device=dev, dtype=torch.float32)
        gap = torch.abs(td[npm.cur_ts] - td[npm.point_ts_create[gids].long()])
        keep = keep & (gap < npm.diff_travel_dist_local)
    if not bool(keep.any()):
        return None
    cells, pts, rows = cells[keep], pts[keep], rows[keep]

    # one empty brick all around the occupied box (ClidBricks.apron): a neighbourhood that can
    # contain points then has its lower-corner brick in [0, dims - 2] on every axis, so the kernels
    # range-test a query once instead of once per brick
    lo = cells.amin(0) - 4
    hi = cells.amax(0) + 4
    lo_c, hi_c = lo.cpu(), hi.cpu()
    dims = [int((hi_c[i] - lo_c[i]) // 4 + 1) for i in range(3)]
    n_bricks = dims[0] * dims[1] * dims[2]
    if n_bricks > MAX_BRICKS:
        return _fallback(f"bounding box of {dims} bricks exceeds the dense header cap of {MAX_BRICKS}")
    rel = cells - lo
    brick = (rel[:, 0] >> 2) + dims[0] * ((rel[:, 1] >> 2) + dims[1] * (rel[:, 2] >> 2))
    bit = (rel[:, 0] & 3) + 4 * (rel[:, 1] & 3) + 16 * (rel[:, 2] & 3)
    order = torch.argsort(brick * 64 + bit)
    brick, bit = brick[order], bit[order]

    records = torch.empty(order.shape[0], 4, dtype=torch.float32, device=dev)
    records[:, :3] = pts[order]
    records[:, 3] = rows[order].view(torch.float32)

    one = torch.ones((), dtype=torch.int64, device=dev)
    mask = torch.zeros(n_bricks, dtype=torch.int64, device=dev)
    mask.scatter_add_(0, brick, one << bit)  # distinct bits per brick: the sum is the OR
    count = torch.zeros(n_bricks, dtype=torch.int64, device=dev)
    count.scatter_add_(0, brick, torch.ones_like(brick))
    base = torch.cumsum(count, 0) - count
    headers = torch.empty(n_bricks, 4, dtype=torch.int32, device=dev)
    headers[:, :2] = mask.view(torch.int32).view(n_bricks, 2)
    headers[:, 2] = base.to(torch.int32)
    headers[:, 3] = count.to(torch.int32)

    stencil = _stencil_table(offsets_cpu, reach, span).to(dev)
    hood = _hood_lines(headers, dims) if USE_HOOD and n_bricks <= MAX_HOOD_BRICKS else None
    return BrickIndex(headers, records, stencil, lo_c.tolist(), dims, span, reach, apron=1, hood=hood)


MAX_HOOD_BRICKS = 1 << 24  # 128 B per brick: 2 GiB of neighbourhood lines at most
USE_HOOD = os.environ.get("CLID_HOOD", "1") != "0"  # False: the kernels read the eight 16-byte headers (tests run both)


def _hood_lines(headers: torch.Tensor, dims) -> torch.Tensor:
    """ClidBricks.hood: line b = the (mask lo, mask hi) pairs and first-record indices of the 2 x 2 x 2 bricks
    whose lower corner is brick b, so that a query reads its whole neighbourhood directory from ONE 128-byte
    line (coop_search.cuh) instead of eight 16-byte headers in eight different lines.  Bricks past the upper
    faces read as empty; the apron keeps every query's lower-corner brick inside [0, dims - 2]."""
    d0, d1, d2 = dims
    hdr = headers.view(d2, d1, d0, 4)
    hood = torch.zeros(d2, d1, d0, 32, dtype=torch.int32, device=headers.device)
    for s in range(8):
        dx, dy, dz = s & 1, (s >> 1) & 1, s >> 2
        src = hdr[dz:, dy:, dx:]
        hood[: d2 - dz, : d1 - dy, : d0 - dx, 2 * s:2 * s + 2] = src[..., 0:2]
        hood[: d2 - dz, : d1 - dy, : d0 - dx, 16 + s] = src[..., 2]
    return hood.view(-1, 32)
