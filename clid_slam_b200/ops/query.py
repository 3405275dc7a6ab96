"""Python side of the query kernels: builds the C-ABI structs from a NeuralPoints object and
launches libclid_sdf.so on the current CUDA stream.  CUDA only (no CPU fallback)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

from .. import _lib

_PRIMES = (73856093, 19349669, 83492791)


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what} lives on {t.device}: the neural-SDF query path is implemented in CUDA only "
            "(libclid_sdf.so, sm_100a); there is no CPU fallback"
        )


def brick_flags(bricks) -> int:
    """Flag bits that select the brick-index kernels for `bricks` (a BrickIndex)."""
    return _lib.USE_BRICKS

# CLID_TC_DECODER=1: the brick-index forward evaluates 64 x 1 decoders on the tensor cores (tcgen05 + TMEM,
# csrc/decoder_tc.cuh).  Parity-green and measured at the speed of the fp32 FMA decoder, not above it (the kernel is
# bound by the latency of its gather chain), so the FMA kernel stays the default.
TC_DECODER = os.environ.get("CLID_TC_DECODER", "0") == "1"
# The tiles of the persistent kernels are dealt round-robin (ClidMap.work_counter = NULL).  The dynamic alternative
# (library built with -DCLID_DYNAMIC_TILES=1 and CLID_STATIC_TILES=0: a ticket counter, one atomicAdd per tile drawn one tile ahead) evens out per-tile cost
# differences, but its atomic RETURNS a value, and in the training kernel that round trip queues behind the ~24
# fire-and-forget reductions every sample sends to the same L2 atomic units: measured 131072 samples, cold L2:
# training kernel 70.9 us with tickets, 62.9 us round-robin (forward, no other atomics: 45.9 us either way).
STATIC_TILES = os.environ.get("CLID_STATIC_TILES", "1") == "1"

_COUNTERS = {}
_COUNTER_POOLS = {}
_POOL_SLOTS = 256


def _work_counter(npm, device: torch.device) -> torch.Tensor:
    """4 zeroed bytes per (device, stream) for the kernels' dynamic tile scheduler; the kernels leave the
    counter at zero, so it is cleared once.  Counters are slices of one pre-zeroed pool per device: a stream
    that is first seen DURING a CUDA-graph capture (the capture stream) must not get a tensor whose
    allocation and zero-fill would belong to that graph's private pool and memset node."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    t = _COUNTERS.get(key)
    if t is None:
        pool = _COUNTER_POOLS.get(device.index)
        if pool is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("the first query on a device must not happen inside a CUDA-graph capture "
                                   "(run one iteration eagerly first)")
            pool = [torch.zeros(_POOL_SLOTS * 32, dtype=torch.int32, device=device), 0]  # one 128-byte line per counter
            torch.cuda.synchronize(device)
            _COUNTER_POOLS[device.index] = pool
        if pool[1] >= _POOL_SLOTS:
            raise RuntimeError("more than 256 distinct CUDA streams launched query kernels on this device")
        t = pool[0][pool[1] * 32: pool[1] * 32 + 1]
        pool[1] += 1
        _COUNTERS[key] = t
    return t


def map_struct(npm, query_locally: bool, certainty_accum: Optional[torch.Tensor] = None) -> Tuple[_lib.ClidMap, int]:
    """ClidMap for `npm` plus the flag bits implied by the map state (locality, time filter,
    layer norm).  All pointers are borrowed from tensors that `npm` keeps alive."""
    cfg = npm.config
    if npm.buffer_pt_index is None:
        raise RuntimeError("the voxel hash is cleared (clear_temp was called); call recreate_hash first")
    _require_cuda(npm.buffer_pt_index, "NeuralPoints")
    m = _lib.ClidMap()
    m.buffer_pt_index = _lib.ptr(npm.buffer_pt_index, torch.int64, "buffer_pt_index")
    m.buffer_size = int(npm.buffer_size)
    for i in range(3):
        m.primes[i] = _PRIMES[i]
    m.neural_points = _lib.ptr(npm.neural_points, torch.float32, "neural_points")
    m.point_ts_create = _lib.ptr(npm.point_ts_create, torch.int32, "point_ts_create")
    m.n_global = npm.neural_points.shape[0]
    flags = 0
    time_filter = bool(npm.temporal_local_map_on and query_locally)
    if time_filter:
        if npm.travel_dist is None:
            raise RuntimeError("travel_dist is not set (the caller injects it every frame, slam.py:160)")
        td = npm.travel_dist
        if td.dtype != torch.float32 or not td.is_contiguous() or td.device != npm.neural_points.device:
            td = td.to(device=npm.neural_points.device, dtype=torch.float32).contiguous()
            npm.travel_dist = td
        m.travel_dist = _lib.ptr(td, torch.float32, "travel_dist")
        m.n_travel = td.shape[0]
        flags |= _lib.TIME_FILTER
    m.cur_ts = int(npm.cur_ts)
    m.diff_travel_dist_local = float(npm.diff_travel_dist_local)
    m.resolution = float(npm.resolution)
    m.max_valid_dist2 = float(npm.max_valid_dist2)
    m.kc = int(npm.neighbor_K)
    m.neighbor_dx = _lib.ptr(npm.neighbor_dx, torch.int64, "neighbor_dx")
    if query_locally:
        flags |= _lib.QUERY_LOCALLY
        m.global2local = _lib.ptr(npm.global2local, torch.int64, "global2local")
        pts, feats = npm.local_neural_points, npm.local_geo_features.data
        cert, ts_upd = npm.local_point_certainties, npm.local_point_ts_update
    else:
        pts, feats = npm.neural_points, npm.geo_features
        cert, ts_upd = npm.point_certainties, None  # the reference updates no global ts (neural_points.py:730-733)
    m.gather_points = _lib.ptr(pts, torch.float32, "gather points")
    m.gather_features = _lib.ptr(feats, torch.float32, "gather features")
    m.gather_certainties = _lib.ptr(cert, torch.float32, "gather certainties")
    m.certainty_accum = _lib.ptr(certainty_accum, torch.float32, "certainty_accum")
    m.gather_ts_update = _lib.ptr(ts_upd, torch.int32, "ts_update")
    m.n_gather = pts.shape[0]
    m.feature_dim = int(npm.geo_feature_dim)
    m.knn = int(cfg.query_nn_k)
    m.work_counter = None if STATIC_TILES else _work_counter(npm, pts.device).data_ptr()
    if cfg.layer_norm_on:
        flags |= _lib.LAYER_NORM
    return m, flags


def _prep_points(x: torch.Tensor, what: str = "query_points") -> torch.Tensor:
    _require_cuda(x, what)
    if x.dim() != 2 or x.shape[1] != 3:
        raise ValueError(f"{what} must be [N,3], got {tuple(x.shape)}")
    xd = x.detach()
    if xd.dtype != torch.float32:
        xd = xd.float()
    return xd.contiguous()


def _prep_ts(ts: Optional[torch.Tensor], n: int) -> Optional[torch.Tensor]:
    if ts is None:
        return None
    _require_cuda(ts, "query_ts")
    if ts.shape[0] != n:
        raise ValueError("query_ts must have one entry per query point")
    return ts.to(torch.int32).contiguous()


def forward(npm, decoder, x: torch.Tensor, ts: Optional[torch.Tensor], training_mode: bool, query_locally: bool,
            want_sdf=False, want_grad=False, want_z=False, want_weights=False, want_idx=False, want_count=False,
            want_certainty=False, certainty_accum: Optional[torch.Tensor] = None,
            use_bricks: Optional[bool] = None):
    """One launch of clid_query_forward.  Returns a dict of the requested outputs.

    use_bricks: None = use the brick index whenever it is exact for this map (default), False =
    probe the reference's hash table, True = require the brick index."""
    lib = _lib.load()
    xd = _prep_points(x)
    n = xd.shape[0]
    tsd = _prep_ts(ts, n)
    dev = xd.device
    k = int(npm.config.query_nn_k)
    if training_mode and certainty_accum is None:
        certainty_accum = npm.local_point_certainties if query_locally else npm.point_certainties
    m, flags = map_struct(npm, query_locally, certainty_accum if training_mode else None)
    if training_mode:
        flags |= _lib.TRAINING_MODE
    if use_bricks is None:
        use_bricks = os.environ.get("CLID_DISABLE_BRICKS", "0") != "1"
        bricks = npm.brick_index(query_locally) if use_bricks else None
    elif use_bricks:
        bricks = npm.brick_index(query_locally)
        if bricks is None:
            raise RuntimeError("the brick index is not available for this map (hash aliasing or empty map)")
    else:
        bricks = None
    if bricks is not None:
        m.bricks = C.pointer(bricks.struct)
        flags |= brick_flags(bricks)
    dec_struct = None
    if decoder is not None:
        dec_struct = decoder.abi_struct()
        if decoder.use_leaky_relu:
            flags |= _lib.LEAKY_RELU
        if TC_DECODER:
            flags |= _lib.TC_DECODER  # taken by the library for 64 x 1 decoders on the brick index
    out = _lib.ClidQueryOut()
    res = {}
    f32 = dict(dtype=torch.float32, device=dev)
    if want_sdf:
        res["sdf"] = torch.empty(n, **f32)
        out.sdf = res["sdf"].data_ptr()
    if want_grad:
        res["grad"] = torch.empty(n, 3, **f32)
        out.grad = res["grad"].data_ptr()
    if want_z:
        res["z"] = torch.empty(n, npm.geo_feature_dim + 3, **f32)
        out.z = res["z"].data_ptr()
    if want_weights:
        res["weights"] = torch.empty(n, k, **f32)
        out.weights = res["weights"].data_ptr()
    if want_idx:
        res["knn_idx"] = torch.empty(n, k, dtype=torch.int32, device=dev)
        out.knn_idx = res["knn_idx"].data_ptr()
    if want_count:
        res["nn_count"] = torch.empty(n, dtype=torch.int32, device=dev)
        out.nn_count = res["nn_count"].data_ptr()
    if want_certainty:
        res["certainty"] = torch.empty(n, **f32)
        out.certainty = res["certainty"].data_ptr()
    with torch.cuda.device(dev):
        rc = lib.clid_query_forward(
            C.byref(m), C.byref(dec_struct) if dec_struct is not None else None,
            xd.data_ptr(), None if tsd is None else tsd.data_ptr(), n, flags, C.byref(out),
            _lib.current_stream(dev),
        )
    _lib.check(rc, "clid_query_forward")
    return res


def query_feature(npm, x: torch.Tensor, ts: Optional[torch.Tensor], training_mode: bool, query_locally: bool):
    """NeuralPoints.query_feature forward (weighted_first).  Returns (z, weights [N,K], nn_counts
    int64, queried_certainty).  Autograd support is attached by ops.autograd when needed."""
    from . import autograd as _ag

    return _ag.query_feature(npm, x, ts, training_mode, query_locally)


def radius_search(npm, x: torch.Tensor, time_filtering: bool):
    lib = _lib.load()
    xd = _prep_points(x, "points")
    n = xd.shape[0]
    m, flags = map_struct(npm, query_locally=False)
    if time_filtering:
        td = npm.travel_dist.to(device=xd.device, dtype=torch.float32).contiguous()
        m.travel_dist = td.data_ptr()
        m.n_travel = td.shape[0]
        flags |= _lib.TIME_FILTER
    d2 = torch.empty(n, m.kc, dtype=torch.float32, device=xd.device)
    idx = torch.empty(n, m.kc, dtype=torch.int64, device=xd.device)
    with torch.cuda.device(xd.device):
        rc = lib.clid_radius_search(C.byref(m), xd.data_ptr(), n, flags, d2.data_ptr(), idx.data_ptr(),
                                    _lib.current_stream(xd.device))
    _lib.check(rc, "clid_radius_search")
    return d2, idx


def query_certainty(npm, x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    xd = _prep_points(x)
    n = xd.shape[0]
    m, _ = map_struct(npm, query_locally=False)
    out = torch.empty(n, dtype=torch.float32, device=xd.device)
    with torch.cuda.device(xd.device):
        rc = lib.clid_query_certainty(C.byref(m), xd.data_ptr(), n,
                                      _lib.ptr(npm.point_certainties, torch.float32, "point_certainties"),
                                      out.data_ptr(), _lib.current_stream(xd.device))
    _lib.check(rc, "clid_query_certainty")
    return out


def decoder_eval(decoder, z: torch.Tensor, want_grad: bool = True, want_mask: bool = False):
    """Decoder.mlp on caller-supplied inputs z [N,11] and d out / d z, on the tensor cores (clid_decoder_eval,
    64 x 1 decoders).  Returns (out [N] un-scaled logit, a [N,11] or None, mask [N,2] int32 or None)."""
    lib = _lib.load()
    _require_cuda(z, "z")
    zd = z.detach().float().contiguous()
    n = zd.shape[0]
    out = torch.empty(n, dtype=torch.float32, device=zd.device)
    a = torch.empty(n, zd.shape[1], dtype=torch.float32, device=zd.device) if want_grad else None
    mask = torch.empty(n, 2, dtype=torch.int32, device=zd.device) if want_mask else None
    dec = decoder.abi_struct()
    flags = _lib.LEAKY_RELU if decoder.use_leaky_relu else 0
    with torch.cuda.device(zd.device):
        rc = lib.clid_decoder_eval(C.byref(dec), zd.data_ptr(), n, flags, out.data_ptr(),
                                   None if a is None else a.data_ptr(), None if mask is None else mask.data_ptr(),
                                   _lib.current_stream(zd.device))
    _lib.check(rc, "clid_decoder_eval")
    return out, a, mask
