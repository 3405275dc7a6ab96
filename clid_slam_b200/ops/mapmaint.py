"""Native per-frame map maintenance: host side of csrc/mapmaint.cuh.

Replaces, for CUDA maps, the eager-torch bodies of ``voxel_down_sample_torch`` (utils/tools.py:639-682),
``NeuralPoints.update`` (model/neural_points.py:324-437), ``reset_local_map`` (:439-536) and
``assign_local_to_global`` (:538-549).  Each operation is a handful of launches and ONE 16-byte read-back (the
size of its result: torch's allocator needs it on the host); the order-dependent results follow the reference's
sequential CPU semantics, so a CUDA map is identical to the CPU fixtures.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib

_I64, _I32, _F32, _U8 = torch.int64, torch.int32, torch.float32, torch.uint8


def _workspace(n: int, device) -> torch.Tensor:
    lib = _lib.load()
    return torch.empty((int(lib.clid_scan_workspace_bytes(int(n))) + 7) // 8, dtype=_I64, device=device)


def _count(ws: torch.Tensor) -> int:
    return int(ws[:2].cpu()[0])  # the one synchronisation of the operation


def voxel_down_sample(points: torch.Tensor, voxel_size: float, value: torch.Tensor | None = None) -> torch.Tensor:
    """Indices of the kept points in ascending voxel order (utils/tools.py:639-724), or None when the voxel key
    does not fit the packed sort key (the caller then uses the eager path)."""
    lib = _lib.load()
    dev = points.device
    pts = points.contiguous()
    if pts.dtype != _F32:
        return None
    n = pts.shape[0]
    stream = _lib.current_stream(dev)
    stats = torch.empty(8, dtype=_I32, device=dev)
    keys = torch.empty(n, dtype=_I64, device=dev)
    val = None if value is None else value.contiguous().to(_F32)
    _lib.check(lib.clid_voxel_keys(_lib.ptr(pts, _F32, "points"), _lib.ptr(val, _F32, "value"), n, float(voxel_size),
                                   stats.data_ptr(), keys.data_ptr(), stream), "clid_voxel_keys")
    skeys, order = torch.sort(keys, stable=True)
    ws = _workspace(n, dev)
    flags = torch.empty(n, dtype=_U8, device=dev)
    selected = torch.empty(n, dtype=_I64, device=dev)
    out = torch.empty(n, dtype=_I64, device=dev)
    _lib.check(lib.clid_voxel_pick(skeys.data_ptr(), order.data_ptr(), n, ws.data_ptr(), ws.numel() * 8, flags.data_ptr(),
                                   selected.data_ptr(), out.data_ptr(), stream), "clid_voxel_pick")
    head = torch.cat((ws[:1], stats[7:8].to(_I64))).cpu()  # {count, overflow} in one read-back
    if int(head[1]) != 0:
        return None
    return out[: int(head[0])]


def insert(npm, cand: torch.Tensor, cur_ts: int) -> int:
    """The body of NeuralPoints.update between the down-sampling and reset_local_map: probe, number the new
    points, store the table, grow the arrays.  Returns the number of new points."""
    lib = _lib.load()
    dev = cand.device
    cand = cand.contiguous()
    n, m = cand.shape[0], npm.count()
    res = float(npm.resolution)
    stream = _lib.current_stream(dev)
    all_fresh = npm.is_empty() or cur_ts == npm.reboot_ts
    slot = torch.empty(n, dtype=_I64, device=dev)
    owner = torch.empty(n, dtype=_I64, device=dev)
    rank = torch.empty(n, dtype=_I64, device=dev)
    fresh = torch.empty(n, dtype=_U8, device=dev)
    ws = _workspace(n, dev)
    a = _lib.ClidInsertArgs()
    a.cand, a.n = _lib.ptr(cand, _F32, "points"), n
    a.buffer_pt_index, a.buffer_size = _lib.ptr(npm.buffer_pt_index, _I64, "buffer_pt_index"), int(npm.buffer_size)
    for i in range(3):
        a.primes[i] = int(npm.primes[i])
    temporal = bool(npm.temporal_local_map_on) and not all_fresh
    td = npm.travel_dist.contiguous() if temporal else None
    a.neural_points = None if all_fresh else _lib.ptr(npm.neural_points, _F32, "neural_points")
    a.ts_update = _lib.ptr(npm.point_ts_update, _I32, "point_ts_update") if temporal else None
    a.travel_dist = _lib.ptr(td, _F32, "travel_dist") if temporal else None
    a.m, a.n_travel = m, (td.shape[0] if temporal else 0)
    a.cur_ts, a.all_fresh = int(cur_ts), int(all_fresh)
    a.resolution = res
    a.far2 = float(torch.tensor(3 * res**2, dtype=_F32))  # the threshold torch compares an fp32 tensor with
    a.diff_travel_dist_local = float(npm.diff_travel_dist_local)
    a.slot, a.owner, a.fresh, a.rank = slot.data_ptr(), owner.data_ptr(), fresh.data_ptr(), rank.data_ptr()
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel() * 8
    _lib.check(lib.clid_map_insert_probe(C.byref(a), stream), "clid_map_insert_probe")
    n_new = _count(ws)

    f32 = npm.dtype
    grown_points = torch.empty((m + n_new, 3), dtype=f32, device=dev)
    grown_create = torch.empty(m + n_new, dtype=_I32, device=dev)
    grown_update = torch.empty(m + n_new, dtype=_I32, device=dev)
    if m:
        grown_points[:m] = npm.neural_points
        grown_create[:m] = npm.point_ts_create
        grown_update[:m] = npm.point_ts_update
    # the commit needs non-NULL row pointers even when nothing is appended
    tail_p = grown_points[m:] if n_new else torch.empty((1, 3), dtype=f32, device=dev)
    tail_c = grown_create[m:] if n_new else torch.empty(1, dtype=_I32, device=dev)
    tail_u = grown_update[m:] if n_new else torch.empty(1, dtype=_I32, device=dev)
    _lib.check(lib.clid_map_insert_commit(C.byref(a), tail_p.data_ptr(), tail_c.data_ptr(), tail_u.data_ptr(), stream),
               "clid_map_insert_commit")
    npm.neural_points, npm.point_ts_create, npm.point_ts_update = grown_points, grown_create, grown_update
    return n_new


def local_window(npm, sensor_position: torch.Tensor, cur_ts: int, use_travel_dist: bool, diff_ts_local: int,
                 reboot_map: bool) -> None:
    """Window selection and the local_* copies of NeuralPoints.reset_local_map (model/neural_points.py:452-536)."""
    lib = _lib.load()
    dev = npm.neural_points.device
    m = npm.count()
    stream = _lib.current_stream(dev)
    flags = torch.empty(m, dtype=_U8, device=dev)
    g2l = torch.empty(m + 1, dtype=_I64, device=dev)
    mask = torch.empty(m + 1, dtype=torch.bool, device=dev)
    gids = torch.empty(m, dtype=_I64, device=dev)
    ws = _workspace(m, dev)
    a = _lib.ClidWindowArgs()
    temporal = bool(npm.temporal_local_map_on)
    td = npm.travel_dist.contiguous() if (temporal and use_travel_dist) else None
    a.neural_points = _lib.ptr(npm.neural_points, _F32, "neural_points")
    a.ts_create = _lib.ptr(npm.point_ts_create, _I32, "point_ts_create")
    a.ts_update = _lib.ptr(npm.point_ts_update, _I32, "point_ts_update")
    a.travel_dist = _lib.ptr(td, _F32, "travel_dist") if td is not None else None
    a.m, a.n_travel = m, (td.shape[0] if td is not None else 0)
    pos = sensor_position.detach().to("cpu", torch.float64)
    for i in range(3):
        a.sensor[i] = float(pos[i])
    a.sensor_is_f64 = int(sensor_position.dtype == torch.float64)
    a.radius2 = float(npm.local_map_radius) ** 2
    a.temporal, a.use_mid_ts = int(temporal), int(bool(npm.config.use_mid_ts))
    a.cur_ts, a.reboot_test, a.reboot_ts = int(cur_ts), int(bool(reboot_map)), int(npm.reboot_ts)
    a.diff_ts_local = int(diff_ts_local)
    a.diff_travel_dist_local = float(npm.diff_travel_dist_local)
    a.flags, a.global2local, a.local_mask, a.gids = flags.data_ptr(), g2l.data_ptr(), mask.data_ptr(), gids.data_ptr()
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel() * 8
    _lib.check(lib.clid_local_window_select(C.byref(a), stream), "clid_local_window_select")
    n_local = _count(ws)

    f32 = npm.dtype
    npm._local_gids = gids[:n_local]
    npm.local_mask, npm.global2local = mask, g2l
    npm.local_neural_points = torch.empty((n_local, 3), dtype=f32, device=dev)
    npm.local_point_orientations = torch.empty((n_local, 4), dtype=f32, device=dev)
    npm.local_point_certainties = torch.empty(n_local, dtype=f32, device=dev)
    npm.local_point_ts_update = torch.empty(n_local, dtype=_I32, device=dev)
    feats = torch.empty((n_local + 1, npm.geo_feature_dim), dtype=f32, device=dev)
    r = _rows(npm, feats)
    _lib.check(lib.clid_local_window_gather(C.byref(r), stream), "clid_local_window_gather")
    npm.local_geo_features = torch.nn.Parameter(feats)


def _rows(npm, local_features: torch.Tensor) -> "_lib.ClidWindowRows":
    r = _lib.ClidWindowRows()
    gids = npm._local_gids
    r.gids, r.n_local, r.m = (gids.data_ptr() if gids.numel() else None), gids.numel(), npm.count()
    r.neural_points = _lib.ptr(npm.neural_points, _F32, "neural_points")
    r.point_orientations = _lib.ptr(npm.point_orientations, _F32, "point_orientations")
    r.point_certainties = _lib.ptr(npm.point_certainties, _F32, "point_certainties")
    r.point_ts_update = _lib.ptr(npm.point_ts_update, _I32, "point_ts_update")
    r.geo_features = _lib.ptr(npm.geo_features, _F32, "geo_features")
    r.local_points = _lib.ptr(npm.local_neural_points, _F32, "local_neural_points")
    r.local_orientations = _lib.ptr(npm.local_point_orientations, _F32, "local_point_orientations")
    r.local_certainties = _lib.ptr(npm.local_point_certainties, _F32, "local_point_certainties")
    r.local_ts_update = _lib.ptr(npm.local_point_ts_update, _I32, "local_point_ts_update")
    r.local_features = _lib.ptr(local_features, _F32, "local_geo_features")
    return r


def assign_local_to_global(npm) -> None:
    """NeuralPoints.assign_local_to_global (model/neural_points.py:538-549) as one scatter launch."""
    lib = _lib.load()
    feats = npm.local_geo_features.data
    r = _rows(npm, feats.contiguous())
    _lib.check(lib.clid_local_window_scatter(C.byref(r), _lib.current_stream(feats.device)), "clid_local_window_scatter")


def table_store(table: torch.Tensor, slots: torch.Tensor, values: torch.Tensor | None, value_base: int = 0) -> None:
    """table[slots] = values (values None: value_base + position), the last occurrence of a repeated slot winning."""
    n = slots.shape[0]
    if n == 0:
        return
    slots = slots.contiguous()
    vals = None if values is None else values.contiguous()
    _lib.check(_lib.load().clid_table_store(_lib.ptr(slots, _I64, "slots"), _lib.ptr(vals, _I64, "values"), n, int(value_base),
                                            _lib.ptr(table, _I64, "buffer_pt_index"), table.shape[0],
                                            _lib.current_stream(table.device)), "clid_table_store")


def pool_filter(global_coord: torch.Tensor, origin: torch.Tensor, radius: float, arrays, n_tail: int, use_norm: bool = False):
    """Replay-pool filter of Mapper.process_frame (utils/mapper.py:420-459): the rows of every array in `arrays`
    (all [n] or [n, 3], 4-byte elements) whose sample lies within `radius` of `origin`, in pool order.
    Returns (kept arrays, kept count, kept count among the last n_tail rows)."""
    lib = _lib.load()
    dev = global_coord.device
    gc = global_coord.contiguous()
    n = gc.shape[0]
    stream = _lib.current_stream(dev)
    flags = torch.empty(n, dtype=_U8, device=dev)
    rank = torch.empty(n, dtype=_I64, device=dev)
    ws = _workspace(n, dev)
    pos = origin.detach().to("cpu", torch.float64)
    sensor = (C.c_double * 3)(float(pos[0]), float(pos[1]), float(pos[2]))
    _lib.check(lib.clid_pool_filter_select(_lib.ptr(gc, _F32, "global_coord_pool"), n, sensor, float(radius),
                                           int(origin.dtype == torch.float64), int(use_norm), flags.data_ptr(), rank.data_ptr(),
                                           ws.data_ptr(), ws.numel() * 8, stream), "clid_pool_filter_select")
    head = torch.cat((ws[:1], flags[n - n_tail:].sum(dtype=_I64).reshape(1))).cpu()  # the one read-back
    n_keep, n_tail_keep = int(head[0]), int(head[1])
    srcs = [a.contiguous() for a in arrays]
    for a in srcs:
        if a.element_size() != 4 or a.shape[0] != n:
            raise TypeError("pool arrays must have n rows of 4-byte elements")
    outs = [torch.empty((n_keep,) + tuple(a.shape[1:]), dtype=a.dtype, device=dev) for a in srcs]
    if n_keep > 0:
        k = len(srcs)
        src_p = (C.c_void_p * k)(*[a.data_ptr() for a in srcs])
        dst_p = (C.c_void_p * k)(*[o.data_ptr() for o in outs])
        words = (C.c_int32 * k)(*[int(a[0].numel()) for a in srcs])
        _lib.check(lib.clid_compact_rows(rank.data_ptr(), n, src_p, dst_p, words, k, stream), "clid_compact_rows")
    return outs, n_keep, n_tail_keep, flags


def ray_samples(cfg, points: torch.Tensor):
    """The ray samples of DataSampler.sample / sample_pin (utils/data_sampler.py:35-140) in ONE launch: returns
    (coord [P*S,3], disp [P*S], weight [P*S]) ray-major.  The random numbers come from torch's generator in the
    reference's order, so a seeded run draws the samples the torch ops draw."""
    lib = _lib.load()
    dev = points.device
    pts = points.contiguous()
    count = pts.shape[0]
    n_surf, n_front, n_behind = int(cfg.surface_sample_n), int(cfg.free_front_n), int(cfg.free_behind_n)
    per_ray = 1 + n_surf + n_front + n_behind
    depth = torch.linalg.norm(pts, dim=1)
    randn_surf = torch.randn(count * n_surf, 1, device=dev)
    rand_front = torch.rand(count * n_front, 1, device=dev)
    rand_behind = torch.rand(count * n_behind, 1, device=dev)
    coord = torch.empty((count * per_ray, 3), dtype=_F32, device=dev)
    disp = torch.empty(count * per_ray, dtype=_F32, device=dev)
    weight = torch.empty(count * per_ray, dtype=_F32, device=dev)
    a = _lib.ClidRaySampleArgs()
    a.points, a.depth = _lib.ptr(pts, _F32, "points"), depth.data_ptr()
    a.randn_surf, a.rand_front, a.rand_behind = randn_surf.data_ptr(), rand_front.data_ptr(), rand_behind.data_ptr()
    a.n_points, a.n_surf, a.n_front, a.n_behind = count, n_surf, n_front, n_behind
    sigma = float(cfg.surface_sample_range_m)
    a.surface_sample_range_m, a.margin = sigma, 2.0 * sigma
    a.free_sample_begin_ratio, a.free_sample_end_dist_m = float(cfg.free_sample_begin_ratio), float(cfg.free_sample_end_dist_m)
    a.weight_top = 1 + float(cfg.dist_weight_scale) * 0.5
    a.dist_weight_scale, a.max_range, a.dist_weight_on = float(cfg.dist_weight_scale), float(cfg.max_range), int(bool(cfg.dist_weight_on))
    a.coord, a.disp, a.weight = coord.data_ptr(), disp.data_ptr(), weight.data_ptr()
    _lib.check(lib.clid_ray_samples(C.byref(a), _lib.current_stream(dev)), "clid_ray_samples")
    return coord, disp, weight, per_ray


def region_labelled_samples(cfg, points: torch.Tensor, local_point_cloud_map, cur_pose_torch):
    """DataSampler.sample on the device (utils/data_sampler.py:260-402): ray samples, region-specific labels of the
    near-surface samples, unreachable samples dropped.  Returns (coord, sdf_label, weight)."""
    from ..utils.tools import transform_torch

    lib = _lib.load()
    dev = points.device
    stream = _lib.current_stream(dev)
    coord, disp, weight, per_ray = ray_samples(cfg, points)
    count, n_surf = points.shape[0], int(cfg.surface_sample_n)
    near = coord.view(count, per_ray, 3)[:, 1:1 + n_surf, :].reshape(-1, 3)
    dist, reachable = local_point_cloud_map.region_specific_sdf_estimation(transform_torch(near, cur_pose_torch))
    reach_u8 = reachable.to(_U8).contiguous()
    dist = dist.contiguous()
    n = count * per_ray
    label = torch.empty(n, dtype=_F32, device=dev)
    keep = torch.empty(n, dtype=_U8, device=dev)
    _lib.check(lib.clid_ray_labels(disp.data_ptr(), dist.data_ptr(), reach_u8.data_ptr(), count, per_ray, n_surf,
                                   label.data_ptr(), keep.data_ptr(), stream), "clid_ray_labels")
    rank = torch.empty(n, dtype=_I64, device=dev)
    ws = _workspace(n, dev)
    _lib.check(lib.clid_flag_ranks(keep.data_ptr(), n, rank.data_ptr(), ws.data_ptr(), ws.numel() * 8, stream), "clid_flag_ranks")
    n_keep = _count(ws)
    outs = [torch.empty((n_keep, 3), dtype=_F32, device=dev), torch.empty(n_keep, dtype=_F32, device=dev),
            torch.empty(n_keep, dtype=_F32, device=dev)]
    if n_keep > 0:
        src_p = (C.c_void_p * 3)(coord.data_ptr(), label.data_ptr(), weight.data_ptr())
        dst_p = (C.c_void_p * 3)(*[o.data_ptr() for o in outs])
        words = (C.c_int32 * 3)(3, 1, 1)
        _lib.check(lib.clid_compact_rows(rank.data_ptr(), n, src_p, dst_p, words, 3, stream), "clid_compact_rows")
    return outs[0], outs[1], outs[2]
