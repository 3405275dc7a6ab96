"""Voxel-hashed neural-point map: mirror of the reference's ``model.neural_points.NeuralPoints``
(model/neural_points.py:27-1089) with the query path running in hand-written sm_100a kernels.

Kept from the reference: constructor, every public method name
and signature, every tensor attribute name / dtype / layout (the GUI and ``save_implicit_map``
read them directly), ``local_geo_features`` as an ``nn.Parameter`` rebuilt on every
``reset_local_map``.  The object holds only tensors and Python scalars (the CUDA library handle
is a module global in ``clid_slam_b200._lib``), so it pickles like the reference's.

Per-frame map maintenance (``update`` / ``reset_local_map`` / ``prune_map`` / ``recreate_hash``)
is host logic expressed in torch ops on the map's device.  Everything per-query
(``query_feature`` and the fused entry points in ``clid_slam_b200.fused``) is CUDA only and
raises on a CPU map: there is no CPU fallback.
"""
from __future__ import annotations

import sys
from typing import Optional

import torch
import torch.nn as nn

from ..ops import query as _q
from ..utils.tools import ieee_div, voxel_down_sample_min_value_torch, voxel_down_sample_torch

PRIMES = (73856093, 19349669, 83492791)  # model/neural_points.py:79-81


class NeuralPoints(nn.Module):
    def __init__(self, config) -> None:
        super().__init__()
        self.config = config
        self.silence = config.silence

        self.geo_feature_dim = config.feature_dim
        self.geo_feature_std = config.feature_std
        self.color_feature_dim = config.feature_dim
        self.color_feature_std = config.feature_std

        self.mean_grid_sampling = False
        self.device = config.device
        self.dtype = config.dtype
        self.idx_dtype = torch.int64

        self.resolution = config.voxel_size_m
        self.buffer_size = config.buffer_size

        self.temporal_local_map_on = True
        self.local_map_radius = config.local_map_radius
        self.diff_travel_dist_local = config.local_map_radius * config.local_map_travel_dist_ratio
        self.diff_ts_local = config.diff_ts_local

        self.reboot_ts = 0
        self.local_orientation = torch.eye(3, device=self.device)
        self.cur_ts = 0
        self.max_ts = 0
        self.travel_dist = None  # [frames] tensor injected by the caller every frame (slam.py:160-162)
        self.est_poses = None
        self.after_pgo = False

        dev, f32, i64, i32 = self.device, self.dtype, self.idx_dtype, torch.int
        self.primes = torch.tensor(PRIMES, dtype=i64, device=dev)
        self.buffer_pt_index = torch.full((self.buffer_size,), -1, dtype=i64, device=dev)

        self.neural_points = torch.empty((0, 3), dtype=f32, device=dev)
        self.point_orientations = torch.empty((0, 4), dtype=f32, device=dev)
        self.geo_features = torch.empty((1, self.geo_feature_dim), dtype=f32, device=dev)
        self.color_on = bool(config.color_on)
        self.color_features = torch.empty((1, self.color_feature_dim), dtype=f32, device=dev) if self.color_on else None
        self.geo_feature_pca = self.color_feature_pca = None

        self.point_ts_create = torch.empty((0,), dtype=i32, device=dev)
        self.point_ts_update = torch.empty((0,), dtype=i32, device=dev)
        self.point_certainties = torch.empty((0,), dtype=f32, device=dev)

        self.local_neural_points = torch.empty((0, 3), dtype=f32, device=dev)
        self.local_point_orientations = torch.empty((0, 4), dtype=f32, device=dev)
        self.local_geo_features = nn.Parameter()
        self.local_color_features = nn.Parameter()
        self.local_point_certainties = torch.empty((0,), dtype=f32, device=dev)
        self.local_point_ts_update = torch.empty((0,), dtype=i32, device=dev)
        self.local_mask = None
        self.global2local = None

        self.set_search_neighborhood(num_nei_cells=config.num_nei_cells, search_alpha=config.search_alpha)

        self.cur_memory_mb = 0.0
        self.memory_footprint = []
        self._map_version = 0      # bumped by every method that changes what a query can return
        self._brick_cache = {}     # query_locally -> (key, BrickIndex | None); never pickled
        self.to(self.device)

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_brick_cache"] = {}
        return state

    def _touch(self) -> None:
        self._map_version = getattr(self, "_map_version", 0) + 1

    def brick_index(self, query_locally: bool):
        """Cached brick index for the current map state (rebuilt lazily after any change), or None
        when the compact index is not exact for this table (see ops/bricks.py)."""
        from ..ops import bricks as _bricks

        def addr(t):
            return 0 if t is None else (t.data_ptr(), tuple(t.shape))

        key = (getattr(self, "_map_version", 0), addr(self.buffer_pt_index), addr(self.neural_points),
               addr(self.neighbor_dx), int(self.cur_ts), float(self.max_valid_dist2),
               addr(self.local_neural_points) if query_locally else 0,
               addr(self.global2local) if query_locally else 0,
               addr(self.travel_dist) if query_locally else 0,
               bool(self.temporal_local_map_on))
        cache = self.__dict__.setdefault("_brick_cache", {})
        hit = cache.get(bool(query_locally))
        if hit is None or hit[0] != key:
            hit = (key, _bricks.build(self, query_locally))
            cache[bool(query_locally)] = hit
        return hit[1]

    # ------------------------------------------------------------------ bookkeeping
    def is_empty(self) -> bool:
        return self.neural_points.shape[0] == 0

    def count(self) -> int:
        return self.neural_points.shape[0]

    def local_count(self) -> int:
        return 0 if self.local_neural_points is None else self.local_neural_points.shape[0]

    def record_memory(self, verbose: bool = True, record_footprint: bool = True) -> None:
        floats_per_point = self.geo_feature_dim + 3 + 4
        if self.color_features is not None:
            floats_per_point += self.color_feature_dim
        self.cur_memory_mb = self.count() * floats_per_point * 4 / 1024 / 1024
        if verbose:
            print("# Global neural point: %d" % self.count())
            print("# Local  neural point: %d" % self.local_count())
            print("Current map memory consumption: {:.3f} MB".format(self.cur_memory_mb))
        if record_footprint:
            self.memory_footprint.append(self.cur_memory_mb)

    def print_memory(self) -> None:  # called by the reference's mapper when not silent (mapper.py:292)
        self.record_memory(verbose=True, record_footprint=False)

    # ------------------------------------------------------------------ hashing
    def _slots_of(self, points: torch.Tensor) -> torch.Tensor:
        cells = ieee_div(points, self.resolution).floor().to(self.primes)
        return torch.fmod((cells * self.primes).sum(-1), int(self.buffer_size))

    def _store_slots(self, slots: torch.Tensor, values: torch.Tensor) -> None:
        """table[slots] = values where the LAST occurrence of a repeated slot wins (what the
        reference's sequential CPU index_put does; CUDA index_put is unordered)."""
        if slots.is_cuda:
            from ..ops import mapmaint as _mm  # bid / commit kernels (csrc/mapmaint.cuh)

            _mm.table_store(self.buffer_pt_index, slots, values)
            return
        slots = torch.remainder(slots, int(self.buffer_size))  # fmod's negative slots wrap to the same entries
        uniq, inverse = torch.unique(slots, return_inverse=True)
        pos = torch.arange(slots.shape[0], device=slots.device)
        last = torch.empty(uniq.shape, dtype=pos.dtype, device=slots.device)
        last.scatter_reduce_(0, inverse, pos, reduce="amax", include_self=False)
        self.buffer_pt_index[uniq] = values[last]

    # ------------------------------------------------------------------ map insert (per frame)
    def update(self, points: torch.Tensor, sensor_position: torch.Tensor, sensor_orientation: torch.Tensor,
               cur_ts: int) -> float:
        """Insert one scan's points: one new neural point per voxel that is empty, owned by a far
        (hash-colliding) point or by a stale one (model/neural_points.py:324-437).  Returns the
        fraction of down-sampled scan points that became new neural points."""
        res = self.resolution
        keep = voxel_down_sample_torch(points, res)
        cand = points[keep]
        if cand.is_cuda and cand.dtype == torch.float32 and cand.shape[0] > 0:
            return self._update_native(cand, sensor_position, sensor_orientation, cur_ts)
        slots = self._slots_of(cand)
        owner = self.buffer_pt_index[slots]

        if not self.is_empty() and cur_ts != self.reboot_ts:
            d2 = ((self.neural_points[owner] - cand) ** 2).sum(-1)
            fresh = (owner == -1) | (d2 > 3 * res**2)
            if self.temporal_local_map_on:
                gap = self.travel_dist[cur_ts] - self.travel_dist[self.point_ts_update[owner]]
                fresh = fresh | (gap > self.diff_travel_dist_local)
        else:
            fresh = torch.ones(owner.shape, dtype=torch.bool, device=self.device)

        added = cand[fresh]
        n_new = added.shape[0]
        new_point_ratio = n_new / cand.shape[0]

        owner = owner.clone()
        owner[fresh] = torch.arange(n_new, dtype=self.idx_dtype, device=self.device) + self.count()
        self._store_slots(slots, owner)

        dev, f32 = self.device, self.dtype
        self.neural_points = torch.cat((self.neural_points, added), 0)
        ident = torch.zeros((n_new, 4), dtype=f32, device=dev)
        ident[:, 0] = 1.0
        self.point_orientations = torch.cat((self.point_orientations, ident), 0)
        stamp = torch.full((n_new,), cur_ts, dtype=torch.int, device=dev)
        self.point_ts_create = torch.cat((self.point_ts_create, stamp), 0)
        self.point_ts_update = torch.cat((self.point_ts_update, stamp), 0)
        # one more row than points: the last row is the padding the reference indexes with -1
        fresh_feat = self.geo_feature_std * torch.randn(n_new + 1, self.geo_feature_dim, device=dev, dtype=f32)
        self.geo_features = torch.cat((self.geo_features[:-1], fresh_feat), 0)
        if self.color_features is not None:
            fresh_col = self.color_feature_std * torch.randn(n_new + 1, self.color_feature_dim, device=dev, dtype=f32)
            self.color_features = torch.cat((self.color_features[:-1], fresh_col), 0)
        self.point_certainties = torch.cat((self.point_certainties, torch.zeros(n_new, device=dev, dtype=f32)), 0)

        self._touch()
        self.reset_local_map(sensor_position, sensor_orientation, cur_ts, reboot_map=True)
        return new_point_ratio

    def _update_native(self, cand: torch.Tensor, sensor_position, sensor_orientation, cur_ts: int) -> float:
        """update() on a CUDA map: probe / numbering / table store in csrc/mapmaint.cuh (ops/mapmaint.py), the
        feature draw stays torch.randn (the reference's generator stream)."""
        from ..ops import mapmaint as _mm

        dev, f32 = self.device, self.dtype
        n_new = _mm.insert(self, cand, cur_ts)
        ident = torch.zeros((n_new, 4), dtype=f32, device=dev)
        ident[:, 0] = 1.0
        self.point_orientations = torch.cat((self.point_orientations, ident), 0)
        fresh_feat = self.geo_feature_std * torch.randn(n_new + 1, self.geo_feature_dim, device=dev, dtype=f32)
        self.geo_features = torch.cat((self.geo_features[:-1], fresh_feat), 0)
        if self.color_features is not None:
            fresh_col = self.color_feature_std * torch.randn(n_new + 1, self.color_feature_dim, device=dev, dtype=f32)
            self.color_features = torch.cat((self.color_features[:-1], fresh_col), 0)
        self.point_certainties = torch.cat((self.point_certainties, torch.zeros(n_new, device=dev, dtype=f32)), 0)
        self._touch()
        self.reset_local_map(sensor_position, sensor_orientation, cur_ts, reboot_map=True)
        return n_new / cand.shape[0]

    def reset_local_map(self, sensor_position: torch.Tensor, sensor_orientation: torch.Tensor, cur_ts: int,
                        use_travel_dist: bool = True, diff_ts_local: int = 50, reboot_map: bool = False) -> None:
        """Select the local window (travel-distance window on the creation stamp AND within
        local_map_radius of the sensor) and rebuild the local_* copies, the global->local remap and
        the trainable ``local_geo_features`` Parameter (model/neural_points.py:439-536)."""
        self._touch()
        self.cur_ts = cur_ts
        self.max_ts = max(self.max_ts, cur_ts)
        dev = self.device
        m = self.count()

        if m > 0 and self.neural_points.is_cuda and self.neural_points.dtype == torch.float32:
            from ..ops import mapmaint as _mm

            _mm.local_window(self, sensor_position, cur_ts, use_travel_dist, diff_ts_local, reboot_map)
            if self.color_features is not None:
                self.local_color_features = nn.Parameter(self.color_features[self.local_mask])
            self.local_orientation = sensor_orientation
            return

        if self.temporal_local_map_on:
            if self.config.use_mid_ts:
                stamp = ((self.point_ts_create + self.point_ts_update) / 2).int()
            else:
                stamp = self.point_ts_create
            if use_travel_dist:
                in_time = torch.abs(self.travel_dist[cur_ts] - self.travel_dist[stamp]) < self.diff_travel_dist_local
            else:
                in_time = torch.abs(cur_ts - stamp) < diff_ts_local
            if reboot_map:
                in_time = in_time & (stamp >= self.reboot_ts)
            if torch.sum(in_time) < 100:  # too few points in the window: take everything
                in_time = torch.ones(m, dtype=torch.bool, device=dev)
        else:
            in_time = torch.ones(m, dtype=torch.bool, device=dev)

        d2 = ((self.neural_points[in_time] - sensor_position) ** 2).sum(-1)
        chosen = torch.nonzero(in_time).squeeze(-1)[d2 < self.local_map_radius**2]
        self._local_gids = chosen  # ascending global ids of the local points (== nonzero(local_mask[:-1]))
        mask = torch.zeros(m, dtype=torch.bool, device=dev)
        mask[chosen] = True

        self.local_neural_points = self.neural_points[mask]
        self.local_point_orientations = self.point_orientations[mask]
        self.local_point_certainties = self.point_certainties[mask]
        self.local_point_ts_update = self.point_ts_update[mask]

        mask = torch.cat((mask, torch.ones(1, dtype=torch.bool, device=dev)))  # padding row is always local
        self.local_mask = mask
        g2l = torch.full((m + 1,), -1, dtype=torch.long, device=dev)
        rows = torch.nonzero(mask).flatten()
        g2l[rows] = torch.arange(rows.numel(), device=dev)
        g2l[-1] = -1
        self.global2local = g2l

        self.local_geo_features = nn.Parameter(self.geo_features[mask])
        if self.color_features is not None:
            self.local_color_features = nn.Parameter(self.color_features[mask])
        self.local_orientation = sensor_orientation

    def assign_local_to_global(self) -> None:
        """Write the trained local window back (model/neural_points.py:538-549)."""
        mask = self.local_mask
        gids = getattr(self, "_local_gids", None)
        if (self.geo_features.is_cuda and gids is not None and gids.is_cuda and gids.numel() == self.local_count()
                and self.local_geo_features.shape[0] == gids.numel() + 1 and self.geo_features.dtype == torch.float32):
            from ..ops import mapmaint as _mm

            _mm.assign_local_to_global(self)
            if self.color_features is not None:
                self.color_features[mask] = self.local_color_features.data
            return
        self.geo_features[mask] = self.local_geo_features.data
        if self.color_features is not None:
            self.color_features[mask] = self.local_color_features.data
        self.point_certainties[mask[:-1]] = self.local_point_certainties
        self.point_ts_update[mask[:-1]] = self.local_point_ts_update

    # ------------------------------------------------------------------ query (CUDA only)
    def query_feature(self, query_points: torch.Tensor, query_ts: Optional[torch.Tensor] = None,
                      training_mode: bool = True, query_locally: bool = True, query_geo_feature: bool = True,
                      query_color_feature: bool = False):
        """kNN over the voxel hash + inverse-distance feature blend (model/neural_points.py:553-769).

        Returns (geo_features_vector [N,F+3], color_features_vector | None, weight_vector [N,K,1],
        nn_counts [N] int64, queried_certainty [N]).  ``geo_features_vector`` is differentiable with
        respect to ``query_points`` and ``local_geo_features`` (including grad-of-grad for the
        analytic eikonal term); certainty / ts side effects happen in training mode."""
        if not query_geo_feature and not query_color_feature:
            sys.exit("you need to at least query one kind of feature")
        if query_color_feature and self.color_features is not None:
            raise NotImplementedError("colour features are outside the neural-SDF hot path")
        if not self.config.weighted_first:
            raise NotImplementedError("weighted_first=False is not used by any shipped configuration")
        z, w, nn_counts, certainty = _q.query_feature(self, query_points, query_ts, training_mode, query_locally)
        return z, None, w.unsqueeze(-1), nn_counts, certainty

    def radius_neighborhood_search(self, points: torch.Tensor, time_filtering: bool = False):
        """(dist2 [N,Kc], idx [N,Kc] global ids, -1 invalid) of the Kc probed cells
        (model/neural_points.py:971-1030).  query_feature does not call this (it probes inside the
        fused kernel); kept for API compatibility."""
        return _q.radius_search(self, points, time_filtering)

    def query_certainty(self, query_points: torch.Tensor) -> torch.Tensor:
        """Max global certainty over the probed cells (model/neural_points.py:1032-1051)."""
        return _q.query_certainty(self, query_points)

    # ------------------------------------------------------------------ map clean-up
    def prune_map(self, prune_certainty_thre, min_prune_count=500, global_prune=False) -> bool:
        """Drop uncertain (and, unless global_prune, inactive) points; the caller must
        recreate_hash() when True is returned (model/neural_points.py:771-812)."""
        drop = self.point_certainties < prune_certainty_thre
        if not global_prune:
            gap = torch.abs(self.travel_dist[self.cur_ts] - self.travel_dist[self.point_ts_update])
            drop = drop & (gap > self.diff_travel_dist_local)
        n_drop = int(drop.sum().item())
        if n_drop <= min_prune_count:
            return False
        if not self.silence:
            print("# Prune neural points: ", n_drop)
        self._touch()
        keep = ~drop
        self.neural_points = self.neural_points[keep]
        self.point_orientations = self.point_orientations[keep]
        self.point_ts_create = self.point_ts_create[keep]
        self.point_ts_update = self.point_ts_update[keep]
        self.point_certainties = self.point_certainties[keep]
        keep_pad = torch.cat((keep, torch.ones(1, dtype=torch.bool, device=keep.device)))
        self.geo_features = self.geo_features[keep_pad]
        if self.color_on:
            self.color_features = self.color_features[keep_pad]
        return True

    def adjust_map(self, pose_diff_torch) -> None:
        """Pose-graph correction of the map (model/neural_points.py:814-838).  CLID-SLAM never
        calls it (loop closure was removed upstream)."""
        raise NotImplementedError("adjust_map has no caller in CLID-SLAM and is outside the hot path")

    def recreate_hash(self, sensor_position, sensor_orientation, kept_points: bool = True, with_ts: bool = True,
                      cur_ts=0) -> None:
        """Refill the voxel hash from scratch, one winner per voxel (closest creation stamp, or
        highest certainty), optionally dropping the losers (model/neural_points.py:840-929)."""
        res = self.resolution
        self._touch()
        self.buffer_pt_index = torch.full((self.buffer_size,), -1, dtype=self.idx_dtype, device=self.device)
        if with_ts:
            if self.config.use_mid_ts:
                stamp = ((self.point_ts_create + self.point_ts_update) / 2).int()
            else:
                stamp = self.point_ts_create
            score = torch.abs(stamp - cur_ts).float()
        else:
            score = self.point_certainties.max() - self.point_certainties
        winners = voxel_down_sample_min_value_torch(self.neural_points, res, score)

        if kept_points:
            self._store_slots(self._slots_of(self.neural_points[winners]), winners)
        else:
            if not self.silence:
                print("Filter duplicated neural points")
            self.neural_points = self.neural_points[winners]
            self.point_orientations = self.point_orientations[winners]
            self.point_ts_create = self.point_ts_create[winners]
            self.point_ts_update = self.point_ts_update[winners]
            self.point_certainties = self.point_certainties[winners]
            pad = torch.cat((winners, torch.tensor([-1], device=winners.device, dtype=winners.dtype)))
            self.geo_features = self.geo_features[pad]
            if self.color_features is not None:
                self.color_features = self.color_features[pad]
            ids = torch.arange(self.count(), dtype=self.idx_dtype, device=self.device)
            self._store_slots(self._slots_of(self.neural_points), ids)

        if sensor_position is not None:
            self.reset_local_map(sensor_position, sensor_orientation, cur_ts)
        if not kept_points:
            self.record_memory(verbose=not self.silence)

    def set_search_neighborhood(self, num_nei_cells: int = 1, search_alpha: float = 1.0) -> None:
        """Cell offsets inside the sphere |d|^2 < (num_nei_cells + search_alpha)^2 and the matching
        validity radius (model/neural_points.py:931-969).  Toggled at run time by the mapper
        (mapper.py:409-423), so the kernels take the table as an argument."""
        r = torch.arange(-num_nei_cells, num_nei_cells + 1, dtype=self.primes.dtype)
        cube = torch.stack(torch.meshgrid(r, r, r, indexing="ij"), dim=-1).reshape(-1, 3)
        inside = (cube**2).sum(-1) < (num_nei_cells + search_alpha) ** 2
        self._neighbor_dx_cpu = cube[inside].contiguous()  # host copy: the brick-index build reads it without a sync
        self.neighbor_dx = self._neighbor_dx_cpu.to(self.primes.device)
        self.neighbor_K = self.neighbor_dx.shape[0]
        self.max_valid_dist2 = 3 * ((num_nei_cells + 1) * self.resolution) ** 2
        self._touch()

    def clear_temp(self, clean_more: bool = False) -> None:
        """Drop everything that is rebuilt on load before pickling (model/neural_points.py:1054-1074)."""
        self._touch()
        self._brick_cache = {}
        self.buffer_pt_index = None
        self.local_neural_points = None
        self.local_point_orientations = None
        self.local_geo_features = nn.Parameter()
        self.local_color_features = nn.Parameter()
        self.local_point_certainties = None
        self.local_point_ts_update = None
        self.local_mask = None
        self.global2local = None
        if clean_more:
            self.point_ts_create = None
            self.point_ts_update = None
            self.point_certainties = None

    # ------------------------------------------------------------------ viz exports (need open3d)
    def get_map_o3d_bbx(self):
        import open3d as o3d

        lo, _ = torch.min(self.neural_points, dim=0)
        hi, _ = torch.max(self.neural_points, dim=0)
        return o3d.geometry.AxisAlignedBoundingBox(lo.cpu().numpy(), hi.cpu().numpy())

    def get_neural_points_o3d(self, query_global: bool = True, color_mode: int = -1, random_down_ratio: int = 1):
        """Point cloud of the (global or local) neural points; colouring modes of the reference's
        GUI other than certainty (3) and random (4) are visualisation-only and not mirrored."""
        import numpy as np
        import open3d as o3d

        pts = self.neural_points if query_global else self.local_neural_points
        pts_np = pts[::random_down_ratio].detach().cpu().numpy().astype(np.float64)
        cloud = o3d.geometry.PointCloud()
        cloud.points = o3d.utility.Vector3dVector(pts_np)
        if color_mode == 3:
            cert = self.point_certainties if query_global else self.local_point_certainties
            grey = 1.0 - cert[::random_down_ratio].detach().cpu().numpy().astype(np.float64) / 1000.0
            cloud.colors = o3d.utility.Vector3dVector(np.repeat(grey.reshape(-1, 1), 3, axis=1))
        elif color_mode == 4:
            cloud.colors = o3d.utility.Vector3dVector(np.random.rand(pts_np.shape[0], 3))
        return cloud
