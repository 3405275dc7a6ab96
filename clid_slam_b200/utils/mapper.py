"""Incremental map optimisation: mirror of the reference's ``utils.mapper.Mapper``
(utils/mapper.py:35-1076) for the neural-SDF hot path.

Same constructor, attributes (replay-pool tensors, ``new_idx``, ``adaptive_iter_offset`` ...) and
method names; ``mapping(iter_count)`` is the training function ``slam.py:200`` calls (``train`` is
an alias for the name BASELINE.json uses).  The loop body runs through the fused CUDA path
(``ops.train.FusedTrainer``: three launches per iteration, no host synchronisation; long calls replay
them as one CUDA graph) whenever the
configuration is covered -- every shipped run file is -- and otherwise through the unfused CUDA
path (our ``query_feature`` autograd Function + the torch decoder + torch.optim), which covers the
sdf losses (bce / zhong / l1 / l2), projective correction, both eikonal modes and its sub-sets, SGD and
multi-level decoders.  The consistency loss and the semantic / colour heads (utils/mapper.py:717-743,
770-830; off in every shipped run file) are NOT implemented: mapping() raises for them instead of
silently training without those terms.  Neither path has a CPU fallback.

Not mirrored (outside the hot path, no caller in CLID-SLAM): ``bundle_adjustment``,
``get_ba_samples``, the Open3D pool export needs open3d at call time.
"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn.functional as F

from ..ops import train as _train
from .data_sampler import DataSampler
from .loss import sdf_bce_loss, sdf_diff_loss, sdf_zhong_loss
from .tools import get_gradient, setup_optimizer, transform_batch_torch, transform_torch


class Mapper:
    def __init__(self, config, dataset, neural_points, local_point_cloud_map, geo_mlp, sem_mlp=None, color_mlp=None):
        self.config = config
        self.silence = config.silence
        self.dataset = dataset
        self.neural_points = neural_points
        self.local_point_cloud_map = local_point_cloud_map
        self.geo_mlp = geo_mlp
        self.sem_mlp = sem_mlp
        self.color_mlp = color_mlp
        self.device = config.device
        self.dtype = config.dtype
        self.used_poses = None
        # analytic gradient is only built when something consumes it (utils/mapper.py:57-69)
        self.require_gradient = bool(config.ekional_loss_on or config.proj_correction_on or config.consistency_loss_on)
        if config.numerical_grad and not config.proj_correction_on and not config.consistency_loss_on:
            self.require_gradient = False
        self.total_iter = 0
        self.sdf_scale = config.logistic_gaussian_ratio * config.sigma_sigmoid_m

        self.sampler = DataSampler(config)
        self.ray_sample_count = 1 + config.surface_sample_n + config.free_behind_n + config.free_front_n

        self.new_idx = None
        self.ba_done_flag = False
        self.adaptive_iter_offset = 0
        self.use_fused = True          # set False to force the unfused CUDA path
        # mapping() calls with at least this many iterations capture [get_batch + iteration] once in a CUDA
        # graph and replay it (host cost per iteration ~15 us instead of ~140 us of Python / ctypes); 0 disables
        self.graph_min_iters = int(os.environ.get("CLID_MAPPING_GRAPH_MIN_ITERS", "24"))
        # mapping() through the native loop (clid_mapping_run); False: Python loop / CUDA graph over get_batch
        self.native_loop = os.environ.get("CLID_NATIVE_LOOP", "1") != "0"
        self.last_losses = None        # [iters,3] device tensor (total, bce, eikonal) of the last mapping() call
        self.last_host_ms = None       # host wall time of the last native-loop mapping() call: set-up / loop enqueue
        # multi-GPU (one process per GPU, set_shards): slab partition of the CURRENT local map and, for a partitioned
        # map, the band-row translation tables; the peer-memory link is kept across mapping() calls
        self.shards = None
        self.shard_peer_rows = None
        self.shard_group = None
        self._peer_link = None

        dev, f32 = self.device, self.dtype
        self.coord_pool = torch.empty((0, 3), device=dev, dtype=f32)
        self.global_coord_pool = torch.empty((0, 3), device=dev, dtype=f32)
        self.sdf_label_pool = torch.empty((0,), device=dev, dtype=f32)
        self.color_pool = torch.empty((0, getattr(config, "color_channel", 0)), device=dev, dtype=f32)
        self.sem_label_pool = torch.empty((0,), device=dev, dtype=torch.int)
        self.normal_label_pool = torch.empty((0, 3), device=dev, dtype=f32)
        self.weight_pool = torch.empty((0,), device=dev, dtype=f32)
        self.time_pool = torch.empty((0,), device=dev, dtype=torch.int)

    # ------------------------------------------------------------------ helpers used per frame
    def dynamic_filter(self, points_torch, type_2_on: bool = True):
        """Mask of static points: uncertain map regions or points not far inside free space
        (utils/mapper.py:99-136).  One fused launch gives sdf, gradient and certainty."""
        from .. import fused

        sdf, grad, _, certainty = fused.sdf_and_gradient(
            self.neural_points, self.geo_mlp, points_torch, training_mode=False, with_gradient=type_2_on)
        cfg = self.config
        static = (certainty < cfg.dynamic_certainty_thre) | (sdf < cfg.dynamic_sdf_ratio_thre * cfg.voxel_size_m)
        if type_2_on:
            steep = grad.norm(dim=-1) > cfg.dynamic_min_grad_norm_thre
            static = static & (steep | (certainty < cfg.dynamic_certainty_thre))
        return static

    def determine_used_pose(self):
        frame = self.dataset.processed_frame
        if self.config.pgo_on:
            src = self.dataset.pgo_poses
        elif self.config.track_on:
            src = self.dataset.odom_poses
        elif self.dataset.gt_pose_provided:
            src = self.dataset.gt_poses
        else:
            return
        self.used_poses = torch.tensor(src[: frame + 1], device=self.device, dtype=torch.float64)

    def process_frame(self, point_cloud_torch, frame_label_torch, cur_pose_torch, frame_id: int,
                      filter_dynamic: bool = False):
        """Per-frame feeder (utils/mapper.py:159-470): sample the scan, insert new neural points,
        append to / filter the replay pool, find the newly observed samples."""
        cfg = self.config
        origin = cur_pose_torch[:3, 3]
        orientation = cur_pose_torch[:3, :3]
        pts = point_cloud_torch[:, :3]

        if not cfg.use_pin_mapper:
            self.local_point_cloud_map.update_map(origin, transform_torch(pts, cur_pose_torch))

        self.static_mask = torch.ones(pts.shape[0], dtype=torch.bool, device=cfg.device)
        if filter_dynamic:
            self.neural_points.reset_local_map(origin, orientation, frame_id)
            self.static_mask = self.dynamic_filter(transform_torch(pts, cur_pose_torch))
            if not self.silence:
                print("# Dynamic points filtered: ", int((self.static_mask == 0).sum().item()))
            pts = pts[self.static_mask]
            if frame_label_torch is not None:
                frame_label_torch = frame_label_torch[self.static_mask]
        self.dataset.static_mask = self.static_mask

        if cfg.use_pin_mapper:
            coord, sdf_label, normal_label, sem_label, color_label, weight = self.sampler.sample_pin(
                pts, None, frame_label_torch, None)
        else:
            normal_label = sem_label = color_label = None
            coord, sdf_label, weight = self.sampler.sample(pts, self.local_point_cloud_map, cur_pose_torch)

        stamp = torch.full((coord.shape[0],), frame_id, dtype=torch.int, device=self.device)
        self.cur_sample_count = sdf_label.shape[0]
        self.pool_sample_count = self.sdf_label_pool.shape[0]

        # which points seed new neural points (utils/mapper.py:258-271)
        if cfg.from_sample_points:
            if cfg.from_all_samples:
                seeds = coord
            else:
                near = torch.abs(sdf_label) < cfg.surface_sample_range_m * cfg.map_surface_ratio
                seeds = transform_torch(coord[near, :], cur_pose_torch)
        else:
            seeds = transform_torch(pts, cur_pose_torch)

        if cfg.prune_map_on and (frame_id + 1) % cfg.prune_freq_frame == 0:
            if self.neural_points.prune_map(cfg.max_prune_certainty):
                self.neural_points.recreate_hash(None, None, True, True, frame_id)

        self.cur_new_point_ratio = self.neural_points.update(seeds, origin, orientation, frame_id)
        if not self.silence:
            self.neural_points.print_memory()

        # replay pool
        self.coord_pool = torch.cat((self.coord_pool, coord), 0)
        self.weight_pool = torch.cat((self.weight_pool, weight), 0)
        self.sdf_label_pool = torch.cat((self.sdf_label_pool, sdf_label), 0)
        self.time_pool = torch.cat((self.time_pool, stamp), 0)
        self.sem_label_pool = torch.cat((self.sem_label_pool, sem_label), 0) if sem_label is not None else None
        self.color_pool = torch.cat((self.color_pool, color_label), 0) if color_label is not None else None
        self.normal_label_pool = (torch.cat((self.normal_label_pool, normal_label), 0)
                                  if normal_label is not None else None)

        self.determine_used_pose()
        if self.ba_done_flag:
            self.global_coord_pool = transform_batch_torch(self.coord_pool, self.used_poses[self.time_pool])
            self.ba_done_flag = False
        else:
            self.global_coord_pool = torch.cat((self.global_coord_pool, transform_torch(coord, cur_pose_torch)), 0)

        if (frame_id + 1) % cfg.pool_filter_freq == 0 and self._filter_pool_native(origin, normal_label, sem_label, color_label):
            pass  # compacted by clid_pool_filter_select / clid_compact_rows (ops/mapmaint.py)
        elif (frame_id + 1) % cfg.pool_filter_freq == 0:
            d2 = ((self.global_coord_pool - origin) ** 2).sum(-1)
            keep = d2 < cfg.window_radius**2
            kept_idx = torch.nonzero(keep).squeeze(-1)
            n_keep = kept_idx.shape[0]
            if n_keep > cfg.pool_capacity:  # random discard down to capacity
                drop = torch.randint(0, n_keep, (n_keep - int(cfg.pool_capacity),), device=self.device)
                keep[kept_idx[drop]] = False
            self.coord_pool = self.coord_pool[keep]
            self.global_coord_pool = self.global_coord_pool[keep]
            self.sdf_label_pool = self.sdf_label_pool[keep]
            self.weight_pool = self.weight_pool[keep]
            self.time_pool = self.time_pool[keep]
            if normal_label is not None:
                self.normal_label_pool = self.normal_label_pool[keep]
            if sem_label is not None:
                self.sem_label_pool = self.sem_label_pool[keep]
            if color_label is not None:
                self.color_pool = self.color_pool[keep]
            self.cur_sample_count = int(keep[-self.cur_sample_count:].sum().item())
            self.pool_sample_count = int(keep.sum().item())
        else:
            self.cur_sample_count = coord.shape[0]
            self.pool_sample_count = self.coord_pool.shape[0]

        if cfg.bs_new_sample > 0:
            # samples of this frame that fall into not-yet-certain voxels get replayed more often
            fresh = self.global_coord_pool[-self.cur_sample_count:]
            fresh_label = self.sdf_label_pool[-self.cur_sample_count:]
            n_fresh = fresh.shape[0]
            certainty = torch.zeros(n_fresh, device=self.device)
            self.neural_points.set_search_neighborhood(num_nei_cells=1, search_alpha=0.0)
            bs = cfg.infer_bs
            for head in range(0, n_fresh, bs):
                certainty[head:head + bs] = self.neural_points.query_certainty(fresh[head:head + bs, :])
            self.neural_points.set_search_neighborhood(num_nei_cells=cfg.num_nei_cells, search_alpha=cfg.search_alpha)

            self.new_idx = torch.where(
                (certainty < cfg.new_certainty_thre) & (torch.abs(fresh_label) < cfg.surface_sample_range_m * 3.0))[0]
            self.new_idx += self.pool_sample_count - self.cur_sample_count
            n_new = self.new_idx.shape[0]

            self.adaptive_iter_offset = 0
            ratio = n_new / max(self.cur_sample_count, 1)
            if cfg.adaptive_iters:
                if ratio < cfg.new_sample_ratio_less:
                    self.adaptive_iter_offset = -5
                elif ratio > cfg.new_sample_ratio_more:
                    self.adaptive_iter_offset = 5
                    if frame_id > cfg.freeze_after_frame and ratio > cfg.new_sample_ratio_restart:
                        self.adaptive_iter_offset = 10

    def _filter_pool_native(self, origin, normal_label, sem_label, color_label) -> bool:
        """The pool filter of process_frame (utils/mapper.py:420-459) as one flag + scan + compaction on the device.
        False (the torch ops below run instead) on CPU pools, with auxiliary label pools, or when the kept samples
        exceed pool_capacity (the reference then discards at random through torch.randint)."""
        cfg = self.config
        gc = self.global_coord_pool
        if (not gc.is_cuda or gc.dtype != torch.float32 or gc.shape[0] == 0
                or normal_label is not None or sem_label is not None or color_label is not None):
            return False
        from ..ops import mapmaint as _mm

        arrays = [self.coord_pool, gc, self.sdf_label_pool, self.weight_pool, self.time_pool]
        outs, n_keep, n_tail_keep, _ = _mm.pool_filter(gc, origin, cfg.window_radius, arrays, self.cur_sample_count)
        if n_keep > cfg.pool_capacity:
            return False
        self.coord_pool, self.global_coord_pool, self.sdf_label_pool, self.weight_pool, self.time_pool = outs
        self.cur_sample_count, self.pool_sample_count = n_tail_keep, n_keep
        return True

    # ------------------------------------------------------------------ batches
    def get_batch(self, global_coord=False):
        """Uniform draw from the replay pool; up to bs_new_sample of the batch come from this
        frame's newly observed samples (utils/mapper.py:473-523).  Returns the reference's 7-tuple."""
        cfg = self.config
        bs = cfg.bs
        use_new = (cfg.bs_new_sample > 0 and self.new_idx is not None and not self.dataset.lose_track
                   and not self.dataset.stop_status and self.new_idx.shape[0] > 0)
        if use_new:
            n_new_pool = self.new_idx.shape[0]
            bs_new = min(n_new_pool, cfg.bs_new_sample)
            history = torch.randint(0, self.pool_sample_count, (bs - bs_new,), device=self.device)
            recent = self.new_idx[torch.randint(0, n_new_pool, (bs_new,), device=self.device)]
            index = torch.cat((history, recent), dim=0)
        else:
            index = torch.randint(0, self.pool_sample_count, (bs,), device=self.device)
        coord = self.global_coord_pool[index, :] if global_coord else self.coord_pool[index, :]
        sdf_label = self.sdf_label_pool[index]
        ts = self.time_pool[index]
        weight = self.weight_pool[index]
        sem_label = self.sem_label_pool[index] if self.sem_label_pool is not None else None
        color_label = self.color_pool[index] if self.color_pool is not None else None
        normal_label = self.normal_label_pool[index, :] if self.normal_label_pool is not None else None
        return coord, sdf_label, ts, normal_label, sem_label, color_label, weight

    # ------------------------------------------------------------------ training
    def mapping(self, iter_count):
        """`iter_count` (+ adaptive offset) Adam iterations on batches from the replay pool, then the
        local window is written back to the global map (utils/mapper.py:620-862)."""
        iter_count = max(1, iter_count + self.adaptive_iter_offset)
        why_not = _train.supported(self.config, self.geo_mlp) if self.use_fused else "use_fused=False"
        if why_not is None:
            self._mapping_fused(iter_count)
        else:
            self._mapping_unfused(iter_count)
        self.neural_points.assign_local_to_global()

    train = mapping  # BASELINE.json's name for the same entry point

    # ------------------------------------------------------------------ multi-GPU
    def set_shards(self, shards, peer_rows=None, group=None) -> None:
        """Run mapping() sharded over the ranks of torch.distributed (clid_slam_b200/dist.py): `shards` is the
        SpatialShards of the current local map; every rank trains on the replay-pool samples of its own slab
        (config.bs // world of them per iteration), band gradients and [decoder gradients | loss] cross NVLink
        peer memory inside the step.  peer_rows: dist.peer_row_tables(...) when every rank holds only its part of
        the map (partitioned), None when the map is replicated.  The shards describe one state of the local map:
        call again (or set_shards(None)) after NeuralPoints.update / reset_local_map."""
        self.shards, self.shard_peer_rows, self.shard_group = shards, peer_rows, group

    def _mapping_sharded(self, iter_count: int) -> None:
        from .. import dist as _dist

        cfg, npm, dev = self.config, self.neural_points, self.device
        rank, world = _dist.world()
        shards, tabs = self.shards, self.shard_peer_rows
        npm.brick_index(True)
        trainer = _train.FusedTrainer(cfg, npm, self.geo_mlp)
        n_small = (trainer.dec_grad.numel() if trainer.dec_grad is not None else 0) + 3
        if self._peer_link is not None and (self._peer_link.capacity_rows < trainer.rows or self._peer_link.stride < n_small):
            self._peer_link.close()
            self._peer_link = None
        if self._peer_link is None:
            self._peer_link = _dist.PeerLink(trainer.rows, npm.geo_feature_dim, n_small, torch.device(dev),
                                             group=self.shard_group, capacity_rows=int(trainer.rows * 1.5) + 1024)
        trainer.attach_peers(shards, group=self.shard_group, peer_rows=tabs, link=self._peer_link)

        pool = self.global_coord_pool[: self.pool_sample_count]
        own = torch.nonzero(shards.owner_of(pool) == rank).flatten()
        bs_rank = max(1, cfg.bs // world)
        n_global = bs_rank * world
        nd_global = world * ((bs_rank + cfg.gradient_decimation - 1) // cfg.gradient_decimation) if trainer.numerical else 0
        cert_before = npm.local_point_certainties.clone()
        history = []
        for _ in range(iter_count):
            if own.numel() > 0:
                index = own[torch.randint(0, own.numel(), (bs_rank,), device=dev)]
            else:  # a slab without samples still takes part in the step's all-reduce
                index = own
            loss = trainer.iteration(pool[index], self.sdf_label_pool[index], self.time_pool[index], self.weight_pool[index],
                                     n_global=n_global, nd_global=nd_global)
            history.append(loss.clone())
            self.total_iter += 1
        trainer.peer.check()
        # per-row side effects of the whole call (neural_points.py:708-733), completed across ranks once
        if tabs is not None:
            delta = npm.local_point_certainties - cert_before
            _dist.exchange_band_values(shards, rank, npm.local_neural_points, delta, "sum", group=self.shard_group)
            npm.local_point_certainties.copy_(cert_before + delta)
            _dist.exchange_band_values(shards, rank, npm.local_neural_points, npm.local_point_ts_update, "max",
                                       group=self.shard_group)
        else:
            _dist.reduce_side_effects(npm.local_point_certainties, cert_before, npm.local_point_ts_update, group=self.shard_group)
            shards.gather_features(npm.local_geo_features.data, rank, group=self.shard_group)
        self.last_losses = torch.stack(history) if history else None
        del trainer.losses[:]
        self._log_losses()

    def _batch_in_global_frame(self):
        coord, sdf_label, ts, _, sem_label, color_label, weight = self.get_batch(global_coord=not self.ba_done_flag)
        if self.ba_done_flag:
            coord = transform_batch_torch(coord, self.used_poses[ts])
        return coord, sdf_label, ts, sem_label, color_label, weight

    def _mapping_fused(self, iter_count: int) -> None:
        import time

        if self.shards is not None:
            from .. import dist as _dist

            if _dist.world()[1] > 1:
                return self._mapping_sharded(iter_count)
        t0 = time.perf_counter()
        self.neural_points.brick_index(True)  # (re)built here when the map changed since the last query: one small read-back
        trainer = _train.FusedTrainer(self.config, self.neural_points, self.geo_mlp)
        t1 = time.perf_counter()

        def body():
            coord, sdf_label, ts, _, _, weight = self._batch_in_global_frame()
            return trainer.iteration(coord, sdf_label, ts, weight)

        stock_batches = getattr(self.get_batch, "__func__", None) is Mapper.get_batch
        # Native loop (the default): the whole call -- replay-pool draw, iteration, optimiser step, x iter_count --
        # is enqueued by ONE C call (clid_mapping_run), ~10 us of host time per iteration instead of ~140 us of
        # Python / ctypes, for the 10-iteration calls of every shipped run file as well.  The draw uses the
        # library's counter-based generator (seeded from torch's), not torch.randint's sequence; a get_batch
        # replaced by the caller (tests feed recorded batches) runs call by call below.
        one_kernel = (not trainer.numerical or trainer.weight_e == 0 or int(self.config.gradient_decimation) == 10)
        if (self.native_loop and stock_batches and not self.ba_done_flag and one_kernel
                and torch.device(self.device).type == "cuda" and self.pool_sample_count > 0):
            seed = int(torch.randint(0, 2**62, (1,)).item())  # CPU generator: follows torch.manual_seed, no device sync
            self.last_losses = trainer.run_loop(self, iter_count, seed, global_coord=True)
            # host-side cost of this call: index rebuild + trainer set-up, and the enqueue of all iterations
            self.last_host_ms = {"setup": (t1 - t0) * 1e3, "loop_enqueue": (time.perf_counter() - t1) * 1e3, "iterations": iter_count}
            self.total_iter += iter_count
            self._log_losses()
            return

        # the replay pool draw (torch.randint on the CUDA generator, fixed shapes) and the iteration are
        # graph-safe; a get_batch replaced by the caller (tests feed recorded batches) is not assumed to be
        graphable = (self.graph_min_iters > 0 and iter_count >= self.graph_min_iters and stock_batches
                     and torch.device(self.device).type == "cuda")
        if not graphable:
            for _ in range(iter_count):
                body()
                self.total_iter += 1
            self.last_losses = torch.stack(trainer.losses) if trainer.losses else None
            self._log_losses()
            return

        dev = torch.device(self.device)
        trainer.ensure_step_state()  # device-side Adam step counter
        history = torch.empty(iter_count, 3, dtype=torch.float32, device=dev)
        history[0].copy_(body())  # first iteration eagerly: builds the brick index, configures the kernels
        trainer._want_overlap()  # the side stream of the forked optimiser step must exist before the capture
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize(dev)
        with torch.cuda.graph(graph):
            loss = body()
        trainer.step -= 1  # the capture advanced the host-side bookkeeping without running anything
        for i in range(1, iter_count):
            graph.replay()
            trainer.step += 1
            history[i].copy_(loss)
        self.total_iter += iter_count
        del trainer.losses[:]
        self.last_losses = history
        self._log_losses()

    def _mapping_unfused(self, iter_count: int) -> None:
        cfg = self.config
        missing = [name for name in ("consistency_loss_on", "semantic_on", "color_on") if getattr(cfg, name, False)]
        if missing:
            raise NotImplementedError(
                f"{', '.join(missing)}: the consistency loss and the semantic / colour heads (utils/mapper.py:717-743, "
                "770-830 of the reference) are outside the neural-SDF hot path and not implemented")
        feat_params = list(self.neural_points.parameters())
        opt = setup_optimizer(
            cfg, feat_params, list(self.geo_mlp.parameters()),
            list(self.sem_mlp.parameters()) if cfg.semantic_on else None,
            list(self.color_mlp.parameters()) if cfg.color_on else None)
        history = []
        for _ in range(iter_count):
            coord, sdf_label, ts, sem_label, color_label, weight = self._batch_in_global_frame()
            if self.require_gradient:
                coord.requires_grad_(True)
            geo_feature, _, weight_knn, _, _ = self.neural_points.query_feature(coord, ts)
            sdf_pred = self.geo_mlp.sdf(geo_feature)

            g = None
            if self.require_gradient:
                g = get_gradient(coord, sdf_pred)
            elif cfg.numerical_grad:
                step = cfg.voxel_size_m * cfg.num_grad_step_ratio
                dec = cfg.gradient_decimation
                g = self.get_numerical_gradient(coord[::dec], sdf_pred[::dec], step)
            if cfg.proj_correction_on:
                origins = self.used_poses[ts][:, :3, 3]
                sdf_label = sdf_label * torch.abs(F.cosine_similarity(g, coord - origins))

            w_abs = torch.abs(weight).detach()
            if cfg.main_loss_type == "bce":
                sdf_loss = sdf_bce_loss(sdf_pred, sdf_label, self.sdf_scale, w_abs, cfg.loss_weight_on)
            elif cfg.main_loss_type == "zhong":
                sdf_loss = sdf_zhong_loss(sdf_pred, sdf_label, None, w_abs, cfg.loss_weight_on)
            elif cfg.main_loss_type == "sdf_l1":
                sdf_loss = sdf_diff_loss(sdf_pred, sdf_label, w_abs, l2_loss=False)
            elif cfg.main_loss_type == "sdf_l2":
                sdf_loss = sdf_diff_loss(sdf_pred, sdf_label, w_abs, l2_loss=True)
            else:
                sys.exit("Please choose a valid loss type")
            total = sdf_loss

            eikonal = torch.zeros((), device=coord.device)
            if cfg.ekional_loss_on and cfg.weight_e > 0 and g is not None:
                near = (torch.abs(sdf_label) < cfg.surface_sample_range_m)[:: cfg.gradient_decimation]
                if cfg.ekional_add_to == "freespace":
                    g_used = g[~near]
                elif cfg.ekional_add_to == "surface":
                    g_used = g[near]
                else:
                    g_used = g
                eikonal = ((g_used.norm(2, dim=-1) - 1.0) ** 2).mean()
                total = total + cfg.weight_e * eikonal

            opt.zero_grad(set_to_none=True)
            total.backward(retain_graph=False)
            opt.step()
            self.total_iter += 1
            history.append(torch.stack((total.detach(), sdf_loss.detach(), eikonal.detach())))
        self.last_losses = torch.stack(history) if history else None
        self._log_losses()

    def _log_losses(self) -> None:
        if not getattr(self.config, "wandb_vis_on", False) or self.last_losses is None:
            return
        import wandb

        rows = self.last_losses.cpu()
        first = self.total_iter - rows.shape[0] + 1
        for i, (total, bce, eik) in enumerate(rows.tolist()):
            wandb.log({"iter": first + i, "loss/total_loss": total, "loss/sdf_loss": bce, "loss/eikonal_loss": eik})

    # ------------------------------------------------------------------ short-hands kept from the reference
    def sdf(self, x, get_std=False):
        """(sdf prediction, None) with the reference's defaults: training-mode query, no timestamps
        (utils/mapper.py:968-982)."""
        geo_feature, _, _, _, _ = self.neural_points.query_feature(x)
        return self.geo_mlp.sdf(geo_feature), None

    def get_numerical_gradient(self, x, sdf_x=None, eps=0.02, two_side=True):
        """Central (or forward) differences of Mapper.sdf, kept in the autograd graph
        (utils/mapper.py:985-1034)."""
        n = x.shape[0]
        basis = torch.eye(3, dtype=x.dtype, device=x.device) * eps
        if two_side:
            probes = torch.cat([x + sgn * basis[a] for a in range(3) for sgn in (1.0, -1.0)], dim=0)
            s = self.sdf(probes)[0].unsqueeze(-1)
            cols = [(s[2 * a * n:(2 * a + 1) * n] - s[(2 * a + 1) * n:(2 * a + 2) * n]) / (2 * eps) for a in range(3)]
        else:
            probes = torch.cat([x + basis[a] for a in range(3)], dim=0)
            s = self.sdf(probes)[0].unsqueeze(-1)
            base = sdf_x.unsqueeze(-1)
            cols = [(s[a * n:(a + 1) * n] - base) / eps for a in range(3)]
        return torch.cat(cols, dim=1)

    # ------------------------------------------------------------------ pool utilities
    def free_pool(self):
        self.coord_pool = self.global_coord_pool = self.weight_pool = None
        self.sdf_label_pool = self.time_pool = None
        self.sem_label_pool = self.color_pool = self.normal_label_pool = None

    def get_data_pool_o3d(self, down_rate=1, only_cur_data=False):
        """Replay pool as an Open3D point cloud coloured by label sign (visualisation only)."""
        import numpy as np
        import open3d as o3d

        pool = self.global_coord_pool if self.global_coord_pool is not None else self.coord_pool
        label = self.sdf_label_pool
        if only_cur_data:
            pool, label = pool[-self.cur_sample_count:], label[-self.cur_sample_count:]
        pts = pool[::down_rate].detach().cpu().numpy().astype(np.float64)
        lab = label[::down_rate].detach().cpu().numpy()
        cloud = o3d.geometry.PointCloud()
        cloud.points = o3d.utility.Vector3dVector(pts)
        colors = np.zeros((pts.shape[0], 3))
        colors[lab > 0, 0] = 1.0
        colors[lab <= 0, 2] = 1.0
        cloud.colors = o3d.utility.Vector3dVector(colors)
        return cloud
