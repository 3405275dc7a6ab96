"""Fused training iteration of the neural-SDF map (the loop body of Mapper.mapping,
utils/mapper.py:642-836 of the reference) on top of the C ABI:

    clid_train_fused           search + IDW blend + decoder (+ closed-form d sdf / d x) + bce / eikonal loss +
                               feature-gradient scatter + certainty / ts side effects, one kernel; one 64-byte
                               decoder-gradient row per evaluated point into scratch
    clid_decoder_grad_reduce   dense reduction of the rows into the flat decoder gradient
    clid_adam_step             Adam on the touched feature rows and the decoder tensors

Three launches per iteration (plus a one-thread kernel when the step counter lives on the device), no host
synchronisation, no intermediate [N, 81, .] tensors.  Configurations the one-kernel path does not cover
(explicit eikonal subsets, decimations other than 10) run as clid_query_forward + clid_sdf_loss +
clid_train_backward.  One `FusedTrainer` lives for one `mapping()` call, exactly like the reference's per-call
Adam (fresh moments, step counter from 1).  `StepPipeline` replays the whole iteration as a CUDA graph.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import torch

from .. import _lib
from . import query as _q

INT32_MIN = -(2**31)


def supported(config, decoder) -> Optional[str]:
    """None when the fused path covers this configuration, else the reason it does not."""
    if not config.weighted_first:
        return "weighted_first=False"
    if config.main_loss_type != "bce":
        return f"main_loss_type={config.main_loss_type}"
    if getattr(config, "proj_correction_on", False) or getattr(config, "consistency_loss_on", False):
        return "projective correction / consistency loss"
    if getattr(config, "semantic_on", False) or getattr(config, "color_on", False):
        return "semantic / colour heads"
    if not config.opt_adam:
        return "SGD optimiser"
    if getattr(config, "ekional_add_to", "all") != "all" and config.ekional_loss_on and config.weight_e > 0:
        return f"ekional_add_to={config.ekional_add_to}"
    shape = (decoder.layers[0].out_features, len(decoder.layers))
    if shape not in ((32, 1), (64, 1), (128, 1), (32, 2)):  # the decoders train_fused_kernel is compiled for
        return f"decoder {shape[0]} x {shape[1]}"
    if decoder.layers[0].in_features != config.feature_dim + 3 or config.feature_dim != 8:
        return "feature_dim != 8 or positional encoding"
    return None


def draw_batch(coord: torch.Tensor, sdf_label: torch.Tensor, weight: torch.Tensor, time: torch.Tensor, count: int,
               n: int, seed: int, offset: int, new_idx: Optional[torch.Tensor] = None, bs_new: int = 0):
    """Mapper.get_batch on the library's generator (clid_draw_batch): n pool rows, the last bs_new of them out of
    new_idx.  Returns (x [n,3], label [n], weight [n], ts [n] int32, index [n] int64)."""
    lib = _lib.load()
    _q._require_cuda(coord, "replay pool")
    dev = coord.device
    pool = _lib.ClidReplayPool()
    pool.coord = _lib.ptr(coord, torch.float32, "coord pool")
    pool.sdf_label = _lib.ptr(sdf_label, torch.float32, "sdf_label_pool")
    pool.weight = _lib.ptr(weight, torch.float32, "weight_pool")
    pool.time = _lib.ptr(time, torch.int32, "time_pool")
    pool.count = int(count)
    if bs_new > 0:
        new_idx = new_idx.to(torch.int64).contiguous()
        pool.new_idx, pool.n_new, pool.bs_new = new_idx.data_ptr(), int(new_idx.shape[0]), int(bs_new)
    x = torch.empty(n, 3, dtype=torch.float32, device=dev)
    label = torch.empty(n, dtype=torch.float32, device=dev)
    w = torch.empty(n, dtype=torch.float32, device=dev)
    ts = torch.empty(n, dtype=torch.int32, device=dev)
    index = torch.empty(n, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.clid_draw_batch(C.byref(pool), n, int(seed) & (2**64 - 1), int(offset), x.data_ptr(), label.data_ptr(),
                                 w.data_ptr(), ts.data_ptr(), index.data_ptr(), _lib.current_stream(dev))
    _lib.check(rc, "clid_draw_batch")
    return x, label, w, ts, index


class FusedTrainer:
    def __init__(self, config, neural_points, decoder):
        why = supported(config, decoder)
        if why is not None:
            raise NotImplementedError(f"fused training does not cover {why}")
        self.cfg, self.npm, self.dec = config, neural_points, decoder
        self.lib = _lib.load()
        feats = neural_points.local_geo_features
        _q._require_cuda(feats, "local_geo_features")
        dev = feats.device
        self.device = dev
        rows = feats.shape[0]
        self.rows = rows
        self.feat_grad = torch.zeros_like(feats.data)
        self.feat_m = torch.zeros_like(feats.data)
        self.feat_v = torch.zeros_like(feats.data)
        self.dense = float(config.weight_decay) != 0.0
        self.touched = None if self.dense else torch.zeros(rows, dtype=torch.uint8, device=dev)
        # The touched flags only save traffic: dense Adam is the identity on a row whose g = m = v = 0, so
        # once the gathers of this mapping() call can have covered the table a couple of times the flags
        # are dropped (no more scattered flag stores in the kernels, no flag reads in Adam).
        self._gathers = 0
        self.dense_switch = 2.0     # switch when gathers so far > dense_switch * rows
        self.train_features = bool(feats.requires_grad)

        self.dec_tensors = decoder.flat_parameters()
        self.train_decoder = any(p is not None and p.requires_grad for p in self.dec_tensors)
        self.dec_numel = [0 if p is None else p.numel() for p in self.dec_tensors]
        # absent biases still own a slot in the flat layout the kernels use
        h = decoder.layers[0].out_features
        expect = [h * decoder.layers[0].in_features, h] + [h * h, h] * (len(decoder.layers) - 1) + [h, 1]
        self.dec_numel = [max(a, b) for a, b in zip(self.dec_numel, expect)]
        total = sum(self.dec_numel)
        if self.train_decoder:
            self.dec_grad = torch.zeros(total, dtype=torch.float32, device=dev)
            self.dec_m = torch.zeros_like(self.dec_grad)
            self.dec_v = torch.zeros_like(self.dec_grad)
        else:
            self.dec_grad = self.dec_m = self.dec_v = None
        self.step = 0
        self.analytic = bool(config.ekional_loss_on and not config.numerical_grad)
        self.numerical = bool(config.ekional_loss_on and config.numerical_grad)
        self.weight_e = float(config.weight_e) if config.ekional_loss_on else 0.0
        self.losses: List[torch.Tensor] = []  # [3] per iteration: total, bce, eikonal (device tensors)
        self.forward_events = None  # bench hook: list that receives (start, end) CUDA events of the forward
        self.backward_events = None
        self.launches = 0           # kernels of libclid_sdf.so launched so far
        # forward + loss + backward run as ONE kernel (clid_train_fused), analytic and numerical eikonal;
        # set False to use the three-launch path (clid_query_forward / clid_sdf_loss / clid_train_backward)
        self.single_kernel = True
        self._scratch = None        # device scratch of clid_train_fused (rows for the decoder-gradient reduction)
        self.use_scratch = True     # False: the warps fold the decoder gradients inside the one kernel
        # device-resident Adam step counter {step, step_size, bc2_sqrt, pad} (ClidAdamArgs.step_state):
        # lets a whole iteration be captured in a CUDA graph (StepPipeline); None = host-side counter
        self.step_state = None
        # single GPU, apply_step=True: the decoder-gradient reduction and the decoder's Adam step run on a side
        # stream concurrently with the (HBM-bound) Adam step of the feature rows, which does not depend on them
        self.overlap_decoder = True
        self._side_stream = None
        self._pending_reduce = None
        self._loop_bufs = None      # [bs] batch scratch of run_loop
        # multi-GPU over peer memory (attach_peers): dist.PeerLink + the slab geometry of this rank
        self.peer = None
        self._peer_geom = None
        self._peer_rows = None

    # ------------------------------------------------------------------
    def _shifted(self, x: torch.Tensor, eik_index: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The six central-difference copies of the decimated batch (mapper.py:985-1003)."""
        eps = self.cfg.voxel_size_m * self.cfg.num_grad_step_ratio
        xd = x[:: self.cfg.gradient_decimation] if eik_index is None else x[eik_index]
        out = xd.repeat(6, 1)
        nd = xd.shape[0]
        for axis in range(3):
            out[(2 * axis) * nd:(2 * axis + 1) * nd, axis] += eps
            out[(2 * axis + 1) * nd:(2 * axis + 2) * nd, axis] -= eps
        return out

    def iteration(self, x: torch.Tensor, label: torch.Tensor, ts: Optional[torch.Tensor], weight: torch.Tensor,
                  apply_step: bool = True, n_global: int = 0, nd_global: int = 0, sync: bool = False,
                  shards=None, eik_index: Optional[torch.Tensor] = None, exchange: bool = True,
                  parity: Optional[int] = None):
        """One mapping iteration on the batch.  With apply_step=False the optimiser step is left to
        a later `adam_step()` call, so the accumulated gradients can be inspected (tests).

        Sharded use (one process per GPU): every rank passes its slice of the batch, the GLOBAL
        batch size `n_global` (and decimated count `nd_global` in numerical mode) as the loss
        normalisers, and sync=True so gradients and losses are all-reduced before the step."""
        cfg, npm, dec, lib, dev = self.cfg, self.npm, self.dec, self.lib, self.device
        x = _q._prep_points(x, "coord")
        n = x.shape[0]
        if n == 0 and not sync and shards is None:
            return
        numerical = self.numerical and self.weight_e > 0
        self._gathers += n * int(cfg.query_nn_k)
        if self.touched is not None and shards is None and self._gathers > self.dense_switch * self.rows:
            # dense from here on (exact, see __init__).  Not with spatial shards: a rank only ever touches
            # the rows of its slab, and the flags keep its Adam step off the other (N-1)/N of the table.
            self.touched = None
        # the one-kernel numerical mode evaluates the x[::10] subset inside the warp tiles; explicit
        # subsets (eik_index, sharded batches) and other decimations take the three-launch path
        one_kernel_ok = (not numerical) or (int(cfg.gradient_decimation) == 10 and eik_index is None)
        if self.peer is not None:
            # peer-memory step: the fused kernel exchanges the band gradients itself, one kernel pair all-reduces the rest
            if not (self.single_kernel and one_kernel_ok):
                raise NotImplementedError("the peer-memory step runs the one-kernel iteration")
            self._parity = int(self.step % 2 if parity is None else parity)
            self.feat_grad = self.peer.grad[self._parity]
            loss = self._iteration_single_kernel(x, label, ts, weight, n_global, nd_global, numerical, defer_reduce=False)
            self._peer_all_reduce(loss)
            self.losses.append(loss)
            if apply_step:
                self.adam_step()
            return loss
        if self.single_kernel and one_kernel_ok:
            defer = bool(apply_step and exchange and shards is None and not sync and self._want_overlap())
            # with the device-side step counter the optimiser step of this iteration is announced at its head
            # (clid_step_begin clears the loss and advances the counter in one launch): the advance is then not a
            # node between the fused kernel and Adam
            self._begin_step = bool(defer and self.step_state is not None)
            loss = self._iteration_single_kernel(x, label, ts, weight, n_global, nd_global, numerical, defer_reduce=defer)
            if not exchange:  # the caller runs pack / all-reduce / unpack / adam_step itself (StepPipeline)
                self.losses.append(loss)
                return loss
            return self._finish_iteration(loss, apply_step, sync, shards)
        nd = 0
        x_all, ts_all = x, ts
        if self.numerical and self.weight_e > 0:
            shifted = self._shifted(x, eik_index)
            nd = shifted.shape[0] // 6
            x_all = torch.cat((x, shifted), 0)
            if ts is not None:  # INT32_MIN makes the ts_update amax a no-op for the shifted copies
                pad = torch.full((shifted.shape[0],), INT32_MIN, dtype=torch.int32, device=dev)
                ts_all = torch.cat((ts.to(torch.int32), pad), 0)
        want_grad = self.analytic and self.weight_e > 0
        if self.forward_events is not None:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        res = _q.forward(npm, dec, x_all, ts_all, training_mode=True, query_locally=True,
                         want_sdf=True, want_grad=want_grad, want_idx=True)
        if self.forward_events is not None:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            self.forward_events.append((ev0, ev1))
        n_all = x_all.shape[0]
        self.launches += 3 if n_all > 0 else 0

        loss = torch.zeros(3, dtype=torch.float32, device=dev)
        dlogit = torch.empty(n_all, dtype=torch.float32, device=dev)
        dgrad = torch.empty(n, 3, dtype=torch.float32, device=dev) if want_grad else None
        la = _lib.ClidLossArgs()
        la.sdf = res["sdf"].data_ptr()
        la.grad = res["grad"].data_ptr() if want_grad else None
        # conversions are bound to names so their storage outlives the (asynchronous) launches
        label_f = label.contiguous().float()
        w = weight.contiguous().float() if weight is not None else None
        la.label = _lib.ptr(label_f, torch.float32, "sdf_label")
        la.weight = _lib.ptr(w, torch.float32, "weight")
        la.dlogit, la.dgrad, la.loss = dlogit.data_ptr(), (dgrad.data_ptr() if want_grad else None), loss.data_ptr()
        la.n, la.nd = n, nd
        la.n_norm, la.nd_norm = int(n_global), int(nd_global)
        la.sdf_scale = float(dec.sdf_scale)
        la.weight_e = self.weight_e
        la.num_eps = float(cfg.voxel_size_m * cfg.num_grad_step_ratio)
        la.weighted = int(bool(cfg.loss_weight_on))
        stream = _lib.current_stream(dev)
        with torch.cuda.device(dev):
            _lib.check(lib.clid_sdf_loss(C.byref(la), stream), "clid_sdf_loss")

            m, flags = _q.map_struct(npm, True)
            if dec.use_leaky_relu:
                flags |= _lib.LEAKY_RELU
            ds = dec.abi_struct()
            if self.backward_events is not None:
                bv0 = torch.cuda.Event(enable_timing=True)
                bv0.record()
            rc = lib.clid_train_backward(
                C.byref(m), C.byref(ds), x_all.data_ptr(), res["knn_idx"].data_ptr(), dlogit.data_ptr(),
                dgrad.data_ptr() if want_grad else None, n_all, n if want_grad else 0, flags,
                self.feat_grad.data_ptr() if self.train_features else None,
                None if self.touched is None else self.touched.data_ptr(),
                None if self.dec_grad is None else self.dec_grad.data_ptr(), stream)
            _lib.check(rc, "clid_train_backward")
            if self.backward_events is not None:
                bv1 = torch.cuda.Event(enable_timing=True)
                bv1.record()
                self.backward_events.append((bv0, bv1))

        return self._finish_iteration(loss, apply_step, sync, shards)

    def _iteration_single_kernel(self, x, label, ts, weight, n_global, nd_global=0, numerical=False,
                                 defer_reduce=False):
        """clid_train_fused: forward + loss + backward in one launch (analytic or numerical eikonal)."""
        npm, dec, lib, dev = self.npm, self.dec, self.lib, self.device
        n = x.shape[0]
        if getattr(self, "_begin_step", False):
            self._begin_step = False
            loss = torch.empty(3, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                rc = lib.clid_step_begin(self.step_state.data_ptr(), float(self.cfg.lr), 0.9, 0.99, loss.data_ptr(),
                                         _lib.current_stream(dev))
            _lib.check(rc, "clid_step_begin")
            self._advanced = True
        else:
            loss = torch.zeros(3, dtype=torch.float32, device=dev)
        label_f = label.contiguous().float()
        w = weight.contiguous().float() if weight is not None else None
        tsd = _q._prep_ts(ts, n)
        m, flags = _q.map_struct(npm, True, npm.local_point_certainties)
        bricks = npm.brick_index(True) if os.environ.get("CLID_DISABLE_BRICKS", "0") != "1" else None
        if bricks is not None:
            m.bricks = C.pointer(bricks.struct)
            flags |= _q.brick_flags(bricks)
        if dec.use_leaky_relu:
            flags |= _lib.LEAKY_RELU
        ds = dec.abi_struct()
        a = _lib.ClidTrainFusedArgs()
        a.x, a.ts = x.data_ptr(), (None if tsd is None else tsd.data_ptr())
        a.label = _lib.ptr(label_f, torch.float32, "sdf_label")
        a.weight = _lib.ptr(w, torch.float32, "weight")
        a.n, a.n_norm, a.nd_norm = n, int(n_global), int(nd_global)
        a.numerical = int(bool(numerical))
        a.num_eps = float(self.cfg.voxel_size_m * self.cfg.num_grad_step_ratio)
        a.weight_e = self.weight_e
        a.weighted = int(bool(self.cfg.loss_weight_on))
        a.gfeat = self.feat_grad.data_ptr() if self.train_features else None
        a.touched = None if self.touched is None else self.touched.data_ptr()
        a.dec_grad = None if self.dec_grad is None else self.dec_grad.data_ptr()
        a.loss = loss.data_ptr()
        if self.peer is not None:
            self._peer_train_args(a, self._parity)
        if (self.dec_grad is not None and self.use_scratch) or len(dec.layers) == 2:
            need = int(lib.clid_train_fused_scratch_bytes_for(C.byref(ds), n, a.numerical))
            if self._scratch is None or self._scratch.numel() < need:
                self._scratch = torch.empty(need, dtype=torch.uint8, device=dev)
            a.scratch, a.scratch_bytes = self._scratch.data_ptr(), need
        if self.forward_events is not None:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        with torch.cuda.device(dev):
            rc = lib.clid_train_fused(C.byref(m), C.byref(ds), C.byref(a), flags, _lib.current_stream(dev))
        _lib.check(rc, "clid_train_fused")
        if self.forward_events is not None:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            self.forward_events.append((ev0, ev1))
        self.launches += 1 if n > 0 else 0
        if a.scratch and n > 0 and self.dec_grad is not None:
            # the per-point decoder-gradient rows sit in scratch: dense reduction into dec_grad -- now, or
            # on the side stream of the following adam_step()
            self._pending_reduce = (ds, a.scratch, n, a.numerical, flags)
            if not defer_reduce:
                self._reduce_pending()
        return loss

    # ------------------------------------------------------------------ multi-GPU over peer memory
    def attach_peers(self, shards, group=None, peer_rows=None, link=None) -> None:
        """Spatially sharded training without NCCL on the data path (dist.PeerLink): the fused kernel adds
        boundary-band gradients straight into the slab neighbours' tables over NVLink, [decoder gradients | loss]
        are all-reduced by a one-shot peer-memory kernel pair, everything stays inside the step's CUDA graph.
        peer_rows: (lower, upper) row-translation tables of a PARTITIONED map (dist.peer_row_tables); None when every
        rank holds the whole map (rows are numbered alike).  link: an existing dist.PeerLink to reuse (Mapper keeps
        one across mapping() calls; its tables are re-shaped to this trainer's rows)."""
        from .. import dist as _dist

        rank, world = _dist.world()
        if not shards.pairwise:
            raise ValueError("slabs narrower than two bands: a band row would have three contributors")
        n_small = (self.dec_grad.numel() if self.dec_grad is not None else 0) + 3
        if link is not None:
            if link.stride < n_small:
                raise ValueError("the PeerLink's all-reduce slots are too short for this decoder")
            link.view_rows(self.rows)
            self.peer = link
        else:
            self.peer = _dist.PeerLink(self.rows, self.feat_grad.shape[1], n_small, self.device, group=group)
        b = shards.boundaries.tolist()
        reach_band = shards.band
        lo = (b[rank - 1] - reach_band, b[rank - 1] + reach_band - 1) if rank > 0 else None
        hi = (b[rank] - reach_band, b[rank] + reach_band - 1) if rank < world - 1 else None
        self._peer_geom = (int(shards.axis), lo, hi, rank, world)
        self._peer_rows = peer_rows  # kept alive: the kernels read them every step
        if self.touched is not None:  # band rows may receive gradient from the neighbour only
            left, right = shards.neighbour_rows(rank)
            for rows in (left, right):
                if rows is not None and rows.numel():
                    self.touched[rows] = 1
        self.feat_grad = self.peer.grad[0]

    def _peer_train_args(self, a, parity: int) -> None:
        axis, lo, hi, rank, world = self._peer_geom
        a.gfeat = self.peer.grad[parity].data_ptr()
        a.peer_axis = axis
        if lo is not None:
            a.peer_grad[0] = self.peer.grad_ptr(rank - 1, parity)
            a.peer_band[0], a.peer_band[1] = lo
        if hi is not None:
            a.peer_grad[1] = self.peer.grad_ptr(rank + 1, parity)
            a.peer_band[2], a.peer_band[3] = hi
        if self._peer_rows is not None:
            for side, tab in enumerate(self._peer_rows):
                if tab is not None:
                    if tab.numel() != self.rows or tab.dtype != torch.int32:
                        raise ValueError("peer_rows tables must be int32 [rows of the local feature table + padding]")
                    a.peer_row[side] = tab.data_ptr()

    def _peer_all_reduce(self, loss: torch.Tensor) -> None:
        """[decoder gradients | loss] summed over ranks through peer memory (also the barrier after which every
        neighbour's remote gradient adds of this step are visible)."""
        n0 = self.dec_grad.numel() if self.dec_grad is not None else 0
        pa = self.peer.args(n0, 3)
        dg = None if self.dec_grad is None else self.dec_grad.data_ptr()
        stream = _lib.current_stream(self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.clid_peer_publish(C.byref(pa), dg, loss.data_ptr(), stream), "clid_peer_publish")
            _lib.check(self.lib.clid_peer_reduce(C.byref(pa), dg, loss.data_ptr(), stream), "clid_peer_reduce")
        self.launches += 2

    def ensure_step_state(self) -> None:
        """Device-resident optimiser step counter {step, step_size, bc2_sqrt, pad} (ClidAdamArgs.step_state)."""
        if self.step_state is None:
            if self.step != 0:
                raise RuntimeError("the device step counter must be created before the first optimiser step")
            self.step_state = torch.zeros(4, dtype=torch.int32, device=self.device)

    def run_loop(self, pool, iters: int, seed: int, global_coord: bool = True) -> torch.Tensor:
        """`iters` mapping iterations enqueued by ONE native call (clid_mapping_run): replay-pool draw, fused
        forward + loss + backward, decoder-gradient reduction, Adam -- no Python between iterations.
        pool: the Mapper (its replay-pool tensors, new_idx, pool_sample_count and dataset flags are read like
        Mapper.get_batch does).  Returns the loss history [iters, 3] (device tensor)."""
        cfg, npm, dec, lib, dev = self.cfg, self.npm, self.dec, self.lib, self.device
        if iters <= 0:
            return torch.empty(0, 3, dtype=torch.float32, device=dev)
        self.ensure_step_state()
        n = int(cfg.bs)
        numerical = self.numerical and self.weight_e > 0
        if numerical and int(cfg.gradient_decimation) != 10:
            raise NotImplementedError("the native loop runs the one-kernel iteration (gradient_decimation 10)")
        # touched flags only save traffic; a whole-call decision (the loop cannot switch in the middle)
        self._gathers += iters * n * int(cfg.query_nn_k)
        if self.touched is not None and self._gathers > self.dense_switch * self.rows:
            self.touched = None

        ma = _lib.ClidMappingArgs()
        coord = pool.global_coord_pool if global_coord else pool.coord_pool
        ma.pool.coord = _lib.ptr(coord, torch.float32, "coord pool")
        ma.pool.sdf_label = _lib.ptr(pool.sdf_label_pool, torch.float32, "sdf_label_pool")
        ma.pool.weight = _lib.ptr(pool.weight_pool, torch.float32, "weight_pool")
        ma.pool.time = _lib.ptr(pool.time_pool, torch.int32, "time_pool")
        ma.pool.count = int(pool.pool_sample_count)
        new_idx = pool.new_idx
        use_new = (cfg.bs_new_sample > 0 and new_idx is not None and not pool.dataset.lose_track
                   and not pool.dataset.stop_status and new_idx.shape[0] > 0)
        if use_new:
            new_idx = new_idx.to(torch.int64).contiguous()
            ma.pool.new_idx, ma.pool.n_new = new_idx.data_ptr(), int(new_idx.shape[0])
            ma.pool.bs_new = min(int(new_idx.shape[0]), int(cfg.bs_new_sample))

        bufs = self._loop_bufs
        if bufs is None or bufs[0].shape[0] != n:
            f32 = dict(dtype=torch.float32, device=dev)
            bufs = (torch.empty(n, 3, **f32), torch.empty(n, **f32), torch.empty(n, **f32),
                    torch.empty(n, dtype=torch.int32, device=dev), torch.zeros(3, **f32))
            self._loop_bufs = bufs
        x, label, weight, ts, loss = bufs
        history = torch.empty(iters, 3, dtype=torch.float32, device=dev)

        m, flags = _q.map_struct(npm, True, npm.local_point_certainties)
        bricks = npm.brick_index(True) if os.environ.get("CLID_DISABLE_BRICKS", "0") != "1" else None
        if bricks is not None:
            m.bricks = C.pointer(bricks.struct)
            flags |= _q.brick_flags(bricks)
        if dec.use_leaky_relu:
            flags |= _lib.LEAKY_RELU
        ds = dec.abi_struct()
        a = ma.train
        a.x, a.ts, a.label, a.weight = x.data_ptr(), ts.data_ptr(), label.data_ptr(), weight.data_ptr()
        a.n, a.n_norm, a.nd_norm = n, 0, 0
        a.numerical = int(bool(numerical))
        a.num_eps = float(cfg.voxel_size_m * cfg.num_grad_step_ratio)
        a.weight_e = self.weight_e
        a.weighted = int(bool(cfg.loss_weight_on))
        a.gfeat = self.feat_grad.data_ptr() if self.train_features else None
        a.touched = None if self.touched is None else self.touched.data_ptr()
        a.dec_grad = None if self.dec_grad is None else self.dec_grad.data_ptr()
        a.loss = loss.data_ptr()
        if self.dec_grad is not None or len(dec.layers) == 2:
            need = int(lib.clid_train_fused_scratch_bytes_for(C.byref(ds), n, a.numerical))
            if self._scratch is None or self._scratch.numel() < need:
                self._scratch = torch.empty(need, dtype=torch.uint8, device=dev)
            a.scratch, a.scratch_bytes = self._scratch.data_ptr(), need
        ma.adam = self._adam_args(True, True, 0)
        ma.iters, ma.seed, ma.offset = int(iters), int(seed) & (2**64 - 1), int(self.step)
        ma.loss_history = history.data_ptr()
        with torch.cuda.device(dev):
            rc = lib.clid_mapping_run(C.byref(m), C.byref(ds), C.byref(ma), flags, _lib.current_stream(dev))
        _lib.check(rc, "clid_mapping_run")
        self.step += iters
        # draw + fused + Adam advance + Adam [+ decoder-gradient reduction] per iteration, + the last loss copy
        self.launches += iters * (3 + (1 if self.dec_grad is not None else 0)) + 1  # draw (+ counter advance), fused, [reduce], Adam
        return history

    def _want_overlap(self) -> bool:
        """Fork the decoder side of the optimiser step onto a second stream?  It costs two extra host calls,
        so: always inside a CUDA-graph capture (the host cost is paid once), otherwise only when the feature
        table is large enough for its Adam step to hide the decoder-gradient reduction."""
        if not self.overlap_decoder or self.dec_grad is None:
            return False
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
        return bool(torch.cuda.is_current_stream_capturing() or self.rows >= 262144)

    def _reduce_pending(self) -> None:
        if self._pending_reduce is None:
            return
        ds, scratch, n, numerical, flags = self._pending_reduce
        self._pending_reduce = None
        with torch.cuda.device(self.device):
            rc = self.lib.clid_decoder_grad_reduce(C.byref(ds), scratch, n, numerical, flags, self.dec_grad.data_ptr(),
                                                   _lib.current_stream(self.device))
        _lib.check(rc, "clid_decoder_grad_reduce")
        self.launches += 1

    def _finish_iteration(self, loss, apply_step, sync, shards):
        if shards is not None:
            # spatial sharding: 3 kB all-reduce of [decoder grads | loss] + neighbour exchange of the band rows
            # (or ONE flat all-reduce of [decoder grads | loss | shared-row gradients] when bands overlap)
            self.sync_spatial(loss, shards)
        elif sync:
            # replicated sharding: flat all-reduce for [decoder grads | loss scalars]; the replicated
            # feature table needs its whole gradient (and which rows were touched) summed as well
            from .. import dist as _dist

            _dist.FlatAllReduce([self.dec_grad, loss])()
            if self.train_features:
                _dist.all_reduce_sum(self.feat_grad)
                if self.touched is not None:
                    _dist.all_reduce_max(self.touched)
        self.losses.append(loss)
        if apply_step:
            self.adam_step()
        return loss

    def pack_spatial(self, loss: torch.Tensor, shards) -> torch.Tensor:
        """[decoder grads | loss | gradients of the shared feature rows] as one flat buffer."""
        parts = []
        if self.dec_grad is not None:
            parts.append(self.dec_grad)
        parts.append(loss)
        if self.train_features:
            parts.append(self.feat_grad[shards.shared_rows].reshape(-1))
            if self.touched is not None and not getattr(self, "_shared_marked", False):
                # shared rows step on every rank with identical state; untouched ones are no-ops
                self.touched[shards.shared_rows] = 1
                self._shared_marked = True
        return torch.cat(parts)

    def unpack_spatial(self, flat: torch.Tensor, loss: torch.Tensor, shards) -> None:
        off = 0
        if self.dec_grad is not None:
            self.dec_grad.copy_(flat[: self.dec_grad.numel()])
            off = self.dec_grad.numel()
        loss.copy_(flat[off:off + 3])
        off += 3
        if self.train_features:
            self.feat_grad[shards.shared_rows] = flat[off:].view(-1, self.feat_grad.shape[1])

    def neighbour_exchange(self, shards, group=None):
        """NeighbourExchange for this trainer (None: single process, or bands overlap -> flat all-reduce).
        group: process group for the send/recv pairs; a group of its own (dist.new_group()) lets them run
        concurrently with the [decoder grads | loss] all-reduce of the default group."""
        from .. import dist as _dist

        rank, world = _dist.world()
        if world == 1 or not shards.pairwise or not self.train_features:
            return None
        ex = getattr(self, "_exchange", None)
        if ex is None:
            ex = _dist.NeighbourExchange(shards, rank, self.feat_grad, group=group)
            if self.touched is not None:  # band rows may receive gradient from the neighbour only
                self.touched[ex.rows()] = 1
            self._exchange = ex
        return ex

    def small_flat(self, loss: torch.Tensor) -> torch.Tensor:
        """[decoder grads | loss]: what every rank contributes to (all-reduce)."""
        return torch.cat([self.dec_grad, loss]) if self.dec_grad is not None else loss.clone()

    def unpack_small(self, flat: torch.Tensor, loss: torch.Tensor) -> None:
        off = 0
        if self.dec_grad is not None:
            off = self.dec_grad.numel()
            self.dec_grad.copy_(flat[:off])
        loss.copy_(flat[off:off + 3])

    def sync_spatial(self, loss: torch.Tensor, shards) -> None:
        import torch.distributed as tdist

        ex = self.neighbour_exchange(shards)
        if ex is not None:
            # [decoder grads | loss] through a 3 kB all-reduce, band rows with the two slab neighbours
            flat = self.small_flat(loss)
            ex.pack(self.feat_grad)
            work = tdist.all_reduce(flat, op=tdist.ReduceOp.SUM, async_op=True)
            ex.exchange()
            work.wait()
            self.unpack_small(flat, loss)
            ex.unpack(self.feat_grad)
            return
        flat = self.pack_spatial(loss, shards)
        if tdist.is_available() and tdist.is_initialized() and tdist.get_world_size() > 1:
            tdist.all_reduce(flat, op=tdist.ReduceOp.SUM)
        self.unpack_spatial(flat, loss, shards)

    def _adam_args(self, features: bool, decoder: bool, step_arg: int) -> "_lib.ClidAdamArgs":
        cfg, npm = self.cfg, self.npm
        aa = _lib.ClidAdamArgs()
        if features and self.train_features:
            aa.feat = npm.local_geo_features.data.data_ptr()
            aa.feat_grad, aa.feat_m, aa.feat_v = (self.feat_grad.data_ptr(), self.feat_m.data_ptr(),
                                                  self.feat_v.data_ptr())
            aa.touched = None if self.touched is None else self.touched.data_ptr()
            aa.rows = self.rows
        else:
            aa.rows = 0
        aa.dec_tensors = len(self.dec_tensors)
        for i, p in enumerate(self.dec_tensors):
            aa.dec_param[i] = None if p is None else p.data.data_ptr()
            aa.dec_numel[i] = self.dec_numel[i]
        if decoder and self.dec_grad is not None:
            aa.dec_grad, aa.dec_m, aa.dec_v = self.dec_grad.data_ptr(), self.dec_m.data_ptr(), self.dec_v.data_ptr()
        aa.lr, aa.beta1, aa.beta2 = float(cfg.lr), 0.9, 0.99
        aa.eps, aa.weight_decay, aa.step = float(cfg.adam_eps), float(cfg.weight_decay), step_arg
        aa.step_state = None if self.step_state is None else self.step_state.data_ptr()
        return aa

    def _adam_call(self, features: bool, decoder: bool, step_arg: int) -> None:
        dev = self.device
        aa = self._adam_args(features, decoder, step_arg)
        with torch.cuda.device(dev):
            _lib.check(self.lib.clid_adam_step(C.byref(aa), _lib.current_stream(dev)), "clid_adam_step")
        self.launches += 1

    def adam_step(self) -> None:
        dev = self.device
        self.step += 1
        advanced = getattr(self, "_advanced", False)  # iteration() advanced the device counter at its head
        self._advanced = False
        if self._pending_reduce is None or self.dec_grad is None:
            self._reduce_pending()
            self._adam_call(True, True, -1 if advanced else self.step)
            return
        # fork: [decoder-gradient reduction -> Adam on the decoder] beside [Adam on the feature rows]
        main = torch.cuda.current_stream(dev)
        side = self._side_stream
        step_arg = self.step
        if advanced:
            step_arg = -1
        elif self.step_state is not None:
            with torch.cuda.device(dev):
                rc = self.lib.clid_adam_advance(self.step_state.data_ptr(), float(self.cfg.lr), 0.9, 0.99, main.cuda_stream)
            _lib.check(rc, "clid_adam_advance")
            step_arg = -1  # both launches below belong to the step that was just advanced
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            self._reduce_pending()
            self._adam_call(False, True, step_arg)
        self._adam_call(True, False, step_arg)
        join = torch.cuda.Event()
        join.record(side)
        main.wait_event(join)


class StepPipeline:
    """Steady-state driver of a FusedTrainer for fixed-size batches.

    The whole iteration (loss clear, fused forward + loss + backward, decoder-gradient reduction,
    Adam) is captured ONCE in a CUDA graph per input buffer and replayed, so a step costs one graph
    launch on the host instead of ~10 Python/ctypes calls.  Batches that live on the host are staged
    through two device buffers on a copy stream: while step i runs, batch i + 1 is already crossing
    PCIe.  The optimiser's step counter lives on the device (ClidAdamArgs.step_state).

        pipe = StepPipeline(trainer, n)
        pipe.stage(0, host_batch0)                 # pinned (x, label, weight, ts)
        for i in range(steps):
            pipe.stage((i + 1) % 2, host_batch[i + 1])
            loss = pipe.run(i % 2)                 # device tensor [3]; .cpu() to read it
    """

    def __init__(self, trainer: FusedTrainer, n: int, n_global: int = 0, nd_global: int = 0, with_ts: bool = True,
                 buffers=None, shards=None, sync: bool = False, p2p_group=None):
        """buffers: optional list of caller-owned device batches (x [n,3] f32, label [n] f32, weight [n]
        f32, ts [n] i32 | None) to capture on directly (no staging copies); default: two staging buffers."""
        if trainer.step != 0 and trainer.step_state is None:
            raise RuntimeError("StepPipeline must own the optimiser from its first step")
        self.trainer, self.n = trainer, int(n)
        dev = trainer.device
        if shards is None and trainer.peer is None and trainer.touched is not None and 2 * n * int(trainer.cfg.query_nn_k) >= trainer.rows:
            trainer.touched = None  # a replayed graph cannot switch later: large batches run dense from the start
        self.device = dev
        trainer.ensure_step_state()
        # multi-GPU (spatial shards): two graphs per buffer with the ONE NCCL all-reduce of the step launched
        # eagerly between them -- [kernels, pack] | all-reduce(flat) | [unpack, Adam]
        if sync and shards is None:
            raise NotImplementedError("StepPipeline covers single-GPU and spatially sharded steps; the replicated "
                                      "sharding (dense feature-gradient all-reduce) runs call by call")
        self.shards = shards
        f32 = dict(dtype=torch.float32, device=dev)
        if trainer.peer is not None and buffers is not None and len(buffers) % 2 != 0:
            raise ValueError("the peer-memory step alternates two gradient tables: use an even number of input buffers")
        if buffers is not None:
            self.bufs = [tuple(b) for b in buffers]
            for x, label, weight, ts in self.bufs:
                if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape) == (n, 3)
                        and label.dtype == torch.float32 and weight.dtype == torch.float32
                        and (ts is None or ts.dtype == torch.int32)):
                    raise ValueError("buffers must be contiguous device tensors: x [n,3] f32, label/weight [n] f32, ts [n] i32")
        else:
            self.bufs = []
            for _ in range(2):
                self.bufs.append((torch.zeros(n, 3, **f32), torch.zeros(n, **f32), torch.zeros(n, **f32),
                                  torch.zeros(n, dtype=torch.int32, device=dev) if with_ts else None))
        nb = len(self.bufs)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ready = [torch.cuda.Event() for _ in range(nb)]   # buffer k holds a staged batch
        self.free = [torch.cuda.Event() for _ in range(nb)]    # the step that read buffer k has finished
        self.graphs, self.losses = [], []
        trainer.npm.brick_index(True)  # build the index outside the capture
        trainer._want_overlap()        # creates the side stream of the forked optimiser step outside the capture
        # warm-up on a side stream (first launches configure kernel attributes), then capture
        state = self._snapshot()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            warm_loss = self._iteration(0, n_global, nd_global)
            if shards is not None:
                ex = trainer.neighbour_exchange(shards, p2p_group)
                if ex is not None:
                    trainer.unpack_small(trainer.small_flat(warm_loss), warm_loss)
                    ex.pack(trainer.feat_grad)
                    ex.unpack(trainer.feat_grad)
                else:
                    trainer.unpack_spatial(trainer.pack_spatial(warm_loss, shards), warm_loss, shards)
                trainer.adam_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.flats, self.post_graphs = [], []
        self.exchange = trainer.neighbour_exchange(shards, p2p_group) if shards is not None else None
        for k in range(nb):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss = self._iteration(k, n_global, nd_global)
                flat = None
                if shards is not None and self.exchange is not None:
                    flat = trainer.small_flat(loss)
                    self.exchange.pack(trainer.feat_grad)
                elif shards is not None:
                    flat = trainer.pack_spatial(loss, shards)
            self.graphs.append(g)
            self.losses.append(loss)
            self.flats.append(flat)
            if shards is not None:
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2):
                    if self.exchange is not None:
                        trainer.unpack_small(flat, loss)
                        self.exchange.unpack(trainer.feat_grad)
                    else:
                        trainer.unpack_spatial(flat, loss, shards)
                    trainer.adam_step()
                self.post_graphs.append(g2)
        torch.cuda.synchronize(dev)
        self._restore(state)  # the warm-up iteration must not count
        for k in range(nb):
            self.free[k].record(torch.cuda.current_stream(dev))
            self.ready[k].record(torch.cuda.current_stream(dev))  # caller-owned buffers are ready as they are

    def _iteration(self, k, n_global, nd_global):
        x, label, weight, ts = self.bufs[k]
        if self.shards is None:
            # peer-memory multi-GPU step: buffer k always runs with gradient table k % 2 (the buffers rotate in order)
            parity = k % 2 if self.trainer.peer is not None else None
            return self.trainer.iteration(x, label, ts, weight, n_global=n_global, nd_global=nd_global, parity=parity)
        return self.trainer.iteration(x, label, ts, weight, n_global=n_global, nd_global=nd_global, shards=self.shards,
                                      exchange=False)

    def _snapshot(self):
        t = self.trainer
        npm = t.npm
        grads = list(t.peer.grad) if t.peer is not None else [t.feat_grad]
        tensors = [npm.local_geo_features.data, npm.local_point_certainties, npm.local_point_ts_update, t.feat_m,
                   t.feat_v, t.step_state] + grads + [p.data for p in t.dec_tensors if p is not None]
        for extra in (t.touched, t.dec_grad, t.dec_m, t.dec_v):
            if extra is not None:
                tensors.append(extra)
        return [(x, x.clone()) for x in tensors], t.step, len(t.losses), t.launches

    def _restore(self, state):
        pairs, step, n_losses, launches = state
        for live, saved in pairs:
            live.copy_(saved)
        t = self.trainer
        t.step = step
        del t.losses[n_losses:]
        t.launches = launches

    def stage(self, k: int, batch) -> None:
        """Copy a batch (x, label, weight, ts), host-pinned or device, into input buffer k on the copy stream."""
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[k])
            for dst, src in zip(self.bufs[k], batch):
                if dst is not None:
                    dst.copy_(src, non_blocking=True)
            self.ready[k].record(self.copy_stream)

    def run(self, k: int) -> torch.Tensor:
        """Replay the iteration on input buffer k (after its staged batch has arrived)."""
        stream = torch.cuda.current_stream(self.device)
        stream.wait_event(self.ready[k])
        self.graphs[k].replay()
        self.free[k].record(stream)
        if self.shards is not None:
            import torch.distributed as tdist

            if tdist.is_available() and tdist.is_initialized() and tdist.get_world_size() > 1:
                # the 3 kB all-reduce and the neighbour send/recv are independent: with the exchange on a
                # process group of its own they overlap on the wire
                work = tdist.all_reduce(self.flats[k], op=tdist.ReduceOp.SUM, async_op=True)
                if self.exchange is not None:
                    self.exchange.exchange()
                work.wait()
            self.post_graphs[k].replay()
        t = self.trainer
        t.step += 1
        t.launches += self.launches_per_step
        return self.losses[k]

    @property
    def launches_per_step(self) -> int:
        # fused kernel + Adam advance + Adam on the features [+ decoder-gradient reduction + Adam on the decoder]
        # [+ peer publish + peer reduce]
        return 3 + (2 if self.trainer.dec_grad is not None else 0) + (2 if self.trainer.peer is not None else 0)
