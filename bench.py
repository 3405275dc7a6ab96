#!/usr/bin/env python3
"""Benchmark of the neural-SDF hot path (BASELINE.json metric: sampled-points/sec through the SDF
decoder + gradient + loss, with % of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one mapping iteration (utils/mapper.py:642-836 of the reference) on one batch:
fused gather + decoder + d sdf/dx forward, bce + eikonal loss, backward into decoder weights and
neural-point features, Adam.  Workload = BASELINE.json configs[2] (the configuration the
north_star target is quoted on): 131072 samples per batch per GPU, ~1.08 M neural points
(4 wavy sheets, 0.4 m voxels, ncd128 decoder 11->64->1, Kc = 81, K = 6), analytic gradient.
Data are synthetic (no dataset is shipped); see clid_slam_b200/synth.py.

The step runs as a CUDA graph per input buffer (clid_slam_b200/ops/train.py StepPipeline).  Three
loops are timed with CUDA events, the L2 flushed before every step: device-resident inputs (`value`),
host-pinned inputs staged every step with the loss read back every step (`e2e`), and a call-by-call
replay with events around the dominant kernel (`roofline`).

For N > 1 launch with torchrun (one rank per GPU): samples and neural points are partitioned by map slab -- every
rank builds and holds only its slab plus the neighbours' halves of its boundary bands (per-rank memory ~ M/N + halo)
and takes 131072 samples of its slab per step (weak scaling).  Boundary-band feature gradients are added straight
into the slab neighbour's table by the fused kernel over NVLink peer memory (rows translated into its numbering),
[decoder gradients | loss] are all-reduced by a one-shot peer-memory kernel pair, the step is one CUDA graph, no NCCL
on the data path (clid_slam_b200/dist.py).  `--sharding peer` keeps the whole map on every rank, `--sharding spatial`
exchanges through NCCL, `--sharding replicated` all-reduces the dense feature gradient.  `--workload subt` is
BASELINE configs[4] (4 M neural points, SubT-MRS shapes).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 131072
SIDE, SHEETS = 520, 4
SUBT_SIDE = 1000  # 4 sheets x 1000^2 = 4 M neural points (BASELINE configs[4])
L2_FLUSH_BYTES = 256 << 20
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--mode", default="analytic", choices=["analytic", "numerical"],
                    help="eikonal gradient mode of the training step")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the reference legs (CPU and eager CUDA)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: 131072 samples per GPU and step (weak, the default the driver measures) or 131072 samples in "
                         "total, split over the GPUs (strong)")
    ap.add_argument("--no-shipped-config", action="store_true", help="skip the configs[1] leg (process_frame + mapping(10))")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed parity gate against the oracle")
    ap.add_argument("--workload", default="configs2", choices=["configs2", "subt"],
                    help="configs2 (default, the configuration the metric is quoted on): 1.08 M neural points, ncd128 shapes; "
                         "subt: BASELINE configs[4], 4 M neural points, SubT-MRS shapes (LayerNorm on the features)")
    ap.add_argument("--sharding", default="partition", choices=["peer", "partition", "spatial", "replicated"],
                    help="N > 1: slab-partitioned samples and neural points; 'peer': every rank holds the whole map, band gradients added straight "
                         "into the slab neighbours' tables by the fused kernel and [decoder grads | loss] all-reduced by a "
                         "one-shot kernel pair, both over NVLink peer memory, one CUDA graph per step, no NCCL on the data "
                         "path; 'partition' (default): the same step on a PARTITIONED map -- every rank builds and holds only its slab "
                         "plus the halves of its boundary bands (per-rank memory ~ M/N + halo), band rows translated into "
                         "the neighbour's numbering; 'spatial': NCCL all-reduce + neighbour send/recv between two graphs; 'replicated': "
                         "any-sample-anywhere with a dense feature-gradient all-reduce")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ native arm
def build_world(device, mode, workload="configs2", partition=None):
    """The synthetic world of the benchmark.  partition = (rank, world): build only this rank's part of the map (its
    slab plus the neighbours' halves of its boundary bands, cut from the hash-slot owners of the one-point-per-voxel
    cloud, clid_slam_b200/dist.py) -- no rank ever holds the whole neural-point table."""
    import torch

    from clid_slam_b200.config import ncd128
    from clid_slam_b200.model.decoder import Decoder
    from clid_slam_b200.model.neural_points import NeuralPoints
    from clid_slam_b200.synth import set_voxel_features, wavy_sheets

    torch.manual_seed(42)
    cfg = ncd128()
    cfg.device = device
    cfg.feature_std = 0.05          # the default 0.0 makes every feature exactly zero (SURVEY 8c gotchas)
    cfg.local_map_radius = 1.0e4    # the whole map is the local window (configs[2])
    cfg.numerical_grad = mode == "numerical"
    cfg.gradient_decimation = 10 if cfg.numerical_grad else 1
    side = SIDE
    if workload == "subt":          # config/run_SubT_MRS.yaml: ncd128 shapes + layer_norm_on, 4 M neural points
        cfg.layer_norm_on = True
        cfg.free_sample_begin_ratio = 0.8
        side = SUBT_SIDE
    dec = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    npm = NeuralPoints(cfg)
    npm.travel_dist = torch.zeros(1, device=device)
    gen = torch.Generator(device=device).manual_seed(1)
    pts = wavy_sheets(side, SHEETS, cfg.voxel_size_m, gen, device=device)
    geom = None
    if partition is not None:
        from clid_slam_b200.dist import axis_cells, hash_owner_mask, partition_mask, slab_boundaries
        from clid_slam_b200.utils.tools import voxel_down_sample_torch

        rank, world = partition
        pts = pts[voxel_down_sample_torch(pts, cfg.voxel_size_m)]
        pts = pts[hash_owner_mask(npm._slots_of(pts), npm.buffer_size)]
        axis = int(torch.argmax(pts.amax(0) - pts.amin(0)).item())
        bnd = slab_boundaries(axis_cells(pts, cfg.voxel_size_m, axis), world)
        geom = (axis, bnd, int(pts.shape[0]))
        pts = pts[partition_mask(pts, cfg.voxel_size_m, axis, bnd, rank, cfg.num_nei_cells + 1)].contiguous()
    npm.update(pts, torch.zeros(3, device=device), torch.eye(3, device=device), 0)
    if partition is not None:
        set_voxel_features(npm, cfg.feature_std)  # a function of the voxel: every rank that holds it starts alike
    return cfg, dec, npm, geom


def _xlwt(batch):
    x, label, weight, ts = batch
    return x, label, weight, ts


def run_native(args):
    global BATCH
    import torch
    import torch.distributed as dist

    from clid_slam_b200 import _lib, fused
    from clid_slam_b200 import build as _build
    from clid_slam_b200.ops.train import FusedTrainer
    from clid_slam_b200.synth import sample_batch

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    if args.scaling == "strong" and world > 1:
        BATCH = BATCH // world  # the global batch stays at 131072 samples
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    if world > 1:
        # library banners (e.g. "NCCL version ...") must not land on stdout: rank 0 prints ONE json line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(device))
            warm = torch.zeros(1, device=device)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    if rank == 0:
        _build.build()
    if world > 1:
        dist.barrier()
    _lib.load()

    partitioned = world > 1 and args.sharding == "partition"
    cfg, dec, npm, geom = build_world(device, args.mode, args.workload, (rank, world) if partitioned else None)
    gen = torch.Generator(device=device).manual_seed(1000 + rank)
    n_batches = 4  # rotate a few batches so no step sees the previous step's exact access pattern
    shards = None
    if world > 1 and args.sharding in ("spatial", "peer", "partition"):
        # slab partition along the longest map axis; a rank's samples are those whose voxel lies in
        # its slab (what a per-rank replay pool would hand out), see clid_slam_b200/dist.py
        from clid_slam_b200.dist import SpatialShards

        if partitioned:
            shards = SpatialShards(npm.local_neural_points, cfg.voxel_size_m, reach=cfg.num_nei_cells, world_size=world,
                                   axis=geom[0], boundaries=geom[1])
        else:
            shards = SpatialShards(npm.local_neural_points, cfg.voxel_size_m, reach=cfg.num_nei_cells, world_size=world)
        own_points = npm.neural_points[shards.row_owner[:-1] == rank]
        batches = []
        for _ in range(n_batches):
            cand = sample_batch(own_points, int(BATCH * 1.25), gen)
            keep = torch.nonzero(shards.owner_of(cand[0]) == rank).flatten()[:BATCH]
            assert keep.numel() == BATCH, "not enough samples inside the slab"
            batches.append(tuple(t[keep].contiguous() for t in cand))
    else:
        batches = [sample_batch(npm.neural_points, BATCH, gen) for _ in range(n_batches)]
    host_batches = [tuple(t.cpu().pin_memory() for t in b) for b in batches]
    flush_w = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=device)
    flush_r = torch.zeros(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=device)
    flush_sink = torch.zeros((), dtype=torch.float32, device=device)

    class _Flush:
        """Evict the step's working set from the 126 MB L2: write a 256 MiB buffer, then read another
        256 MiB one, so the lines the step evicts are clean (a write-only flush leaves the L2 full of
        dirty lines whose write-back is charged to whatever runs next)."""

        @staticmethod
        def zero_():
            flush_w.zero_()
            flush_sink.copy_(flush_r.sum())

    flush = _Flush()

    # ---- untimed parity gate at the benchmark configuration: forward + gradient and one training
    # iteration of the CUDA path against the CPU oracle on a 16384-sample slice of batch 0, on this very world
    parity = None
    if world == 1 and not args.no_parity:
        from oracle.bridge import parity_gate  # checker only: never timed, never on the product path

        parity = parity_gate(npm, dec, cfg, *_xlwt(batches[0]))

    trainer = FusedTrainer(cfg, npm, dec)
    peer_mode = world > 1 and args.sharding in ("peer", "partition")
    if partitioned:
        from clid_slam_b200.dist import peer_row_tables

        trainer.attach_peers(shards, peer_rows=peer_row_tables(shards, rank, npm.local_neural_points, trainer.rows))
    elif peer_mode:
        trainer.attach_peers(shards)
    n_global = BATCH * world
    nd_global = 0
    if cfg.numerical_grad:
        nd_global = world * ((BATCH + cfg.gradient_decimation - 1) // cfg.gradient_decimation)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    multi = world > 1
    graphed = os.environ.get("CLID_BENCH_NO_GRAPH", "0") != "1"
    if graphed:
        # the iteration (and, for N > 1, the NCCL all-reduce inside it) is one CUDA graph per input
        # buffer (ops/train.py StepPipeline)
        from clid_slam_b200.ops.train import StepPipeline

        # neighbour send/recv on a process group of its own: overlaps the [decoder grads | loss] all-reduce
        p2p_group = dist.new_group() if multi and shards is not None and not peer_mode else None
        kw = dict(n_global=n_global, nd_global=nd_global, shards=None if peer_mode else shards,
                  sync=multi and shards is None, p2p_group=p2p_group)
        pipe_dev = StepPipeline(trainer, BATCH, buffers=batches, **kw)
        pipe_host = StepPipeline(trainer, BATCH, **kw)

        def step(i):
            return pipe_dev.run(i % n_batches)
    else:
        def step(i):
            x, label, weight, ts = batches[i % n_batches]
            return trainer.iteration(x, label, ts, weight, n_global=n_global, nd_global=nd_global,
                                     sync=multi and shards is None, shards=None if peer_mode else shards)

    for i in range(max(args.warmup, 3)):
        flush.zero_()
        step(i)
    sync_all()

    # ---- timed region: K steps, device-resident inputs, L2 flushed (untimed) before every step
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = trainer.launches
    events = []
    sync_all()
    torch.cuda.nvtx.range_push("timed")
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(i)
        e1.record()
        events.append((e0, e1))
    sync_all()
    wall_dev_ms = (time.perf_counter() - wall0) * 1e3
    torch.cuda.nvtx.range_pop()
    total_ms = sum(a.elapsed_time(b) for a, b in events)
    launches = trainer.launches - launches0

    # ---- end to end: host (pinned) batch -> device, step, loss back to the host, every step.
    # Graphed: a software pipeline of depth two -- batch i + 1 crosses PCIe on the copy stream while step i
    # runs (StepPipeline), and the host reads the loss of step i - 1 (pinned, copied right behind its graph)
    # after it has enqueued step i, so the device never waits for the host.  Every step's inputs are copied
    # and every step's loss is read inside the timed region.
    e2e_events = []
    loss_pinned = [torch.zeros(3, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]
    losses_read = 0
    loss_host = None
    sync_all()
    wall1 = time.perf_counter()
    if graphed:
        pipe_host.stage(0, host_batches[0])
    for i in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if graphed:
            pipe_host.stage((i + 1) % 2, host_batches[(i + 1) % n_batches])
            loss = pipe_host.run(i % 2)
            if multi and not peer_mode:
                # N > 1 through NCCL: the host side of a step (two graph launches + NCCL enqueues from Python) is about as
                # long as the step itself, so the loss is simply read back here (blocking)
                loss_host = loss.cpu()
                losses_read += 1
                e1.record()
            else:
                e1.record()
                # device -> host read of this step's result, on the copy stream behind the step
                pipe_host.copy_stream.wait_event(e1)
                with torch.cuda.stream(pipe_host.copy_stream):
                    loss_pinned[i % 2].copy_(loss, non_blocking=True)
                    loss_ready[i % 2].record(pipe_host.copy_stream)
                if i >= 1:
                    loss_ready[(i - 1) % 2].synchronize()
                    loss_host = loss_pinned[(i - 1) % 2].clone()
                    losses_read += 1
        else:
            x, label, weight, ts = (t.to(device, non_blocking=True) for t in host_batches[i % n_batches])
            loss = trainer.iteration(x, label, ts, weight, n_global=n_global, nd_global=nd_global,
                                     sync=shards is None, shards=None if peer_mode else shards)
            loss_host = loss.cpu()  # device -> host read of the step's result (synchronises)
            losses_read += 1
            e1.record()
        e2e_events.append((e0, e1))
    if graphed and (peer_mode or not multi):
        loss_ready[(args.steps - 1) % 2].synchronize()
        loss_host = loss_pinned[(args.steps - 1) % 2].clone()
        losses_read += 1
    assert losses_read == args.steps
    torch.cuda.synchronize()
    wall_e2e_ms = (time.perf_counter() - wall1) * 1e3
    sync_all()
    e2e_events_ms = sum(a.elapsed_time(b) for a, b in e2e_events)
    clocks = sampler.stop() if rank == 0 else None
    # the end-to-end time is the host wall clock of that loop (copies, launches, loss reads, everything) minus the
    # untimed L2 flushes, whose wall time is measured by running the same number of flushes alone
    sync_all()
    wall2 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()
    torch.cuda.synchronize()
    wall_flush_ms = (time.perf_counter() - wall2) * 1e3
    e2e_ms = max(wall_e2e_ms - wall_flush_ms, e2e_events_ms)

    # ---- kernel-only timing for the roofline: the same step launched call by call (no graph) with
    # CUDA events around the dominant kernel, same L2 hygiene
    trainer.forward_events, trainer.backward_events = [], []
    torch.cuda._sleep(int(2e7))  # ~10 ms spin kernel: the host queues ahead, so the events see device time, not launch gaps
    for i in range(min(args.steps, 50)):
        flush.zero_()
        x, label, weight, ts = batches[i % n_batches]
        trainer.iteration(x, label, ts, weight, n_global=n_global, nd_global=nd_global,
                          sync=multi and shards is None, shards=None if peer_mode else shards)
    sync_all()
    if peer_mode:
        trainer.peer.check()
    fwd_ms = [a.elapsed_time(b) for a, b in trainer.forward_events]
    bwd_ms = [a.elapsed_time(b) for a, b in trainer.backward_events]
    trainer.forward_events = trainer.backward_events = None

    # ---- inference-only forward (what the 40 % roofline target is stated on), same L2 hygiene
    inf_events = []
    x0 = batches[0][0]
    for i in range(3):
        fused.sdf_and_gradient(npm, dec, x0)
    torch.cuda._sleep(int(2e7))
    for i in range(50):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, _, nn_count, _ = fused.sdf_and_gradient(npm, dec, batches[i % n_batches][0])
        e1.record()
        inf_events.append((e0, e1))
    torch.cuda.synchronize()
    # the event clock of this part ticks every ~1.9 us: a median sits on that grid, the mean (the estimator of
    # kernel_ms_avg above) resolves finer -- both are reported, the fraction uses the mean
    inf_all = [a.elapsed_time(b) for a, b in inf_events]
    inf_ms, inf_median_ms = statistics.mean(inf_all), statistics.median(inf_all)
    mean_nn = float(nn_count.float().mean().item())

    # the same kernel at the reference's own inference chunk (config.infer_bs = 1 048 576 queries: what Mesher.query_points
    # and the tracker feed it): 14 tile rounds instead of 1.7, i.e. its steady-state rate.  Not the headline.
    inf_big = None
    if not multi:
        big_n = 1 << 20
        xb = torch.cat([b[0] for b in batches] * ((big_n + BATCH * n_batches - 1) // (BATCH * n_batches)))[:big_n].contiguous()
        fused.sdf_and_gradient(npm, dec, xb)
        big_events = []
        for i in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _, _, nn_big, _ = fused.sdf_and_gradient(npm, dec, xb)
            e1.record()
            big_events.append((e0, e1))
        torch.cuda.synchronize()
        inf_big = (big_n, statistics.mean(a.elapsed_time(b) for a, b in big_events), float(nn_big.float().mean().item()))
        del xb

    # ---- the shipped default of every run file is the NUMERICAL eikonal gradient (config.py:204-206): second value,
    # same map, same batches, its own trainer and graphs (N = 1 only; `--mode numerical` makes it the headline)
    numerical_line = None
    if world == 1 and graphed and args.mode == "analytic":
        import copy

        cfg_n = copy.copy(cfg)
        cfg_n.numerical_grad, cfg_n.gradient_decimation = True, 10
        trainer_n = FusedTrainer(cfg_n, npm, dec)
        nd_n = (BATCH + 9) // 10
        pipe_n = StepPipeline(trainer_n, BATCH, buffers=batches, n_global=BATCH, nd_global=nd_n)
        for i in range(3):
            flush.zero_()
            pipe_n.run(i % n_batches)
        ev = []
        for i in range(min(args.steps, 30)):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pipe_n.run(i % n_batches)
            e1.record()
            ev.append((e0, e1))
        torch.cuda.synchronize()
        ms_n = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
        numerical_line = {"value": BATCH / (ms_n * 1e-3), "unit": "samples/s", "ms_per_step": ms_n, "steps": len(ev),
                          "evaluations_per_step": BATCH + 6 * nd_n,
                          "note": "same step with get_numerical_gradient (6 shifted evaluations of every 10th sample)"}
        del pipe_n, trainer_n

    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])
    map_rows_per_rank = [int(npm.count())]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, int(npm.count()))
        map_rows_per_rank = gathered
    map_rows_total = geom[2] if partitioned else int(npm.count())

    if rank == 0:
        peak, peak_src = measured_peak()
        samples_per_step = BATCH * world
        value = samples_per_step * args.steps / (total_ms * 1e-3)
        e2e_value = samples_per_step * args.steps / (e2e_ms * 1e-3)
        # algorithmic bytes of the fused forward kernel per sample (SURVEY.md 8d / DESIGN.md section 5)
        b_fwd = 576.0 + 16.0 * mean_nn
        b_bwd = 540.0  # labels + feature-grad RMW + certainty/ts RMW + saved neighbour rows
        evals = BATCH + (6 * ((BATCH + 9) // 10) if cfg.numerical_grad else 0)
        # 10 %-trimmed mean of the per-launch event times: one host hiccup (a launch that reaches the queue late shows up
        # as a 100+ us "kernel") must not move the roofline figure; the plain mean and the spread are reported beside it
        fwd_sorted = sorted(fwd_ms)
        trim = len(fwd_sorted) // 10
        fwd_core = fwd_sorted[trim:len(fwd_sorted) - trim] if len(fwd_sorted) > 2 * trim else fwd_sorted
        fwd_avg_ms = statistics.mean(fwd_core)
        one_kernel = not bwd_ms  # analytic mode: forward + loss + backward are one launch (clid_train_fused)
        b_kernel = b_fwd + b_bwd if one_kernel else b_fwd
        kernel_name = ("train_fused_l1_kernel<64,6,bricks,rows> (forward + loss + backward of the step; its per-point "
                       "decoder-gradient rows are reduced by decoder_grad_kernel afterwards)" if one_kernel
                       else "query_forward_kernel<64,1,6,bricks> (training forward)")
        achieved = b_kernel * evals / (fwd_avg_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                tj = json.load(fh)
                traffic = tj.get("train_fused_l1_kernel_dram_bytes_per_launch" if one_kernel
                                 else "query_forward_kernel_dram_bytes_per_launch")
        h2d = sum(tt.numel() * tt.element_size() for tt in host_batches[0])
        line = {
            "metric": "sampled-points/sec through SDF decoder+grad+loss (train step: fused gather+MLP+grad forward, "
                      "bce+eikonal loss, backward, Adam)",
            "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": ("BASELINE configs[2]: 131072 samples/batch/GPU, 1.08M neural points (4 wavy sheets, "
                             "0.4 m voxels), ncd128 decoder 11->64->1, Kc=81, K=6, brick index, "
                             f"{args.mode} eikonal gradient" if args.workload == "configs2" else
                             f"BASELINE configs[4]: SubT-MRS shapes (ncd128 decoder 11->64->1, LayerNorm on the features, "
                             f"0.4 m voxels, Kc=81, K=6), {map_rows_total} neural points in total (4 wavy sheets), "
                             f"{BATCH} samples/batch/GPU, brick index, {args.mode} eikonal gradient"),
                "samples_per_step": samples_per_step, "neural_points": map_rows_total,
                "neural_points_per_rank": map_rows_per_rank,
                "mean_valid_candidates": mean_nn, "l2": "flushed before every timed step (256 MiB write, then 256 MiB read so evictions are clean)",
                "parallelism": ("single GPU" if world == 1 else
                                f"x{world}: PARTITIONED map -- every rank builds and holds only its slab plus the neighbours' "
                                f"halves of its boundary bands ({map_rows_per_rank} rows per rank of {map_rows_total} in "
                                f"total), samples belong to the rank of their slab; boundary-band feature gradients are "
                                f"added into the slab neighbour's table (rows translated into its numbering) by the fused "
                                f"kernel over NVLink peer memory, [decoder grads | loss] all-reduced by a one-shot "
                                f"peer-memory kernel pair, one CUDA graph per step, no NCCL on the data path"
                                if partitioned else
                                f"x{world}: samples and neural points sharded by map slab; boundary-band feature gradients "
                                f"({int(shards.shared_rows.numel())} band rows in total) are added into the slab neighbour's "
                                f"table by the fused kernel over NVLink peer memory, [decoder grads | loss] all-reduced by a "
                                f"one-shot peer-memory kernel pair, one CUDA graph per step, no NCCL on the data path"
                                if peer_mode else
                                f"x{world}: samples and neural points sharded by map slab; per step one 3 kB NCCL "
                                f"all-reduce [decoder grads | loss] and a neighbour send/recv of the boundary-band "
                                f"feature gradients ({int(shards.shared_rows.numel())} band rows in total)"
                                if shards is not None else
                                f"x{world}: batch-sharded, replicated map, dense feature-gradient all-reduce"),
            },
            "roofline": {
                "kernel": kernel_name,
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_sample": b_kernel, "samples_per_launch": evals,
                "kernel_ms_avg": fwd_avg_ms, "kernel_ms_plain_mean": statistics.mean(fwd_ms),
                "kernel_ms_min_max": [min(fwd_ms), max(fwd_ms)], "kernel_launches_timed": len(fwd_ms),
                "timing": "CUDA events around the kernel in a call-by-call replay of the step (the timed region itself "
                          "runs as one CUDA graph per step), same inputs and L2 flush; kernel_ms_avg is the 10 %-trimmed "
                          "mean of the launches, the plain mean and min / max are given beside it",
                "backward_kernel_ms_avg": statistics.mean(bwd_ms) if bwd_ms else None,
                "inference_forward": {
                    "kernel": "query_forward_kernel<64,1,6,bricks> (sdf + grad, no side effects; the kernel the "
                              "north_star's 40 % target is stated on)",
                    "algorithmic_bytes_per_sample": b_fwd, "ms_avg": inf_ms, "ms_median": inf_median_ms, "launches": len(inf_all),
                    "samples_per_s": BATCH / (inf_ms * 1e-3),
                    "achieved_gbs": b_fwd * BATCH / (inf_ms * 1e-3) / 1e9,
                    "frac": b_fwd * BATCH / (inf_ms * 1e-3) / 1e9 / peak,
                    "at_infer_bs": None if inf_big is None else {
                        "queries": inf_big[0], "ms_avg": inf_big[1], "samples_per_s": inf_big[0] / (inf_big[1] * 1e-3),
                        "algorithmic_bytes_per_sample": 576 + 16 * inf_big[2],
                        "achieved_gbs": (576 + 16 * inf_big[2]) * inf_big[0] / (inf_big[1] * 1e-3) / 1e9,
                        "frac": (576 + 16 * inf_big[2]) * inf_big[0] / (inf_big[1] * 1e-3) / 1e9 / peak,
                        "note": "same kernel on config.infer_bs queries (the reference's inference chunk): 14 tile rounds "
                                "instead of 1.7; at 131072 queries the kernel is two rounds of one tile's latency",
                    },
                },
            },
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
                    "ms_per_step": e2e_ms / args.steps,
                    # host wall clock of the two loops INCLUDING the untimed L2 flush kernels between steps
                    # (identical in both): their difference is what the host traffic costs end to end
                    "wall_ms_per_step_incl_flush": wall_e2e_ms / args.steps,
                    "wall_ms_per_step_incl_flush_device_resident_loop": wall_dev_ms / args.steps,
                    "pipeline": ("depth 2: batch i+1 staged host->device on a copy stream during step i; the loss of "
                                 "step i-1 is read on the host after step i has been enqueued") if (graphed and (peer_mode or not multi))
                    else "batch i+1 staged host->device on a copy stream during step i; blocking loss read per step"},
            "gpu_launches": launches,
            "clocks": clocks,
            "final_loss": [float(v) for v in loss_host.tolist()],
        }
        if parity is not None:
            line["parity"] = parity
        if numerical_line is not None:
            line["numerical_mode"] = numerical_line
        if world == 1 and not args.no_shipped_config:
            # the configuration the reference ships (configs[1]), both gradient modes; untimed w.r.t. `value`
            line["shipped_config"] = {mode: shipped_config(device, mode, with_reference=not args.no_cpu_baseline)
                                      for mode in ("numerical", "analytic")}
        if world == 1 and not args.no_cpu_baseline:
            # free the benchmark's device memory before the reference builds its own world on the same GPU
            line["cpu_baseline"] = cpu_reference(args.mode, steps=3, warmup=1, n=BATCH)
            line["eager_cuda_baseline"] = eager_cuda_reference(args.mode, device, steps=5, warmup=2, n=BATCH)
            if line["eager_cuda_baseline"].get("value"):
                line["speedup_over_eager_cuda"] = {"step": value / line["eager_cuda_baseline"]["value"],
                                                   "inference_forward": (BATCH / (inf_ms * 1e-3)) /
                                                   line["eager_cuda_baseline"]["inference_forward"]["value"]}
        print(json.dumps(line), flush=True)
        if parity is not None and not parity["ok"]:
            raise SystemExit("parity gate FAILED: " + json.dumps(parity))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ shipped configuration
SHIP_SIDE, SHIP_SHEETS, SHIP_RADIUS, SHIP_SCAN = 600, 1, 72.0, 30000


def shipped_config(device, mode, frames=4, with_reference=True):
    """BASELINE configs[1]: the configuration the reference ships (config/run_ncd128.yaml: 16384-sample batches,
    10 iterations per frame), driven the way slam.py:135-208 drives it -- per frame Mapper.process_frame (sampler,
    map insert, replay pool, local window; the brick index is rebuilt lazily by the first query) then
    Mapper.mapping(10) with a fresh trainer.  ~100 k local neural points (one wavy sheet, local_map_radius 72 m),
    30 k-point synthetic scans.  Times are host wall clock with a device synchronise on both sides."""
    import numpy as np
    import torch

    from clid_slam_b200.config import ncd128
    from clid_slam_b200.model.decoder import Decoder
    from clid_slam_b200.model.local_point_cloud_map import LocalPointCloudMap
    from clid_slam_b200.model.neural_points import NeuralPoints
    from clid_slam_b200.synth import wavy_sheets
    from clid_slam_b200.utils.mapper import Mapper

    class _DS:
        lose_track = False
        stop_status = False
        processed_frame = 0
        gt_pose_provided = True
        pgo_poses = None
        static_mask = None

    torch.manual_seed(42)
    cfg = ncd128()
    cfg.device = device
    cfg.feature_std = 0.05
    cfg.local_map_radius = SHIP_RADIUS
    cfg.numerical_grad = mode == "numerical"
    cfg.gradient_decimation = 10 if cfg.numerical_grad else 1
    dec = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    npm = NeuralPoints(cfg)
    gen = torch.Generator(device=device).manual_seed(1)
    world = wavy_sheets(SHIP_SIDE, SHIP_SHEETS, cfg.voxel_size_m, gen, device=device)
    n_total = frames + 1
    ds = _DS()
    ds.gt_poses = ds.odom_poses = np.tile(np.eye(4), (n_total, 1, 1))
    npm.travel_dist = torch.zeros(1, device=device)
    npm.update(world, torch.zeros(3, device=device), torch.eye(3, device=device), 0)
    mapper = Mapper(cfg, ds, npm, LocalPointCloudMap(cfg), dec)
    rows = {"process_frame_ms": [], "mapping_ms": [], "mapping_enqueue_ms": [], "local_points": [], "setup_ms": [], "loop_enqueue_ms": []}
    for frame in range(n_total):
        ds.processed_frame = frame
        pose = torch.eye(4, device=device, dtype=torch.float64)
        pose[0, 3], pose[2, 3] = 0.5 * frame, 2.0
        ds.gt_poses[frame] = pose.cpu().numpy()
        npm.travel_dist = torch.arange(frame + 1, device=device, dtype=torch.float32) * 0.5
        # a scan: 30 k world points within range of the sensor, slightly noisy, in the sensor frame
        origin = pose[:3, 3].float()
        near = world[(world - origin).norm(dim=1) < cfg.max_range]
        pick = torch.randint(0, near.shape[0], (SHIP_SCAN,), generator=gen, device=device)
        scan = near[pick] + 0.01 * torch.randn(SHIP_SCAN, 3, generator=gen, device=device) - origin
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mapper.process_frame(scan, None, pose, frame)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        mapper.mapping(cfg.iters)
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        if frame >= 1:  # frame 0 pays one-off costs (kernel attributes, first allocations)
            rows["process_frame_ms"].append((t1 - t0) * 1e3)
            rows["mapping_enqueue_ms"].append((t2 - t1) * 1e3)
            rows["mapping_ms"].append((t3 - t1) * 1e3)
            rows["local_points"].append(int(npm.local_count()))
            if mapper.last_host_ms is not None:
                rows["setup_ms"].append(mapper.last_host_ms["setup"])
                rows["loop_enqueue_ms"].append(mapper.last_host_ms["loop_enqueue"])
    iters = max(1, cfg.iters + mapper.adaptive_iter_offset)
    med = {k: (statistics.median(v) if v else None) for k, v in rows.items()}
    out = {"workload": f"BASELINE configs[1]: run_ncd128 shapes, batch {cfg.bs}, {iters} iterations/frame ({mode} gradient), "
                       f"{int(med['local_points'])} local neural points, {SHIP_SCAN}-point scans; per frame process_frame + mapping",
           "frames_timed": frames, "iterations_per_frame": iters,
           "process_frame_ms": med["process_frame_ms"], "mapping_ms": med["mapping_ms"],
           "mapping_host_enqueue_ms": med["mapping_enqueue_ms"],
           "mapping_us_per_iteration": med["mapping_ms"] * 1e3 / iters,
           # inside mapping(): index rebuild (one 28-byte read-back) + trainer set-up, then ONE native call enqueues all iterations
           "mapping_setup_ms": med["setup_ms"],
           "loop_host_us_per_iteration": None if med["loop_enqueue_ms"] is None else med["loop_enqueue_ms"] * 1e3 / iters,
           "samples_per_s_in_mapping": iters * cfg.bs / (med["mapping_ms"] * 1e-3),
           "final_loss": [float(v) for v in mapper.last_losses[-1].tolist()]}
    if with_reference:
        try:
            from baseline import ref_runner

            if ref_runner.available():
                del mapper, npm
                torch.cuda.empty_cache()
                r = ref_runner.run(device, mode, cfg.bs, SHIP_SIDE, SHIP_SHEETS, steps=10, warmup=10, inference_passes=2,
                                   local_map_radius=SHIP_RADIUS)
                out["reference_eager_cuda_us_per_iteration"] = r["ms_per_step"] * 1e3
                out["reference_local_points"] = r["local_points"]
                out["speedup_over_eager_cuda"] = r["ms_per_step"] * 1e3 / out["mapping_us_per_iteration"]
        except Exception as exc:
            out["reference_eager_cuda_us_per_iteration"] = None
            out["reference_error"] = f"{type(exc).__name__}: {exc}"[:200]
    return out


# ------------------------------------------------------------------------------------------ reference arm
def cpu_reference(mode, steps, warmup, n):
    """The reference's own implementation of the step on the host cores: the UNMODIFIED reference shipped under
    baseline/_ref (baseline/ref_runner.py: NeuralPoints.update, Mapper.mapping with its stock get_batch) when it is
    there, else the oracle port (torch CPU restatement, oracle/sdf_oracle.py)."""
    import torch

    threads = os.cpu_count() or 1
    from baseline import ref_runner

    if ref_runner.available():
        r = ref_runner.run("cpu", mode, n, SIDE, SHEETS, steps=steps, warmup=warmup, threads=threads)
        return {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["threads"], "kind": "reference",
                "sample": f"Mapper.mapping({steps}) of the unmodified reference (baseline/_ref) with {n}-sample batches drawn by its "
                          f"stock get_batch from a {4 * n}-sample pool, 1.08M-point world ({mode} gradient), after mapping({warmup})",
                "ms_per_step": r["ms_per_step"],
                "inference_forward": {"value": r["inference_samples_per_s"], "unit": "samples/s",
                                      "sample": f"query_feature + Decoder.sdf + get_gradient on {n} queries, mean of 2 after 1 warm-up"}}
    from oracle import sdf_oracle as oc

    torch.set_num_threads(threads)
    cfg = oc.OracleConfig(local_map_radius=1.0e4, numerical_grad=(mode == "numerical"),
                          gradient_decimation=10 if mode == "numerical" else 1)
    gen = torch.Generator().manual_seed(1)
    m = oc.empty_map(cfg)
    pts = oc.wavy_sheets(SIDE, SHEETS, cfg.voxel_size_m, gen)
    oc.map_insert(m, pts, torch.zeros(3), 0, generator=gen)
    params = oc.init_decoder(cfg, gen)
    opt = oc.make_adam(cfg, [m.local_features], params)
    batches = [oc.sample_batch(m.points, n, gen) for _ in range(2)]
    times = []
    for i in range(warmup + steps):
        x, label, weight, ts = batches[i % 2]
        t0 = time.perf_counter()
        oc.train_iteration(m, params, opt, x.clone(), label, ts, weight)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    fwd_times = []
    for i in range(3):
        x = batches[i % 2][0].clone().requires_grad_(True)
        t0 = time.perf_counter()
        z, _, _, _ = oc.query_feature(m, x, None, training_mode=False, query_locally=True)
        oc.sdf_gradient(x, oc.decoder_sdf(params, z, cfg.sdf_scale))
        if i >= 1:
            fwd_times.append(time.perf_counter() - t0)
    return {"value": n * steps / total, "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": f"{steps} steps of {n} samples on the same 1.08M-point world ({mode} gradient), after {warmup} warm-up",
            "ms_per_step": total / steps * 1e3,
            "inference_forward": {"value": n * len(fwd_times) / sum(fwd_times), "unit": "samples/s",
                                  "sample": f"{len(fwd_times)} forward + gradient passes of {n} queries, after 1 warm-up"}}


def eager_cuda_reference(mode, device, steps, warmup, n):
    """The unmodified reference on the SAME B200 with device="cuda": its eager PyTorch path (what the reference
    actually deploys, README.md:33), the meaningful baseline for the speed-up."""
    import torch

    from baseline import ref_runner

    if not ref_runner.available():
        return {"unavailable": "baseline/_ref is missing (python -m baseline.install_ref in the build container)"}
    try:
        torch.cuda.empty_cache()
        r = ref_runner.run(device, mode, n, SIDE, SHEETS, steps=steps, warmup=warmup)
    except Exception as exc:  # the baseline must never take the benchmark line down
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
    return {"value": r["samples_per_s"], "unit": "samples/s", "kind": "reference-eager-cuda", "ms_per_step": r["ms_per_step"],
            "sample": f"Mapper.mapping({steps}) of the unmodified reference with device={device}, {n}-sample batches, after mapping({warmup})",
            "inference_forward": {"value": r["inference_samples_per_s"], "unit": "samples/s", "ms": r["inference_ms"]}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warmup = max(1, min(args.warmup, 2))
    base = cpu_reference(args.mode, steps=steps, warmup=warmup, n=BATCH)
    kind = "the unmodified reference (baseline/_ref)" if base["kind"] == "reference" else "reference algorithm (oracle port, torch CPU)"
    line = {
        "impl": "reference",
        "metric": "sampled-points/sec through SDF decoder+grad+loss (train step: fused gather+MLP+grad forward, "
                  "bce+eikonal loss, backward, Adam)",
        "value": base["value"], "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[2]: 131072 samples/batch, 1.08M neural points, ncd128 decoder, "
                               f"{args.mode} eikonal gradient; {kind} on host cores"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
