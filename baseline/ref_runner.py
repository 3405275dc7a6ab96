"""Run the UNMODIFIED reference (baseline/_ref) on the benchmark workload through its own public API:
`NeuralPoints.update` builds the map, `Mapper.mapping(k)` (utils/mapper.py:620-862) runs k training
iterations drawing 131072-sample batches from its replay pool with its stock `get_batch`, and the
inference variant is `query_feature` -> `Decoder.sdf` -> `get_gradient` (utils/error_state_iekf.py:203-231).
Nothing of this repository's kernels or engine is on that path; only the synthetic input generators
(clid_slam_b200/synth.py) are shared, so both arms see the same world and the same sample distribution.

The reference imports visualisation / IO packages at module level that it never touches on this path
(open3d, matplotlib, wandb, ...); those that are not installed are replaced by MagicMock stubs.
"""
from __future__ import annotations

import os
import sys
import time
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")
_STUBS = ["open3d", "matplotlib", "matplotlib.cm", "matplotlib.pyplot", "roma", "skimage", "skimage.measure", "pypose",
          "natsort", "plyfile", "laspy", "evo", "pyquaternion", "rerun", "dtyper", "wandb"]


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


def _load():
    for name in _STUBS:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = MagicMock(name=name)
    for taken in ("model", "utils"):
        mod = sys.modules.get(taken)
        if mod is not None and not str(getattr(mod, "__file__", "")).startswith(REF_ROOT):
            raise RuntimeError(f"module name {taken!r} already imported from elsewhere")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from types import SimpleNamespace

    from model.decoder import Decoder
    from model.local_point_cloud_map import LocalPointCloudMap
    from model.neural_points import NeuralPoints
    from utils import tools
    from utils.config import Config
    from utils.mapper import Mapper

    return SimpleNamespace(Config=Config, Decoder=Decoder, NeuralPoints=NeuralPoints, Mapper=Mapper,
                           LocalPointCloudMap=LocalPointCloudMap, tools=tools)


class _FakeDataset:  # the attributes utils/mapper.py reads from SLAMDataset
    lose_track = False
    stop_status = False
    processed_frame = 0
    gt_pose_provided = True
    gt_poses = odom_poses = np.eye(4)[None]
    pgo_poses = None
    static_mask = None


def run(device: str, mode: str, batch: int, side: int, sheets: int, steps: int, warmup: int, threads: int = 0,
        pool_batches: int = 4, inference_passes: int = 3, local_map_radius: float = 1.0e4):
    """Times `Mapper.mapping(steps)` (after `mapping(warmup)`) and the inference forward + gradient.
    Returns a dict (samples/s of both, ms per step, the world size, the final losses are not exposed by
    the reference)."""
    import torch

    sys.path.insert(0, os.path.dirname(HERE))
    from clid_slam_b200.synth import sample_batch, wavy_sheets

    ref = _load()
    if threads > 0:
        torch.set_num_threads(threads)
    torch.manual_seed(42)
    cfg = ref.Config()
    cfg.load(os.path.join(REF_ROOT, "config", "run_ncd128.yaml"))
    cfg.device = device
    cfg.silence = True
    cfg.o3d_vis_on = False
    cfg.wandb_vis_on = False
    cfg.feature_std = 0.05
    cfg.local_map_radius = local_map_radius
    cfg.numerical_grad = mode == "numerical"
    cfg.gradient_decimation = 10 if cfg.numerical_grad else 1
    cfg.bs = batch
    dec = ref.Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    npm = ref.NeuralPoints(cfg)
    npm.travel_dist = torch.zeros(1, device=device)
    gen = torch.Generator(device=device).manual_seed(1)
    pts = wavy_sheets(side, sheets, cfg.voxel_size_m, gen, device=device)
    npm.update(pts, torch.zeros(3, device=device), torch.eye(3, device=device), 0)

    mapper = ref.Mapper(cfg, _FakeDataset(), npm, ref.LocalPointCloudMap(cfg), dec)
    mapper.used_poses = torch.eye(4, dtype=torch.float64, device=device)[None]
    mapper.adaptive_iter_offset = 0
    # replay pool = a few benchmark batches; the stock get_batch draws cfg.bs of them uniformly per iteration
    parts = [sample_batch(npm.neural_points, batch, gen) for _ in range(pool_batches)]
    x = torch.cat([p[0] for p in parts], 0)
    mapper.coord_pool = x
    mapper.global_coord_pool = x
    mapper.sdf_label_pool = torch.cat([p[1] for p in parts], 0)
    mapper.weight_pool = torch.cat([p[2] for p in parts], 0)
    mapper.time_pool = torch.cat([p[3] for p in parts], 0)
    mapper.sem_label_pool = mapper.color_pool = mapper.normal_label_pool = None
    mapper.pool_sample_count = x.shape[0]
    mapper.new_idx = None

    def sync():
        if device.startswith("cuda"):
            torch.cuda.synchronize()

    if warmup > 0:
        mapper.mapping(warmup)
    sync()
    t0 = time.perf_counter()
    mapper.mapping(steps)
    sync()
    step_s = (time.perf_counter() - t0) / steps

    fwd = []
    for i in range(inference_passes):
        xq = parts[i % pool_batches][0].clone().requires_grad_(True)
        sync()
        t0 = time.perf_counter()
        feat, _, _, _, _ = npm.query_feature(xq, training_mode=False, query_locally=True)
        sdf = dec.sdf(feat)
        ref.tools.get_gradient(xq, sdf)
        sync()
        if i >= 1 or inference_passes == 1:
            fwd.append(time.perf_counter() - t0)
    fwd_s = sum(fwd) / len(fwd)
    return {"samples_per_s": batch / step_s, "ms_per_step": step_s * 1e3, "neural_points": int(npm.count()),
            "local_points": int(npm.local_count()),
            "inference_samples_per_s": batch / fwd_s, "inference_ms": fwd_s * 1e3,
            "threads": torch.get_num_threads(), "device": device, "steps": steps, "warmup": warmup}
