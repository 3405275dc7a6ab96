"""Builders shared by the GPU parity tests: put an oracle / fixture map state into the product
classes (clid_slam_b200) on a CUDA device."""
from __future__ import annotations

import torch
import torch.nn as nn

from clid_slam_b200.config import Config
from clid_slam_b200.model.decoder import Decoder
from clid_slam_b200.model.neural_points import NeuralPoints
from oracle import sdf_oracle as oc

# Parity tolerances (BASELINE.json north_star: 1e-4 rel on SDF values/gradients, 1e-3 on loss).
# The absolute floors cover fp32 cancellation noise where the value itself is ~0:
SDF_RTOL, SDF_ATOL = 1e-4, 2e-7     # metres; sdf_scale is 0.055 m, so the floor is ~4e-6 of scale
GRAD_RTOL, GRAD_ATOL = 1e-4, 5e-6   # d sdf / d x is O(1): the floor is 5e-6 of the gradient norm
Z_RTOL, Z_ATOL = 1e-5, 1e-6
LOSS_RTOL = 1e-3


def product_config(ocfg: oc.OracleConfig, device="cuda") -> Config:
    cfg = Config()
    for name in ocfg.__dataclass_fields__:
        setattr(cfg, name, getattr(ocfg, name))
    cfg.device = device
    cfg.local_map_radius = ocfg.local_map_radius
    return cfg


def product_map(m: oc.OracleMap, device="cuda") -> NeuralPoints:
    """NeuralPoints carrying exactly the state of the oracle map `m`."""
    cfg = product_config(m.cfg, device)
    small = Config.__new__(Config)
    small.__dict__.update(cfg.__dict__)
    small.buffer_size = 1  # avoid allocating a throw-away 400 MB table in the constructor
    npm = NeuralPoints(small)
    npm.config = cfg
    npm.buffer_size = int(m.cfg.buffer_size)
    npm.buffer_pt_index = m.table.to(device)
    npm.neural_points = m.points.to(device)
    npm.point_orientations = torch.zeros(m.points.shape[0], 4, device=device)
    npm.point_orientations[:, 0] = 1
    npm.point_ts_create = m.ts_create.to(device)
    npm.point_ts_update = m.ts_update.to(device)
    npm.point_certainties = m.certainties.to(device).clone()
    npm.geo_features = m.features.detach().to(device).clone()
    npm.travel_dist = m.travel_dist.to(device)
    npm.cur_ts = int(m.cur_ts)
    npm.reboot_ts = int(m.reboot_ts)
    npm.neighbor_dx = m.offsets.to(device).contiguous()
    npm.neighbor_K = npm.neighbor_dx.shape[0]
    npm.max_valid_dist2 = float(m.max_valid_dist2)
    npm.temporal_local_map_on = bool(m.cfg.temporal_local_map_on)
    npm.diff_travel_dist_local = m.cfg.diff_travel_dist_local
    npm.local_map_radius = m.cfg.local_map_radius
    if m.local_points is not None:
        npm.local_neural_points = m.local_points.to(device)
        npm.local_point_orientations = torch.zeros(m.local_points.shape[0], 4, device=device)
        npm.local_geo_features = nn.Parameter(m.local_features.detach().to(device).clone())
        npm.local_point_certainties = m.local_certainties.to(device).clone()
        npm.local_point_ts_update = m.local_ts_update.to(device).clone()
        npm.local_mask = m.local_mask.to(device)
        npm.global2local = m.global2local.to(device)
    return npm


def product_decoder(ocfg: oc.OracleConfig, params, device="cuda") -> Decoder:
    cfg = product_config(ocfg, device)
    dec = Decoder(cfg, ocfg.geo_mlp_hidden_dim, ocfg.geo_mlp_level, 1)
    flat = dec.flat_parameters()
    assert len(flat) == len(params)
    with torch.no_grad():
        for dst, src in zip(flat, params):
            dst.copy_(src.detach().to(device))
    return dec


def build_oracle_world(n_side: int, n_sheets: int, seed: int, cfg: oc.OracleConfig = None):
    """Seeded synthetic world inserted through the oracle's map_insert."""
    cfg = cfg or oc.OracleConfig()
    gen = torch.Generator().manual_seed(seed)
    m = oc.empty_map(cfg)
    pts = oc.wavy_sheets(n_side, n_sheets, cfg.voxel_size_m, gen)
    oc.map_insert(m, pts, torch.zeros(3), 0, generator=gen)
    params = oc.init_decoder(cfg, gen)
    return m, params, gen
