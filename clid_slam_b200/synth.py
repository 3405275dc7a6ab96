"""Seeded synthetic LiDAR-like worlds and sample batches on any device.

Used by bench.py, the smoke test and the GPU tests to create inputs of the benchmark shapes
without a dataset: wavy sheets of surface points (inserted through NeuralPoints.update) and
batches that mirror the sampler's 5:2:1 surface / free-front / free-behind mixture
(config/run_ncd128.yaml:17-21).
"""
from __future__ import annotations

import torch


def wavy_sheets(n_side: int, n_sheets: int, pitch: float, generator: torch.Generator, device="cpu") -> torch.Tensor:
    """Points of `n_sheets` sheets z = 3 s + 0.8 sin(x/5) cos(y/7) + N(0, 0.05^2) on an n_side^2 grid."""
    half = n_side * pitch / 2
    ax = torch.arange(n_side, dtype=torch.float32, device=device) * pitch - half + pitch / 2
    gx, gy = torch.meshgrid(ax, ax, indexing="ij")
    sheets = []
    for s in range(n_sheets):
        z = 3.0 * s + 0.8 * torch.sin(gx / 5) * torch.cos(gy / 7)
        z = z + 0.05 * torch.randn(gx.shape, generator=generator, device=device)
        sheets.append(torch.stack((gx, gy, z), -1).reshape(-1, 3))
    return torch.cat(sheets, 0)


def sample_batch(anchor_points: torch.Tensor, n: int, generator: torch.Generator, max_range: float = 60.0):
    """(x [n,3], sdf label [n], signed weight [n], ts [n] int32) anchored at random map points."""
    dev = anchor_points.device
    pick = torch.randint(0, anchor_points.shape[0], (n,), generator=generator, device=dev)
    x = anchor_points[pick].clone()
    x[:, :2] += 0.1 * torch.randn(n, 2, generator=generator, device=dev)
    kind = torch.rand(n, generator=generator, device=dev)
    disp = 0.25 * torch.randn(n, generator=generator, device=dev)
    front = -(0.5 + 8.0 * torch.rand(n, generator=generator, device=dev))
    behind = 0.5 + 0.7 * torch.rand(n, generator=generator, device=dev)
    is_front = (kind >= 0.625) & (kind < 0.875)
    is_behind = kind >= 0.875
    disp = torch.where(is_front, front, disp)
    disp = torch.where(is_behind, behind, disp)
    x[:, 2] += disp
    label = -disp
    weight = 1.0 + 0.4 - 0.8 * x.norm(dim=-1) / max_range
    weight = torch.where(is_front | is_behind, -weight, weight)
    ts = torch.zeros(n, dtype=torch.int32, device=dev)
    return x.contiguous(), label, weight, ts


def voxel_features(points: torch.Tensor, resolution: float, dim: int, std: float) -> torch.Tensor:
    """[n, dim] pseudo-random features that are a function of each point's VOXEL only (not of its row or of the
    generator state): ranks of a partitioned map that hold the same voxel initialise it alike."""
    from .utils.tools import ieee_div

    c = torch.floor(ieee_div(points, float(resolution)))
    ch = torch.arange(dim, device=points.device, dtype=torch.float32)
    phase = c[:, 0:1] * 12.9898 + c[:, 1:2] * 78.233 + c[:, 2:3] * 37.719 + ch * 1.618
    return std * 1.41421 * torch.sin(phase * 43.0)


def set_voxel_features(npm, std: float) -> None:
    """Overwrite the map's (global and local-window) geometric features with voxel_features."""
    d = npm.geo_feature_dim
    with torch.no_grad():
        npm.geo_features[:-1] = voxel_features(npm.neural_points, npm.resolution, d, std)
        npm.local_geo_features.data[:-1] = voxel_features(npm.local_neural_points, npm.resolution, d, std)
    npm._touch()
