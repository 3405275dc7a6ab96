"""Hot-path subset of the reference's ``utils.tools`` (same names, same results).

Only the helpers the neural-SDF path needs are mirrored: gradient and
optimiser glue, voxel down-sampling used by the map insert, rigid transforms, timing.
Plot / wandb / Open3D / PCA helpers of the reference file are out of scope.
"""
from __future__ import annotations

import time

import torch
import torch.nn as nn
from torch import optim


def get_time() -> float:
    """Wall clock after draining the device (utils/tools.py:385-392)."""
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return time.time()


def get_gradient(inputs: torch.Tensor, outputs: torch.Tensor) -> torch.Tensor:
    """d outputs / d inputs by autograd, graph kept (utils/tools.py:298-311)."""
    seed = torch.ones_like(outputs, requires_grad=False)
    (g,) = torch.autograd.grad(outputs, inputs, seed, create_graph=True, retain_graph=True, only_inputs=True)
    return g


def freeze_model(model: nn.Module) -> None:  # utils/tools.py:314-317
    for child in model.children():
        for p in child.parameters():
            p.requires_grad = False


def unfreeze_model(model: nn.Module) -> None:
    for child in model.children():
        for p in child.parameters():
            p.requires_grad = True


def setup_optimizer(config, neural_point_feat, mlp_geo_param=None, mlp_sem_param=None,
                    mlp_color_param=None, poses=None, lr_ratio=1.0):
    """Adam(betas 0.9/0.99, eps adam_eps) or SGD(momentum 0.9); decoder groups carry no weight
    decay, the feature group carries config.weight_decay (utils/tools.py:205-255)."""
    lr = config.lr * lr_ratio
    groups = []
    extra = [(mlp_geo_param, True), (mlp_sem_param, getattr(config, "semantic_on", False)),
             (mlp_color_param, getattr(config, "color_on", False))]
    for params, enabled in extra:
        if params is not None and enabled:
            groups.append({"params": params, "lr": lr, "weight_decay": 0.0})
    if poses is not None:
        groups.append({"params": poses, "lr": config.lr_pose, "weight_decay": config.weight_decay})
    groups.append({"params": neural_point_feat, "lr": lr, "weight_decay": config.weight_decay})
    if config.opt_adam:
        return optim.Adam(groups, betas=(0.9, 0.99), eps=config.adam_eps)
    return optim.SGD(groups, momentum=0.9)


def transform_torch(points: torch.Tensor, transformation: torch.Tensor) -> torch.Tensor:
    """Rigid transform of [N,3] points by a 4x4 matrix (utils/tools.py:590-609)."""
    tf = transformation.to(points)
    return points @ tf[:3, :3].T + tf[:3, 3]


def transform_batch_torch(points: torch.Tensor, transformation: torch.Tensor) -> torch.Tensor:
    """Per-point rigid transform, transformation [N,4,4] (utils/tools.py:612-636)."""
    rot = transformation[:, :3, :3].to(points)
    trans = transformation[:, :3, 3].to(points)
    return torch.einsum("nij,nj->ni", rot, points) + trans


def ieee_div(x: torch.Tensor, divisor: float) -> torch.Tensor:
    """x / divisor with IEEE division on every device.  torch's CUDA kernels turn `tensor / python_scalar` into a
    multiplication by the reciprocal, which moves a coordinate that sits within one ulp of a voxel face into the
    neighbouring cell; the CPU kernels (what the reference fixtures pin) and the CUDA kernels of this library
    (cell_of: floorf(__fdiv_rn(x, res))) divide.  A 0-dim tensor divisor makes torch divide too, so the host logic,
    the kernels and the CPU reference agree on every point's voxel."""
    return x / torch.full((), divisor, dtype=x.dtype, device=x.device)


def _voxel_keys(points: torch.Tensor, voxel_size: float):
    lowest = torch.floor(ieee_div(points.min(dim=0)[0], voxel_size)).long()
    cell_f = torch.floor(ieee_div(points, voxel_size))
    cell = cell_f.long() - lowest
    span = cell.max()  # the reference uses one span for all axes (utils/tools.py:662-663)
    key = cell[:, 0] + cell[:, 1] * span + cell[:, 2] * span * span
    return cell_f, key


def _argmin_per_voxel(key: torch.Tensor, rank: torch.Tensor) -> torch.Tensor:
    """For every distinct key (ascending) the index with the smallest (rank, index)."""
    n = key.shape[0]
    uniq, inverse = torch.unique(key, return_inverse=True)
    order = torch.arange(n, dtype=torch.int64, device=key.device)
    packed = rank * n + order
    best = torch.empty(uniq.shape, dtype=torch.int64, device=key.device)
    best.scatter_reduce_(0, inverse, packed, reduce="amin", include_self=False)
    return best % n


def voxel_down_sample_torch(points: torch.Tensor, voxel_size: float) -> torch.Tensor:
    """Indices of the point closest to each occupied voxel's centre, distance quantised to
    1000 levels with ties going to the smaller index (utils/tools.py:639-682)."""
    if points.is_cuda and points.dtype == torch.float32 and points.shape[0] > 0:
        from ..ops import mapmaint as _mm  # keys -> stable sort -> run heads (csrc/mapmaint.cuh)

        native = _mm.voxel_down_sample(points, voxel_size)
        if native is not None:
            return native
    levels = 1000
    cell_f, key = _voxel_keys(points, voxel_size)
    centre = (cell_f + 0.5) * voxel_size
    dist = ((points - centre) ** 2).sum(dim=1) ** 0.5
    rank = (dist / dist.max() * (levels - 1)).long()
    return _argmin_per_voxel(key, rank)


def voxel_down_sample_min_value_torch(points: torch.Tensor, voxel_size: float, value: torch.Tensor) -> torch.Tensor:
    """Indices of the point with the smallest `value` in each voxel (utils/tools.py:685-724)."""
    if points.is_cuda and points.dtype == torch.float32 and value.dtype == torch.float32 and points.shape[0] > 0:
        from ..ops import mapmaint as _mm

        native = _mm.voxel_down_sample(points, voxel_size, value)
        if native is not None:
            return native
    levels = 1000
    _, key = _voxel_keys(points, voxel_size)
    rank = (value / value.max() * (levels - 1)).long()
    return _argmin_per_voxel(key, rank)
