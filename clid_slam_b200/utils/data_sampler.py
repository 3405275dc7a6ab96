"""Training-sample generation along LiDAR rays: mirror of the reference's
``utils.data_sampler.DataSampler`` (feeder of the hot path).

``sample_pin`` (utils/data_sampler.py:16-258) is reproduced with the same random-number call
order (one randn for the surface samples, one rand each for the front / behind free-space
samples), so that with the same torch seed and device it draws the same samples as the
reference.  Output is ray-major: for every scan point its 1 + surface_n + front_n + behind_n
samples are contiguous.
"""
from __future__ import annotations

import torch

from .tools import transform_torch


class DataSampler:
    def __init__(self, config):
        self.config = config
        self.dev = config.device

    def sample_pin(self, points_torch, normal_torch=None, sem_label_torch=None, color_torch=None):
        """points_torch [P,3] in the sensor frame -> (coord [P*S,3], sdf_label [P*S], normal | None,
        sem | None, color | None, weight [P*S]).  Label = -(displacement along the ray); weight
        sign flags free-space samples (negative)."""
        cfg, dev = self.config, self.dev
        if (points_torch.is_cuda and points_torch.dtype == torch.float32 and points_torch.shape[0] > 0
                and normal_torch is None and sem_label_torch is None and color_torch is None
                and not getattr(cfg, "behind_dropoff_on", False)):
            from ..ops import mapmaint as _mm  # one launch instead of ~30 eager ops (csrc/mapmaint.cuh)

            coord, disp, weight, _ = _mm.ray_samples(cfg, points_torch)
            return coord, -disp, None, None, None, weight
        return self._sample_pin_torch(points_torch, normal_torch, sem_label_torch, color_torch)

    def _sample_pin_torch(self, points_torch, normal_torch=None, sem_label_torch=None, color_torch=None):
        """sample_pin() with torch ops (host logic; what the CPU tests pin against the reference's fixtures)."""
        cfg, dev = self.config, self.dev
        sigma = cfg.surface_sample_range_m
        n_surf, n_front, n_behind = cfg.surface_sample_n, cfg.free_front_n, cfg.free_behind_n
        per_ray = 1 + n_surf + n_front + n_behind
        count = points_torch.shape[0]
        depth = torch.linalg.norm(points_torch, dim=1, keepdim=True)  # [P,1]

        # displacement along the ray and the matching depth ratio, block by block
        disp_hit = torch.zeros_like(depth)
        ratio_hit = torch.ones_like(depth)

        disp_surf = torch.randn(count * n_surf, 1, device=dev) * sigma
        ratio_surf = disp_surf / depth.repeat(n_surf, 1) + 1.0

        margin = 2.0  # free-space samples keep 2 sigma away from the surface
        d_front = depth.repeat(n_front, 1)
        hi = 1.0 - margin * sigma / d_front
        lo = cfg.free_sample_begin_ratio
        ratio_front = torch.rand(count * n_front, 1, device=dev) * (hi - lo) + lo
        disp_front = (ratio_front - 1.0) * d_front

        d_behind = depth.repeat(n_behind, 1)
        hi = cfg.free_sample_end_dist_m / d_behind + 1.0
        lo = 1.0 + margin * sigma / d_behind
        ratio_behind = torch.rand(count * n_behind, 1, device=dev) * (hi - lo) + lo
        disp_behind = (ratio_behind - 1.0) * d_behind

        disp = torch.cat((disp_hit, disp_surf, disp_front, disp_behind), 0)
        ratio = torch.cat((ratio_hit, ratio_surf, ratio_front, ratio_behind), 0)
        coord = points_torch.repeat(per_ray, 1) * ratio
        depth_all = depth.repeat(per_ray, 1)

        weight = torch.ones_like(depth_all)
        n_near = count * (n_surf + 1)
        if cfg.dist_weight_on:  # far surface samples weigh less: [1 - s/2, 1 + s/2]
            weight[:n_near] = 1 + cfg.dist_weight_scale * 0.5 - (depth_all[:n_near] / cfg.max_range) * cfg.dist_weight_scale
        if getattr(cfg, "behind_dropoff_on", False):
            top = cfg.free_sample_end_dist_m
            bottom = 0.2 * top
            fade = torch.clamp((top - disp) / (top - bottom), min=0.0, max=1.0)
            weight = weight * (fade * 0.8 + 0.2)
        weight[n_near:] *= -1.0

        def ray_major(t, width=None):
            if width is None:
                return t.reshape(per_ray, -1).transpose(0, 1).reshape(-1)
            return t.reshape(per_ray, -1, width).transpose(0, 1).reshape(-1, width)

        coord = ray_major(coord, 3)
        sdf_label = -ray_major(disp.squeeze(1))
        weight = ray_major(weight)

        normal = None if normal_torch is None else ray_major(normal_torch.repeat(per_ray, 1), 3)
        sem = None
        if sem_label_torch is not None:
            free = torch.zeros(count * (n_front + n_behind), device=dev, dtype=sem_label_torch.dtype)
            sem = ray_major(torch.cat((sem_label_torch.repeat(1 + n_surf), free), 0).int())
        color = None
        if color_torch is not None:
            ch = color_torch.shape[1]
            free = torch.zeros(count * (n_front + n_behind), ch, device=dev, dtype=color_torch.dtype)
            color = ray_major(torch.cat((color_torch.repeat(1 + n_surf, 1), free), 0), ch)
        return coord, sdf_label, normal, sem, color, weight

    def sample(self, points_torch, local_point_cloud_map, cur_pose_torch):
        """CLID-SLAM's sampler (utils/data_sampler.py:260-402): the same ray samples as sample_pin,
        but the labels of the near-surface samples come from the local point-cloud map
        (``region_specific_sdf_estimation``: point-to-plane distance where a plane fits, nearest
        point distance otherwise), signed by the side of the surface the sample was drawn on;
        samples with no stored point in reach are dropped.  Returns (coord, sdf_label, weight)."""
        cfg, dev = self.config, self.dev
        if points_torch.is_cuda and points_torch.dtype == torch.float32 and points_torch.shape[0] > 0:
            from ..ops import mapmaint as _mm  # ray samples, labels and compaction as kernels (csrc/mapmaint.cuh)

            return _mm.region_labelled_samples(cfg, points_torch, local_point_cloud_map, cur_pose_torch)
        return self._sample_torch(points_torch, local_point_cloud_map, cur_pose_torch)

    def _sample_torch(self, points_torch, local_point_cloud_map, cur_pose_torch):
        """sample() with torch ops (host logic; what the CPU tests pin against the reference's fixtures)."""
        cfg, dev = self.config, self.dev
        sigma = cfg.surface_sample_range_m
        n_surf, n_front, n_behind = cfg.surface_sample_n, cfg.free_front_n, cfg.free_behind_n
        per_ray = 1 + n_surf + n_front + n_behind
        count = points_torch.shape[0]
        depth = torch.linalg.norm(points_torch, dim=1, keepdim=True)

        disp_surf = torch.randn(count * n_surf, 1, device=dev) * sigma
        ratio_surf = disp_surf / depth.repeat(n_surf, 1) + 1.0
        margin = 2.0
        d_front = depth.repeat(n_front, 1)
        hi = 1.0 - margin * sigma / d_front
        lo = cfg.free_sample_begin_ratio
        ratio_front = torch.rand(count * n_front, 1, device=dev) * (hi - lo) + lo
        disp_front = (ratio_front - 1.0) * d_front
        d_behind = depth.repeat(n_behind, 1)
        hi = cfg.free_sample_end_dist_m / d_behind + 1.0
        lo = 1.0 + margin * sigma / d_behind
        ratio_behind = torch.rand(count * n_behind, 1, device=dev) * (hi - lo) + lo
        disp_behind = (ratio_behind - 1.0) * d_behind

        disp = torch.cat((torch.zeros_like(depth), disp_surf, disp_front, disp_behind), 0)
        ratio = torch.cat((torch.ones_like(depth), ratio_surf, ratio_front, ratio_behind), 0)
        coord = points_torch.repeat(per_ray, 1) * ratio
        depth_all = depth.repeat(per_ray, 1)

        # region-specific labels for the near-surface samples (block 1 .. n_surf of the block order)
        n_near = count * (n_surf + 1)
        sign = torch.where(disp_surf.squeeze(1) < 0, 1, -1)
        keep = torch.ones(count * per_ray, dtype=torch.bool, device=dev)
        sdf_label = -1 * disp.squeeze(1)
        near_world = transform_torch(coord[count:n_near], cur_pose_torch)
        dist, reachable = local_point_cloud_map.region_specific_sdf_estimation(near_world)
        keep[count:n_near] = reachable
        sdf_label[count:n_near] = sign * dist

        weight = torch.ones_like(depth_all)
        if cfg.dist_weight_on:
            weight[:n_near] = 1 + cfg.dist_weight_scale * 0.5 - (depth_all[:n_near] / cfg.max_range) * cfg.dist_weight_scale
        weight[n_near:] *= -1.0

        coord = coord.reshape(per_ray, -1, 3).transpose(0, 1).reshape(-1, 3)
        sdf_label = sdf_label.reshape(per_ray, -1).transpose(0, 1).reshape(-1)
        weight = weight.reshape(per_ray, -1).transpose(0, 1).reshape(-1)
        keep = keep.reshape(per_ray, -1).transpose(0, 1).reshape(-1)
        return coord[keep], sdf_label[keep], weight[keep]
