"""Local raw-point map and region-specific SDF labels: mirror of the reference's
``model.local_point_cloud_map.LocalPointCloudMap`` (model/local_point_cloud_map.py:11-201), the
label generator CLID-SLAM adds on top of PIN-SLAM (first "next" row).

A second voxel hash (own prime triple, 0.2 m voxels, 5e6 slots) over the raw scan points near the
sensor.  For a batch of surface samples it finds the 4 nearest stored points among the 7 probed
cells, fits a plane to them (SVD), and returns the point-to-plane distance where the fit is good
and the nearest-point distance otherwise.  Host logic in torch ops on the map's device (runs once
per frame on 4 x #scan-points samples); duplicate-slot writes are made deterministic (last writer
wins, what the reference's CPU index_put does).
"""
from __future__ import annotations


import torch

from ..utils.tools import ieee_div, voxel_down_sample_torch

PRIMES_LOCAL = (73856093, 19349663, 83492791)  # model/local_point_cloud_map.py:27-29 (2nd prime differs)


def _store_last_wins(table: torch.Tensor, slots: torch.Tensor, values: torch.Tensor) -> None:
    if table.is_cuda:
        from ..ops import mapmaint as _mm  # bid / commit kernels (csrc/mapmaint.cuh)

        _mm.table_store(table, slots, values)
        return
    slots = torch.remainder(slots, table.shape[0])
    uniq, inverse = torch.unique(slots, return_inverse=True)
    pos = torch.arange(slots.shape[0], device=slots.device)
    last = torch.empty(uniq.shape, dtype=pos.dtype, device=slots.device)
    last.scatter_reduce_(0, inverse, pos, reduce="amax", include_self=False)
    table[uniq] = values[last]


class LocalPointCloudMap:
    def __init__(self, config) -> None:
        self.config = config
        self.idx_dtype = torch.int64
        self.dtype = config.dtype
        self.device = config.device
        self.resolution = config.local_voxel_size_m
        self.buffer_size = int(config.local_buffer_size)
        self.buffer_pt_index = torch.full((self.buffer_size,), -1, dtype=self.idx_dtype, device=self.device)
        self.local_point_cloud_map = torch.empty((0, 3), dtype=torch.float32, device=self.device)
        self.primes = torch.tensor(PRIMES_LOCAL, dtype=self.idx_dtype, device=self.device)
        self.neighbor_idx = None
        self.max_valid_range = None
        self.set_search_neighborhood()
        self.map_size = config.local_map_size

    def voxel_hash(self, points: torch.Tensor) -> torch.Tensor:
        cells = ieee_div(points, self.resolution).floor().to(self.primes)
        return torch.fmod((cells * self.primes).sum(-1), self.buffer_size)

    def insert_points(self, points: torch.Tensor) -> None:
        """One stored point per voxel that is still empty (model/local_point_cloud_map.py:40-56)."""
        cand = points[voxel_down_sample_torch(points, self.resolution)]
        slots = self.voxel_hash(cand)
        empty = self.buffer_pt_index[slots] == -1
        fresh = cand[empty]
        ids = torch.arange(fresh.shape[0], device=self.device) + self.local_point_cloud_map.shape[0]
        _store_last_wins(self.buffer_pt_index, slots[empty], ids)
        self.local_point_cloud_map = torch.cat((self.local_point_cloud_map, fresh), 0)

    def update_map(self, sensor_position: torch.Tensor, points: torch.Tensor) -> None:
        """Insert a scan, drop points farther than map_size from the sensor, rebuild the hash
        (model/local_point_cloud_map.py:58-72)."""
        self.insert_points(points)
        cloud = self.local_point_cloud_map
        if cloud.is_cuda and cloud.dtype == torch.float32 and cloud.shape[0] > 0:
            from ..ops import mapmaint as _mm  # flag + scan + compaction on the device, one read-back

            (self.local_point_cloud_map,), _, _, _ = _mm.pool_filter(cloud, sensor_position, self.map_size, [cloud], 0, use_norm=True)
        else:
            near = torch.norm(cloud - sensor_position, dim=-1) < self.map_size
            self.local_point_cloud_map = cloud[near]
        table = torch.full((self.buffer_size,), -1, dtype=self.idx_dtype, device=self.device)
        ids = torch.arange(self.local_point_cloud_map.shape[0], device=self.device)
        _store_last_wins(table, self.voxel_hash(self.local_point_cloud_map), ids)
        self.buffer_pt_index = table

    def set_search_neighborhood(self, num_nei_cells: int = 1, search_alpha: float = 0.2) -> None:
        r = torch.arange(-num_nei_cells, num_nei_cells + 1, device=self.primes.device, dtype=self.primes.dtype)
        cube = torch.stack(torch.meshgrid(r, r, r, indexing="ij"), dim=-1).reshape(-1, 3)
        self.neighbor_idx = cube[(cube**2).sum(-1) < (num_nei_cells + search_alpha) ** 2]
        self.max_valid_range = 1.732 * (num_nei_cells + 1) * self.resolution

    def region_specific_sdf_estimation(self, points: torch.Tensor):
        """(|sdf| [N], surface_mask [N]) for surface samples in the world frame
        (model/local_point_cloud_map.py:98-152).  On a CUDA map: one launch of clid_region_sdf."""
        if points.is_cuda and self.buffer_pt_index.is_cuda:
            return self._region_sdf_native(points)
        return self._region_sdf_torch(points)

    def _region_sdf_native(self, points: torch.Tensor):
        import ctypes as C

        from .. import _lib

        pts = points.detach().float().contiguous()
        n = pts.shape[0]
        c = _lib.ClidLocalCloud()
        c.table = _lib.ptr(self.buffer_pt_index, torch.int64, "local buffer_pt_index")
        c.buffer_size = self.buffer_size
        for i, pr in enumerate(PRIMES_LOCAL):
            c.primes[i] = pr
        cloud = self.local_point_cloud_map.contiguous()
        c.points = _lib.ptr(cloud, torch.float32, "local_point_cloud_map") if cloud.shape[0] else None
        c.n_points = cloud.shape[0]
        nidx = self.neighbor_idx.contiguous()
        c.neighbor_idx, c.kc = _lib.ptr(nidx, torch.int64, "neighbor_idx"), nidx.shape[0]
        c.resolution, c.max_valid_range = float(self.resolution), float(self.max_valid_range)
        sdf_abs = torch.empty(n, dtype=torch.float32, device=pts.device)
        mask = torch.empty(n, dtype=torch.uint8, device=pts.device)
        with torch.cuda.device(pts.device):
            rc = _lib.load().clid_region_sdf(C.byref(c), pts.data_ptr(), n, sdf_abs.data_ptr(), mask.data_ptr(),
                                            _lib.current_stream(pts.device))
        _lib.check(rc, "clid_region_sdf")
        mask = mask.bool()
        if not self.config.silence:
            print(mask.sum().item() / max(mask.numel(), 1))
        return sdf_abs, mask

    def _region_sdf_torch(self, points: torch.Tensor):
        """The same with torch ops (host logic; what the CPU tests pin against the reference's fixtures)."""
        n = points.shape[0]
        far = self.max_valid_range
        sdf_abs = torch.full((n,), far, device=points.device, dtype=torch.float32)
        surface_mask = torch.ones(n, dtype=torch.bool, device=points.device)
        chunk = 262144
        for head in range(0, n, chunk):
            pts = points[head:head + chunk, :]
            cells = ieee_div(pts, self.resolution).floor().to(self.primes)
            cells = cells[..., None, :] + self.neighbor_idx
            slots = torch.fmod((cells * self.primes).sum(-1), self.buffer_size)
            idx = self.buffer_pt_index[slots]
            cand = self.local_point_cloud_map[idx]  # idx == -1 reads the last point, masked below
            dist = torch.norm(cand - pts.view(-1, 1, 3), dim=-1)
            dist = torch.where(idx == -1, far, dist)
            near_d, near_i = torch.topk(dist, 4, largest=False, dim=1)
            knn = torch.gather(cand, 1, near_i.unsqueeze(-1).expand(-1, -1, 3))
            four = near_d[:, 3] < far  # four real neighbours: a plane can be fitted

            normal = torch.zeros_like(pts)
            offset = torch.zeros(pts.shape[0], device=pts.device)
            good = torch.zeros(pts.shape[0], dtype=torch.bool, device=pts.device)
            n_fit, c_fit, ok_fit = estimate_plane(knn[four])
            normal[four] = n_fit
            offset[four] = c_fit
            good[four] = ok_fit
            good &= four
            surface_mask[head:head + chunk] &= near_d[:, 0] < far
            plane_dist = torch.abs((normal * pts).sum(dim=1) + offset)
            sdf_abs[head:head + chunk] = torch.where(good, plane_dist, near_d[:, 0])
        if not self.config.silence:
            print(surface_mask.sum().item() / max(surface_mask.numel(), 1))
        return sdf_abs, surface_mask


def estimate_plane(points: torch.Tensor, eta_threshold: float = 0.2, threshold: float = 0.1):
    """Least-squares planes through [M,4,3] point sets (model/local_point_cloud_map.py:155-201).
    Returns (unit normal [M,3] or 0, plane constant [M], success [M]): success needs a flat
    neighbourhood (smallest / middle singular value <= eta_threshold) and every point within
    `threshold` of the plane."""
    centroid = points.mean(dim=1, keepdim=True)
    _, sing, vh = torch.linalg.svd(points - centroid, full_matrices=False)
    flat = sing[:, -1] / (sing[:, 1] + 1e-6) <= eta_threshold
    normal = torch.where(flat.unsqueeze(1), vh[:, -1, :], torch.zeros_like(vh[:, -1, :]))
    constant = -(normal * centroid.squeeze(1)).sum(dim=1)
    resid = torch.abs(torch.bmm(points, normal.unsqueeze(-1)).squeeze(-1) + constant.unsqueeze(-1))
    return normal, constant, (resid.max(dim=1).values <= threshold) & flat
