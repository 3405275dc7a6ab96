"""Measurement model of the scan-to-map registration: mirror of ``IEKFOM.h_model`` and of the normal equations
``IEKFOM.update_iterated`` builds from it (utils/error_state_iekf.py:176-264, :303-309 of the reference).

The reference evaluates, per registration iteration (up to 50 per frame), query_feature -> Decoder.sdf ->
get_gradient (an autograd backward with create_graph) on the whole scan, then ~25 eager ops and a boolean-mask
compaction (device -> host synchronisation) to form H and R_inv, and two [18, N] x [N, 18] products.  Here:

    h_model(...)           one fused launch (sdf + closed-form gradient + counts), then the reference's outputs
                           (sdf_residual, H [Nv, 18], valid_points, R_inv) for drop-in use -- compacts, so it syncs
    normal_equations(...)  the same launch + clid_registration_terms: H^T R^-1 H, H^T R^-1 z and the valid count
                           reduced on the device in fp64, no compaction, no synchronisation; update_iterated's
                           K z and K H are K_front @ g and K_front @ S (K = K_front H^T R^-1)

The filter state (StateIkfom, boxplus / boxminus, predict) is 18 x 18 host-side algebra and stays with the
reference's tracker; only this measurement path touches the neural map.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import torch

from .. import _lib, fused
from .tools import transform_torch


def _pose(rot: torch.Tensor, pos: torch.Tensor, device) -> torch.Tensor:
    T = torch.eye(4, device=device)  # fp32 like the reference's torch.eye(4) (error_state_iekf.py:182-184)
    T[:3, :3] = rot.to(device=device, dtype=torch.float32)
    T[:3, 3] = pos.to(device=device, dtype=torch.float32)
    return T


def _thresholds(config):
    max_sdf_std = config.surface_sample_range_m * getattr(config, "max_sdf_std_ratio", 1.0)
    if not max_sdf_std > 0:
        raise ValueError("max_sdf_std must be positive (sdf_std is identically 0 with weighted_first)")
    return (int(getattr(config, "track_mask_query_nn_k", config.query_nn_k)),
            float(getattr(config, "reg_min_grad_norm", 0.5)), float(getattr(config, "reg_max_grad_norm", 1.5)))


def _forward(config, neural_points, decoder, pc_imu, rot, pos):
    if not config.weighted_first:
        raise NotImplementedError("weighted_first=False is not used by any shipped configuration")
    pc_imu = pc_imu.float().contiguous()
    pc_map = transform_torch(pc_imu, _pose(rot, pos, pc_imu.device))
    sdf, grad, nn, _ = fused.sdf_and_gradient(neural_points, decoder, pc_map, training_mode=False, query_locally=True,
                                              with_gradient=True, with_certainty=False)
    return pc_imu, pc_map, sdf, grad, nn


def normal_equations(config, neural_points, decoder, pc_imu: torch.Tensor, rot: torch.Tensor, pos: torch.Tensor
                     ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(S [18,18] f64 = H^T R^-1 H, g [18] f64 = H^T R^-1 z, n_valid [] f64) of the scan pc_imu [N,3] (imu frame) at
    the pose (rot [3,3], pos [3]); device tensors, nothing is read back."""
    pc_imu, _, sdf, grad, nn = _forward(config, neural_points, decoder, pc_imu, rot, pos)
    min_nn, gmin, gmax = _thresholds(config)
    dev = pc_imu.device
    out = torch.zeros(28, dtype=torch.float64, device=dev)
    r9 = (C.c_float * 9)(*[float(v) for v in rot.to(dtype=torch.float32).reshape(-1).tolist()])
    with torch.cuda.device(dev):
        rc = _lib.load().clid_registration_terms(pc_imu.data_ptr(), sdf.data_ptr(), grad.data_ptr(), nn.data_ptr(),
                                                pc_imu.shape[0], r9, min_nn, gmin, gmax, out.data_ptr(), None,
                                                _lib.current_stream(dev))
    _lib.check(rc, "clid_registration_terms")
    iu = torch.triu_indices(6, 6, device=dev)
    S6 = torch.zeros(6, 6, dtype=torch.float64, device=dev)
    S6[iu[0], iu[1]] = out[:21]
    S6 = S6 + S6.T - torch.diag(torch.diagonal(S6))
    S = torch.zeros(18, 18, dtype=torch.float64, device=dev)
    S[:6, :6] = S6
    g = torch.zeros(18, dtype=torch.float64, device=dev)
    g[:6] = out[21:27]
    return S, g, out[27]


def h_model(config, neural_points, decoder, pc_imu: torch.Tensor, rot: torch.Tensor, pos: torch.Tensor):
    """The reference's return values: (sdf_residual [Nv] f64, H [Nv,18] f64, valid_points [Nv,3], R_inv [Nv] f64)."""
    pc_imu, pc_map, sdf, grad, nn = _forward(config, neural_points, decoder, pc_imu, rot, pos)
    min_nn, gmin, gmax = _thresholds(config)
    gn = grad.norm(dim=-1)
    valid = (nn >= min_nn) & (gn < gmax) & (gn > gmin)
    p, g, s, gn = pc_imu[valid], grad[valid], sdf[valid], gn[valid]
    n = p.shape[0]
    tran = getattr(config, "tran_dtype", torch.float64)
    H = torch.zeros((n, 18), device=p.device, dtype=tran)
    hat = torch.zeros(n, 3, 3, device=p.device)
    hat[:, 0, 1], hat[:, 0, 2] = -p[:, 2], p[:, 1]
    hat[:, 1, 0], hat[:, 1, 2] = p[:, 2], -p[:, 0]
    hat[:, 2, 0], hat[:, 2, 1] = -p[:, 1], p[:, 0]
    A = torch.bmm(rot.to(device=p.device, dtype=torch.float32).unsqueeze(0).repeat(n, 1, 1), hat)
    H[:, 0:3] = -torch.bmm(g.unsqueeze(1), A).squeeze(1)
    H[:, 3:6] = g
    z = s.to(tran)
    an = (gn - 1.0).to(tran)
    R_inv = (1 / (1 + an**2)) * (0.4 / (0.4 + z**2)) * 1000
    return z, H, pc_map[valid], R_inv
