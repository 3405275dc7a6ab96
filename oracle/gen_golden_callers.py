"""Golden fixtures for the INFERENCE CALLERS of the hot path, produced by running the
UNMODIFIED reference (/root/reference) on CPU:

  caller_iekf    utils/error_state_iekf.py:176-264  IEKFOM.h_model(pc_imu) on a briefly trained map
                 -> sdf_residual, H[:, :6], R_inv, the valid mask, and the normal equations update_iterated builds
                    from them (H^T R^-1 H, H^T R^-1 z, :303-309)
  caller_mesher  utils/mesher.py:38-163             Mesher.query_points(grid, bs) -> sdf_pred, mc_mask

TEST INFRASTRUCTURE ONLY.  Run in the build container:  python -m oracle.gen_golden_callers
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import gen_golden as gg  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle import sdf_oracle as oc  # noqa: E402


def trained_world(ref, seed=3, iters=60):
    """One wavy sheet, decoder + features trained for a few iterations by the reference's own Mapper so that the SDF
    and its gradient are meaningful (|grad| ~ 1 near the surface)."""
    cfg = gg.make_ref_config(ref, buffer_size=500_009, numerical_grad=False)
    cfg.local_map_radius = 60.0
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    dec = ref.Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    pts = oc.wavy_sheets(72, 1, 0.4, gen)
    npm = gg.populate(ref, cfg, [(pts, torch.zeros(3), 0)], [0.0])
    cfg.bs = 4096
    mapper = ref.Mapper(cfg, gg.FakeDataset(), npm, ref.LocalPointCloudMap(cfg), dec)
    mapper.used_poses = torch.eye(4, dtype=torch.float64)[None]
    mapper.adaptive_iter_offset = 0
    x, label, weight, ts = oc.sample_batch(npm.neural_points, 4 * cfg.bs, gen)
    mapper.coord_pool = mapper.global_coord_pool = x
    mapper.sdf_label_pool, mapper.weight_pool, mapper.time_pool = label, weight, ts
    mapper.sem_label_pool = mapper.color_pool = mapper.normal_label_pool = None
    mapper.pool_sample_count = x.shape[0]
    mapper.new_idx = None
    mapper.mapping(iters)
    return cfg, dec, npm, gen, pts


def iekf_case(ref):
    sys.path.insert(0, ref_loader.REFERENCE_ROOT)
    from utils.error_state_iekf import IEKFOM

    cfg, dec, npm, gen, pts = trained_world(ref)
    ekf = IEKFOM(cfg, npm, dec)
    # a slightly wrong pose: 1.5 deg about z, 6 cm off
    ang = np.deg2rad(1.5)
    rot = torch.tensor([[np.cos(ang), -np.sin(ang), 0.0], [np.sin(ang), np.cos(ang), 0.0], [0.0, 0.0, 1.0]], dtype=cfg.tran_dtype)
    pos = torch.tensor([0.04, -0.03, 0.03], dtype=cfg.tran_dtype)
    ekf.x.rot, ekf.x.pos = rot, pos
    # a scan in the sensor (imu) frame: noisy surface points, plus some far points that must be rejected
    pick = torch.randint(0, pts.shape[0], (6000,), generator=gen)
    scan = pts[pick] + 0.02 * torch.randn(6000, 3, generator=gen)
    scan[:300] += torch.tensor([0.0, 0.0, 6.0])
    pc_imu = ((scan - pos.float()) @ rot.float()).contiguous()  # R^T (p - t)
    before = gg.map_state_arrays(npm)
    z, H, valid_points = ekf.h_model(pc_imu.clone())
    R_inv = ekf.R_inv
    # which scan points survived (h_model returns the compacted set)
    T = torch.eye(4)
    T[:3, :3] = rot
    T[:3, 3] = pos
    pc_map = ref.tools.transform_torch(pc_imu, T)
    valid = torch.zeros(pc_imu.shape[0], dtype=torch.bool)
    vp = {tuple(np.round(p, 5)) for p in valid_points.detach().numpy().tolist()}
    for i, p in enumerate(pc_map.detach().numpy().tolist()):
        valid[i] = tuple(np.round(p, 5)) in vp
    assert int(valid.sum()) == valid_points.shape[0], (int(valid.sum()), valid_points.shape[0])
    HtRinv = H.T * R_inv
    S = HtRinv @ H
    g = HtRinv @ z
    out = dict(before)
    out.update(gg.decoder_arrays(dec))
    out.update(cfg=gg.cfg_json(cfg), pc_imu=pc_imu.numpy(), rot=rot.numpy(), pos=pos.numpy(),
               track_mask_query_nn_k=cfg.track_mask_query_nn_k, reg_min_grad_norm=cfg.reg_min_grad_norm,
               reg_max_grad_norm=cfg.reg_max_grad_norm, max_sdf_std=cfg.surface_sample_range_m * cfg.max_sdf_std_ratio,
               out_z=z.numpy(), out_H6=H[:, :6].numpy(), out_valid=valid.numpy(), out_R_inv=R_inv.numpy(),
               out_S=S.numpy(), out_g=g.numpy(), out_valid_points=valid_points.detach().numpy())
    assert float(H[:, 6:].abs().max()) == 0.0
    path = os.path.join(gg.GOLDEN_DIR, "caller_iekf.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "valid", int(valid.sum()), "of", pc_imu.shape[0])


def mesher_case(ref):
    sys.path.insert(0, ref_loader.REFERENCE_ROOT)
    from utils.mesher import Mesher

    cfg, dec, npm, gen, pts = trained_world(ref, seed=4, iters=30)
    mesher = Mesher(cfg, npm, {"sdf": dec, "semantic": None, "color": None})
    # a dense grid crossing the sheet (the mesher's query pattern), 0.2 m pitch
    ax = torch.arange(-6.0, 6.0, 0.2)
    az = torch.arange(-2.0, 2.0, 0.2)
    grid = torch.stack(torch.meshgrid(ax, ax, az, indexing="ij"), -1).reshape(-1, 3).contiguous()
    before = gg.map_state_arrays(npm)
    sdf_pred, _, _, mc_mask = mesher.query_points(grid, 20000, query_sdf=True, query_mask=True, query_locally=False,
                                                  mask_min_nn_count=cfg.mesh_min_nn, out_torch=True)
    out = dict(before)
    out.update(gg.decoder_arrays(dec))
    out.update(cfg=gg.cfg_json(cfg), grid=grid.numpy(), bs=20000, mesh_min_nn=cfg.mesh_min_nn,
               out_sdf=sdf_pred.numpy(), out_mask=mc_mask.numpy())
    path = os.path.join(gg.GOLDEN_DIR, "caller_mesher.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "grid", grid.shape[0], "masked-in", int(mc_mask.sum()))


def main():
    ref = ref_loader.load()
    iekf_case(ref)
    mesher_case(ref)


if __name__ == "__main__":
    main()
