"""Inference callers of the hot path against fixtures produced by the reference's own code
(oracle/gen_golden_callers.py): IEKFOM.h_model / the normal equations of update_iterated
(utils/error_state_iekf.py:176-264, :303-309) and Mesher.query_points (utils/mesher.py:38-163)."""
import numpy as np
import pytest
import torch

import golden_io as gio
import helpers as hp

pytestmark = pytest.mark.gpu


def _world(fx):
    m = gio.oracle_map(fx)
    npm = hp.product_map(m)
    dec = hp.product_decoder(m.cfg, gio.decoder_params(fx))
    cfg = hp.product_config(m.cfg)
    return m, npm, dec, cfg


def test_h_model_and_normal_equations_match_the_reference_tracker():
    from clid_slam_b200.utils import registration as reg

    fx = gio.load("caller", "iekf")
    m, npm, dec, cfg = _world(fx)
    cfg.track_mask_query_nn_k = int(fx["track_mask_query_nn_k"])
    cfg.reg_min_grad_norm, cfg.reg_max_grad_norm = float(fx["reg_min_grad_norm"]), float(fx["reg_max_grad_norm"])
    cfg.tran_dtype = torch.float64
    pc_imu = gio.t(fx["pc_imu"]).cuda()
    rot, pos = gio.t(fx["rot"]), gio.t(fx["pos"])

    z, H, valid_points, R_inv = reg.h_model(cfg, npm, dec, pc_imu, rot, pos)
    assert z.dtype == torch.float64 and H.shape[1] == 18
    assert z.shape[0] == int(fx["out_valid"].sum()), "the same scan points survive the validity tests"
    gio.assert_close(valid_points, fx["out_valid_points"], 1e-6, 1e-6, "valid points")
    gio.assert_close(z, fx["out_z"], hp.SDF_RTOL, hp.SDF_ATOL, "sdf residual")
    # H[:, 0:3] = -(grad^T R [p]x): the gradient tolerance (1e-4 rel + 5e-6 abs) scaled by the lever arm |p| (< 25 m)
    gio.assert_close(H[:, 3:6], fx["out_H6"][:, 3:6], hp.GRAD_RTOL, hp.GRAD_ATOL, "measurement Jacobian (translation)")
    gio.assert_close(H[:, 0:3], fx["out_H6"][:, 0:3], hp.GRAD_RTOL, 25 * hp.GRAD_ATOL + 2e-4, "measurement Jacobian (rotation)")
    assert float(H[:, 6:].abs().max()) == 0.0
    gio.assert_close(R_inv, fx["out_R_inv"], 1e-4, 1e-6, "R_inv")

    S, g, n_valid = reg.normal_equations(cfg, npm, dec, pc_imu, rot, pos)
    assert int(n_valid.item()) == z.shape[0]
    S_ref, g_ref = gio.t(fx["out_S"]), gio.t(fx["out_g"])
    scale = S_ref.abs().max()
    assert float((S.cpu() - S_ref).abs().max() / scale) < 1e-5, "H^T R^-1 H"
    assert float((g.cpu() - g_ref).abs().max() / g_ref.abs().max()) < 1e-4, "H^T R^-1 z"
    # what update_iterated does with them (:303-309): K z == K_front g, K H == K_front S
    P_inv = torch.eye(18, dtype=torch.float64)
    Hc, Rc, zc = H.cpu(), R_inv.cpu(), z.cpu()
    K_front = torch.linalg.inv(Hc.T * Rc @ Hc + P_inv)
    K = K_front @ (Hc.T * Rc)
    Kf2 = torch.linalg.inv(S.cpu() + P_inv)
    torch.testing.assert_close(Kf2 @ g.cpu(), K @ zc, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(Kf2 @ S.cpu(), K @ Hc, rtol=1e-4, atol=1e-7)


def test_mesher_query_points_matches_the_reference_mesher():
    from clid_slam_b200.utils.mesher import Mesher

    fx = gio.load("caller", "mesher")
    m, npm, dec, cfg = _world(fx)
    mesher = Mesher(cfg, npm, {"sdf": dec, "semantic": None, "color": None})
    grid = gio.t(fx["grid"])
    sdf, sem, col, mask = mesher.query_points(grid, int(fx["bs"]), query_sdf=True, query_mask=True, query_locally=False,
                                              mask_min_nn_count=int(fx["mesh_min_nn"]), out_torch=True)
    assert sem is None and col is None
    assert torch.equal(mask, gio.t(fx["out_mask"])), "marching-cubes mask must be exact (candidate counts)"
    gio.assert_close(sdf, fx["out_sdf"], hp.SDF_RTOL, hp.SDF_ATOL, "grid sdf")
    sdf_np, _, _, mask_np = mesher.query_points(grid, 7000, mask_min_nn_count=int(fx["mesh_min_nn"]))
    assert isinstance(sdf_np, np.ndarray) and sdf_np.dtype == np.float64 and mask_np.shape == sdf_np.shape
    np.testing.assert_allclose(sdf_np, sdf.numpy().astype(np.float64), rtol=0, atol=0)


def test_native_region_sdf_matches_the_torch_mirror():
    """clid_region_sdf against the torch restatement of LocalPointCloudMap.region_specific_sdf_estimation (which the
    CPU suite pins to the reference's fixtures): same surface mask, same |sdf| labels; the plane-acceptance tests
    (sigma ratio <= 0.2, residual <= 0.1 m) may flip for a handful of borderline neighbourhoods (fp32 LAPACK SVD
    there, double Jacobi here)."""
    from clid_slam_b200.config import ncd128
    from clid_slam_b200.model.local_point_cloud_map import LocalPointCloudMap

    cfg = ncd128()
    cfg.device = "cuda"
    cfg.local_buffer_size = 2_000_003
    lpm = LocalPointCloudMap(cfg)
    gen = torch.Generator(device="cuda").manual_seed(11)
    xy = (torch.rand(60000, 2, generator=gen, device="cuda") - 0.5) * 60
    floor = torch.cat((xy, 0.4 * torch.sin(xy[:, :1] / 4) + 0.01 * torch.randn(60000, 1, generator=gen, device="cuda")), 1)
    wall = torch.cat((torch.full((8000, 1), 9.0, device="cuda"), (torch.rand(8000, 2, generator=gen, device="cuda") - 0.5) * 12), 1)
    lpm.update_map(torch.zeros(3, device="cuda"), torch.cat((floor, wall), 0))
    assert lpm.local_point_cloud_map.shape[0] > 20000
    pick = torch.randint(0, floor.shape[0], (50000,), generator=gen, device="cuda")
    q = torch.cat((floor[pick] + 0.15 * torch.randn(50000, 3, generator=gen, device="cuda"),
                   (torch.rand(3000, 3, generator=gen, device="cuda") - 0.5) * 100,       # mostly far from everything
                   -floor[pick[:2000]]), 0).contiguous()                                    # negative cells too
    sdf_n, mask_n = lpm._region_sdf_native(q)
    sdf_t, mask_t = lpm._region_sdf_torch(q)
    assert torch.equal(mask_n, mask_t)
    bad = ((sdf_n - sdf_t).abs() > 1e-5 + 1e-4 * sdf_t.abs()).double().mean().item()
    assert bad < 2e-3, f"{bad:.2e} of the labels differ"
    assert 0.3 < float(mask_n.float().mean()) < 1.0


@pytest.mark.parametrize("locally", [True, False], ids=["local", "global"])
def test_native_brick_build_equals_the_torch_build(locally):
    """clid_brick_keep / keys / fill (one read-back) against the torch-op build of the same index: identical records,
    headers (mask, count, first record of every occupied brick) and neighbourhood lines, with a time-filtered
    second frame and hash collisions in play."""
    import oracle.sdf_oracle as oc
    from clid_slam_b200.ops import bricks as B

    cfg = oc.OracleConfig(buffer_size=400_009, local_map_radius=30.0)
    m, params, gen = hp.build_oracle_world(90, 2, seed=3, cfg=cfg)
    m.travel_dist = torch.tensor([0.0, 500.0])
    pts2 = oc.wavy_sheets(40, 1, cfg.voxel_size_m, gen) + torch.tensor([3.0, 2.0, 0.2])
    oc.map_insert(m, pts2, torch.zeros(3), 1, generator=gen)
    oc.reset_local_window(m, torch.zeros(3), 0)
    npm = hp.product_map(m)
    off = npm.neighbor_dx.detach().cpu()
    a = B._build_native(npm, locally, off, 2, 2)
    b = B._build_torch(npm, locally, off, 2, 2)
    assert a is not None and b is not None
    assert list(a.struct.origin) == list(b.struct.origin) and list(a.struct.dims) == list(b.struct.dims)
    assert torch.equal(a.records, b.records)
    ha, hb = a.headers.view(-1, 4), b.headers.view(-1, 4)
    assert torch.equal(ha[:, [0, 1, 3]], hb[:, [0, 1, 3]]), "occupancy masks and counts"
    occupied = hb[:, 3] > 0
    assert torch.equal(ha[occupied, 2], hb[occupied, 2]), "first record of every occupied brick"
    assert torch.equal(a.hood.view(-1, 32)[:, :16], b.hood.view(-1, 32)[:, :16])
    # first-record words of the neighbourhood lines matter for occupied bricks only
    occ8 = (b.hood.view(-1, 32)[:, 0:16:2] != 0) | (b.hood.view(-1, 32)[:, 1:16:2] != 0)
    assert torch.equal(a.hood.view(-1, 32)[:, 16:24][occ8], b.hood.view(-1, 32)[:, 16:24][occ8])
