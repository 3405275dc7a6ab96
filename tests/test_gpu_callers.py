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
