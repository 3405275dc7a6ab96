"""TEST INFRASTRUCTURE ONLY (see oracle/sdf_oracle.py): carry the state of a NeuralPoints-like object -- the
reference's or this repository's mirror, on any device -- into the CPU oracle, and the parity gate built on
it: the same inputs through the CUDA path and through the oracle, compared with SURVEY.md 8(d)'s tolerances.

Used by oracle/gen_golden.py, tests/ and the untimed parity block of bench.py (checker, never measured).
"""
from __future__ import annotations

from typing import Dict

import torch

from oracle import sdf_oracle as oc


def oracle_config_from(cfg) -> oc.OracleConfig:
    o = oc.OracleConfig()
    for name in o.__dataclass_fields__:
        if hasattr(cfg, name):
            setattr(o, name, getattr(cfg, name))
    return o


def oracle_map_from(npm, ocfg: oc.OracleConfig) -> oc.OracleMap:
    def c(t):
        return t.detach().cpu().clone()

    m = oc.OracleMap(
        cfg=ocfg, table=c(npm.buffer_pt_index), points=c(npm.neural_points), ts_create=c(npm.point_ts_create),
        ts_update=c(npm.point_ts_update), certainties=c(npm.point_certainties), features=c(npm.geo_features),
        travel_dist=c(npm.travel_dist), cur_ts=int(npm.cur_ts), reboot_ts=int(npm.reboot_ts),
    )
    m.offsets = c(npm.neighbor_dx)
    m.max_valid_dist2 = float(npm.max_valid_dist2)
    m.local_points = c(npm.local_neural_points)
    m.local_features = c(npm.local_geo_features).requires_grad_(True)
    m.local_certainties = c(npm.local_point_certainties)
    m.local_ts_update = c(npm.local_point_ts_update)
    m.local_mask = c(npm.local_mask)
    m.global2local = c(npm.global2local)
    return m


# SURVEY.md 8(d) "parity gate run in the same job" / BASELINE.json north_star
TOL = {"sdf_rel": 1e-4, "grad_rel": 1e-4, "loss_rel": 1e-3, "feat_grad_rel": 1e-3, "dec_grad_rel": 1e-3}


def parity_gate(npm, dec, cfg, x, label, weight, ts, n_slice: int = 16384) -> Dict:
    """Forward + gradient and one training iteration (no optimiser step) of the CUDA path against the oracle
    on the first n_slice samples of the batch, on the CURRENT map state.  Side effects of the training
    iteration on the map (certainty / ts) are undone.  Returns the max errors and ok = all within TOL."""
    from clid_slam_b200 import fused
    from clid_slam_b200.ops.train import FusedTrainer

    n = min(n_slice, x.shape[0])
    xs, ls, ws, tss = x[:n].contiguous(), label[:n].contiguous(), weight[:n].contiguous(), ts[:n].contiguous()
    ocfg = oracle_config_from(cfg)
    m = oracle_map_from(npm, ocfg)
    params = [p.detach().cpu().clone().requires_grad_(True) for p in dec.flat_parameters()]

    # ---- inference forward + gradient
    sdf_g, grad_g, nn_g, _ = fused.sdf_and_gradient(npm, dec, xs)
    xo = xs.detach().cpu().clone().requires_grad_(True)
    z, _, nn_o, _ = oc.query_feature(m, xo, None, training_mode=False, query_locally=True)
    sdf_o = oc.decoder_sdf(params, z, ocfg.sdf_scale, leaky=ocfg.mlp_leaky_relu)
    grad_o = oc.sdf_gradient(xo, sdf_o)
    sdf_o, grad_o = sdf_o.detach(), grad_o.detach()
    floor_sdf = 1e-3 * ocfg.sdf_scale
    sdf_rel = ((sdf_g.cpu() - sdf_o).abs() / sdf_o.abs().clamp_min(floor_sdf)).max().item()
    grad_rel = ((grad_g.cpu() - grad_o).norm(dim=-1) / grad_o.norm(dim=-1).clamp_min(1e-3)).max().item()
    nn_equal = bool(torch.equal(nn_g.cpu().long(), nn_o.long()))

    # ---- one training iteration, gradients only
    saved = (npm.local_point_certainties.clone(), npm.local_point_ts_update.clone())
    trainer = FusedTrainer(cfg, npm, dec)
    loss_g = trainer.iteration(xs, ls, tss, ws, apply_step=False).cpu()
    fg_g = trainer.feat_grad.detach().cpu().clone()
    dg_g = None if trainer.dec_grad is None else trainer.dec_grad.detach().cpu().clone()
    npm.local_point_certainties.copy_(saved[0])
    npm.local_point_ts_update.copy_(saved[1])
    del trainer

    total, l_bce, l_eik, _, _ = oc.training_loss(m, params, xs.cpu(), ls.cpu(), tss.cpu(), ws.cpu())
    grads = torch.autograd.grad(total, [m.local_features] + params)
    loss_o = torch.stack((total.detach(), l_bce.detach(), l_eik.detach()))
    loss_rel = ((loss_g - loss_o).abs() / loss_o.abs().clamp_min(1e-12)).max().item()
    fg_o = grads[0]
    # atomics reorder fp32 sums (and the +- eps evaluations of the numerical mode nearly cancel inside an entry):
    # the error of a gradient tensor is measured against its largest entry (max-norm relative error)
    feat_rel = ((fg_g - fg_o).abs().max() / fg_o.abs().max().clamp_min(1e-30)).item()
    dec_rel = None
    if dg_g is not None:
        dg_o = torch.cat([g.flatten() for g in grads[1:]])
        dec_rel = ((dg_g - dg_o).abs().max() / dg_o.abs().max().clamp_min(1e-30)).item()

    out = {"samples": int(n), "neural_points": int(npm.count()), "nn_counts_equal": nn_equal, "sdf_max_rel": sdf_rel,
           "grad_max_rel": grad_rel, "loss_max_rel": loss_rel, "feat_grad_max_rel": feat_rel, "dec_grad_max_rel": dec_rel,
           "loss_cuda": [float(v) for v in loss_g], "loss_oracle": [float(v) for v in loss_o], "tolerances": TOL}
    out["ok"] = bool(nn_equal and sdf_rel <= TOL["sdf_rel"] and grad_rel <= TOL["grad_rel"] and loss_rel <= TOL["loss_rel"]
                     and feat_rel <= TOL["feat_grad_rel"] and (dec_rel is None or dec_rel <= TOL["dec_grad_rel"]))
    return out
