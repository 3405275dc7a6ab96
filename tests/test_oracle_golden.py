"""Pin the CPU oracle (oracle/sdf_oracle.py) against fixtures produced by the UNMODIFIED
reference (oracle/gen_golden.py).  CPU only."""
import pytest
import torch

import golden_io as gio
from oracle import sdf_oracle as oc


@pytest.mark.parametrize("name", gio.names("query"))
def test_query_matches_reference(name):
    fx = gio.load("query", name)
    m = gio.oracle_map(fx)
    cfg = m.cfg
    training, locally = bool(fx["training_mode"]), bool(fx["query_locally"])
    x = gio.t(fx["x"]).requires_grad_(True)
    ts = gio.t(fx["ts"]) if bool(fx["has_ts"]) else None
    params = gio.decoder_params(fx)

    d2, idx = oc.radius_search(m, x.detach(), cfg.temporal_local_map_on and locally)
    assert torch.equal(idx[:256], gio.t(fx["out_idx"]))
    gio.assert_close(d2[:256], fx["out_dist2"], 0, 0, "dist2")
    gio.assert_close(oc.query_certainty(m, x.detach()), fx["out_query_certainty"], 0, 0, "query_certainty")

    z, w, nn, cert = oc.query_feature(m, x, ts, training, locally)
    sdf = oc.decoder_sdf(params, z, cfg.sdf_scale, cfg.mlp_leaky_relu)
    grad = oc.sdf_gradient(x, sdf)
    assert torch.equal(nn, gio.t(fx["out_nn"]))
    # same torch build => same bits; a different torch/BLAS may differ in the last ulps
    gio.assert_close(z, fx["out_z"], 1e-6, 1e-7, "z")
    gio.assert_close(w, fx["out_w"], 1e-6, 1e-9, "weights")
    gio.assert_close(cert, fx["out_certainty"], 1e-6, 1e-7, "queried certainty")
    gio.assert_close(sdf, fx["out_sdf"], 1e-5, 1e-8, "sdf")
    gio.assert_close(grad, fx["out_grad"], 1e-5, 1e-7, "grad")
    if locally:
        gio.assert_close(m.local_certainties, fx["after_local_certainties"], 1e-6, 1e-6, "certainty side effect")
        assert torch.equal(m.local_ts_update, gio.t(fx["after_local_ts_update"]))
    else:
        gio.assert_close(m.certainties, fx["after_certainties"], 1e-6, 1e-6, "certainty side effect")


@pytest.mark.parametrize("name", gio.names("train"))
def test_training_matches_reference(name):
    fx = gio.load("train", name)
    m = gio.oracle_map(fx)
    cfg = m.cfg
    frozen = bool(fx["freeze_decoder"])
    params = gio.decoder_params(fx, requires_grad=not frozen)
    opt = oc.make_adam(cfg, [m.local_features], None if frozen else params)
    for it in range(int(fx["n_iters"])):
        x, label = gio.t(fx["batch_x"][it]), gio.t(fx["batch_label"][it])
        ts, weight = gio.t(fx["batch_ts"][it]), gio.t(fx["batch_weight"][it])
        total, l_bce, l_eik = oc.train_iteration(m, params, opt, x, label, ts, weight)
        gio.assert_close(total, fx["loss_total"][it], 1e-5, 0, f"total loss it{it}")
        gio.assert_close(l_bce, fx["loss_bce"][it], 1e-5, 0, f"bce it{it}")
        gio.assert_close(l_eik, fx["loss_eikonal"][it], 1e-5, 0, f"eikonal it{it}")
        gio.assert_close(m.local_features.grad, fx["feat_grads"][it], 1e-4, 1e-9, f"feature grad it{it}", 1e-3)
        if not frozen:
            for j, p in enumerate(params):
                gio.assert_close(p.grad, fx[f"dec_grad_it{it}_{j}"], 1e-4, 1e-8, f"decoder grad {j} it{it}")
    oc.write_back_local(m)
    gio.assert_close(m.local_features, fx["after_local_features"], 1e-3, 1e-5, "features after Adam", 2e-3)
    gio.assert_close(m.features, fx["after_features"], 1e-3, 1e-5, "global features", 2e-3)
    gio.assert_close(m.local_certainties, fx["after_local_certainties"], 1e-5, 1e-6, "certainties")
    assert torch.equal(m.local_ts_update, gio.t(fx["after_local_ts_update"]))
    after = gio.decoder_params(fx, prefix="after_dec_", requires_grad=False)
    for a, b in zip(params, after):
        gio.assert_close(a, b, 1e-3, 1e-5, "decoder after Adam", 2e-3)


@pytest.mark.parametrize("name", gio.names("map"))
def test_map_insert_matches_reference(name):
    fx = gio.load("map", name)
    cfg = gio.config_of(fx)
    m = oc.empty_map(cfg)
    m.travel_dist = gio.t(fx["travel_dist"])
    for i in range(int(fx["n_frames"])):
        ratio = oc.map_insert(m, gio.t(fx[f"frame{i}_points"]), gio.t(fx[f"frame{i}_sensor"]), int(fx[f"frame{i}_ts"]))
        pre = f"frame{i}_map_"
        assert ratio == float(fx[f"frame{i}_ratio"])
        assert torch.equal(m.table, gio.dense_table(fx, pre))
        assert torch.equal(m.points, gio.t(fx[pre + "points"]))
        assert torch.equal(m.ts_create, gio.t(fx[pre + "ts_create"]))
        assert torch.equal(m.local_mask, gio.t(fx[pre + "local_mask"]))
        assert torch.equal(m.global2local, gio.t(fx[pre + "global2local"]))
        assert torch.equal(m.local_points, gio.t(fx[pre + "local_points"]))


def test_neighborhood_sizes():
    # counts listed at neural_points.py:955-965
    assert oc.neighborhood_offsets(2, 0.5).shape[0] == 81
    assert oc.neighborhood_offsets(2, 0.2).shape[0] == 33
    assert oc.neighborhood_offsets(1, 0.2).shape[0] == 7
    assert oc.neighborhood_offsets(1, 0.0).shape[0] == 1
    assert oc.neighborhood_offsets(2, 1.0).shape[0] == 93


def test_negative_hash_wraps_like_modulus():
    cells = torch.tensor([[-5, -7, -11], [3, -2, 9], [-100000, 4, -6]])
    b = 1000003
    h = oc.voxel_hash(cells, b)
    table = torch.arange(b)
    mathematical = ((cells * torch.tensor(oc.PRIMES_NEURAL_POINTS)).sum(-1)) % b
    assert torch.equal(table[h], mathematical)
