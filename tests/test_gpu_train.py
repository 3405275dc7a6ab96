"""GPU parity of the training step (loss, gradients, Adam, side effects) against fixtures that
were produced by driving the reference's own Mapper.mapping() (oracle/gen_golden.py)."""
import pytest
import torch

import golden_io as gio
import helpers as hp

pytestmark = pytest.mark.gpu

L1_CASES = [n for n in gio.names("train") if "l2h32" not in n]
L2_CASES = [n for n in gio.names("train") if "l2h32" in n]
ALL_CASES = gio.names("train")


class _Dataset:
    lose_track = False
    stop_status = False
    processed_frame = 0
    gt_pose_provided = True
    pgo_poses = None
    static_mask = None


def _setup(name):
    fx = gio.load("train", name)
    m = gio.oracle_map(fx)
    cfg = hp.product_config(m.cfg)
    npm = hp.product_map(m)
    frozen = bool(fx["freeze_decoder"])
    dec = hp.product_decoder(m.cfg, gio.decoder_params(fx))
    if frozen:
        from clid_slam_b200.utils.tools import freeze_model

        freeze_model(dec)
    return fx, m, cfg, npm, dec, frozen


def _batch(fx, it):
    return (gio.t(fx["batch_x"][it]).cuda(), gio.t(fx["batch_label"][it]).cuda(),
            gio.t(fx["batch_ts"][it]).cuda(), gio.t(fx["batch_weight"][it]).cuda())


def _check_final_state(fx, npm, dec):
    gio.assert_close(npm.local_geo_features, fx["after_local_features"], 1e-3, 1e-5, "features after Adam", 2e-3)
    gio.assert_close(npm.local_point_certainties, fx["after_local_certainties"], 1e-5, 1e-5, "certainties")
    assert torch.equal(npm.local_point_ts_update.cpu(), gio.t(fx["after_local_ts_update"]))
    after = gio.decoder_params(fx, prefix="after_dec_", requires_grad=False)
    for a, b in zip(dec.flat_parameters(), after):
        gio.assert_close(a, b, 1e-3, 1e-5, "decoder after Adam", 2e-3)


# rows: decoder-gradient rows to scratch + reduction kernel (default); infold: the warps fold them inside
# the one kernel; headers: brick index without the neighbourhood lines; three-kernels: forward / loss / backward launches
@pytest.mark.parametrize("variant", ["rows", "infold", "headers", "three-kernels"])
@pytest.mark.parametrize("name", L1_CASES)
def test_fused_training_matches_reference(name, variant, monkeypatch):
    from clid_slam_b200.ops import bricks as bk
    from clid_slam_b200.ops.train import FusedTrainer

    monkeypatch.setattr(bk, "USE_HOOD", variant != "headers")
    fx, m, cfg, npm, dec, frozen = _setup(name)
    trainer = FusedTrainer(cfg, npm, dec)
    trainer.single_kernel = variant != "three-kernels"
    trainer.use_scratch = variant != "infold"
    for it in range(int(fx["n_iters"])):
        x, label, ts, weight = _batch(fx, it)
        loss = trainer.iteration(x, label, ts, weight, apply_step=False)
        gio.assert_close(loss[0], fx["loss_total"][it], hp.LOSS_RTOL, 0, f"total loss it{it}")
        gio.assert_close(loss[1], fx["loss_bce"][it], 1e-4, 0, f"bce loss it{it}")
        gio.assert_close(loss[2], fx["loss_eikonal"][it], 1e-4, 0, f"eikonal loss it{it}")
        gio.assert_close(trainer.feat_grad, fx["feat_grads"][it], 1e-3, 2e-9, f"dL/dfeatures it{it}", 1e-3)
        if not frozen:
            flat = torch.cat([gio.t(fx[f"dec_grad_it{it}_{j}"]).flatten() for j in range(4)])
            gio.assert_close(trainer.dec_grad, flat, 1e-3, 1e-7, f"dL/ddecoder it{it}")
        trainer.adam_step()
    _check_final_state(fx, npm, dec)


@pytest.mark.parametrize("name", L2_CASES)
def test_fused_training_of_the_two_level_decoder_matches_reference(name):
    """32 x 2 decoder (BASELINE configs[0]) through train_fused_kernel<32, 2> + decoder_grad_l2_kernel: losses,
    dL/dfeatures, dL/d{W0, b0, W1, b1, wout, bout} per iteration and the post-Adam state against the reference."""
    from clid_slam_b200.ops.train import FusedTrainer

    fx, m, cfg, npm, dec, frozen = _setup(name)
    assert len(dec.layers) == 2
    trainer = FusedTrainer(cfg, npm, dec)
    for it in range(int(fx["n_iters"])):
        x, label, ts, weight = _batch(fx, it)
        loss = trainer.iteration(x, label, ts, weight, apply_step=False)
        gio.assert_close(loss[0], fx["loss_total"][it], hp.LOSS_RTOL, 0, f"total loss it{it}")
        gio.assert_close(loss[1], fx["loss_bce"][it], 1e-4, 0, f"bce loss it{it}")
        gio.assert_close(loss[2], fx["loss_eikonal"][it], 1e-4, 0, f"eikonal loss it{it}")
        gio.assert_close(trainer.feat_grad, fx["feat_grads"][it], 1e-3, 2e-9, f"dL/dfeatures it{it}", 1e-3)
        if not frozen:
            flat = torch.cat([gio.t(fx[f"dec_grad_it{it}_{j}"]).flatten() for j in range(6)])
            gio.assert_close(trainer.dec_grad, flat, 1e-3, 1e-7, f"dL/ddecoder it{it}", 1e-3)
        trainer.adam_step()
    _check_final_state(fx, npm, dec)


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "unfused"])
@pytest.mark.parametrize("name", ALL_CASES)
def test_mapper_mapping_matches_reference(name, fused):
    """Mapper.mapping() end to end on the recorded batches: fused kernels, and the unfused path
    (query_feature autograd incl. double backward + torch decoder + torch Adam)."""
    from clid_slam_b200.ops import train as tr
    from clid_slam_b200.utils.mapper import Mapper

    fx, m, cfg, npm, dec, frozen = _setup(name)
    if fused and tr.supported(cfg, dec) is not None:
        pytest.skip("configuration not covered by the fused path: " + tr.supported(cfg, dec))
    n_iters = int(fx["n_iters"])
    cfg.bs = fx["batch_x"].shape[1]
    mapper = Mapper(cfg, _Dataset(), npm, None, dec)
    mapper.use_fused = fused
    mapper.adaptive_iter_offset = 0
    feed = iter(range(n_iters))

    def replay(global_coord=False):
        x, label, ts, weight = _batch(fx, next(feed))
        return x, label, ts, None, None, None, weight

    mapper.get_batch = replay
    mapper.mapping(n_iters)
    assert mapper.total_iter == n_iters
    losses = mapper.last_losses.cpu()
    gio.assert_close(losses[:, 0], fx["loss_total"], hp.LOSS_RTOL, 0, "total loss")
    gio.assert_close(losses[:, 1], fx["loss_bce"], 1e-4, 0, "bce loss")
    gio.assert_close(losses[:, 2], fx["loss_eikonal"], 1e-4, 0, "eikonal loss")
    _check_final_state(fx, npm, dec)
    gio.assert_close(npm.geo_features, fx["after_features"], 1e-3, 1e-5, "global features after write-back", 2e-3)


def test_unfused_gradients_match_reference():
    """Analytic mode through torch autograd: first- and second-order kernels of query_feature."""
    from clid_slam_b200.utils.loss import sdf_bce_loss
    from clid_slam_b200.utils.tools import get_gradient

    fx, m, cfg, npm, dec, _ = _setup("analytic_l1h64")
    x, label, ts, weight = _batch(fx, 0)
    x.requires_grad_(True)
    z, _, _, _, _ = npm.query_feature(x, ts)
    sdf = dec.sdf(z)
    g = get_gradient(x, sdf)
    loss = sdf_bce_loss(sdf, label, dec.sdf_scale, weight.abs(), cfg.loss_weight_on)
    loss = loss + cfg.weight_e * ((g.norm(2, dim=-1) - 1.0) ** 2).mean()
    loss.backward()
    gio.assert_close(loss, fx["loss_total"][0], 1e-4, 0, "loss")
    gio.assert_close(npm.local_geo_features.grad, fx["feat_grads"][0], 1e-3, 2e-9, "dL/dfeatures", 1e-3)
    for j, p in enumerate(dec.flat_parameters()):
        gio.assert_close(p.grad, fx[f"dec_grad_it0_{j}"], 1e-3, 1e-7, f"dL/ddecoder[{j}]")


@pytest.mark.parametrize("name", L1_CASES)
def test_step_pipeline_matches_reference(name):
    """The graphed iteration (CUDA graph per input buffer, device-side Adam step counter, host batches
    staged on the copy stream) reproduces the reference's losses and post-Adam state."""
    from clid_slam_b200.ops.train import FusedTrainer, StepPipeline

    fx, m, cfg, npm, dec, frozen = _setup(name)
    trainer = FusedTrainer(cfg, npm, dec)
    n = fx["batch_x"].shape[1]
    pipe = StepPipeline(trainer, n)
    n_iters = int(fx["n_iters"])

    def host_batch(it):
        x, label, ts, weight = _batch(fx, it)
        return tuple(t.cpu().pin_memory() for t in (x, label, weight, ts.to(torch.int32)))

    pipe.stage(0, host_batch(0))
    for it in range(n_iters):
        if it + 1 < n_iters:
            pipe.stage((it + 1) % 2, host_batch(it + 1))
        loss = pipe.run(it % 2).cpu()
        gio.assert_close(loss[0], fx["loss_total"][it], hp.LOSS_RTOL, 0, f"total loss it{it}")
        gio.assert_close(loss[1], fx["loss_bce"][it], 1e-4, 0, f"bce loss it{it}")
        gio.assert_close(loss[2], fx["loss_eikonal"][it], 1e-4, 0, f"eikonal loss it{it}")
    assert trainer.step == n_iters
    assert int(trainer.step_state[0].item()) == n_iters
    _check_final_state(fx, npm, dec)


def test_decoder_grad_reduce_matches_infold():
    """Rows + dense reduction kernel against the in-kernel fold on a batch large enough for many tiles."""
    from clid_slam_b200.ops.train import FusedTrainer
    import oracle.sdf_oracle as oc

    cfg_o = oc.OracleConfig(buffer_size=2_000_003, local_map_radius=80.0)
    mo, params, gen = hp.build_oracle_world(200, 2, seed=5, cfg=cfg_o)
    x, label, weight, ts = oc.sample_batch(mo.points, 40000, gen)
    grads = []
    for use_scratch in (True, False):
        npm = hp.product_map(mo)
        dec = hp.product_decoder(cfg_o, params)
        trainer = FusedTrainer(hp.product_config(cfg_o), npm, dec)
        trainer.use_scratch = use_scratch
        trainer.iteration(x.cuda(), label.cuda(), ts.cuda(), weight.cuda(), apply_step=False)
        grads.append((trainer.dec_grad.clone(), trainer.feat_grad.clone()))
    gio.assert_close(grads[0][0], grads[1][0], 1e-4, 1e-7, "decoder gradients: rows vs in-kernel fold")
    gio.assert_close(grads[0][1], grads[1][1], 1e-4, 1e-9, "feature gradients")


@pytest.mark.parametrize("variant", ["rows", "headers"])
@pytest.mark.parametrize("hidden,numerical,leaky", [(32, False, False), (128, False, False), (32, True, False),
                                                     (128, True, True), (64, False, True)])
def test_training_gradients_match_oracle_other_widths(hidden, numerical, leaky, variant, monkeypatch):
    """Decoder widths / activation the reference fixtures do not cover: loss, dL/dfeatures and dL/ddecoder of
    one iteration against torch autograd through the oracle (double backward in analytic mode)."""
    import oracle.sdf_oracle as oc
    from clid_slam_b200.ops import bricks as bk
    from clid_slam_b200.ops.train import FusedTrainer

    monkeypatch.setattr(bk, "USE_HOOD", variant != "headers")
    cfg_o = oc.OracleConfig(buffer_size=2_000_003, local_map_radius=80.0, geo_mlp_hidden_dim=hidden,
                            numerical_grad=numerical, gradient_decimation=10 if numerical else 1, mlp_leaky_relu=leaky)
    mo, params, gen = hp.build_oracle_world(120, 2, seed=7, cfg=cfg_o)
    x, label, weight, ts = oc.sample_batch(mo.points, 6000, gen)
    npm = hp.product_map(mo)
    dec = hp.product_decoder(cfg_o, params)

    feats = mo.local_features
    feats.requires_grad_(True)
    for p in params:
        p.requires_grad_(True)
    total, l_bce, l_eik, _, _ = oc.training_loss(mo, params, x, label, ts, weight)
    grads = torch.autograd.grad(total, [feats] + list(params))

    trainer = FusedTrainer(hp.product_config(cfg_o), npm, dec)
    loss = trainer.iteration(x.cuda(), label.cuda(), ts.cuda(), weight.cuda(), apply_step=False)
    gio.assert_close(loss[0], total.detach(), hp.LOSS_RTOL, 0, "total loss")
    gio.assert_close(loss[1], l_bce.detach(), 1e-4, 0, "bce loss")
    gio.assert_close(loss[2], l_eik.detach(), 1e-4, 0, "eikonal loss")
    gio.assert_close(trainer.feat_grad, grads[0], 1e-3, 2e-9, "dL/dfeatures", 1e-3)
    flat = torch.cat([g.flatten() for g in grads[1:]])
    gio.assert_close(trainer.dec_grad, flat, 1e-3, 1e-7, "dL/ddecoder")
