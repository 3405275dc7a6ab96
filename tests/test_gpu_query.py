"""GPU parity of the query path (through the C ABI) against the reference-generated golden
fixtures and against the CPU oracle on seeded inputs."""
import pytest
import torch

import golden_io as gio
import helpers as hp
from oracle import sdf_oracle as oc

pytestmark = pytest.mark.gpu


def _cuda(t):
    return None if t is None else t.cuda()


@pytest.mark.parametrize("name", gio.names("query"))
def test_query_feature_matches_reference_fixture(name):
    from clid_slam_b200.utils.tools import get_gradient

    fx = gio.load("query", name)
    m = gio.oracle_map(fx)
    npm = hp.product_map(m)
    dec = hp.product_decoder(m.cfg, gio.decoder_params(fx))
    training, locally = bool(fx["training_mode"]), bool(fx["query_locally"])
    x = gio.t(fx["x"]).cuda().requires_grad_(True)
    ts = _cuda(gio.t(fx["ts"])) if bool(fx["has_ts"]) else None

    d2, idx = npm.radius_neighborhood_search(x.detach(), npm.temporal_local_map_on and locally)
    assert torch.equal(idx[:256].cpu(), gio.t(fx["out_idx"]))
    gio.assert_close(d2[:256], fx["out_dist2"], 0, 0, "dist2 (bit exact)")
    gio.assert_close(npm.query_certainty(x.detach()), fx["out_query_certainty"], 0, 0, "query_certainty")

    z, _, w, nn, cert = npm.query_feature(x, ts, training_mode=training, query_locally=locally)
    sdf = dec.sdf(z)
    grad = get_gradient(x, sdf)
    assert nn.dtype == torch.int64 and torch.equal(nn.cpu(), gio.t(fx["out_nn"])), "nn_counts must be exact"
    assert w.shape == (x.shape[0], m.cfg.query_nn_k, 1)
    gio.assert_close(z, fx["out_z"], hp.Z_RTOL, hp.Z_ATOL, "z")
    gio.assert_close(w, fx["out_w"], 1e-5, 1e-7, "weights")
    gio.assert_close(cert, fx["out_certainty"], 1e-5, 1e-6, "queried certainty")
    gio.assert_close(sdf, fx["out_sdf"], hp.SDF_RTOL, hp.SDF_ATOL, "sdf (torch decoder on our z)")
    gio.assert_close(grad, fx["out_grad"], hp.GRAD_RTOL, hp.GRAD_ATOL, "grad (autograd through our backward)")
    if locally:
        gio.assert_close(npm.local_point_certainties, fx["after_local_certainties"], 1e-5, 1e-5, "certainty side effect")
        assert torch.equal(npm.local_point_ts_update.cpu(), gio.t(fx["after_local_ts_update"]))
    else:
        gio.assert_close(npm.point_certainties, fx["after_certainties"], 1e-5, 1e-5, "certainty side effect")


def _kernel_family(monkeypatch, family):
    """hashed: reference hash table; bricks: brick index with the 128-byte neighbourhood lines (the default);
    headers: brick index read through its eight 16-byte headers; bricks-tc: the brick index with 64 x 1 decoders
    on the tensor cores (tcgen05).  Returns the use_bricks argument."""
    from clid_slam_b200.ops import bricks as b
    from clid_slam_b200.ops import query as qy

    monkeypatch.setattr(b, "USE_HOOD", family != "headers")
    monkeypatch.setattr(qy, "TC_DECODER", family == "bricks-tc")
    return family != "hashed"


FAMILIES = ["hashed", "bricks", "headers", "bricks-tc"]


@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("name", gio.names("query"))
def test_fused_forward_matches_reference_fixture(name, family, monkeypatch):
    from clid_slam_b200 import fused

    bricks = _kernel_family(monkeypatch, family)

    fx = gio.load("query", name)
    m = gio.oracle_map(fx)
    npm = hp.product_map(m)
    dec = hp.product_decoder(m.cfg, gio.decoder_params(fx))
    locally = bool(fx["query_locally"])
    x = gio.t(fx["x"]).cuda()
    if bricks:
        assert npm.brick_index(locally) is not None, "brick index unexpectedly unavailable"
    sdf, grad, nn, cert = fused.sdf_and_gradient(npm, dec, x, None, training_mode=False, query_locally=locally,
                                                 use_bricks=bricks)
    assert torch.equal(nn.cpu().long(), gio.t(fx["out_nn"]))
    gio.assert_close(sdf, fx["out_sdf"], hp.SDF_RTOL, hp.SDF_ATOL, "fused sdf")
    gio.assert_close(grad, fx["out_grad"], hp.GRAD_RTOL, hp.GRAD_ATOL, "fused grad")
    gio.assert_close(cert, fx["out_certainty"], 1e-5, 1e-6, "fused certainty")
    # inference mode leaves the map untouched
    gio.assert_close(npm.local_point_certainties, m.local_certainties, 0, 0, "no side effect")


@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("layer_norm,levels,hidden", [(False, 1, 64), (True, 1, 64), (False, 2, 32), (False, 1, 32), (False, 1, 128)])
def test_fused_forward_matches_oracle_large(layer_norm, levels, hidden, family, monkeypatch):
    from clid_slam_b200 import fused

    bricks = _kernel_family(monkeypatch, family)

    cfg = oc.OracleConfig(buffer_size=2_000_003, layer_norm_on=layer_norm, geo_mlp_level=levels,
                          geo_mlp_hidden_dim=hidden, local_map_radius=80.0)
    m, params, gen = hp.build_oracle_world(300, 2, seed=11, cfg=cfg)  # 180 k points
    m.local_certainties.copy_(torch.rand(m.local_certainties.shape, generator=gen))
    n = 65536
    x, _, _, ts = oc.sample_batch(m.points, n, gen)
    npm = hp.product_map(m)
    dec = hp.product_decoder(cfg, params)

    xo = x.clone().requires_grad_(True)
    z, w, nn, cert = oc.query_feature(m, xo, None, training_mode=False, query_locally=True)
    sdf_o = oc.decoder_sdf(params, z, cfg.sdf_scale)
    grad_o = oc.sdf_gradient(xo, sdf_o)

    sdf, grad, nn_g, cert_g = fused.sdf_and_gradient(npm, dec, x.cuda(), use_bricks=bricks)
    assert torch.equal(nn_g.cpu().long(), nn)
    gio.assert_close(sdf, sdf_o, hp.SDF_RTOL, hp.SDF_ATOL, "sdf")
    gio.assert_close(grad, grad_o, hp.GRAD_RTOL, hp.GRAD_ATOL, "grad")
    gio.assert_close(cert_g, cert, 1e-5, 1e-6, "certainty")


def test_empty_and_tiny_batches():
    from clid_slam_b200 import fused

    m, params, gen = hp.build_oracle_world(40, 1, seed=3, cfg=oc.OracleConfig(buffer_size=100_003))
    npm = hp.product_map(m)
    dec = hp.product_decoder(m.cfg, params)
    sdf, grad, nn, cert = fused.sdf_and_gradient(npm, dec, torch.empty(0, 3, device="cuda"))
    assert sdf.shape == (0,) and grad.shape == (0, 3) and nn.shape == (0,)
    for n in (1, 31, 129):
        x, _, _, _ = oc.sample_batch(m.points, n, gen)
        xo = x.clone().requires_grad_(True)
        z, _, nn_o, _ = oc.query_feature(m, xo, None, False, True)
        sdf_o = oc.decoder_sdf(params, z, m.cfg.sdf_scale)
        sdf, grad, nn, _ = fused.sdf_and_gradient(npm, dec, x.cuda())
        assert torch.equal(nn.cpu().long(), nn_o)
        gio.assert_close(sdf, sdf_o, hp.SDF_RTOL, hp.SDF_ATOL, f"sdf n={n}")


def test_far_queries_have_no_neighbours():
    """Zero-neighbour rows: z = 0, grad = 0, sdf = s (w_out . relu(b1) + b_out)  (SURVEY appendix B4)."""
    from clid_slam_b200 import fused

    m, params, gen = hp.build_oracle_world(40, 1, seed=4, cfg=oc.OracleConfig(buffer_size=100_003))
    npm = hp.product_map(m)
    dec = hp.product_decoder(m.cfg, params)
    x = torch.tensor([[500.0, -400.0, 300.0], [-1234.5, 10.0, 0.0]], device="cuda")
    sdf, grad, nn, cert = fused.sdf_and_gradient(npm, dec, x)
    assert int(nn.abs().sum()) == 0 and float(grad.abs().sum()) == 0.0 and float(cert.abs().sum()) == 0.0
    expect = oc.decoder_sdf(params, torch.zeros(2, 11), m.cfg.sdf_scale)
    gio.assert_close(sdf, expect, 1e-6, 1e-8, "sdf of empty neighbourhood")


def test_cpu_tensors_are_rejected_loudly():
    m, params, gen = hp.build_oracle_world(40, 1, seed=5, cfg=oc.OracleConfig(buffer_size=100_003))
    npm = hp.product_map(m)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        npm.query_feature(torch.zeros(4, 3))
