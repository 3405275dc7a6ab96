"""slam.py-shaped flow on the GPU: a few synthetic scans through Mapper.process_frame + mapping
(sample_pin sampler), the way slam.py:135-208 drives the modules."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class FakeDataset:
    lose_track = False
    stop_status = False
    processed_frame = 0
    gt_pose_provided = True
    pgo_poses = None
    static_mask = None

    def __init__(self, n):
        self.gt_poses = self.odom_poses = np.tile(np.eye(4), (n, 1, 1))


def _scan(gen, device, n=6000):
    """Points of a wavy floor + a wall, in the sensor frame."""
    xy = (torch.rand(n, 2, generator=gen, device=device) - 0.5) * 40
    floor = torch.cat((xy, -1.5 + 0.3 * torch.sin(xy[:, :1] / 3)), dim=1)
    yz = (torch.rand(n // 3, 2, generator=gen, device=device) - 0.5) * torch.tensor([30.0, 4.0], device=device)
    wall = torch.cat((torch.full((n // 3, 1), 12.0, device=device), yz), dim=1)
    pts = torch.cat((floor, wall), 0)
    return pts[pts.norm(dim=1) > 1.0]


@pytest.mark.parametrize("sampler", ["pin", "clid"])
@pytest.mark.parametrize("mode", ["numerical", "analytic"])
def test_three_frames_train_and_write_back(mode, sampler):
    from clid_slam_b200.config import ncd128
    from clid_slam_b200.model.decoder import Decoder
    from clid_slam_b200.model.local_point_cloud_map import LocalPointCloudMap
    from clid_slam_b200.model.neural_points import NeuralPoints
    from clid_slam_b200.utils.mapper import Mapper
    from clid_slam_b200.utils.tools import freeze_model

    torch.manual_seed(42)
    cfg = ncd128()
    cfg.device = "cuda"
    cfg.use_pin_mapper = sampler == "pin"  # "clid": region-specific labels from the local point-cloud map
    cfg.local_buffer_size = 500_009
    cfg.buffer_size = 2_000_003
    cfg.feature_std = 0.0  # reference default: features start at exactly zero
    if mode == "analytic":
        cfg.numerical_grad, cfg.gradient_decimation = False, 1
    dec = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    npm = NeuralPoints(cfg)
    frames = 3
    ds = FakeDataset(frames)
    mapper = Mapper(cfg, ds, npm, LocalPointCloudMap(cfg), dec)
    gen = torch.Generator(device="cuda").manual_seed(5)
    first_loss = last_loss = None
    for frame in range(frames):
        ds.processed_frame = frame
        pose = torch.eye(4, device="cuda", dtype=torch.float64)
        pose[0, 3] = 0.5 * frame
        ds.gt_poses[frame, 0, 3] = 0.5 * frame
        npm.travel_dist = torch.arange(frame + 1, device="cuda", dtype=torch.float32) * 0.5
        mapper.process_frame(_scan(gen, "cuda"), None, pose, frame)
        assert npm.count() > 0 and npm.local_count() > 0
        assert mapper.pool_sample_count == mapper.coord_pool.shape[0]
        if frame == 2:
            freeze_model(dec)  # slam.py:193-196 freezes the decoder after freeze_after_frame
        before = npm.geo_features.clone()
        dec_before = [p.detach().clone() for p in dec.parameters()]
        mapper.mapping(12)
        losses = mapper.last_losses.cpu()
        assert torch.isfinite(losses).all()
        first_loss = losses[0, 0] if first_loss is None else first_loss
        last_loss = losses[-1, 0]
        assert not torch.equal(before, npm.geo_features), "trained features must be written back to the global map"
        changed = any(not torch.equal(a, b.detach()) for a, b in zip(dec_before, dec.parameters()))
        assert changed == (frame < 2), "decoder trains until frozen"
        assert float(npm.point_certainties.sum()) > 0
        assert int(npm.point_ts_update.max()) == frame
    assert last_loss < first_loss, (first_loss, last_loss)
    # the map answers queries afterwards
    from clid_slam_b200 import fused

    probe = npm.neural_points[::50].contiguous()
    sdf, grad, nn, cert = fused.sdf_and_gradient(npm, dec, probe)
    assert torch.isfinite(sdf).all() and torch.isfinite(grad).all() and int(nn.min()) >= 1
    assert float(sdf.abs().mean()) < 0.2  # neural points sit on the surface: |sdf| should already be small


def test_dynamic_filter_and_certainty_queries_run():
    from clid_slam_b200.config import ncd128
    from clid_slam_b200.model.decoder import Decoder
    from clid_slam_b200.model.neural_points import NeuralPoints
    from clid_slam_b200.utils.mapper import Mapper

    torch.manual_seed(1)
    cfg = ncd128()
    cfg.device, cfg.use_pin_mapper, cfg.buffer_size = "cuda", True, 500_009
    dec = Decoder(cfg, 64, 1, 1)
    npm = NeuralPoints(cfg)
    ds = FakeDataset(2)
    mapper = Mapper(cfg, ds, npm, None, dec)
    gen = torch.Generator(device="cuda").manual_seed(9)
    npm.travel_dist = torch.zeros(2, device="cuda")
    pose = torch.eye(4, device="cuda", dtype=torch.float64)
    mapper.process_frame(_scan(gen, "cuda", 3000), None, pose, 0)
    mapper.mapping(5)
    ds.processed_frame = 1
    mapper.process_frame(_scan(gen, "cuda", 3000), None, pose, 1, filter_dynamic=True)
    assert mapper.static_mask.dtype == torch.bool and mapper.static_mask.numel() > 0
    assert mapper.new_idx is not None


@pytest.mark.parametrize("mode", ["numerical", "analytic"])
def test_graphed_mapping_tracks_the_eager_loop(mode):
    """mapping() with many iterations replays one CUDA graph of [replay-pool draw + iteration]; it must
    train like the call-by-call loop (different random batches, so the comparison is statistical)."""
    from clid_slam_b200.config import ncd128
    from clid_slam_b200.model.decoder import Decoder
    from clid_slam_b200.model.neural_points import NeuralPoints
    from clid_slam_b200.utils.mapper import Mapper

    iters = 40
    results = {}
    for graphed in (True, False):
        torch.manual_seed(3)
        cfg = ncd128()
        cfg.device, cfg.use_pin_mapper, cfg.buffer_size = "cuda", True, 2_000_003
        if mode == "analytic":
            cfg.numerical_grad, cfg.gradient_decimation = False, 1
        dec = Decoder(cfg, 64, 1, 1)
        npm = NeuralPoints(cfg)
        ds = FakeDataset(1)
        mapper = Mapper(cfg, ds, npm, None, dec)
        mapper.graph_min_iters = 8 if graphed else 0
        gen = torch.Generator(device="cuda").manual_seed(9)
        npm.travel_dist = torch.zeros(1, device="cuda")
        mapper.process_frame(_scan(gen, "cuda"), None, torch.eye(4, device="cuda", dtype=torch.float64), 0)
        mapper.adaptive_iter_offset = 0
        mapper.mapping(iters)
        losses = mapper.last_losses.cpu()
        assert losses.shape == (iters, 3) and torch.isfinite(losses).all()
        assert mapper.total_iter == iters
        assert float(npm.point_certainties.sum()) > 0
        results[graphed] = (losses, npm.geo_features.clone(), [p.detach().clone() for p in dec.parameters()])
    lg, le = results[True][0], results[False][0]
    assert lg[-5:, 0].mean() < 0.8 * lg[:3, 0].mean(), "the graphed loop must reduce the loss"
    # same trajectory up to batch noise: the tail losses of the two loops agree within 15 %
    assert abs(float(lg[-10:, 0].mean()) / float(le[-10:, 0].mean()) - 1.0) < 0.15, (lg[-10:, 0].mean(), le[-10:, 0].mean())
    # and the decoders moved the same way
    for a, b in zip(results[True][2], results[False][2]):
        assert torch.nn.functional.cosine_similarity(a.flatten() - 0, b.flatten() - 0, dim=0) > 0.99


def test_native_batch_draw_follows_get_batch_semantics():
    """clid_draw_batch: uniform rows of the pool, the last bs_new of the batch out of new_idx, values gathered from
    the drawn rows, reproducible for (seed, offset) and different across offsets (utils/mapper.py:473-523)."""
    from clid_slam_b200.ops.train import draw_batch

    gen = torch.Generator(device="cuda").manual_seed(3)
    P, n, bs_new = 50_000, 16384, 1000
    coord = torch.randn(P + 77, 3, generator=gen, device="cuda")  # rows beyond pool_sample_count must never be drawn
    label = torch.randn(P + 77, generator=gen, device="cuda")
    weight = torch.randn(P + 77, generator=gen, device="cuda")
    time = torch.randint(0, 9, (P + 77,), generator=gen, device="cuda", dtype=torch.int32)
    new_idx = torch.randperm(P, generator=gen, device="cuda")[:3000]
    x, lb, w, ts, idx = draw_batch(coord, label, weight, time, P, n, seed=1234, offset=7, new_idx=new_idx, bs_new=bs_new)
    assert int(idx.min()) >= 0 and int(idx.max()) < P
    assert torch.equal(x, coord[idx]) and torch.equal(lb, label[idx]) and torch.equal(w, weight[idx]) and torch.equal(ts, time[idx])
    assert bool(torch.isin(idx[n - bs_new:], new_idx).all()), "the tail of the batch comes from the new samples"
    # uniform over the pool: 16 equal bins of the history part, chi-square with 15 dof (99.9 % quantile 37.7)
    hist = torch.bincount((idx[: n - bs_new] * 16 // P), minlength=16).double()
    expect = (n - bs_new) / 16
    assert float(((hist - expect) ** 2 / expect).sum()) < 45.0
    again = draw_batch(coord, label, weight, time, P, n, seed=1234, offset=7, new_idx=new_idx, bs_new=bs_new)[4]
    other = draw_batch(coord, label, weight, time, P, n, seed=1234, offset=8, new_idx=new_idx, bs_new=bs_new)[4]
    assert torch.equal(idx, again) and not torch.equal(idx, other)


@pytest.mark.parametrize("mode", ["numerical", "analytic"])
def test_native_mapping_loop_matches_the_python_loop(mode):
    """Mapper.mapping through clid_mapping_run (one native call for all iterations) against the call-by-call Python
    loop fed with the very batches the native draw produces: same losses, same trained map."""
    from clid_slam_b200.config import ncd128
    from clid_slam_b200.model.decoder import Decoder
    from clid_slam_b200.model.local_point_cloud_map import LocalPointCloudMap
    from clid_slam_b200.model.neural_points import NeuralPoints
    from clid_slam_b200.ops.train import draw_batch
    from clid_slam_b200.utils.mapper import Mapper

    iters = 10
    results = []
    for native in (True, False):
        torch.manual_seed(42)
        cfg = ncd128()
        cfg.device = "cuda"
        cfg.use_pin_mapper = True
        cfg.buffer_size = 2_000_003
        cfg.feature_std = 0.05
        cfg.bs = 4096
        if mode == "analytic":
            cfg.numerical_grad, cfg.gradient_decimation = False, 1
        dec = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
        npm = NeuralPoints(cfg)
        ds = FakeDataset(1)
        mapper = Mapper(cfg, ds, npm, LocalPointCloudMap(cfg), dec)
        gen = torch.Generator(device="cuda").manual_seed(5)
        npm.travel_dist = torch.zeros(1, device="cuda")
        mapper.process_frame(_scan(gen, "cuda"), None, torch.eye(4, device="cuda", dtype=torch.float64), 0)
        mapper.adaptive_iter_offset = 0
        mapper.native_loop = native
        torch.manual_seed(7)  # the native loop seeds its generator from torch's CPU generator
        if not native:
            seed = int(torch.randint(0, 2**62, (1,)).item())
            n_new = 0 if mapper.new_idx is None else int(mapper.new_idx.shape[0])
            bs_new = min(n_new, cfg.bs_new_sample) if cfg.bs_new_sample > 0 else 0
            feed = iter(range(iters))

            def replay(global_coord=False):
                x, lb, w, ts, _ = draw_batch(mapper.global_coord_pool, mapper.sdf_label_pool, mapper.weight_pool,
                                             mapper.time_pool, mapper.pool_sample_count, cfg.bs, seed, next(feed),
                                             new_idx=mapper.new_idx, bs_new=bs_new)
                return x, lb, ts, None, None, None, w

            mapper.get_batch = replay
        mapper.mapping(iters)
        results.append((mapper.last_losses.cpu(), npm.geo_features.clone(), npm.point_certainties.clone(),
                        [p.detach().clone() for p in dec.parameters()]))
    (l_n, f_n, c_n, d_n), (l_p, f_p, c_p, d_p) = results
    assert l_n.shape == (iters, 3)
    torch.testing.assert_close(l_n, l_p, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(c_n, c_p, rtol=1e-4, atol=1e-5)
    bad = ((f_n - f_p).abs() > 1e-5 + 1e-3 * f_p.abs()).double().mean().item()
    assert bad < 5e-3, f"{bad:.2e} of the feature entries differ"  # Adam turns ~0 gradients' signs into +-lr steps
    for a, b in zip(d_n, d_p):
        torch.testing.assert_close(a, b, rtol=2e-3, atol=2e-5)
