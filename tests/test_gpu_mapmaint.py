"""Native per-frame map maintenance (csrc/mapmaint.cuh) on the GPU: bit-exact against the fixtures
the reference produced on CPU (model/neural_points.py:324-549) and against the torch host logic of the product
on seeded worlds, including the reference's edge cases (hash collisions in a small table, repeated slots, the
travel-distance window, fewer than 100 points in the window, use_mid_ts, a float64 sensor position)."""
import pytest
import torch

import golden_io as gio
from clid_slam_b200.config import Config, ncd128
from clid_slam_b200.model.neural_points import NeuralPoints
from clid_slam_b200.utils import tools
from oracle import sdf_oracle as oc

pytestmark = pytest.mark.gpu


def _config(ocfg, device) -> Config:
    cfg = Config()
    for name in ocfg.__dataclass_fields__:
        setattr(cfg, name, getattr(ocfg, name))
    cfg.device = device
    return cfg


@pytest.mark.parametrize("name", gio.names("map"))
def test_native_insert_and_window_match_the_reference_fixture(name):
    fx = gio.load("map", name)
    cfg = _config(gio.config_of(fx), "cuda")
    npm = NeuralPoints(cfg)
    npm.travel_dist = gio.t(fx["travel_dist"]).cuda()
    for i in range(int(fx["n_frames"])):
        pre = f"frame{i}_map_"
        ratio = npm.update(gio.t(fx[f"frame{i}_points"]).cuda(), gio.t(fx[f"frame{i}_sensor"]).cuda(),
                           torch.eye(3, device="cuda"), int(fx[f"frame{i}_ts"]))
        assert ratio == float(fx[f"frame{i}_ratio"])
        assert torch.equal(npm.buffer_pt_index.cpu(), gio.dense_table(fx, pre))
        assert torch.equal(npm.neural_points.cpu(), gio.t(fx[pre + "points"]))
        assert torch.equal(npm.point_ts_create.cpu(), gio.t(fx[pre + "ts_create"]))
        assert torch.equal(npm.point_ts_update.cpu(), gio.t(fx[pre + "ts_update"]))
        assert torch.equal(npm.local_mask.cpu(), gio.t(fx[pre + "local_mask"]))
        assert torch.equal(npm.global2local.cpu(), gio.t(fx[pre + "global2local"]))
        assert torch.equal(npm.local_neural_points.cpu(), gio.t(fx[pre + "local_points"]))
        assert npm.local_mask.dtype == torch.bool and npm.global2local.dtype == torch.int64
        assert isinstance(npm.local_geo_features, torch.nn.Parameter)
        assert torch.equal(npm.local_geo_features.data, npm.geo_features[npm.local_mask])
        assert torch.equal(npm._local_gids, torch.nonzero(npm.local_mask[:-1]).flatten())
    before = npm.geo_features.clone()
    npm.assign_local_to_global()
    assert torch.equal(before, npm.geo_features)


def _twin_maps(buffer_size=200_003, **over):
    maps = []
    for device in ("cpu", "cuda"):
        cfg = ncd128()
        cfg.device, cfg.buffer_size = device, buffer_size
        for k, v in over.items():
            setattr(cfg, k, v)
        npm = NeuralPoints(cfg)
        maps.append(npm)
    return maps


def _same_state(a, b):
    for name in ("buffer_pt_index", "neural_points", "point_ts_create", "point_ts_update", "point_orientations",
                 "local_mask", "global2local", "local_neural_points", "local_point_orientations",
                 "local_point_ts_update", "local_point_certainties"):
        ta, tb = getattr(a, name), getattr(b, name)
        assert ta.dtype == tb.dtype and ta.shape == tb.shape, name
        assert torch.equal(ta, tb.cpu()), name


@pytest.mark.parametrize("case", ["plain", "collisions", "mid_ts", "f64_sensor", "tiny_window"])
def test_native_map_maintenance_equals_the_host_logic_over_frames(case):
    over = {}
    buffer_size = 200_003
    if case == "collisions":
        buffer_size = 4099  # far fewer slots than voxels: repeated slots and far owners on every frame
    if case == "mid_ts":
        over["use_mid_ts"] = True
    if case == "tiny_window":
        over["local_map_radius"] = 6.0
    cpu, gpu = _twin_maps(buffer_size, **over)
    gen = torch.Generator().manual_seed(7)
    travel = torch.cumsum(torch.full((8,), 4.0 if case != "tiny_window" else 400.0), 0) - 4.0
    cpu.travel_dist, gpu.travel_dist = travel, travel.cuda()
    dtype = torch.float64 if case == "f64_sensor" else torch.float32
    for ts in range(6):
        centre = torch.tensor([3.0 * ts, 1.0 * ts, 0.0])
        n_pts = 20_000 if case != "tiny_window" else (60 if ts else 3000)  # later frames see < 100 points in the window
        pts = (torch.rand(n_pts, 3, generator=gen) - 0.5) * torch.tensor([40.0, 30.0, 3.0]) + centre
        sensor = centre.to(dtype)
        r_cpu = cpu.update(pts, sensor, torch.eye(3), ts)
        r_gpu = gpu.update(pts.cuda(), sensor.cuda(), torch.eye(3, device="cuda"), ts)
        assert r_cpu == r_gpu
        _same_state(cpu, gpu)
        # a window move without insert (slam.py:178-181), then a trained write-back
        cpu.reset_local_map(sensor + 1.0, None, ts)
        gpu.reset_local_map((sensor + 1.0).cuda(), None, ts)
        _same_state(cpu, gpu)
        gpu.geo_features = cpu.geo_features.cuda()  # the feature draws come from different generators
        gpu.reset_local_map((sensor + 1.0).cuda(), None, ts)
        for npm in (cpu, gpu):
            npm.local_geo_features.data.mul_(1.5)
            npm.local_point_certainties += 2.0
            npm.local_point_ts_update.fill_(ts)
            npm.assign_local_to_global()
        assert torch.equal(cpu.geo_features, gpu.geo_features.cpu())
        assert torch.equal(cpu.point_certainties, gpu.point_certainties.cpu())
        assert torch.equal(cpu.point_ts_update, gpu.point_ts_update.cpu())


@pytest.mark.parametrize("n,voxel", [(20_000, 0.4), (200_000, 0.08), (1, 0.4), (777, 5.0)])
def test_native_voxel_down_sampling_equals_the_oracle(n, voxel):
    gen = torch.Generator().manual_seed(n)
    pts = torch.rand(n, 3, generator=gen) * torch.tensor([30.0, 25.0, 4.0]) - 10.0
    want = oc.voxel_downsample_indices(pts, voxel)
    got = tools.voxel_down_sample_torch(pts.cuda(), voxel)
    assert got.dtype == torch.int64 and torch.equal(got.cpu(), want)
    # duplicates of one point: ties go to the smaller index
    dup = pts[:50].repeat(4, 1)
    assert torch.equal(tools.voxel_down_sample_torch(dup.cuda(), voxel).cpu(), oc.voxel_downsample_indices(dup, voxel))


def test_native_min_value_down_sampling_equals_the_host_logic():
    gen = torch.Generator().manual_seed(5)
    pts = torch.rand(50_000, 3, generator=gen) * 20.0 - 5.0
    val = torch.rand(50_000, generator=gen)
    want = tools.voxel_down_sample_min_value_torch(pts, 0.4, val)
    got = tools.voxel_down_sample_min_value_torch(pts.cuda(), 0.4, val.cuda())
    assert torch.equal(got.cpu(), want)


def test_large_map_window_is_exact():
    """1 M points: every block of the three-launch scan carries an offset."""
    cfg = ncd128()
    cfg.device = "cuda"
    cfg.local_map_radius = 60.0
    npm = NeuralPoints(cfg)
    npm.travel_dist = torch.zeros(2, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(1)
    from clid_slam_b200.synth import wavy_sheets

    pts = wavy_sheets(520, 4, cfg.voxel_size_m, gen, device="cuda")
    npm.update(pts, torch.zeros(3, device="cuda"), torch.eye(3, device="cuda"), 0)
    assert npm.count() > 1_000_000
    d2 = ((npm.neural_points - torch.zeros(3, device="cuda")) ** 2).sum(-1)
    want = d2 < cfg.local_map_radius**2
    differ = npm.local_mask[:-1] != want  # torch's CUDA reduction may round d2 differently within an ulp of the radius
    assert int(differ.sum()) <= 2 and bool(((d2 - cfg.local_map_radius**2).abs()[differ] < 1e-2).all())
    want = npm.local_mask[:-1]
    assert 100_000 < int(want.sum()) < npm.count()
    rows = torch.nonzero(npm.local_mask).flatten()
    g2l = torch.full((npm.count() + 1,), -1, dtype=torch.long, device="cuda")
    g2l[rows] = torch.arange(rows.numel(), device="cuda")
    g2l[-1] = -1
    assert torch.equal(npm.global2local, g2l)
    assert torch.equal(npm.local_neural_points, npm.neural_points[want])
    # every point owns the slot of its voxel (one point per voxel on this world)
    slots = npm._slots_of(npm.neural_points) % int(npm.buffer_size)
    owners = npm.buffer_pt_index[slots]
    assert (owners >= 0).all()
