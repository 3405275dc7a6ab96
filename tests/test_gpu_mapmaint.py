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


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32_pose", "f64_pose"])
def test_native_pool_filter_equals_the_torch_ops(dtype):
    from clid_slam_b200.ops import mapmaint as mm

    gen = torch.Generator(device="cuda").manual_seed(11)
    n, n_tail = 300_001, 7000
    gc = (torch.rand(n, 3, generator=gen, device="cuda") - 0.5) * 150.0
    coord = torch.rand(n, 3, generator=gen, device="cuda")
    label = torch.rand(n, generator=gen, device="cuda")
    weight = torch.rand(n, generator=gen, device="cuda") - 0.5
    stamp = torch.randint(0, 90, (n,), generator=gen, device="cuda", dtype=torch.int32)
    origin = torch.tensor([3.0, -2.0, 0.5], device="cuda", dtype=dtype)
    radius = 50.0
    outs, n_keep, n_tail_keep, flags = mm.pool_filter(gc, origin, radius, [coord, gc, label, weight, stamp], n_tail)
    d2 = ((gc - origin) ** 2).sum(-1)  # utils/mapper.py:421-423 (float64 when the pose is float64)
    keep = d2 < radius**2
    if dtype == torch.float64:
        assert torch.equal(flags.bool(), keep)
    keep = flags.bool()  # fp32: torch's CUDA reduction may round a d2 within an ulp of the radius differently
    assert int((keep != (d2 < radius**2)).sum()) <= 2
    assert n_keep == int(keep.sum()) and n_tail_keep == int(keep[-n_tail:].sum()) and 0 < n_keep < n
    for got, src in zip(outs, [coord, gc, label, weight, stamp]):
        assert got.dtype == src.dtype and torch.equal(got, src[keep])


def test_process_frame_with_the_native_pool_filter_equals_the_torch_filter(monkeypatch):
    """Two mappers on the same scans, one with the native filter switched off: identical pools and counters."""
    from clid_slam_b200.model.decoder import Decoder
    from clid_slam_b200.model.local_point_cloud_map import LocalPointCloudMap
    from clid_slam_b200.utils.mapper import Mapper
    from test_gpu_mapper_flow import FakeDataset, _scan

    pools = []
    for native in (True, False):
        torch.manual_seed(42)
        cfg = ncd128()
        cfg.device, cfg.use_pin_mapper = "cuda", True
        cfg.buffer_size, cfg.local_buffer_size = 2_000_003, 500_009
        cfg.pool_filter_freq, cfg.window_radius = 1, 14.0  # every frame drops the far samples
        dec = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
        npm = NeuralPoints(cfg)
        ds = FakeDataset(3)
        mapper = Mapper(cfg, ds, npm, LocalPointCloudMap(cfg), dec)
        if not native:
            monkeypatch.setattr(mapper, "_filter_pool_native", lambda *a, **k: False)
        gen = torch.Generator(device="cuda").manual_seed(5)
        torch.cuda.manual_seed(9)
        for frame in range(3):
            ds.processed_frame = frame
            pose = torch.eye(4, device="cuda", dtype=torch.float64)
            pose[0, 3] = 2.0 * frame
            ds.gt_poses[frame, 0, 3] = 2.0 * frame
            npm.travel_dist = torch.arange(frame + 1, device="cuda", dtype=torch.float32) * 2.0
            mapper.process_frame(_scan(gen, "cuda"), None, pose, frame)
        pools.append((mapper.coord_pool, mapper.global_coord_pool, mapper.sdf_label_pool, mapper.weight_pool,
                      mapper.time_pool, mapper.pool_sample_count, mapper.cur_sample_count, mapper.new_idx))
    a, b = pools
    assert a[5] == b[5] and a[6] == b[6] and 0 < a[5]
    for x, y in zip(a[:5], b[:5]):
        assert torch.equal(x, y)
    assert torch.equal(a[7], b[7])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32_pose", "f64_pose"])
def test_local_point_cloud_map_on_the_gpu_equals_the_host_logic(dtype):
    """LocalPointCloudMap.update_map (model/local_point_cloud_map.py:35-72): insert, range filter, hash rebuild --
    the CUDA map (native down-sampling, table store, compaction) holds the same points in the same order and the
    same table as the torch ops on CPU; a small table forces repeated slots."""
    from clid_slam_b200.model.local_point_cloud_map import LocalPointCloudMap

    clouds = []
    for device in ("cpu", "cuda"):
        cfg = ncd128()
        cfg.device, cfg.local_buffer_size, cfg.local_map_size = device, 20_011, 18.0
        clouds.append(LocalPointCloudMap(cfg))
    cpu, gpu = clouds
    gen = torch.Generator().manual_seed(3)
    for frame in range(5):
        centre = torch.tensor([4.0 * frame, -1.0 * frame, 0.2])
        pts = (torch.rand(25_000, 3, generator=gen) - 0.5) * torch.tensor([50.0, 40.0, 3.0]) + centre
        cpu.update_map(centre.to(dtype), pts)
        gpu.update_map(centre.to(dtype).cuda(), pts.cuda())
        assert gpu.local_point_cloud_map.shape[0] > 1000
        assert torch.equal(cpu.local_point_cloud_map, gpu.local_point_cloud_map.cpu())
        assert torch.equal(cpu.buffer_pt_index, gpu.buffer_pt_index.cpu())


def test_recreate_hash_on_the_gpu_equals_the_host_logic():
    """NeuralPoints.recreate_hash (model/neural_points.py:840-929) after a prune: min-value down-sampling and the
    last-writer-wins table store run natively on the GPU."""
    cpu, gpu = _twin_maps(200_003)
    gen = torch.Generator().manual_seed(0)
    pts = oc.wavy_sheets(60, 2, cpu.resolution, gen)
    cpu.travel_dist, gpu.travel_dist = torch.zeros(4), torch.zeros(4, device="cuda")
    cpu.update(pts, torch.zeros(3), torch.eye(3), 0)
    gpu.update(pts.cuda(), torch.zeros(3, device="cuda"), torch.eye(3, device="cuda"), 0)
    cert = torch.rand(cpu.count(), generator=gen) * 4
    cpu.point_certainties, gpu.point_certainties = cert.clone(), cert.cuda()
    gpu.geo_features = cpu.geo_features.cuda()
    for with_ts, kept in ((True, True), (False, True), (False, False)):
        assert cpu.prune_map(1.0, min_prune_count=10, global_prune=True) == gpu.prune_map(1.0, min_prune_count=10, global_prune=True)
        cpu.recreate_hash(torch.zeros(3), torch.eye(3), kept_points=kept, with_ts=with_ts, cur_ts=0)
        gpu.recreate_hash(torch.zeros(3, device="cuda"), torch.eye(3, device="cuda"), kept_points=kept, with_ts=with_ts, cur_ts=0)
        _same_state(cpu, gpu)
        assert torch.equal(cpu.geo_features, gpu.geo_features.cpu())
        cpu.point_certainties *= 0.7
        gpu.point_certainties *= 0.7


@pytest.mark.parametrize("run_file", ["ncd128", "subt"])
def test_native_ray_sampler_draws_the_samples_of_the_torch_ops(run_file):
    """DataSampler.sample_pin / sample on the GPU (utils/data_sampler.py:16-402): with the same seed the kernels
    return, bit for bit, what the chain of torch ops returns on the same device."""
    from clid_slam_b200.model.local_point_cloud_map import LocalPointCloudMap
    from clid_slam_b200.utils.data_sampler import DataSampler
    from test_gpu_mapper_flow import _scan

    cfg = ncd128()
    cfg.device, cfg.local_buffer_size = "cuda", 500_009
    if run_file == "subt":  # config/run_SubT_MRS.yaml:19: other begin ratio, more samples per ray
        cfg.free_sample_begin_ratio, cfg.surface_sample_n, cfg.free_front_n, cfg.free_sample_end_dist_m = 0.8, 3, 3, 0.8
    sampler = DataSampler(cfg)
    gen = torch.Generator(device="cuda").manual_seed(2)
    scan = _scan(gen, "cuda", 20_000)

    torch.manual_seed(7)
    want = sampler._sample_pin_torch(scan)
    torch.manual_seed(7)
    got = sampler.sample_pin(scan)
    for name, a, b in zip(("coord", "label", "normal", "sem", "color", "weight"), got, want):
        assert (a is None and b is None) or torch.equal(a, b), name

    lpcm = LocalPointCloudMap(cfg)
    pose = torch.eye(4, device="cuda", dtype=torch.float64)
    pose[:3, 3] = torch.tensor([1.0, -2.0, 0.5])
    lpcm.update_map(pose[:3, 3], tools.transform_torch(scan, pose))
    torch.manual_seed(8)
    want = sampler._sample_torch(scan, lpcm, pose)
    torch.manual_seed(8)
    got = sampler.sample(scan, lpcm, pose)
    assert 0 < got[0].shape[0] < scan.shape[0] * 8 + 1
    for name, a, b in zip(("coord", "label", "weight"), got, want):
        assert a.shape == b.shape and torch.equal(a, b), name
