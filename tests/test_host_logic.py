"""CPU tests of the host-side logic: per-frame map maintenance of the product NeuralPoints (torch ops,
device-agnostic) against reference-generated fixtures, the sampler, config loading, and the
batch-sharding / all-reduce plumbing with a 2-process gloo group."""
import json
import os
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

import golden_io as gio
from clid_slam_b200 import dist as cdist
from clid_slam_b200.config import Config, ncd128
from clid_slam_b200.model.neural_points import NeuralPoints
from clid_slam_b200.utils.data_sampler import DataSampler
from clid_slam_b200.utils.tools import voxel_down_sample_torch
from oracle import sdf_oracle as oc


def _cpu_config(ocfg) -> Config:
    cfg = Config()
    for name in ocfg.__dataclass_fields__:
        setattr(cfg, name, getattr(ocfg, name))
    cfg.device = "cpu"
    return cfg


@pytest.mark.parametrize("name", gio.names("map"))
def test_map_insert_and_local_window_match_reference(name):
    fx = gio.load("map", name)
    cfg = _cpu_config(gio.config_of(fx))
    npm = NeuralPoints(cfg)
    npm.travel_dist = gio.t(fx["travel_dist"])
    for i in range(int(fx["n_frames"])):
        pre = f"frame{i}_map_"
        ratio = npm.update(gio.t(fx[f"frame{i}_points"]), gio.t(fx[f"frame{i}_sensor"]), torch.eye(3),
                           int(fx[f"frame{i}_ts"]))
        assert ratio == float(fx[f"frame{i}_ratio"])
        assert torch.equal(npm.buffer_pt_index, gio.dense_table(fx, pre))
        assert torch.equal(npm.neural_points, gio.t(fx[pre + "points"]))
        assert torch.equal(npm.point_ts_create, gio.t(fx[pre + "ts_create"]))
        assert torch.equal(npm.point_ts_update, gio.t(fx[pre + "ts_update"]))
        assert torch.equal(npm.local_mask, gio.t(fx[pre + "local_mask"]))
        assert torch.equal(npm.global2local, gio.t(fx[pre + "global2local"]))
        assert torch.equal(npm.local_neural_points, gio.t(fx[pre + "local_points"]))
        assert npm.geo_features.shape == (npm.count() + 1, cfg.feature_dim)
        assert isinstance(npm.local_geo_features, torch.nn.Parameter)
        assert npm.local_geo_features.shape[0] == npm.local_count() + 1
    # write-back is the identity when nothing was trained
    before = npm.geo_features.clone()
    npm.assign_local_to_global()
    assert torch.equal(before, npm.geo_features)


def test_prune_and_recreate_hash_keep_the_table_consistent():
    cfg = ncd128()
    cfg.device, cfg.buffer_size = "cpu", 200_003
    npm = NeuralPoints(cfg)
    npm.travel_dist = torch.zeros(4)
    gen = torch.Generator().manual_seed(0)
    pts = oc.wavy_sheets(40, 1, cfg.voxel_size_m, gen)
    npm.update(pts, torch.zeros(3), torch.eye(3), 0)
    n0 = npm.count()
    npm.point_certainties = torch.rand(n0, generator=gen) * 4
    assert npm.prune_map(2.0, min_prune_count=10, global_prune=True)
    assert npm.count() < n0 and npm.geo_features.shape[0] == npm.count() + 1
    npm.recreate_hash(torch.zeros(3), torch.eye(3), kept_points=True, with_ts=True, cur_ts=0)
    # every slot points at a point whose voxel hashes to that slot
    slots = torch.nonzero(npm.buffer_pt_index >= 0).flatten()
    owners = npm.buffer_pt_index[slots]
    cells = (npm.neural_points[owners] / npm.resolution).floor().long()
    assert torch.equal(oc.voxel_hash(cells, cfg.buffer_size) % cfg.buffer_size, slots)
    assert owners.unique().numel() == owners.numel()


def test_pickle_round_trip_drops_the_cuda_cache():
    import pickle

    cfg = ncd128()
    cfg.device, cfg.buffer_size = "cpu", 50_021
    npm = NeuralPoints(cfg)
    npm.travel_dist = torch.zeros(1)
    npm.update(torch.rand(2000, 3) * 20, torch.zeros(3), torch.eye(3), 0)
    npm._brick_cache = {True: ("key", object())}
    npm.clear_temp()
    clone = pickle.loads(pickle.dumps(npm))
    assert clone.count() == npm.count() and clone._brick_cache == {}
    assert clone.buffer_pt_index is None  # like the reference, rebuilt by recreate_hash on load


def test_voxel_downsample_matches_oracle():
    gen = torch.Generator().manual_seed(3)
    pts = torch.rand(20000, 3, generator=gen) * torch.tensor([30.0, 25.0, 4.0]) - 10.0
    assert torch.equal(voxel_down_sample_torch(pts, 0.4), oc.voxel_downsample_indices(pts, 0.4))


def test_sample_pin_matches_reference_fixture():
    fx = gio.load("sampler", "ncd128")
    cfg = ncd128()
    cfg.device = "cpu"
    for k, v in json.loads(str(fx["cfg_sampler"])).items():
        setattr(cfg, k, v)
    torch.manual_seed(int(fx["seed"]))
    coord, label, normal, sem, color, weight = DataSampler(cfg).sample_pin(gio.t(fx["scan"]), None, None, None)
    assert normal is None and sem is None and color is None
    assert torch.equal(coord, gio.t(fx["coord"]))
    assert torch.equal(label, gio.t(fx["label"]))
    assert torch.equal(weight, gio.t(fx["weight"]))


def test_config_load_reads_hot_path_sections():
    text = """
process: {min_range_m: 1.0, max_range_m: 60.0, vox_down_m: 0.1}
sampler: {surface_sample_range_m: 0.25, surface_sample_n: 4, free_sample_begin_ratio: 0.5,
          free_sample_end_dist_m: 1.2, free_front_sample_n: 2}
neuralpoints: {voxel_size_m: 0.4, num_nei_cells: 2, search_alpha: 0.5, weighted_first: True}
loss: {sigma_sigmoid_m: 0.1, loss_weight_on: True, dist_weight_scale: 0.8}
continual: {batch_size_new_sample: 1000, pool_capacity: 1e7}
optimizer: {iters: 10, batch_size: 16384, learning_rate: 0.01, adaptive_iters: True}
"""
    with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as fh:
        fh.write(text)
    try:
        cfg = Config()
        cfg.load(fh.name)
    finally:
        os.unlink(fh.name)
    ref = ncd128()
    for key in ("voxel_size_m", "num_nei_cells", "search_alpha", "surface_sample_n", "free_front_n", "bs", "iters",
                "loss_weight_on", "bs_new_sample", "pool_capacity", "local_map_radius", "infer_bs", "sigma_sigmoid_m"):
        assert getattr(cfg, key) == getattr(ref, key), key
    assert cfg.numerical_grad and cfg.gradient_decimation == 10


def test_shard_bounds_tile_the_batch():
    for n in (0, 1, 7, 16384, 131072, 131075):
        for world in (1, 2, 3, 8):
            edges = [cdist.shard_bounds(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1
            for dec in (1, 10):
                assert sum(cdist.decimated_count(b, e, dec) for b, e in edges) == len(range(0, n, dec))


def _gloo_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dec_grad = torch.full((833,), float(rank + 1))
        loss = torch.tensor([1.0, 2.0, 3.0]) * (rank + 1)
        cdist.FlatAllReduce([dec_grad, None, loss])()
        feat = torch.zeros(10, 8)
        feat[rank] = 1.0
        cdist.all_reduce_sum(feat)
        touched = torch.zeros(10, dtype=torch.uint8)
        touched[rank] = 1
        cdist.all_reduce_max(touched)
        before = torch.arange(6, dtype=torch.float32)
        cert = before + (rank + 1) * 0.5
        ts = torch.tensor([rank, 5 - rank, 2], dtype=torch.int32)
        cdist.reduce_side_effects(cert, before, ts)
        torch.save({"dec": dec_grad, "loss": loss, "feat": feat, "touched": touched, "cert": cert, "ts": ts},
                   os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        torch.distributed.destroy_process_group()


def test_gradient_sync_with_two_gloo_ranks():
    world = 2
    port = 29500 + os.getpid() % 2000
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_gloo_worker, args=(world, port, tmp), nprocs=world, join=True)
        outs = [torch.load(os.path.join(tmp, f"r{r}.pt")) for r in range(world)]
    for o in outs:
        assert torch.equal(o["dec"], torch.full((833,), 3.0))
        assert torch.equal(o["loss"], torch.tensor([3.0, 6.0, 9.0]))
        assert o["feat"].sum().item() == 16.0 and o["feat"][0].sum() == 8 and o["feat"][1].sum() == 8
        assert o["touched"].tolist()[:3] == [1, 1, 0]
        assert torch.allclose(o["cert"], torch.arange(6, dtype=torch.float32) + 1.5)
        assert o["ts"].tolist() == [1, 5, 2]
    assert torch.equal(outs[0]["dec"], outs[1]["dec"])


def test_clid_sampler_and_local_map_match_reference_fixture():
    """LocalPointCloudMap.update_map / region_specific_sdf_estimation and DataSampler.sample against
    the reference (same torch seed on CPU => same random draws)."""
    from clid_slam_b200.model.local_point_cloud_map import LocalPointCloudMap
    from clid_slam_b200.utils.tools import transform_torch

    fx = gio.load("clidsampler", "ncd128")
    cfg = ncd128()
    cfg.device = "cpu"
    for k, v in json.loads(str(fx["cfg_sampler"])).items():
        setattr(cfg, k, v)
    lmap = LocalPointCloudMap(cfg)
    for i in range(2):
        pose = gio.t(fx[f"pose{i}"])
        lmap.update_map(pose[:3, 3].float(), transform_torch(gio.t(fx[f"scan{i}"]), pose))
        assert torch.equal(lmap.local_point_cloud_map, gio.t(fx[f"map{i}_points"]))
        table = torch.full((cfg.local_buffer_size,), -1, dtype=torch.int64)
        table[gio.t(fx[f"map{i}_slots"])] = gio.t(fx[f"map{i}_vals"])
        assert torch.equal(lmap.buffer_pt_index, table)
    d, mask = lmap.region_specific_sdf_estimation(gio.t(fx["probe"]))
    assert torch.equal(mask, gio.t(fx["probe_mask"]))
    gio.assert_close(d, fx["probe_dist"], 1e-4, 1e-6, "region-specific |sdf|", 2e-3)
    torch.manual_seed(int(fx["seed"]))
    coord, label, weight = DataSampler(cfg).sample(gio.t(fx["scan1"]), lmap, gio.t(fx["pose1"]))
    assert coord.shape == tuple(fx["coord"].shape)
    assert torch.equal(coord, gio.t(fx["coord"]))
    gio.assert_close(label, fx["label"], 1e-4, 1e-6, "labels", 2e-3)
    assert torch.equal(weight, gio.t(fx["weight"]))


def _exchange_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 3 slabs along x over a 60-cell line of points (one point per cell, res 1.0)
        pts = torch.stack((torch.arange(60, dtype=torch.float32) + 0.5, torch.zeros(60), torch.zeros(60)), 1)
        shards = cdist.SpatialShards(pts, 1.0, reach=2, world_size=world, axis=0)
        assert shards.pairwise
        grad = torch.zeros(61, 8)  # + padding row
        mine = (shards.row_owner == rank) | shards.shared_mask
        grad[mine] = float(rank + 1)  # every rank contributes to its slab and to all bands it can reach ...
        left, right = shards.neighbour_rows(rank)
        reach_rows = torch.zeros(61, dtype=torch.bool)
        reach_rows[shards.row_owner == rank] = True
        for rows in (left, right):
            if rows is not None:
                reach_rows[rows] = True
        grad[~reach_rows] = 0.0       # ... but only the bands of its own two boundaries
        ex = cdist.NeighbourExchange(shards, rank, grad)
        ex.pack(grad)
        ex.exchange()
        ex.unpack(grad)
        torch.save({"grad": grad, "owner": shards.row_owner, "bands": [b for b in shards.band_rows]},
                   os.path.join(out_dir, f"x{rank}.pt"))
    finally:
        torch.distributed.destroy_process_group()


def test_neighbour_exchange_with_three_gloo_ranks():
    """Band rows end up with the sum of both contributors on both ranks; private rows are untouched."""
    world = 3
    port = 31500 + os.getpid() % 2000
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_exchange_worker, args=(world, port, tmp), nprocs=world, join=True)
        outs = [torch.load(os.path.join(tmp, f"x{r}.pt")) for r in range(world)]
    bands = outs[0]["bands"]
    assert len(bands) == 2
    for j, rows in enumerate(bands):  # boundary j separates ranks j and j + 1
        want = float((j + 1) + (j + 2))
        for r in (j, j + 1):
            assert torch.all(outs[r]["grad"][rows] == want)
    owner = outs[0]["owner"]
    in_band = torch.zeros(61, dtype=torch.bool)
    for rows in bands:
        in_band[rows] = True
    for r in range(world):
        private = (owner == r) & ~in_band
        private[-1] = False
        assert torch.all(outs[r]["grad"][private] == float(r + 1))


def test_spatial_shards_bands_and_pairwise_flag():
    """Band rows partition correctly; slabs narrower than two bands switch the exchange to the flat all-reduce."""
    pts = torch.stack((torch.arange(80, dtype=torch.float32) + 0.5, torch.zeros(80), torch.zeros(80)), 1)
    wide = cdist.SpatialShards(pts, 1.0, reach=2, world_size=4, axis=0)
    assert wide.pairwise and len(wide.band_rows) == 3
    union = torch.cat(wide.band_rows).sort().values
    assert torch.equal(union, wide.shared_rows[wide.shared_rows < 80])  # the padding row is never shared
    for j, rows in enumerate(wide.band_rows):  # band j hugs boundary j: reach + margin = 3 cells either side
        b = int(wide.boundaries[j])
        assert rows.tolist() == list(range(b - 3, b + 3))
    left, right = wide.neighbour_rows(0)
    assert left is None and torch.equal(right, wide.band_rows[0])
    left, right = wide.neighbour_rows(3)
    assert right is None and torch.equal(left, wide.band_rows[2])
    narrow = cdist.SpatialShards(pts, 1.0, reach=2, world_size=20, axis=0)  # 4-cell slabs, 6-cell bands
    assert not narrow.pairwise
    with pytest.raises(ValueError):
        cdist.NeighbourExchange(narrow, 1, torch.zeros(81, 8))


@pytest.mark.parametrize("yaml_name", ["run_ncd128.yaml", "run_SubT_MRS.yaml", "run_quad.yaml"])
def test_config_load_matches_the_reference_loader(yaml_name):
    """Every attribute our Config carries gets the value the reference's Config.load gives it for the shipped run
    files (utils/config.py:410-910) -- in particular use_pin_mapper / track_on, which select the sampler and the
    pose source of Mapper.process_frame."""
    import os

    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("reference tree not available")
    ref = ref_loader.load()
    from clid_slam_b200.config import Config

    path = os.path.join(ref_loader.REFERENCE_ROOT, "config", yaml_name)
    theirs = ref.Config()
    theirs.load(path)
    ours = Config()
    ours.load(path)
    skipped = {"device", "dtype", "tran_dtype", "idx_dtype", "silence"}
    checked = 0
    for name, mine in vars(ours).items():
        if name.startswith("_") or name in skipped or not hasattr(theirs, name):
            continue
        want = getattr(theirs, name)
        if name == "track_on":
            want = bool(want)  # the reference stores the tracker section itself (truthy), utils/config.py:676
        if isinstance(mine, float) or isinstance(want, float):
            assert float(mine) == pytest.approx(float(want), rel=1e-12), name
        else:
            assert mine == want, name
        checked += 1
    assert checked > 50
    assert ours.use_pin_mapper is False and ours.track_on is True


def test_slab_boundaries_of_a_clustered_map_stay_strictly_increasing():
    """90 % of the points in one cell: the quantiles repeat, the shards must still have world-1 boundaries and
    every rank must find its neighbour bands (IndexError on ranks 2 and 3 otherwise)."""
    from clid_slam_b200.dist import SpatialShards

    gen = torch.Generator().manual_seed(0)
    pts = torch.cat((torch.rand(900, 3, generator=gen) * 0.3, torch.rand(100, 3, generator=gen) * 40.0))
    shards = SpatialShards(pts, 0.4, reach=2, world_size=4)
    b = shards.boundaries.tolist()
    assert len(b) == 3 and all(b[i] < b[i + 1] for i in range(2))
    assert len(shards.band_rows) == 3
    owner = shards.owner_of(pts)
    assert int(owner.min()) >= 0 and int(owner.max()) <= 3


def test_hash_owner_mask_keeps_the_last_writer_of_every_slot():
    slots = torch.tensor([5, 3, 5, 7, 3, 3, -2, 8])
    keep = cdist.hash_owner_mask(slots, 10)  # -2 wraps to slot 8: the later point wins it
    assert keep.tolist() == [False, False, True, True, False, True, False, True]


def _partition_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = torch.Generator().manual_seed(3)
        # one point per voxel of a 60 x 9 x 2 slab of space, in scrambled insertion order (the same on every rank)
        gx, gy, gz = torch.meshgrid(torch.arange(60), torch.arange(9), torch.arange(2), indexing="ij")
        cells = torch.stack((gx, gy, gz), -1).reshape(-1, 3).float()
        pts = (cells + 0.1 + 0.8 * torch.rand(cells.shape, generator=gen))[torch.randperm(cells.shape[0], generator=gen)]
        res, reach = 1.0, 2
        axis = 0
        bnd = cdist.slab_boundaries(cdist.axis_cells(pts, res, axis), world)
        mine = pts[cdist.partition_mask(pts, res, axis, bnd, rank, reach + 1)]
        shards = cdist.SpatialShards(mine, res, reach=reach, world_size=world, axis=axis, boundaries=bnd)
        tabs = cdist.peer_row_tables(shards, rank, mine, mine.shape[0] + 1)
        # certainty-like side effect: every rank adds (rank + 1) to the band rows it shares, then the bands are completed
        delta = torch.zeros(mine.shape[0])
        for rows in shards.neighbour_rows(rank):
            if rows is not None:
                delta[rows] += float(rank + 1)
        cdist.exchange_band_values(shards, rank, mine, delta, "sum")
        torch.save({"points": mine, "tabs": tabs, "owner": shards.row_owner, "bnd": bnd, "delta": delta,
                    "bands": shards.neighbour_rows(rank)}, os.path.join(out_dir, f"p{rank}.pt"))
    finally:
        torch.distributed.destroy_process_group()


def test_partitioned_map_tables_and_row_translation_with_three_gloo_ranks():
    """Every rank holds its slab plus the halves of its bands; the owned rows tile the map; the translation tables map a
    band row to the row of the SAME voxel in the neighbour's table and are -1 everywhere else."""
    world = 3
    port = 33500 + os.getpid() % 2000
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_partition_worker, args=(world, port, tmp), nprocs=world, join=True)
        outs = [torch.load(os.path.join(tmp, f"p{r}.pt")) for r in range(world)]
    total = 60 * 9 * 2
    owned = [o["points"][o["owner"][:-1] == r] for r, o in enumerate(outs)]
    keys = torch.cat([cdist.voxel_keys(p, 1.0) for p in owned])
    assert keys.numel() == total and torch.unique(keys).numel() == total  # tiles the map, nothing twice
    assert all(o["points"].shape[0] < 0.6 * total for o in outs)          # nobody holds the whole map
    for r, o in enumerate(outs):
        lo, hi = o["tabs"]
        assert (lo is None) == (r == 0) and (hi is None) == (r == world - 1)
        for tab, peer in ((lo, r - 1), (hi, r + 1)):
            if tab is None:
                continue
            assert tab.dtype == torch.int32 and tab.numel() == o["points"].shape[0] + 1 and int(tab[-1]) == -1
            rows = torch.nonzero(tab >= 0).flatten()
            band = 3 * 2 * 9 * 2  # (reach + margin) cells either side of the boundary, 9 x 2 voxels per cell layer
            assert rows.numel() == band
            mine_keys = cdist.voxel_keys(o["points"][rows], 1.0)
            theirs_keys = cdist.voxel_keys(outs[peer]["points"][tab[rows].long()], 1.0)
            assert torch.equal(mine_keys, theirs_keys)
        # band side effects: both contributors' increments on both sides, nothing on private rows
        left, right = o["bands"]
        expect = torch.zeros(o["points"].shape[0])
        if left is not None:
            expect[left] += float(r + 1) + float(r)
        if right is not None:
            expect[right] += float(r + 1) + float(r + 2)
        assert torch.equal(o["delta"], expect)
