"""Multi-GPU check (torchrun, N >= 2): Mapper.mapping() sharded over the ranks (Mapper.set_shards) on a REPLICATED map
trains like the single-process call on the same replay pool -- same loss level after the same number of iterations
(the draws differ: every rank draws config.bs // N samples of its own slab), the trained features of every rank's slab
and the certainty / ts_update side effects are re-replicated, so all ranks end with the same map."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clid_slam_b200.config import ncd128
from clid_slam_b200.model.decoder import Decoder
from clid_slam_b200.model.local_point_cloud_map import LocalPointCloudMap
from clid_slam_b200.model.neural_points import NeuralPoints
from clid_slam_b200.utils.mapper import Mapper
from clid_slam_b200.dist import SpatialShards

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); device = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(device))


class FakeDataset:
    lose_track = False; stop_status = False; processed_frame = 0; gt_pose_provided = True; pgo_poses = None; static_mask = None
    def __init__(self, n):
        self.gt_poses = self.odom_poses = np.tile(np.eye(4), (n, 1, 1))


def scan(gen, n=30000):
    xy = (torch.rand(n, 2, generator=gen, device=device) - 0.5) * 80
    floor = torch.cat((xy, -1.5 + 0.3 * torch.sin(xy[:, :1] / 3)), dim=1)
    return floor[floor.norm(dim=1) > 1.0]


def run(sharded, mode):
    torch.manual_seed(42)
    cfg = ncd128(); cfg.device = device; cfg.use_pin_mapper = True; cfg.feature_std = 0.0
    cfg.local_buffer_size = 500_009; cfg.buffer_size = 2_000_003
    if mode == "analytic":
        cfg.numerical_grad, cfg.gradient_decimation = False, 1
    dec = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    npm = NeuralPoints(cfg)
    ds = FakeDataset(1)
    mapper = Mapper(cfg, ds, npm, LocalPointCloudMap(cfg), dec)
    gen = torch.Generator(device=device).manual_seed(5)  # the same scan, hence the same map and pool, on every rank
    npm.travel_dist = torch.zeros(1, device=device)
    mapper.process_frame(scan(gen), None, torch.eye(4, device=device, dtype=torch.float64), 0)
    if sharded:
        mapper.set_shards(SpatialShards(npm.local_neural_points, cfg.voxel_size_m, reach=cfg.num_nei_cells, world_size=world))
    for _ in range(3):
        mapper.mapping(20)
    torch.cuda.synchronize()
    return mapper.last_losses[-5:].mean(0).cpu(), npm.geo_features.clone(), npm.point_certainties.clone(), npm.point_ts_update.clone(), npm.local_count()


for mode in ("analytic", "numerical"):
    l1, f1, c1, t1, rows = run(False, mode)
    ls, fs, cs, ts, _ = run(True, mode)
    # every rank must hold the same map afterwards
    ref = [fs.clone(), cs.clone()]
    dist.broadcast(ref[0], 0); dist.broadcast(ref[1], 0)
    same = bool(torch.equal(ref[0], fs) and torch.equal(ref[1], cs))
    flags = [None] * world
    dist.all_gather_object(flags, same)
    if rank == 0:
        print(f"{mode}: {rows} local neural points, N={world}; mean of the last 5 losses [total, bce, eikonal]: single {l1.tolist()} sharded {ls.tolist()}; "
              f"all ranks hold identical features / certainties afterwards: {all(flags)}; "
              f"certainty mass single {float(c1.sum()):.1f} sharded {float(cs.sum()):.1f}; trained rows single {int((f1.abs().sum(1) > 0).sum())} "
              f"sharded {int((fs.abs().sum(1) > 0).sum())}", flush=True)
        assert all(flags)
        assert abs(float(ls[0]) - float(l1[0])) < 0.1 * float(l1[0])
        assert abs(float(cs.sum()) - float(c1.sum())) < 0.02 * float(c1.sum())
dist.barrier(); dist.destroy_process_group()
