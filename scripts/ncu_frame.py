"""Dev tool: one frame of the shipped configuration (process_frame + mapping(10)) between cudaProfilerStart / Stop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/frame_launches.csv python scripts/ncu_frame.py
"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clid_slam_b200.config import ncd128
from clid_slam_b200.model.decoder import Decoder
from clid_slam_b200.model.local_point_cloud_map import LocalPointCloudMap
from clid_slam_b200.model.neural_points import NeuralPoints
from clid_slam_b200.synth import wavy_sheets
from clid_slam_b200.utils.mapper import Mapper

device = "cuda:0"


class _DS:
    lose_track = False; stop_status = False; processed_frame = 0; gt_pose_provided = True; pgo_poses = None; static_mask = None


torch.manual_seed(42)
cfg = ncd128(); cfg.device = device; cfg.feature_std = 0.05; cfg.local_map_radius = 72.0
dec = Decoder(cfg, 64, 1, 1); npm = NeuralPoints(cfg)
gen = torch.Generator(device=device).manual_seed(1)
world = wavy_sheets(600, 1, 0.4, gen, device=device)
ds = _DS(); ds.gt_poses = ds.odom_poses = np.tile(np.eye(4), (6, 1, 1))
npm.travel_dist = torch.zeros(1, device=device)
npm.update(world, torch.zeros(3, device=device), torch.eye(3, device=device), 0)
mapper = Mapper(cfg, ds, npm, LocalPointCloudMap(cfg), dec)
for frame in range(4):
    ds.processed_frame = frame
    pose = torch.eye(4, device=device, dtype=torch.float64); pose[0, 3], pose[2, 3] = 0.5 * frame, 2.0
    ds.gt_poses[frame] = pose.cpu().numpy()
    npm.travel_dist = torch.arange(frame + 1, device=device, dtype=torch.float32) * 0.5
    origin = pose[:3, 3].float()
    near = world[(world - origin).norm(dim=1) < cfg.max_range]
    pick = torch.randint(0, near.shape[0], (30000,), generator=gen, device=device)
    scan = near[pick] + 0.01 * torch.randn(30000, 3, generator=gen, device=device) - origin
    torch.cuda.synchronize()
    if frame == 3:
        torch.cuda.profiler.start()
    mapper.process_frame(scan, None, pose, frame)
    mapper.mapping(10)
    torch.cuda.synchronize()
    if frame == 3:
        torch.cuda.profiler.stop()
print("local points", npm.local_count(), "pool", mapper.pool_sample_count)
