import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from clid_slam_b200.config import ncd128
from clid_slam_b200.model.decoder import Decoder
from clid_slam_b200.model.local_point_cloud_map import LocalPointCloudMap
from clid_slam_b200.model.neural_points import NeuralPoints
from clid_slam_b200.synth import wavy_sheets
from clid_slam_b200.utils.mapper import Mapper
from clid_slam_b200.utils import tools
device="cuda:0"
class _DS:
    lose_track=False; stop_status=False; processed_frame=0; gt_pose_provided=True; pgo_poses=None; static_mask=None
torch.manual_seed(42)
cfg=ncd128(); cfg.device=device; cfg.feature_std=0.05; cfg.local_map_radius=72.0
dec=Decoder(cfg,64,1,1); npm=NeuralPoints(cfg)
gen=torch.Generator(device=device).manual_seed(1)
world=wavy_sheets(600,1,0.4,gen,device=device)
ds=_DS(); ds.gt_poses=ds.odom_poses=np.tile(np.eye(4),(6,1,1))
npm.travel_dist=torch.zeros(1,device=device)
npm.update(world, torch.zeros(3,device=device), torch.eye(3,device=device), 0)
lpm=LocalPointCloudMap(cfg)
mapper=Mapper(cfg,ds,npm,lpm,dec)
def tm(f):
    torch.cuda.synchronize(); t=time.perf_counter(); r=f(); torch.cuda.synchronize(); return (time.perf_counter()-t)*1e3, r
for frame in range(4):
    ds.processed_frame=frame
    pose=torch.eye(4,device=device,dtype=torch.float64); pose[0,3],pose[2,3]=0.5*frame,2.0
    ds.gt_poses[frame]=pose.cpu().numpy()
    npm.travel_dist=torch.arange(frame+1,device=device,dtype=torch.float32)*0.5
    origin=pose[:3,3].float()
    near=world[(world-origin).norm(dim=1)<cfg.max_range]
    pick=torch.randint(0,near.shape[0],(30000,),generator=gen,device=device)
    scan=near[pick]+0.01*torch.randn(30000,3,generator=gen,device=device)-origin
    t_lpm,_=tm(lambda: lpm.update_map(origin, tools.transform_torch(scan,pose)))
    t_s,(coord,label,weight)=tm(lambda: mapper.sampler.sample(scan,lpm,pose))
    seeds=tools.transform_torch(scan,pose)
    t_u,_=tm(lambda: npm.update(seeds, origin, pose[:3,:3], frame))
    t_idx,_=tm(lambda: npm.brick_index(True))
    npm.set_search_neighborhood(num_nei_cells=1, search_alpha=0.0)
    t_c,_=tm(lambda: npm.query_certainty(tools.transform_torch(coord,pose)))
    npm.set_search_neighborhood(num_nei_cells=cfg.num_nei_cells, search_alpha=cfg.search_alpha)
    t_pf,_=tm(lambda: mapper.process_frame(scan,None,pose,frame))
    print(f"frame {frame}: local-pc-map update {t_lpm:.2f} | sampler.sample {t_s:.2f} | npm.update {t_u:.2f} (local {npm.local_count()}) | brick index {t_idx:.2f} | query_certainty {t_c:.2f} | whole process_frame {t_pf:.2f} ms", flush=True)
