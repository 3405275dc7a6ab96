"""Dev tool: phase times of the spatially sharded step (run under torchrun, N >= 2)."""
import os, sys, statistics
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from clid_slam_b200.ops.train import FusedTrainer, StepPipeline
from clid_slam_b200.dist import SpatialShards
from clid_slam_b200.synth import sample_batch

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); device = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(device))
cfg, dec, npm, _ = B.build_world(device, "analytic")
gen = torch.Generator(device=device).manual_seed(1000 + rank)
shards = SpatialShards(npm.local_neural_points, cfg.voxel_size_m, reach=cfg.num_nei_cells, world_size=world)
own = npm.neural_points[shards.row_owner[:-1] == rank]
batches = []
for _ in range(4):
    cand = sample_batch(own, int(B.BATCH * 1.25), gen)
    keep = torch.nonzero(shards.owner_of(cand[0]) == rank).flatten()[:B.BATCH]
    batches.append(tuple(t[keep].contiguous() for t in cand))
trainer = FusedTrainer(cfg, npm, dec)
p2p_group = dist.new_group() if os.environ.get("CLID_P2P_GROUP", "1") == "1" else None
pipe = StepPipeline(trainer, B.BATCH, n_global=B.BATCH * world, buffers=batches, shards=shards, p2p_group=p2p_group)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
fr = torch.zeros(64 << 20, dtype=torch.float32, device=device)
stream = torch.cuda.current_stream()
ev = []
import time
host = []
for i in range(30):
    flush.zero_(); fr.sum()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    k = i % 4
    h0 = time.perf_counter()
    e[0].record(); pipe.graphs[k].replay(); e[1].record()
    h1 = time.perf_counter()
    work = dist.all_reduce(pipe.flats[k], async_op=True)
    h2 = time.perf_counter()
    if pipe.exchange is not None:
        pipe.exchange.exchange()
    work.wait()
    h3 = time.perf_counter()
    e[2].record()
    pipe.post_graphs[k].replay(); e[3].record()
    h4 = time.perf_counter()
    host.append((h1 - h0, h2 - h1, h3 - h2, h4 - h3))
    ev.append(e)
torch.cuda.synchronize()
ev = ev[5:]
med = lambda a, b: statistics.median(x[a].elapsed_time(x[b]) * 1e3 for x in ev)
hm = [statistics.median(h[j] for h in host[5:]) * 1e6 for j in range(4)]
print(f"rank {rank}: host enqueue us: graphA {hm[0]:.0f} | all_reduce {hm[1]:.0f} | exchange {hm[2]:.0f} | graphB {hm[3]:.0f} | sum {sum(hm):.0f}", flush=True)
print(f"rank {rank}: pre-graph {med(0,1):.1f} us | all-reduce {med(1,2):.1f} us ({pipe.flats[0].numel()*4/1e3:.0f} kB) | post-graph {med(2,3):.1f} us | total {med(0,3):.1f} us", flush=True)
dist.barrier(); dist.destroy_process_group()
