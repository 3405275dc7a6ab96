"""Multi-GPU check (torchrun, N >= 2): the peer-memory step (band gradients added into the slab neighbour's table
by the fused kernel, one-shot peer all-reduce of [decoder grads | loss]), the NCCL neighbour send/recv exchange
and the flat all-reduce over all bands give the same training trajectory, and all agree with a single-process
run on the union batch (up to fp32 summation order)."""
import copy, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clid_slam_b200.config import ncd128
from clid_slam_b200.model.decoder import Decoder
from clid_slam_b200.model.neural_points import NeuralPoints
from clid_slam_b200.ops.train import FusedTrainer
from clid_slam_b200.dist import SpatialShards
from clid_slam_b200.synth import wavy_sheets, sample_batch

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); device = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(device))


def build():
    torch.manual_seed(42)
    cfg = ncd128(); cfg.device = device; cfg.feature_std = 0.05; cfg.local_map_radius = 1e4
    cfg.numerical_grad = False; cfg.gradient_decimation = 1
    dec = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    npm = NeuralPoints(cfg); npm.travel_dist = torch.zeros(1, device=device)
    gen = torch.Generator(device=device).manual_seed(1)
    npm.update(wavy_sheets(240, 2, cfg.voxel_size_m, gen, device=device), torch.zeros(3, device=device),
               torch.eye(3, device=device), 0)
    return cfg, dec, npm


n = 16384
results = {}
for mode in ("peer", "p2p", "flat", "single"):
    cfg, dec, npm = build()
    shards = SpatialShards(npm.local_neural_points, cfg.voxel_size_m, reach=cfg.num_nei_cells, world_size=world)
    if mode == "flat":
        shards.pairwise = False
    trainer = FusedTrainer(cfg, npm, dec)
    if mode == "peer":
        trainer.attach_peers(shards)
    losses = []
    for it in range(6):
        # identical global batch on every rank, split by slab ownership
        gen = torch.Generator(device=device).manual_seed(100 + it)
        x, label, weight, ts = sample_batch(npm.neural_points, n * world, gen)
        if mode == "single":
            loss = trainer.iteration(x, label, ts, weight)
        else:
            mine = shards.owner_of(x) == rank
            loss = trainer.iteration(x[mine], label[mine], ts[mine], weight[mine], n_global=n * world,
                                     shards=None if mode == "peer" else shards)
        losses.append(loss.clone())
    feats = npm.local_geo_features.data.clone()
    if mode != "single":
        shards.gather_features(feats, rank)
    results[mode] = (torch.stack(losses), feats, torch.cat([p.data.flatten() for p in dec.flat_parameters()]))
    if mode == "peer":
        trainer.peer.check()
    assert mode in ("single", "flat", "peer") or trainer.neighbour_exchange(shards) is not None

torch.cuda.synchronize()
if rank == 0:
    def cmp(a, b, what):
        la, fa, da = results[a]; lb, fb, db = results[b]
        dl = ((la - lb).abs() / lb.abs().clamp_min(1e-12)).max().item()
        df = (fa - fb).abs()
        dd = ((da - db).abs()).max().item()
        bad = (df > 1e-5 + 1e-3 * fb.abs()).double().mean().item()
        print(f"{what}: max rel loss diff {dl:.2e}; features max abs diff {df.max().item():.2e}, outside 1e-3 rel + 1e-5: {bad * 100:.3f} %; "
              f"decoder max abs diff {dd:.2e}")
    cmp("peer", "single", f"N={world} peer-memory step vs single process")
    cmp("peer", "p2p", f"N={world} peer-memory step vs NCCL neighbour exchange")
    cmp("p2p", "flat", f"N={world} neighbour exchange vs flat all-reduce")
    cmp("p2p", "single", f"N={world} neighbour exchange vs single process")
    cmp("flat", "single", f"N={world} flat all-reduce vs single process")
dist.barrier(); dist.destroy_process_group()
