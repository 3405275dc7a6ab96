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
from clid_slam_b200.utils.tools import voxel_down_sample_torch
from clid_slam_b200.dist import hash_owner_mask, SpatialShards, axis_cells, partition_mask, peer_row_tables, slab_boundaries, voxel_keys
from clid_slam_b200.synth import wavy_sheets, sample_batch, set_voxel_features

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); device = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(device))


def build(partitioned=False):
    """The whole map on every rank, or (partitioned) only this rank's slab plus the halves of its boundary bands;
    features are a function of the voxel, so every table that holds a voxel starts from the same values."""
    torch.manual_seed(42)
    cfg = ncd128(); cfg.device = device; cfg.feature_std = 0.05; cfg.local_map_radius = 1e4
    cfg.numerical_grad = False; cfg.gradient_decimation = 1
    dec = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    npm = NeuralPoints(cfg); npm.travel_dist = torch.zeros(1, device=device)
    gen = torch.Generator(device=device).manual_seed(1)
    pts = wavy_sheets(240, 2, cfg.voxel_size_m, gen, device=device)
    # one point per voxel, and only the points that own their hash slot (see dist.hash_owner_mask), in every mode
    pts = pts[voxel_down_sample_torch(pts, cfg.voxel_size_m)]
    pts = pts[hash_owner_mask(npm._slots_of(pts), npm.buffer_size)]
    geom = None
    if partitioned:
        axis = int(torch.argmax(pts.amax(0) - pts.amin(0)).item())
        bnd = slab_boundaries(axis_cells(pts, cfg.voxel_size_m, axis), world)
        pts = pts[partition_mask(pts, cfg.voxel_size_m, axis, bnd, rank, cfg.num_nei_cells + 1)]
        geom = (axis, bnd)
    npm.update(pts, torch.zeros(3, device=device), torch.eye(3, device=device), 0)
    set_voxel_features(npm, cfg.feature_std)
    return cfg, dec, npm, geom


def by_voxel(points, feats, res):
    key = voxel_keys(points, res)
    order = torch.argsort(key)
    return key[order], feats[order]


n = 16384
results = {}
anchors = build()[2].neural_points.clone()  # sample anchors: the full map's points, in every mode
for mode in ("peer", "partition", "p2p", "flat", "single"):
    cfg, dec, npm, geom = build(partitioned=mode == "partition")
    if mode == "partition":
        shards = SpatialShards(npm.local_neural_points, cfg.voxel_size_m, reach=cfg.num_nei_cells, world_size=world,
                               axis=geom[0], boundaries=geom[1])
    else:
        shards = SpatialShards(npm.local_neural_points, cfg.voxel_size_m, reach=cfg.num_nei_cells, world_size=world)
    if mode == "flat":
        shards.pairwise = False
    trainer = FusedTrainer(cfg, npm, dec)
    if mode == "peer":
        trainer.attach_peers(shards)
    if mode == "partition":
        tabs = peer_row_tables(shards, rank, npm.local_neural_points, trainer.rows)
        trainer.attach_peers(shards, peer_rows=tabs)
        print(f"rank {rank}: partitioned table {npm.local_neural_points.shape[0]} of {anchors.shape[0]} neural points, "
              f"band rows {[None if t is None else int((t >= 0).sum()) for t in tabs]}", flush=True)
    losses = []
    for it in range(6):
        # identical global batch on every rank, split by slab ownership
        gen = torch.Generator(device=device).manual_seed(100 + it)
        x, label, weight, ts = sample_batch(anchors, n * world, gen)
        if mode == "single":
            loss = trainer.iteration(x, label, ts, weight)
        else:
            mine = shards.owner_of(x) == rank
            loss = trainer.iteration(x[mine], label[mine], ts[mine], weight[mine], n_global=n * world,
                                     shards=None if mode in ("peer", "partition") else shards)
        losses.append(loss.clone())
    feats = npm.local_geo_features.data.clone()
    if mode == "partition":
        # every rank contributes the rows it owns, named by voxel; rank 0 puts them into the single-process order
        own = shards.row_owner[:-1] == rank
        parts = [None] * world
        dist.all_gather_object(parts, (voxel_keys(npm.local_neural_points[own], cfg.voxel_size_m).cpu(), feats[:-1][own].cpu()))
        keys = torch.cat([p[0] for p in parts]); vals = torch.cat([p[1] for p in parts])
        order = torch.argsort(keys)
        ref_keys = voxel_keys(anchors, cfg.voxel_size_m)
        assert torch.equal(keys[order], torch.sort(ref_keys.cpu()).values), "the partitions do not tile the map"
        full = torch.zeros(anchors.shape[0] + 1, feats.shape[1], device=device)
        full[torch.argsort(ref_keys)] = vals[order].to(device)
        feats = full
    elif mode != "single":
        shards.gather_features(feats, rank)
    results[mode] = (torch.stack(losses), feats, torch.cat([p.data.flatten() for p in dec.flat_parameters()]))
    if mode in ("peer", "partition"):
        trainer.peer.check()
    assert mode in ("single", "flat", "peer", "partition") or trainer.neighbour_exchange(shards) is not None

torch.cuda.synchronize()
if rank == 0:
    def cmp(a, b, what):
        la, fa, da = results[a]; lb, fb, db = results[b]
        if "partition" in (a, b):  # the padding row is not a neural point: no partition owns it
            fa, fb = fa[:-1], fb[:-1]
        dl = ((la - lb).abs() / lb.abs().clamp_min(1e-12)).max().item()
        df = (fa - fb).abs()
        dd = ((da - db).abs()).max().item()
        bad = (df > 1e-5 + 1e-3 * fb.abs()).double().mean().item()
        print(f"{what}: max rel loss diff {dl:.2e}; features max abs diff {df.max().item():.2e}, outside 1e-3 rel + 1e-5: {bad * 100:.3f} %; "
              f"decoder max abs diff {dd:.2e}")
    cmp("peer", "single", f"N={world} peer-memory step vs single process")
    cmp("partition", "single", f"N={world} PARTITIONED map (per-rank tables, row translation) vs single process")
    cmp("peer", "p2p", f"N={world} peer-memory step vs NCCL neighbour exchange")
    cmp("p2p", "flat", f"N={world} neighbour exchange vs flat all-reduce")
    cmp("p2p", "single", f"N={world} neighbour exchange vs single process")
    cmp("flat", "single", f"N={world} flat all-reduce vs single process")
dist.barrier(); dist.destroy_process_group()
