"""Dev tool: time the fused forward (sdf + grad) kernel on a synthetic world.  GPU only."""
import argparse, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clid_slam_b200.config import ncd128
from clid_slam_b200.model.decoder import Decoder
from clid_slam_b200.model.neural_points import NeuralPoints
from clid_slam_b200 import fused
from clid_slam_b200.synth import wavy_sheets, sample_batch

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=131072)
ap.add_argument("--side", type=int, default=520)
ap.add_argument("--sheets", type=int, default=4)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--bricks", type=int, default=0)
ap.add_argument("--sort", type=int, default=0)
args = ap.parse_args()

torch.manual_seed(42)
cfg = ncd128(); cfg.device = "cuda"; cfg.feature_std = 0.05; cfg.local_map_radius = 1e4
dec = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
npm = NeuralPoints(cfg)
npm.travel_dist = torch.zeros(1, device="cuda")
gen = torch.Generator(device="cuda").manual_seed(1)
pts = wavy_sheets(args.side, args.sheets, cfg.voxel_size_m, gen, device="cuda")
t0 = time.time(); npm.update(pts, torch.zeros(3, device="cuda"), torch.eye(3, device="cuda"), 0); torch.cuda.synchronize()
print(f"map: {npm.count()} points, local {npm.local_count()}, update {time.time()-t0:.2f}s")
x, label, weight, ts = sample_batch(npm.neural_points, args.n, gen)
if args.sort:
    key = ((x / (cfg.voxel_size_m * 8)).floor().long() * torch.tensor([1, 4096, 4096 * 4096], device="cuda")).sum(-1)
    x = x[torch.argsort(key)].contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for flush_l2 in (True, False):
    times = []
    for it in range(args.iters + 3):
        if flush_l2: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sdf, grad, nn, cert = fused.sdf_and_gradient(npm, dec, x, use_bricks=bool(args.bricks))
        e1.record(); torch.cuda.synchronize()
        if it >= 3: times.append(e0.elapsed_time(e1))
    times.sort(); med = times[len(times) // 2]
    nv = nn.float().mean().item()
    b_alg = 576 + 16 * nv
    print(f"flush_l2={flush_l2}: median {med*1e3:.1f} us  {args.n/med/1e3:.1f} M samples/s  mean nn {nv:.2f} "
          f"alg {b_alg:.0f} B/sample -> {args.n*b_alg/med/1e6:.1f} GB/s = {args.n*b_alg/med/1e6/6551.7*100:.1f}% of 6551.7")

# back-to-back launches, no host sync in between (steady state of a mapping loop, map L2-resident)
for reps in (20, 100):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fused.sdf_and_gradient(npm, dec, x, use_bricks=bool(args.bricks))
    e1.record(); torch.cuda.synchronize()
    print(f"back-to-back x{reps}: {e0.elapsed_time(e1)/reps*1e3:.1f} us per launch")

# diagnostics: where does the time go?  (a) far queries: no candidates at all (search ~free, MLP still runs)
# (b) all queries identical: every load hits L1 (instruction-bound time of the full path)
def bb(xq, reps=20):
    # the launches are replayed from a CUDA graph: the Python / ctypes cost of a call (~20-50 us) is not in the number
    for _ in range(3):
        fused.sdf_and_gradient(npm, dec, xq, use_bricks=bool(args.bricks))
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            out = fused.sdf_and_gradient(npm, dec, xq, use_bricks=bool(args.bricks))
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
far = x + 1000.0
same = x[:1].repeat(args.n, 1).contiguous()
srt = x[torch.argsort(((x / 1.6).floor().long() * torch.tensor([1, 4096, 4096 * 4096], device="cuda")).sum(-1))].contiguous()
print(f"diag: normal {bb(x):.1f} us | brick-sorted queries {bb(srt):.1f} us | far (no candidates) {bb(far):.1f} us | identical queries {bb(same):.1f} us")
