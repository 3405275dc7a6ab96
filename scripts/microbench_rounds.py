"""Dev tool: forward time against the number of queries around the grid's round boundaries (tile quantisation)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clid_slam_b200.config import ncd128
from clid_slam_b200.model.decoder import Decoder
from clid_slam_b200.model.neural_points import NeuralPoints
from clid_slam_b200 import fused
from clid_slam_b200.synth import wavy_sheets, sample_batch

torch.manual_seed(42)
cfg = ncd128(); cfg.device = "cuda"; cfg.feature_std = 0.05; cfg.local_map_radius = 1e4
dec = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
npm = NeuralPoints(cfg); npm.travel_dist = torch.zeros(1, device="cuda")
gen = torch.Generator(device="cuda").manual_seed(1)
npm.update(wavy_sheets(520, 4, cfg.voxel_size_m, gen, device="cuda"), torch.zeros(3, device="cuda"), torch.eye(3, device="cuda"), 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
far = "--far" in sys.argv  # queries outside the map: no candidates, the decoder still runs
args = [a for a in sys.argv[1:] if a != "--far"]
sizes = [int(v) for v in (args or "37888 75776 98304 113664 131072 151552 196608 227328".split())]
big = sample_batch(npm.neural_points, max(sizes), gen)[0]
if far:
    big = big + 1000.0
out = []
for n in sizes:
    x = big[:n].contiguous()
    for _ in range(3):
        fused.sdf_and_gradient(npm, dec, x)
    torch.cuda.synchronize(); torch.cuda._sleep(int(2e7))
    evs = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fused.sdf_and_gradient(npm, dec, x); e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)[10]
    out.append(f"n {n:7d} tiles {n // 32:5d}: {t:6.1f} us  {n / t:7.0f} M/s")
print("\n".join(out))
