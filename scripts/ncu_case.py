"""Dev tool: launch one hot-path kernel a few times on the bench world, for ncu captures.

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 3 -c 1 -o gpurun_out/x \
        python scripts/ncu_case.py --case fwd
"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clid_slam_b200.config import ncd128
from clid_slam_b200.model.decoder import Decoder
from clid_slam_b200.model.neural_points import NeuralPoints
from clid_slam_b200 import fused
from clid_slam_b200.ops.train import FusedTrainer
from clid_slam_b200.synth import wavy_sheets, sample_batch

ap = argparse.ArgumentParser()
ap.add_argument("--case", default="fwd", choices=["fwd", "fwd_far", "fwd_same", "fused", "fused_far", "fused_num", "step"])
ap.add_argument("--n", type=int, default=131072)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
torch.manual_seed(42)
cfg = ncd128(); cfg.device = "cuda"; cfg.feature_std = 0.05; cfg.local_map_radius = 1e4
cfg.numerical_grad = args.case == "fused_num"
cfg.gradient_decimation = 10 if cfg.numerical_grad else 1
dec = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
npm = NeuralPoints(cfg)
npm.travel_dist = torch.zeros(1, device="cuda")
gen = torch.Generator(device="cuda").manual_seed(1)
pts = wavy_sheets(520, 4, cfg.voxel_size_m, gen, device="cuda")
npm.update(pts, torch.zeros(3, device="cuda"), torch.eye(3, device="cuda"), 0)
x, label, weight, ts = sample_batch(npm.neural_points, args.n, gen)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
if args.case.endswith("_far"):
    x = x + 1000.0
if args.case.endswith("_same"):
    x = x[:1].repeat(args.n, 1).contiguous()
trainer = FusedTrainer(cfg, npm, dec) if args.case.startswith("fused") or args.case == "step" else None
if trainer is not None:
    trainer.touched = None  # what StepPipeline / bench.py run at this batch size: dense Adam, no touched-flag stores
for i in range(args.reps):
    flush.zero_()
    if args.case.startswith("fwd"):
        fused.sdf_and_gradient(npm, dec, x)
    else:
        trainer.iteration(x, label, ts, weight, apply_step=args.case == "step")
torch.cuda.synchronize()
