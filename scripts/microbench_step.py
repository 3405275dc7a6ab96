"""Dev tool: GPU-bound timings of the hot-path kernels on the bench world (GPU only).

Every measurement queues its launches behind a spin kernel (torch.cuda._sleep) so the host runs
ahead of the device and CUDA events see device time only; the L2 is either flushed before every
launch (cold, what bench.py reports) or left warm.
"""
import argparse, os, sys, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clid_slam_b200.config import ncd128
from clid_slam_b200.model.decoder import Decoder
from clid_slam_b200.model.neural_points import NeuralPoints
from clid_slam_b200 import fused
from clid_slam_b200.ops.train import FusedTrainer
from clid_slam_b200.synth import wavy_sheets, sample_batch

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=131072)
ap.add_argument("--side", type=int, default=520)
ap.add_argument("--sheets", type=int, default=4)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--mode", default="analytic")
ap.add_argument("--diag", type=int, default=1)
ap.add_argument("--mean", type=int, default=0, help="report the 10 %-trimmed mean instead of the median")
args = ap.parse_args()

torch.manual_seed(42)
cfg = ncd128(); cfg.device = "cuda"; cfg.feature_std = 0.05; cfg.local_map_radius = 1e4
cfg.numerical_grad = args.mode == "numerical"
cfg.gradient_decimation = 10 if cfg.numerical_grad else 1
dec = Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
npm = NeuralPoints(cfg)
npm.travel_dist = torch.zeros(1, device="cuda")
gen = torch.Generator(device="cuda").manual_seed(1)
pts = wavy_sheets(args.side, args.sheets, cfg.voxel_size_m, gen, device="cuda")
npm.update(pts, torch.zeros(3, device="cuda"), torch.eye(3, device="cuda"), 0)
batches = [sample_batch(npm.neural_points, args.n, gen) for _ in range(4)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=args.reps, cold=True):
    """median device time of fn(i) in us"""
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    torch.cuda._sleep(int(2e7))  # ~10 ms: the host gets ahead of the device
    evs = []
    for i in range(reps):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(i); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
    if args.mean:  # CUDA event stamps tick every ~1.9 us here: a median sits on that grid, the trimmed mean resolves finer
        k = len(ts) // 10
        core = ts[k:len(ts) - k] if len(ts) > 2 * k else ts
        return sum(core) / len(core)
    return ts[len(ts) // 2]


def fwd(xq):
    return lambda i: fused.sdf_and_gradient(npm, dec, xq)


x0 = batches[0][0]
nn = fused.sdf_and_gradient(npm, dec, x0)[2].float().mean().item()
b_fwd = 576 + 16 * nn
t = timed(lambda i: fused.sdf_and_gradient(npm, dec, batches[i % 4][0]))
print(f"forward sdf+grad cold: {t:.1f} us  {args.n / t:.0f} M/s  {args.n * b_fwd / t / 1e3:.0f} GB/s = {args.n * b_fwd / t / 1e3 / 6546.6 * 100:.1f}%")
t = timed(lambda i: fused.sdf_and_gradient(npm, dec, batches[i % 4][0]), cold=False)
print(f"forward sdf+grad warm: {t:.1f} us")
if args.diag:
    far = x0 + 1000.0
    same = x0[:1].repeat(args.n, 1).contiguous()
    key = ((x0 / 1.6).floor().long() * torch.tensor([1, 4096, 4096 * 4096], device="cuda")).sum(-1)
    srt = x0[torch.argsort(key)].contiguous()
    print(f"diag warm: far {timed(fwd(far), cold=False):.1f} | identical {timed(fwd(same), cold=False):.1f} | "
          f"brick-sorted {timed(fwd(srt), cold=False):.1f} us ; cold: brick-sorted {timed(fwd(srt)):.1f}")
    t = timed(lambda i: fused.sdf_and_gradient(npm, dec, x0, with_gradient=False, with_certainty=False), cold=False)
    print(f"forward sdf only warm: {t:.1f} us")

trainer = FusedTrainer(cfg, npm, dec)


def fused_only(i):
    x, label, weight, ts = batches[i % 4]
    trainer.iteration(x, label, ts, weight, apply_step=False)


def full_step(i):
    x, label, weight, ts = batches[i % 4]
    trainer.iteration(x, label, ts, weight)


print(f"train_fused ({args.mode}) cold: {timed(fused_only):.1f} us | warm: {timed(fused_only, cold=False):.1f} us")
print(f"adam cold: {timed(lambda i: trainer.adam_step()):.1f} us | warm: {timed(lambda i: trainer.adam_step(), cold=False):.1f} us")
print(f"full step cold: {timed(full_step):.1f} us | warm: {timed(full_step, cold=False):.1f} us")
if args.diag:
    xf = x0 + 1000.0
    lab, wgt, tss = batches[0][1], batches[0][2], batches[0][3]
    print(f"train_fused far warm: {timed(lambda i: trainer.iteration(xf, lab, tss, wgt, apply_step=False), cold=False):.1f} us")
    trainer.dec_grad_saved = trainer.dec_grad
    trainer.dec_grad = None
    print(f"train_fused frozen decoder warm: {timed(fused_only, cold=False):.1f} us")
    trainer.dec_grad = trainer.dec_grad_saved
