"""Dev tool: host-side overhead per call and CUDA-graph replay time of the hot-path entry points."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clid_slam_b200.config import ncd128
from clid_slam_b200.model.decoder import Decoder
from clid_slam_b200.model.neural_points import NeuralPoints
from clid_slam_b200 import fused
from clid_slam_b200.ops.train import FusedTrainer
from clid_slam_b200.synth import wavy_sheets, sample_batch

torch.manual_seed(42)
cfg = ncd128(); cfg.device = "cuda"; cfg.feature_std = 0.05; cfg.local_map_radius = 1e4
cfg.numerical_grad, cfg.gradient_decimation = False, 1
dec = Decoder(cfg, 64, 1, 1); npm = NeuralPoints(cfg); npm.travel_dist = torch.zeros(1, device="cuda")
gen = torch.Generator(device="cuda").manual_seed(1)
npm.update(wavy_sheets(520, 4, 0.4, gen, device="cuda"), torch.zeros(3, device="cuda"), torch.eye(3, device="cuda"), 0)

def host_us(fn, reps=300):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize()
    return (t1 - t0) / reps * 1e6

def graph_us(fn, reps=20, replays=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * replays) * 1e3

for n in (32, 16384, 131072):
    x, label, weight, ts = sample_batch(npm.neural_points, n, gen)
    far = x + 1000.0
    same = x[:1].repeat(n, 1).contiguous()
    f = lambda xq=x: fused.sdf_and_gradient(npm, dec, xq)
    print(f"n={n}: forward host {host_us(f):.1f} us/call | graph replay: normal {graph_us(f):.1f} us, "
          f"far {graph_us(lambda: fused.sdf_and_gradient(npm, dec, far)):.1f} us, identical {graph_us(lambda: fused.sdf_and_gradient(npm, dec, same)):.1f} us")
    tr = FusedTrainer(cfg, npm, dec)
    it = lambda: tr.iteration(x, label, ts, weight)
    print(f"n={n}: train iteration host {host_us(it, 100):.1f} us/call | graph replay {graph_us(it, 5, 10):.1f} us")
