"""Parity at the benchmark configuration (BASELINE.json configs[2]): the 1.08 M-point world of bench.py with its
5e7-slot hash table on the GPU, 16384 samples of a benchmark batch through the CUDA path and through the CPU
oracle -- forward + gradient, candidate counts, losses, dL/dfeatures, dL/ddecoder (oracle/bridge.py)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["analytic", "numerical"])
def test_parity_gate_on_the_benchmark_world(mode):
    sys.path.insert(0, ROOT)
    import bench
    from clid_slam_b200.synth import sample_batch
    from oracle.bridge import TOL, parity_gate

    cfg, dec, npm, _ = bench.build_world("cuda:0", mode)
    assert npm.count() > 1_000_000 and int(npm.buffer_size) == 50_000_000
    gen = torch.Generator(device="cuda:0").manual_seed(1000)
    x, label, weight, ts = sample_batch(npm.neural_points, 16384, gen)
    res = parity_gate(npm, dec, cfg, x, label, weight, ts)
    assert res["nn_counts_equal"], "candidate counts must be exact"
    assert res["sdf_max_rel"] <= TOL["sdf_rel"], res
    assert res["grad_max_rel"] <= TOL["grad_rel"], res
    assert res["loss_max_rel"] <= TOL["loss_rel"], res
    assert res["feat_grad_max_rel"] <= TOL["feat_grad_rel"], res
    assert res["dec_grad_max_rel"] <= TOL["dec_grad_rel"], res
    assert res["ok"]
