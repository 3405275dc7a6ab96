"""Spatially sharded training == single-GPU training on the same global batch.

Two ranks are simulated on one GPU: each "rank" owns a full copy of the map and a FusedTrainer,
processes the samples of its slab, and the all-reduce of the flat [decoder grads | loss | shared-row
gradients] buffer is done by hand (sum of the two packed buffers).  What NCCL does on the box is
exactly that sum."""
import pytest
import torch

import golden_io as gio
import helpers as hp
from oracle import sdf_oracle as oc

pytestmark = pytest.mark.gpu


def _world(numerical):
    cfg = oc.OracleConfig(buffer_size=1_000_003, local_map_radius=200.0, numerical_grad=numerical,
                          gradient_decimation=10 if numerical else 1)
    m, params, gen = hp.build_oracle_world(120, 2, seed=21, cfg=cfg)  # 28.8 k points, 48 m wide
    return cfg, m, params, gen


@pytest.mark.parametrize("numerical", [False, True], ids=["analytic", "numerical"])
def test_two_spatial_shards_match_single_rank(numerical):
    from clid_slam_b200.dist import SpatialShards
    from clid_slam_b200.ops.train import FusedTrainer

    cfg, m, params, gen = _world(numerical)
    pcfg = hp.product_config(cfg)
    n, world, iters = 8192, 2, 3
    batches = [oc.sample_batch(m.points, n, gen) for _ in range(iters)]

    def fresh():
        npm = hp.product_map(m)
        dec = hp.product_decoder(cfg, params)
        return npm, dec, FusedTrainer(pcfg, npm, dec)

    npm1, dec1, single = fresh()
    ranks = [fresh() for _ in range(world)]
    shards = SpatialShards(npm1.local_neural_points, cfg.voxel_size_m, reach=2, world_size=world)
    assert shards.boundaries.numel() == world - 1
    assert 0 < shards.shared_rows.numel() < npm1.local_count() // 4

    for it, (x, label, weight, ts) in enumerate(batches):
        x, label, weight, ts = x.cuda(), label.cuda(), weight.cuda(), ts.cuda()
        loss1 = single.iteration(x, label, ts, weight, apply_step=False)
        owner = shards.owner_of(x)
        nd_global = len(range(0, n, cfg.gradient_decimation)) if numerical else 0
        packed, losses = [], []
        for r, (npm_r, dec_r, tr) in enumerate(ranks):
            sel = owner == r
            assert int(sel.sum()) > n // 4
            idx = torch.nonzero(sel).flatten()
            eik = None
            if numerical:  # the eikonal subset is the global x[::10]: each rank takes its members of it
                eik = torch.nonzero(idx % cfg.gradient_decimation == 0).flatten()
            lr = tr.iteration(x[idx], label[idx], ts[idx], weight[idx], apply_step=False,
                              n_global=n, nd_global=nd_global, eik_index=eik)
            losses.append(lr)
            packed.append(tr.pack_spatial(lr, shards))
        flat = packed[0] + packed[1]
        for r, (npm_r, dec_r, tr) in enumerate(ranks):
            tr.unpack_spatial(flat.clone(), losses[r], shards)
            gio.assert_close(losses[r], loss1, 1e-5, 1e-7, f"loss rank{r} it{it}")
            gio.assert_close(tr.dec_grad, single.dec_grad, 1e-4, 1e-8, f"decoder grad rank{r} it{it}")
            mine = (shards.row_owner == r) | shards.shared_mask
            gio.assert_close(tr.feat_grad[mine], single.feat_grad[mine], 1e-3, 1e-9, f"feature grad rank{r} it{it}", 1e-3)
            other_private = (shards.row_owner != r) & ~shards.shared_mask
            assert float(tr.feat_grad[other_private].abs().sum()) == 0.0, "a rank must not touch foreign private rows"
        single.adam_step()
        for _, _, tr in ranks:
            tr.adam_step()
        for r, (npm_r, dec_r, tr) in enumerate(ranks):
            mine = (shards.row_owner == r) | shards.shared_mask
            # Adam (eps 1e-15) turns the sign of a ~0 gradient into a +-lr step: allow a few outliers
            gio.assert_close(npm_r.local_geo_features[mine], npm1.local_geo_features[mine], 1e-3, 1e-5,
                             f"features rank{r} it{it}", 5e-3)
            for a, b in zip(dec_r.flat_parameters(), dec1.flat_parameters()):
                gio.assert_close(a, b, 1e-3, 1e-5, f"decoder rank{r} it{it}", 5e-3)

    # re-replication: every row taken from its owner reproduces the single-rank table
    merged = torch.zeros_like(npm1.local_geo_features.data)
    for r, (npm_r, _, _) in enumerate(ranks):
        sel = (shards.row_owner == r).unsqueeze(1)
        merged += torch.where(sel, npm_r.local_geo_features.data, torch.zeros_like(merged))
    gio.assert_close(merged, npm1.local_geo_features, 1e-3, 1e-5, "gathered feature table", 5e-3)
