"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference, CPU) on seeded synthetic inputs, and pin oracle/sdf_oracle.py against it.

TEST INFRASTRUCTURE ONLY.  Run in the build container:  python -m oracle.gen_golden
(the reference tree does not exist on the GPU box; the fixtures travel instead).

Each fixture stores the full map state *before* the operation (hash table as sparse
(slot, value) pairs), the operation's inputs and the reference's outputs.
"""
from __future__ import annotations

import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle import sdf_oracle as oc  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


class FakeDataset:  # the attributes utils/mapper.py reads from SLAMDataset
    lose_track = False
    stop_status = False
    processed_frame = 0
    gt_pose_provided = True
    gt_poses = odom_poses = np.eye(4)[None]
    pgo_poses = None
    static_mask = None


def make_ref_config(ref, **over):
    cfg = ref.Config()
    cfg.load(os.path.join(ref_loader.REFERENCE_ROOT, "config", "run_ncd128.yaml"))
    cfg.device = "cpu"
    cfg.silence = True
    cfg.feature_std = 0.05
    cfg.o3d_vis_on = False
    for k, v in over.items():
        if k == "numerical_grad" and v is False:
            cfg.gradient_decimation = 1  # what Config.load does for numerical_grad_on: False
        setattr(cfg, k, v)
    return cfg


from oracle.bridge import oracle_config_from as oracle_cfg_from_ref  # noqa: E402
from oracle.bridge import oracle_map_from as oracle_map_from_ref  # noqa: E402


def map_state_arrays(npm, prefix="map_"):
    tbl = npm.buffer_pt_index
    slots = torch.nonzero(tbl >= 0).flatten()
    out = {
        "table_slots": slots.numpy(),
        "table_vals": tbl[slots].numpy(),
        "buffer_size": np.int64(tbl.numel()),
        "points": npm.neural_points.numpy(),
        "ts_create": npm.point_ts_create.numpy(),
        "ts_update": npm.point_ts_update.numpy(),
        "certainties": npm.point_certainties.numpy(),
        "features": npm.geo_features.detach().numpy(),
        "travel_dist": npm.travel_dist.numpy(),
        "cur_ts": np.int64(npm.cur_ts),
        "reboot_ts": np.int64(npm.reboot_ts),
        "offsets": npm.neighbor_dx.numpy(),
        "max_valid_dist2": np.float64(npm.max_valid_dist2),
        "local_points": npm.local_neural_points.numpy(),
        "local_features": npm.local_geo_features.detach().numpy(),
        "local_certainties": npm.local_point_certainties.numpy(),
        "local_ts_update": npm.local_point_ts_update.numpy(),
        "local_mask": npm.local_mask.numpy(),
        "global2local": npm.global2local.numpy(),
    }
    return {prefix + k: np.array(v, copy=True) for k, v in out.items()}


def decoder_arrays(dec, prefix="dec_"):
    out = {}
    for i, layer in enumerate(dec.layers):
        out[f"{prefix}W{i}"] = layer.weight.detach().numpy().copy()
        out[f"{prefix}b{i}"] = layer.bias.detach().numpy().copy()
    out[prefix + "Wout"] = dec.lout.weight.detach().numpy().copy()
    out[prefix + "bout"] = dec.lout.bias.detach().numpy().copy()
    return out


def decoder_param_list(dec):
    ps = []
    for layer in dec.layers:
        ps += [layer.weight, layer.bias]
    ps += [dec.lout.weight, dec.lout.bias]
    return ps


def cfg_json(cfg) -> str:
    o = oracle_cfg_from_ref(cfg)
    return json.dumps({k: getattr(o, k) for k in o.__dataclass_fields__})


def close(a, b, rtol, atol, what, max_bad_frac=0.0):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    err = (a - b).abs()
    bound = atol + rtol * b.abs()
    bad = err > bound
    if bad.double().mean().item() > max_bad_frac:
        raise AssertionError(
            f"oracle != reference for {what}: max err {err.max().item():.3e}, "
            f"{int(bad.sum())}/{bad.numel()} outside rtol={rtol} atol={atol}"
        )


# --------------------------------------------------------------------------------------
def sheet_world(gen, n_side=72, n_sheets=1, pitch=0.4):
    return oc.wavy_sheets(n_side, n_sheets, pitch, gen)


def populate(ref, cfg, frames, travel):
    """Insert scans through the reference's own NeuralPoints.update."""
    npm = ref.NeuralPoints(cfg)
    npm.travel_dist = torch.tensor(travel, dtype=torch.float32)
    for pts, sensor, ts in frames:
        npm.update(pts, sensor, torch.eye(3), ts)
    return npm


def query_case(ref, name, cfg, frames, travel, n_query, training_mode, query_locally,
               seed, with_ts=True, neighborhood=None):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    dec = ref.Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    npm = populate(ref, cfg, frames, travel)
    if neighborhood is not None:
        npm.set_search_neighborhood(*neighborhood)
    # certainties that are not all zero make queried_certainty meaningful
    npm.point_certainties = torch.rand(npm.count(), generator=gen) * 3.0
    npm.reset_local_map(frames[-1][1], torch.eye(3), frames[-1][2])
    x, _, _, _ = oc.sample_batch(npm.neural_points, n_query, gen)
    x[: n_query // 16] += 40.0 * torch.randn(n_query // 16, 3, generator=gen)  # far / empty queries
    ts = torch.randint(0, len(travel), (n_query,), generator=gen).int() if with_ts else None

    ocfg = oracle_cfg_from_ref(cfg)
    before = map_state_arrays(npm)
    omap = oracle_map_from_ref(npm, ocfg)

    # ---- reference
    d2_all, idx_all = npm.radius_neighborhood_search(x, time_filtering=npm.temporal_local_map_on and query_locally)
    qc = npm.query_certainty(x)  # before query_feature mutates the certainties
    xr = x.clone().requires_grad_(True)
    z, _, w, nn, cert = npm.query_feature(xr, ts, training_mode=training_mode, query_locally=query_locally)
    sdf = dec.sdf(z)
    grad = ref.tools.get_gradient(xr, sdf)

    # ---- oracle (live pin)
    d2o, idxo = oc.radius_search(omap, x, ocfg.temporal_local_map_on and query_locally)
    assert torch.equal(idxo, idx_all)
    close(d2o, d2_all, 0, 0, name + ":dist2")
    close(oc.query_certainty(omap, x), qc, 0, 0, name + ":query_certainty")
    xo = x.clone().requires_grad_(True)
    params = [p.detach().clone().requires_grad_(True) for p in decoder_param_list(dec)]
    zo, wo, nno, certo = oc.query_feature(omap, xo, ts, training_mode, query_locally)
    sdfo = oc.decoder_sdf(params, zo, ocfg.sdf_scale, ocfg.mlp_leaky_relu)
    grado = oc.sdf_gradient(xo, sdfo)
    assert torch.equal(nn, nno), "nn_counts"
    close(zo, z, 0, 0, name + ":z")
    close(wo, w, 0, 0, name + ":w")
    close(certo, cert, 0, 0, name + ":certainty")
    close(sdfo, sdf, 0, 0, name + ":sdf")
    close(grado, grad, 0, 0, name + ":grad")
    if query_locally:
        close(omap.local_certainties, npm.local_point_certainties, 0, 0, name + ":cert side effect")
        assert torch.equal(omap.local_ts_update, npm.local_point_ts_update)
    else:
        close(omap.certainties, npm.point_certainties, 0, 0, name + ":cert side effect")

    out = dict(before)
    out.update(decoder_arrays(dec))
    out.update(
        cfg=np.array(cfg_json(cfg)),
        training_mode=np.bool_(training_mode), query_locally=np.bool_(query_locally),
        x=x.numpy(), ts=(ts.numpy() if ts is not None else np.zeros((0,), np.int32)),
        has_ts=np.bool_(ts is not None),
        out_z=z.detach().numpy(), out_w=w.detach().numpy(), out_nn=nn.numpy(),
        out_certainty=cert.numpy(), out_sdf=sdf.detach().numpy(), out_grad=grad.detach().numpy(),
        out_dist2=d2_all[:256].numpy(), out_idx=idx_all[:256].numpy(),  # first 256 rows only (size)
        out_query_certainty=qc.numpy(),
        after_local_certainties=npm.local_point_certainties.numpy().copy(),
        after_local_ts_update=npm.local_point_ts_update.numpy().copy(),
        after_certainties=npm.point_certainties.numpy().copy(),
    )
    path = os.path.join(GOLDEN_DIR, f"query_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"  wrote {path}: N={n_query} M={npm.count()} local={npm.local_count()} "
          f"mean nn={nn.float().mean():.2f} zero-nn={(nn == 0).float().mean():.3f}")


def train_case(ref, name, cfg, frames, travel, n_batch, n_iters, seed, freeze_decoder=False):
    """Drive the reference's own Mapper.mapping() with recorded batches."""
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    dec = ref.Decoder(cfg, cfg.geo_mlp_hidden_dim, cfg.geo_mlp_level, 1)
    npm = populate(ref, cfg, frames, travel)
    cfg.bs = n_batch
    cfg.wandb_vis_on = True
    mapper = ref.Mapper(cfg, FakeDataset(), npm, ref.LocalPointCloudMap(cfg), dec)
    mapper.used_poses = torch.eye(4, dtype=torch.float64)[None].repeat(len(travel), 1, 1)
    mapper.adaptive_iter_offset = 0
    if freeze_decoder:
        ref.tools.freeze_model(dec)

    batches = []
    for _ in range(n_iters):
        x, label, weight, _ = oc.sample_batch(npm.neural_points, n_batch, gen)
        x[: n_batch // 16] += 40.0 * torch.randn(n_batch // 16, 3, generator=gen)
        ts = torch.randint(0, len(travel), (n_batch,), generator=gen).int()
        batches.append((x, label, ts, weight))
    feed = iter(batches)
    mapper.get_batch = lambda global_coord=False: (lambda b: (b[0].clone(), b[1], b[2], None, None, None, b[3]))(next(feed))

    logs = []
    import utils.mapper as ref_mapper_mod
    ref_mapper_mod.wandb = SimpleNamespace(log=lambda d: logs.append(
        {k: float(v) for k, v in d.items() if k.startswith("loss/")}))

    ocfg = oracle_cfg_from_ref(cfg)
    before = map_state_arrays(npm)
    dec_before = decoder_arrays(dec)
    omap = oracle_map_from_ref(npm, ocfg)
    oparams = [p.detach().clone().requires_grad_(True) for p in decoder_param_list(dec)]

    # the reference creates its Adam inside mapping(); iterate one call at a time is NOT the same
    # (fresh optimiser state per call), so run all iterations in one mapping() call
    feat_grads, dec_grads = [], []
    # hook: capture grads after every backward via optimizer step wrapper
    orig_setup = ref_mapper_mod.setup_optimizer

    def setup_and_spy(*a, **k):
        opt = orig_setup(*a, **k)
        step = opt.step

        def spy_step(*sa, **sk):
            feat_grads.append(npm.local_geo_features.grad.detach().clone())
            dec_grads.append([None if p.grad is None else p.grad.detach().clone() for p in decoder_param_list(dec)])
            return step(*sa, **sk)

        opt.step = spy_step
        return opt

    ref_mapper_mod.setup_optimizer = setup_and_spy
    try:
        mapper.mapping(n_iters)
    finally:
        ref_mapper_mod.setup_optimizer = orig_setup
    assert len(logs) == n_iters and len(feat_grads) == n_iters

    # ---- oracle (live pin)
    opt = oc.make_adam(ocfg, [omap.local_features], None if freeze_decoder else oparams)
    if freeze_decoder:
        for p in oparams:
            p.requires_grad_(False)
    o_losses, o_fg, o_dg = [], [], []
    for (x, label, ts, weight) in batches:
        tot, lb, le = oc.train_iteration(omap, oparams, opt, x.clone(), label, ts, weight)
        o_losses.append((float(tot), float(lb), float(le)))
        o_fg.append(omap.local_features.grad.detach().clone())
        o_dg.append([None if p.grad is None else p.grad.detach().clone() for p in oparams])
    oc.write_back_local(omap)
    for it in range(n_iters):
        close(o_losses[it][0], logs[it]["loss/total_loss"], 1e-6, 0, f"{name}: total loss it{it}")
        close(o_losses[it][1], logs[it]["loss/sdf_loss"], 1e-6, 0, f"{name}: bce loss it{it}")
        close(o_losses[it][2], logs[it]["loss/eikonal_loss"], 1e-6, 0, f"{name}: eikonal it{it}")
        # scatter-accumulate order differs run to run on a threaded CPU: tolerance, not bits
        close(o_fg[it], feat_grads[it], 1e-4, 1e-9, f"{name}: feature grad it{it}", 1e-3)
        for a, b in zip(o_dg[it], dec_grads[it]):
            if b is not None:
                close(a, b, 1e-4, 1e-8, f"{name}: decoder grad it{it}")
    # Adam with eps=1e-15 turns a sign flip of a ~0 gradient into a full +-lr step
    close(omap.local_features, npm.local_geo_features, 1e-3, 1e-5, name + ": features after Adam", 2e-3)
    close(omap.features, npm.geo_features, 1e-3, 1e-5, name + ": global features after write-back", 2e-3)
    close(omap.local_certainties, npm.local_point_certainties, 1e-5, 1e-6, name + ": certainties")
    assert torch.equal(omap.local_ts_update, npm.local_point_ts_update)
    for a, b in zip(oparams, decoder_param_list(dec)):
        close(a, b, 1e-3, 1e-5, name + ": decoder after Adam", 2e-3)

    out = dict(before)
    out.update(dec_before)
    out.update(decoder_arrays(dec, prefix="after_dec_"))
    out.update(
        cfg=np.array(cfg_json(cfg)), n_iters=np.int64(n_iters), freeze_decoder=np.bool_(freeze_decoder),
        batch_x=np.stack([b[0].numpy() for b in batches]),
        batch_label=np.stack([b[1].numpy() for b in batches]),
        batch_ts=np.stack([b[2].numpy() for b in batches]),
        batch_weight=np.stack([b[3].numpy() for b in batches]),
        loss_total=np.array([l["loss/total_loss"] for l in logs]),
        loss_bce=np.array([l["loss/sdf_loss"] for l in logs]),
        loss_eikonal=np.array([l["loss/eikonal_loss"] for l in logs]),
        feat_grads=np.stack([g.numpy() for g in feat_grads]),
        after_local_features=npm.local_geo_features.detach().numpy().copy(),
        after_features=npm.geo_features.detach().numpy().copy(),
        after_local_certainties=npm.local_point_certainties.numpy().copy(),
        after_local_ts_update=npm.local_point_ts_update.numpy().copy(),
    )
    if not freeze_decoder:
        for it, gl in enumerate(dec_grads):
            for j, g in enumerate(gl):
                out[f"dec_grad_it{it}_{j}"] = g.numpy()
    path = os.path.join(GOLDEN_DIR, f"train_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"  wrote {path}: N={n_batch} iters={n_iters} M={npm.count()} loss={logs[-1]}")


def map_case(ref, name, cfg, frames, travel, seed):
    """NeuralPoints.update x len(frames) + assign_local_to_global: host-logic fixture."""
    torch.manual_seed(seed)
    npm = ref.NeuralPoints(cfg)
    npm.travel_dist = torch.tensor(travel, dtype=torch.float32)
    ocfg = oracle_cfg_from_ref(cfg)
    ocfg.feature_std = 0.0  # randn streams differ between implementations; features start at 0
    npm.geo_feature_std = 0.0
    omap = oc.empty_map(ocfg)
    omap.travel_dist = npm.travel_dist.clone()
    out = {"cfg": np.array(json.dumps({k: getattr(ocfg, k) for k in ocfg.__dataclass_fields__})),
           "travel_dist": np.array(travel, np.float32), "n_frames": np.int64(len(frames))}
    for i, (pts, sensor, ts) in enumerate(frames):
        r_ref = npm.update(pts, sensor, torch.eye(3), ts)
        r_orc = oc.map_insert(omap, pts, sensor, ts)
        assert r_ref == r_orc, (r_ref, r_orc)
        assert torch.equal(npm.buffer_pt_index, omap.table)
        assert torch.equal(npm.neural_points, omap.points)
        assert torch.equal(npm.local_mask, omap.local_mask)
        assert torch.equal(npm.global2local, omap.global2local)
        assert torch.equal(npm.point_ts_create, omap.ts_create)
        out[f"frame{i}_points"] = pts.numpy()
        out[f"frame{i}_sensor"] = sensor.numpy()
        out[f"frame{i}_ts"] = np.int64(ts)
        out[f"frame{i}_ratio"] = np.float64(r_ref)
        out.update(map_state_arrays(npm, prefix=f"frame{i}_map_"))
    path = os.path.join(GOLDEN_DIR, f"map_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"  wrote {path}: frames={len(frames)} M={npm.count()} local={npm.local_count()}")


def sampler_case(ref, name, cfg, seed):
    """DataSampler.sample_pin of the reference under a fixed torch seed (CPU generator)."""
    from utils.data_sampler import DataSampler

    gen = torch.Generator().manual_seed(seed)
    scan = torch.randn(300, 3, generator=gen) * torch.tensor([15.0, 15.0, 2.0]) + torch.tensor([0.0, 0.0, 1.0])
    torch.manual_seed(seed)
    coord, label, _, _, _, weight = DataSampler(cfg).sample_pin(scan, None, None, None)
    path = os.path.join(GOLDEN_DIR, f"sampler_{name}.npz")
    np.savez_compressed(path, cfg_sampler=np.array(json.dumps({
        k: getattr(cfg, k) for k in ("surface_sample_range_m", "surface_sample_n", "free_front_n", "free_behind_n",
                                      "free_sample_begin_ratio", "free_sample_end_dist_m", "dist_weight_on",
                                      "dist_weight_scale", "max_range", "behind_dropoff_on")})),
        seed=np.int64(seed), scan=scan.numpy(), coord=coord.numpy(), label=label.numpy(), weight=weight.numpy())
    print(f"  wrote {path}: {scan.shape[0]} rays -> {coord.shape[0]} samples")


def clid_sampler_case(ref, name, cfg, seed):
    """LocalPointCloudMap.update_map x2 + DataSampler.sample (region-specific SDF labels)."""
    from utils.data_sampler import DataSampler

    gen = torch.Generator().manual_seed(seed)

    def scan(n):
        xy = (torch.rand(n, 2, generator=gen) - 0.5) * 30
        floor = torch.cat((xy, -1.5 + 0.2 * torch.sin(xy[:, :1] / 3)), dim=1)
        yz = (torch.rand(n // 3, 2, generator=gen) - 0.5) * torch.tensor([20.0, 3.0])
        wall = torch.cat((torch.full((n // 3, 1), 8.0), yz), dim=1)
        pts = torch.cat((floor, wall), 0)
        return pts[pts.norm(dim=1) > 1.0]

    scans = [scan(1500), scan(1500)]
    poses = [torch.eye(4, dtype=torch.float64), torch.eye(4, dtype=torch.float64)]
    poses[1][0, 3] = 0.4
    lmap = ref.LocalPointCloudMap(cfg)
    out = {"cfg_sampler": np.array(json.dumps({
        k: getattr(cfg, k) for k in ("surface_sample_range_m", "surface_sample_n", "free_front_n", "free_behind_n",
                                      "free_sample_begin_ratio", "free_sample_end_dist_m", "dist_weight_on",
                                      "dist_weight_scale", "max_range", "local_voxel_size_m", "local_buffer_size",
                                      "local_map_size")})), "seed": np.int64(seed)}
    for i, (pts, pose) in enumerate(zip(scans, poses)):
        lmap.update_map(pose[:3, 3].float(), ref.tools.transform_torch(pts, pose))
        out[f"scan{i}"] = pts.numpy()
        out[f"pose{i}"] = pose.numpy()
        out[f"map{i}_points"] = lmap.local_point_cloud_map.numpy().copy()
        slots = torch.nonzero(lmap.buffer_pt_index >= 0).flatten()
        out[f"map{i}_slots"] = slots.numpy()
        out[f"map{i}_vals"] = lmap.buffer_pt_index[slots].numpy()
    torch.manual_seed(seed)
    coord, label, weight = DataSampler(cfg).sample(scans[1], lmap, poses[1])
    probe = ref.tools.transform_torch(scans[1][:512] + 0.05, poses[1])
    d, mask = lmap.region_specific_sdf_estimation(probe)
    out.update(coord=coord.numpy(), label=label.numpy(), weight=weight.numpy(),
               probe=probe.numpy(), probe_dist=d.numpy(), probe_mask=mask.numpy())
    path = os.path.join(GOLDEN_DIR, f"clidsampler_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"  wrote {path}: {coord.shape[0]} samples kept, map {lmap.local_point_cloud_map.shape[0]} points")


def main():
    """python -m oracle.gen_golden [query|train|train_widths|map|sampler]   (no argument = everything)"""
    only = sys.argv[1] if len(sys.argv) > 1 else None
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref = ref_loader.load()
    torch.set_num_threads(8)
    gen = torch.Generator().manual_seed(7)
    world = sheet_world(gen, n_side=72)  # 5184 points, one sheet, 28.8 m
    origin = torch.zeros(3)
    one_frame = [(world, origin, 0)]
    # two scans of shifted halves: second scan is far in travel distance from the first
    left = world[world[:, 0] < 2.0]
    right = world[world[:, 0] > -2.0] + torch.tensor([0.07, -0.05, 0.03])
    two_frames = [(left, origin, 0), (right, torch.tensor([1.0, 0.0, 0.0]), 2)]

    if only in (None, "query"):
        print("query fixtures")
        _query_cases(ref, gen, one_frame, two_frames, origin)
    if only in (None, "train"):
        print("train fixtures")
        _train_cases(ref, one_frame, two_frames)
    if only in (None, "train_widths"):
        print("train fixtures, other decoder widths")
        _train_width_cases(ref, one_frame)
    if only in (None, "map"):
        print("map fixtures")
        _map_cases(ref, two_frames)
    if only in (None, "sampler"):
        print("sampler fixtures")
        sampler_case(ref, "ncd128", make_ref_config(ref), 41)
        clid_sampler_case(ref, "ncd128", make_ref_config(ref), 43)
    print("done")


def _query_cases(ref, gen, one_frame, two_frames, origin):
    query_case(ref, "ncd128_train", make_ref_config(ref), one_frame, [0.0], 2048, True, True, 1)
    query_case(ref, "ncd128_infer", make_ref_config(ref), one_frame, [0.0], 2048, False, True, 2, with_ts=False)
    query_case(ref, "smallbuf_collisions", make_ref_config(ref, buffer_size=4001), one_frame, [0.0], 2048, True, True, 3)
    query_case(ref, "layernorm", make_ref_config(ref, layer_norm_on=True), one_frame, [0.0], 2048, True, True, 4)
    query_case(ref, "two_frames_timefilter", make_ref_config(ref), two_frames, [0.0, 100.0, 400.0], 2048, True, True, 5)
    query_case(ref, "two_frames_near", make_ref_config(ref), two_frames, [0.0, 1.0, 2.0], 2048, True, True, 6)
    query_case(ref, "global_infer", make_ref_config(ref), two_frames, [0.0, 100.0, 400.0], 2048, False, False, 7, with_ts=False)
    query_case(ref, "global_train", make_ref_config(ref), one_frame, [0.0], 1024, True, False, 8)
    query_case(ref, "kc33", make_ref_config(ref, search_alpha=0.2), one_frame, [0.0], 1024, True, True, 9)
    query_case(ref, "kc7_k4", make_ref_config(ref, num_nei_cells=1, search_alpha=0.2, query_nn_k=4), one_frame, [0.0], 1024, True, True, 10)
    query_case(ref, "l2h32", make_ref_config(ref, geo_mlp_level=2, geo_mlp_hidden_dim=32), one_frame, [0.0], 1024, False, True, 11, with_ts=False)
    query_case(ref, "leaky", make_ref_config(ref, mlp_leaky_relu=True), one_frame, [0.0], 1024, False, True, 12, with_ts=False)
    query_case(ref, "res02", make_ref_config(ref, voxel_size_m=0.2, sigma_sigmoid_m=0.05), [(sheet_world(gen, 96, 1, 0.2), origin, 0)], [0.0], 1024, True, True, 13)


def _train_cases(ref, one_frame, two_frames):
    train_case(ref, "analytic_l1h64", make_ref_config(ref, numerical_grad=False), one_frame, [0.0], 2048, 3, 21)
    train_case(ref, "numerical_l1h64", make_ref_config(ref), one_frame, [0.0], 2048, 3, 22)
    train_case(ref, "analytic_l2h32", make_ref_config(ref, numerical_grad=False, geo_mlp_level=2, geo_mlp_hidden_dim=32), one_frame, [0.0], 2048, 2, 23)
    train_case(ref, "numerical_l2h32", make_ref_config(ref, geo_mlp_level=2, geo_mlp_hidden_dim=32), one_frame, [0.0], 2048, 2, 24)
    train_case(ref, "numerical_layernorm", make_ref_config(ref, layer_norm_on=True), two_frames, [0.0, 1.0, 2.0], 2048, 2, 25)
    train_case(ref, "analytic_layernorm", make_ref_config(ref, layer_norm_on=True, numerical_grad=False), one_frame, [0.0], 2048, 2, 26)
    train_case(ref, "numerical_frozen", make_ref_config(ref), one_frame, [0.0], 2048, 2, 27, freeze_decoder=True)
    train_case(ref, "analytic_unweighted", make_ref_config(ref, numerical_grad=False, loss_weight_on=False), one_frame, [0.0], 1024, 2, 28)


def _train_width_cases(ref, one_frame):
    """One hidden level at the other widths the fused kernels are compiled for (H = 32, 128)."""
    train_case(ref, "analytic_l1h32", make_ref_config(ref, numerical_grad=False, geo_mlp_hidden_dim=32), one_frame, [0.0], 2048, 2, 29)
    train_case(ref, "numerical_l1h128", make_ref_config(ref, geo_mlp_hidden_dim=128), one_frame, [0.0], 2048, 2, 30)


def _map_cases(ref, two_frames):
    map_case(ref, "two_frames", make_ref_config(ref), two_frames, [0.0, 1.0, 2.0], 31)
    map_case(ref, "two_frames_far", make_ref_config(ref), two_frames, [0.0, 100.0, 400.0], 32)
    map_case(ref, "smallbuf", make_ref_config(ref, buffer_size=4001), two_frames, [0.0, 1.0, 2.0], 33)


if __name__ == "__main__":
    main()
