"""Load tests/golden/*.npz fixtures (made by oracle/gen_golden.py from the reference)."""
from __future__ import annotations

import glob
import json
import os

import numpy as np
import torch

from oracle import sdf_oracle as oc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(kind: str):
    files = sorted(glob.glob(os.path.join(GOLDEN_DIR, f"{kind}_*.npz")))
    return [os.path.basename(f)[len(kind) + 1 : -4] for f in files]


def load(kind: str, name: str):
    return np.load(os.path.join(GOLDEN_DIR, f"{kind}_{name}.npz"), allow_pickle=False)


def config_of(fx) -> oc.OracleConfig:
    cfg = oc.OracleConfig()
    for k, v in json.loads(str(fx["cfg"])).items():
        setattr(cfg, k, v)
    return cfg


def t(a) -> torch.Tensor:
    return torch.from_numpy(np.array(a, copy=True))


def dense_table(fx, prefix="map_") -> torch.Tensor:
    table = torch.full((int(fx[prefix + "buffer_size"]),), -1, dtype=torch.int64)
    table[t(fx[prefix + "table_slots"])] = t(fx[prefix + "table_vals"])
    return table


def oracle_map(fx, prefix="map_") -> oc.OracleMap:
    cfg = config_of(fx)
    m = oc.OracleMap(
        cfg=cfg,
        table=dense_table(fx, prefix),
        points=t(fx[prefix + "points"]),
        ts_create=t(fx[prefix + "ts_create"]),
        ts_update=t(fx[prefix + "ts_update"]),
        certainties=t(fx[prefix + "certainties"]),
        features=t(fx[prefix + "features"]),
        travel_dist=t(fx[prefix + "travel_dist"]),
        cur_ts=int(fx[prefix + "cur_ts"]),
        reboot_ts=int(fx[prefix + "reboot_ts"]),
    )
    m.offsets = t(fx[prefix + "offsets"])
    m.max_valid_dist2 = float(fx[prefix + "max_valid_dist2"])
    m.local_points = t(fx[prefix + "local_points"])
    m.local_features = t(fx[prefix + "local_features"]).requires_grad_(True)
    m.local_certainties = t(fx[prefix + "local_certainties"])
    m.local_ts_update = t(fx[prefix + "local_ts_update"])
    m.local_mask = t(fx[prefix + "local_mask"])
    m.global2local = t(fx[prefix + "global2local"])
    return m


def decoder_params(fx, prefix="dec_", requires_grad=True):
    ps = []
    i = 0
    while f"{prefix}W{i}" in fx:
        ps += [t(fx[f"{prefix}W{i}"]), t(fx[f"{prefix}b{i}"])]
        i += 1
    ps += [t(fx[prefix + "Wout"]), t(fx[prefix + "bout"])]
    return [p.requires_grad_(requires_grad) for p in ps]


def assert_close(a, b, rtol, atol, what, max_bad_frac=0.0):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs()
    bad = err > (atol + rtol * b.abs())
    frac = bad.double().mean().item() if bad.numel() else 0.0
    assert frac <= max_bad_frac, (
        f"{what}: max abs err {err.max().item():.3e}; {int(bad.sum())}/{bad.numel()} "
        f"elements outside rtol={rtol} atol={atol} (allowed fraction {max_bad_frac})"
    )
