"""Host-side check of the brick index (clid_slam_b200/ops/bricks.py) without a GPU: the candidate set a
kernel obtains by walking the index (emulated here in numpy, cell by cell like search_bricks) is exactly
the set of valid candidates of the reference's hash probe (oracle.radius_search, model/neural_points.py:
971-1030), including the one-brick apron geometry, the stencil table and the at-most-12-half-bricks bound."""
import numpy as np
import pytest
import torch

import helpers as hp
import oracle.sdf_oracle as oc
from clid_slam_b200.ops import bricks as B


def _emulate(idx, res, max_d2, x):
    s = idx.struct
    H = idx.headers.cpu().numpy().view(np.uint32).reshape(-1, 4)
    st = idx.stencil.cpu().numpy().view(np.uint64).reshape(64, 8)
    rec = idx.records.cpu().numpy()
    rows_of_rec = rec[:, 3].copy().view(np.int32)
    origin, dims = np.array(list(s.origin)), list(s.dims)
    cells = np.floor(x.numpy() / np.float32(res)).astype(np.int64)  # floor(x / res) in fp32, like cell_of
    out, max_fill = [], 0
    for q in range(x.shape[0]):
        r = cells[q] - origin - s.reach
        b0 = r >> 2
        if not all(0 <= b0[a] < dims[a] - 1 for a in range(3)):
            out.append(set())
            continue
        pos = ((r[2] & 3) * 4 + (r[1] & 3)) * 4 + (r[0] & 3)
        found, fill = set(), 0
        for slot in range(8):
            hi = ((b0[2] + (slot >> 2)) * dims[1] + b0[1] + ((slot >> 1) & 1)) * dims[0] + b0[0] + (slot & 1)
            mask = int(H[hi, 0]) | (int(H[hi, 1]) << 32)
            want = mask & int(st[pos, slot])
            fill += int((want & 0xFFFFFFFF) != 0) + int((want >> 32) != 0)
            base = int(H[hi, 2])
            while want:
                bit = (want & -want).bit_length() - 1
                want &= want - 1
                k = base + bin(mask & ((1 << bit) - 1)).count("1")
                d = rec[k, :3] - x[q].numpy()
                d2 = np.float32(np.float32(d[0] * d[0]) + np.float32(d[1] * d[1])) + np.float32(d[2] * d[2])
                if not d2 > np.float32(max_d2):
                    found.add(int(rows_of_rec[k]))
        max_fill = max(max_fill, fill)
        out.append(found)
    return out, max_fill


@pytest.mark.parametrize("case", ["plain", "two_frames_timefilter", "global"])
def test_brick_walk_returns_the_hash_probe_candidates(case):
    cfg = oc.OracleConfig(buffer_size=2_000_003, local_map_radius=30.0 if case != "plain" else 80.0)
    m, params, gen = hp.build_oracle_world(60, 2, seed=3, cfg=cfg)
    if case == "two_frames_timefilter":
        # a second scan far along the trajectory: its points fail the travel-distance window of frame 0 queries
        m.travel_dist = torch.tensor([0.0, 500.0])
        pts2 = oc.wavy_sheets(40, 1, cfg.voxel_size_m, gen) + torch.tensor([3.0, 2.0, 0.2])
        oc.map_insert(m, pts2, torch.zeros(3), 1, generator=gen)
        oc.reset_local_window(m, torch.zeros(3), 0)
    locally = case != "global"
    npm = hp.product_map(m, device="cpu")
    idx = B.build(npm, locally)
    assert idx is not None and idx.struct.span == 2 and idx.struct.apron == 1
    x, _, _, _ = oc.sample_batch(m.points, 1500, gen)
    x = torch.cat((x, x[:50] + 500.0))  # far queries: outside the aproned grid
    d2, gidx = oc.radius_search(m, x, time_filtering=locally and m.cfg.temporal_local_map_on if hasattr(m.cfg, "temporal_local_map_on") else locally)
    if locally:
        lid = m.global2local[gidx]
        lid[gidx < 0] = -1
    else:
        lid = gidx
    want = [set(int(v) for v in row[row >= 0].tolist()) for row in lid]
    got, max_fill = _emulate(idx, m.cfg.voxel_size_m, float(npm.max_valid_dist2), x)
    assert max_fill <= 12, "a 5-cell neighbourhood touches at most 3 x 2 x 2 half-bricks"
    bad = [q for q in range(x.shape[0]) if want[q] != got[q]]
    assert not bad, f"{len(bad)} queries differ, first: {bad[0]} want {want[bad[0]]} got {got[bad[0]]}"
    assert all(len(got[q]) == 0 for q in range(1500, 1550))


def test_aliasing_table_sizes_fall_back_to_the_hash_probe():
    assert B.hash_is_alias_free(50_000_000, 2)
    assert B.hash_is_alias_free(2_000_003, 2)
    assert B.hash_is_alias_free(4001, 2)      # the collision fixtures' table: orphans yes, near aliases no
    assert not B.hash_is_alias_free(1009, 2)  # tiny table: cells within reach of one query share slots
    cfg = oc.OracleConfig(buffer_size=1009)
    m, _, _ = hp.build_oracle_world(20, 1, seed=1, cfg=cfg)
    assert B.build(hp.product_map(m, device="cpu"), True) is None
