"""CPU oracle for the CLID-SLAM neural-SDF hot path.  TEST INFRASTRUCTURE ONLY.

This module is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The shipped path (``clid_slam_b200``) never does, and
raises when its CUDA library is missing.

It restates, with dense eager torch ops on the CPU, the algorithm of these reference
functions (paths relative to the upstream repo DUTRobot/CLID-SLAM @ 5c4f9e7):

  model/neural_points.py:931-969   set_search_neighborhood      -> neighborhood_offsets
  model/neural_points.py:971-1030  radius_neighborhood_search   -> radius_search
  model/neural_points.py:553-769   query_feature                -> query_feature
  model/neural_points.py:1032-1051 query_certainty              -> query_certainty
  model/neural_points.py:324-437   update                       -> map_insert
  model/neural_points.py:439-536   reset_local_map              -> reset_local_window
  model/neural_points.py:538-549   assign_local_to_global       -> write_back_local
  model/decoder.py:58-82           Decoder.mlp / Decoder.sdf    -> decoder_mlp / decoder_sdf
  utils/tools.py:298-311           get_gradient                 -> sdf_gradient
  utils/tools.py:639-682           voxel_down_sample_torch      -> voxel_downsample_indices
  utils/loss.py:44-62              sdf_bce_loss                 -> bce_loss
  utils/mapper.py:780-798          eikonal term                 -> eikonal_loss
  utils/mapper.py:985-1034         get_numerical_gradient       -> numerical_gradient
  utils/mapper.py:642-836          one mapping iteration        -> train_iteration
  utils/tools.py:205-255           setup_optimizer              -> make_adam

The arithmetic itself lives in PyTorch ATen (sort, layer_norm, scatter_add_, linear,
autograd, BCEWithLogits, Adam); the installed torch is therefore the arithmetic of both
the reference and this oracle.

Pinning: the reference ships no tests or golden vectors.  The oracle
is pinned against outputs of the reference's own modules imported from /root/reference in
the build container; ``oracle/gen_golden.py`` is the generating script and the vectors are
committed under ``tests/golden/`` (``tests/test_oracle_golden.py`` checks them).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

PRIMES_NEURAL_POINTS = (73856093, 19349669, 83492791)  # neural_points.py:79-81
IDW_EPS = 1e-15  # neural_points.py:688
FAR_DIST2 = 9e3  # neural_points.py:606


# --------------------------------------------------------------------------------------
# configuration subset (names follow utils/config.py)
# --------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    voxel_size_m: float = 0.4
    num_nei_cells: int = 2
    search_alpha: float = 0.5
    query_nn_k: int = 6
    feature_dim: int = 8
    feature_std: float = 0.05
    buffer_size: int = int(5e7)
    layer_norm_on: bool = False
    weighted_first: bool = True
    local_map_radius: float = 62.0
    local_map_travel_dist_ratio: float = 5.0
    use_mid_ts: bool = False
    temporal_local_map_on: bool = True
    # decoder
    geo_mlp_level: int = 1
    geo_mlp_hidden_dim: int = 64
    mlp_leaky_relu: bool = False
    logistic_gaussian_ratio: float = 0.55
    sigma_sigmoid_m: float = 0.1
    # loss / optimiser
    loss_weight_on: bool = True
    ekional_loss_on: bool = True
    weight_e: float = 0.5
    numerical_grad: bool = True
    gradient_decimation: int = 10
    num_grad_step_ratio: float = 0.2
    lr: float = 0.01
    weight_decay: float = 0.0
    adam_eps: float = 1e-15

    @property
    def sdf_scale(self) -> float:  # decoder.py:51-53, mapper.py:71
        return self.logistic_gaussian_ratio * self.sigma_sigmoid_m

    @property
    def diff_travel_dist_local(self) -> float:  # neural_points.py:59-61
        return self.local_map_radius * self.local_map_travel_dist_ratio


# --------------------------------------------------------------------------------------
# map state
# --------------------------------------------------------------------------------------
@dataclass
class OracleMap:
    cfg: OracleConfig
    table: torch.Tensor  # [B] int64, -1 = empty        (buffer_pt_index)
    points: torch.Tensor  # [M,3] f32                    (neural_points)
    ts_create: torch.Tensor  # [M] int32
    ts_update: torch.Tensor  # [M] int32
    certainties: torch.Tensor  # [M] f32
    features: torch.Tensor  # [M+1,F] f32, last row = padding
    travel_dist: torch.Tensor  # [frames] f32
    cur_ts: int = 0
    reboot_ts: int = 0
    offsets: Optional[torch.Tensor] = None  # [Kc,3] int64   (neighbor_dx)
    max_valid_dist2: float = 0.0
    # local window
    local_points: Optional[torch.Tensor] = None
    local_features: Optional[torch.Tensor] = None  # leaf, requires_grad
    local_certainties: Optional[torch.Tensor] = None
    local_ts_update: Optional[torch.Tensor] = None
    local_mask: Optional[torch.Tensor] = None  # [M+1] bool
    global2local: Optional[torch.Tensor] = None  # [M+1] int64

    def set_neighborhood(self, num_nei_cells: int, search_alpha: float) -> None:
        self.offsets = neighborhood_offsets(num_nei_cells, search_alpha)
        self.max_valid_dist2 = 3.0 * ((num_nei_cells + 1) * self.cfg.voxel_size_m) ** 2


def empty_map(cfg: OracleConfig) -> OracleMap:
    m = OracleMap(
        cfg=cfg,
        table=torch.full((int(cfg.buffer_size),), -1, dtype=torch.int64),
        points=torch.empty((0, 3)),
        ts_create=torch.empty((0,), dtype=torch.int32),
        ts_update=torch.empty((0,), dtype=torch.int32),
        certainties=torch.empty((0,)),
        features=torch.empty((1, cfg.feature_dim)),
        travel_dist=torch.zeros(1),
    )
    m.set_neighborhood(cfg.num_nei_cells, cfg.search_alpha)
    return m


def neighborhood_offsets(num_nei_cells: int, search_alpha: float) -> torch.Tensor:
    """Integer cell offsets inside the sphere |d|^2 < (n + alpha)^2, ij-meshgrid order."""
    r = torch.arange(-num_nei_cells, num_nei_cells + 1, dtype=torch.int64)
    cube = torch.stack(torch.meshgrid(r, r, r, indexing="ij"), dim=-1).reshape(-1, 3)
    keep = (cube * cube).sum(-1) < (num_nei_cells + search_alpha) ** 2
    return cube[keep]


def voxel_hash(cells: torch.Tensor, buffer_size: int, primes=PRIMES_NEURAL_POINTS) -> torch.Tensor:
    """C-style remainder of the prime-weighted cell sum (may be negative: a negative index
    into the table wraps, i.e. equals the mathematical modulus)."""
    p = torch.tensor(primes, dtype=torch.int64)
    return torch.fmod((cells * p).sum(-1), int(buffer_size))


# --------------------------------------------------------------------------------------
# neighbourhood search + feature query
# --------------------------------------------------------------------------------------
def radius_search(m: OracleMap, x: torch.Tensor, time_filtering: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (dist2 [N,Kc] f32, idx [N,Kc] int64 global ids, -1 = invalid)."""
    res = m.cfg.voxel_size_m
    cell = (x / res).floor().to(torch.int64)
    cells = cell[:, None, :] + m.offsets  # [N,Kc,3]
    slot = voxel_hash(cells, m.cfg.buffer_size)
    idx = m.table[slot]  # negative slots wrap
    if time_filtering:
        gap = (m.travel_dist[m.cur_ts] - m.travel_dist[m.ts_create[idx]]).abs()
        idx[~(gap < m.cfg.diff_travel_dist_local)] = -1
    delta = m.points[idx] - x.view(-1, 1, 3)  # idx == -1 reads the last point
    d2 = (delta**2).sum(-1)
    d2[idx == -1] = m.max_valid_dist2
    idx[d2 > m.max_valid_dist2] = -1
    return d2, idx


def query_feature(
    m: OracleMap,
    x: torch.Tensor,
    query_ts: Optional[torch.Tensor] = None,
    training_mode: bool = True,
    query_locally: bool = True,
):
    """Returns (z, weights [N,K,1], nn_counts [N] int64, queried_certainty [N]).

    z is [N,F+3] if weighted_first else [N,K,F+3].  Mutates the certainty / ts_update
    arrays when training_mode (under no_grad), exactly like the reference."""
    cfg = m.cfg
    K = cfg.query_nn_k
    N = x.shape[0]
    d2, idx = radius_search(m, x, time_filtering=cfg.temporal_local_map_on and query_locally)
    if query_locally:
        idx = m.global2local[idx]
    nn_counts = (idx >= 0).sum(-1)

    d2[idx == -1] = FAR_DIST2
    d2_sorted, order = torch.sort(d2, dim=1)
    idx = idx.gather(1, order)[:, :K]
    d2 = d2_sorted[:, :K]
    valid = idx >= 0

    feat_table = m.local_features if query_locally else m.features
    pts_table = m.local_points if query_locally else m.points
    cert_table = m.local_certainties if query_locally else m.certainties

    feats = torch.zeros(N, K, cfg.feature_dim)
    feats[valid] = feat_table[idx[valid]]
    if cfg.layer_norm_on:
        feats = F.layer_norm(feats, [cfg.feature_dim])

    cert = cert_table[idx]
    rel = x.view(-1, 1, 3) - pts_table[idx]
    rel[~valid] = 0.0
    q = torch.cat((feats, rel), dim=2)  # [N,K,F+3]

    w = 1.0 / (d2 + IDW_EPS)
    w[~valid] = 0.0
    w[nn_counts == 0] = IDW_EPS
    w = w / w.sum(1, keepdim=True)
    w[~valid] = 0.0

    with torch.no_grad():
        if training_mode:
            idx = idx.clone()
            idx[~valid] = 0
            if query_locally:
                m.local_certainties.scatter_add_(0, idx.flatten(), w.detach().flatten())
                if query_ts is not None:
                    ts_rep = query_ts.view(-1, 1).repeat(1, K)
                    ts_rep[~valid] = 0
                    m.local_ts_update.scatter_reduce_(
                        0, idx.flatten(), ts_rep.flatten(), reduce="amax", include_self=True
                    )
            else:
                m.certainties.scatter_add_(0, idx.flatten(), w.detach().flatten())
        cert[~valid] = 0.0
        queried_certainty = (cert * w).sum(1)

    w = w.unsqueeze(-1)
    z = (q * w).sum(1) if cfg.weighted_first else q
    return z, w, nn_counts, queried_certainty


def query_certainty(m: OracleMap, x: torch.Tensor) -> torch.Tensor:
    """Max of the *global* certainties over the candidates (no time filter)."""
    _, idx = radius_search(m, x, time_filtering=False)
    c = m.certainties[idx]
    c[idx < 0] = 0.0
    return c.max(dim=-1)[0]


# --------------------------------------------------------------------------------------
# decoder
# --------------------------------------------------------------------------------------
def init_decoder(cfg: OracleConfig, generator: Optional[torch.Generator] = None) -> List[torch.Tensor]:
    """Flat parameter list [W1,b1,(W2,b2,...),Wout,bout] with nn.Linear's default init."""
    dims = [cfg.feature_dim + 3] + [cfg.geo_mlp_hidden_dim] * cfg.geo_mlp_level + [1]
    params: List[torch.Tensor] = []
    for fan_in, fan_out in zip(dims[:-1], dims[1:]):
        bound = 1.0 / math.sqrt(fan_in)
        w = (torch.rand(fan_out, fan_in, generator=generator) * 2 - 1) * bound
        b = (torch.rand(fan_out, generator=generator) * 2 - 1) * bound
        params += [w.requires_grad_(True), b.requires_grad_(True)]
    return params


def decoder_mlp(params: Sequence[torch.Tensor], z: torch.Tensor, leaky: bool = False) -> torch.Tensor:
    h = z
    n_hidden = len(params) // 2 - 1
    for level in range(n_hidden):
        h = F.linear(h, params[2 * level], params[2 * level + 1])
        h = F.leaky_relu(h) if leaky else F.relu(h)
    return F.linear(h, params[-2], params[-1])


def decoder_sdf(params, z, sdf_scale: float, leaky: bool = False) -> torch.Tensor:
    return decoder_mlp(params, z, leaky).squeeze(1) * sdf_scale


def sdf_gradient(x: torch.Tensor, sdf: torch.Tensor) -> torch.Tensor:
    """d sdf / d x by autograd, kept in the graph (second-order capable)."""
    return torch.autograd.grad(
        sdf, x, torch.ones_like(sdf), create_graph=True, retain_graph=True, only_inputs=True
    )[0]


def predict_sdf(m: OracleMap, params, x, query_ts=None, training_mode=True, query_locally=True):
    z, w, nn, cert = query_feature(m, x, query_ts, training_mode, query_locally)
    s = decoder_sdf(params, z, m.cfg.sdf_scale, m.cfg.mlp_leaky_relu)
    if not m.cfg.weighted_first:
        s = (s * w).sum(1).squeeze(1)
    return s, nn, cert


def numerical_gradient(m: OracleMap, params, x: torch.Tensor, eps: float) -> torch.Tensor:
    """Two-sided central differences; the six shifted copies are one concatenated query in
    training mode without timestamps (mapper.py:968-982 defaults)."""
    n = x.shape[0]
    shifts = []
    for axis in range(3):
        e = torch.zeros(3)
        e[axis] = eps
        shifts += [x + e, x - e]
    s = predict_sdf(m, params, torch.cat(shifts, 0))[0].unsqueeze(-1)
    cols = [(s[(2 * a) * n : (2 * a + 1) * n] - s[(2 * a + 1) * n : (2 * a + 2) * n]) / (2 * eps) for a in range(3)]
    return torch.cat(cols, dim=1)


# --------------------------------------------------------------------------------------
# losses / optimiser / one training iteration
# --------------------------------------------------------------------------------------
def bce_loss(pred, label, sigma: float, weight, weighted: bool) -> torch.Tensor:
    target = torch.sigmoid(label / sigma)
    return F.binary_cross_entropy_with_logits(
        pred / sigma, target, weight=weight if weighted else None, reduction="mean"
    )


def eikonal_loss(g: torch.Tensor) -> torch.Tensor:
    return ((g.norm(2, dim=-1) - 1.0) ** 2).mean()


def make_adam(cfg: OracleConfig, feature_params, decoder_params) -> torch.optim.Optimizer:
    groups = []
    if decoder_params is not None:
        groups.append({"params": list(decoder_params), "lr": cfg.lr, "weight_decay": 0.0})
    groups.append({"params": list(feature_params), "lr": cfg.lr, "weight_decay": cfg.weight_decay})
    return torch.optim.Adam(groups, betas=(0.9, 0.99), eps=cfg.adam_eps)


def training_loss(m: OracleMap, params, x, label, ts, weight):
    """Forward part of one mapping iteration.  Returns (total, bce, eikonal, sdf_pred, g)."""
    cfg = m.cfg
    analytic = cfg.ekional_loss_on and not cfg.numerical_grad
    if analytic:
        x = x.detach().clone().requires_grad_(True)
    sdf_pred, _, _ = predict_sdf(m, params, x, ts)
    g = None
    if analytic:
        g = sdf_gradient(x, sdf_pred)
    elif cfg.numerical_grad:
        step = cfg.voxel_size_m * cfg.num_grad_step_ratio
        g = numerical_gradient(m, params, x[:: cfg.gradient_decimation], step)
    l_bce = bce_loss(sdf_pred, label, cfg.sdf_scale, weight.abs().detach(), cfg.loss_weight_on)
    total = l_bce
    l_eik = torch.zeros(())
    if cfg.ekional_loss_on and cfg.weight_e > 0 and g is not None:
        l_eik = eikonal_loss(g)
        total = total + cfg.weight_e * l_eik
    return total, l_bce, l_eik, sdf_pred, g


def train_iteration(m: OracleMap, params, opt, x, label, ts, weight):
    total, l_bce, l_eik, sdf_pred, g = training_loss(m, params, x, label, ts, weight)
    opt.zero_grad(set_to_none=True)
    total.backward()
    opt.step()
    return total.detach(), l_bce.detach(), l_eik.detach()


# --------------------------------------------------------------------------------------
# map maintenance (per frame, host logic)
# --------------------------------------------------------------------------------------
def voxel_downsample_indices(points: torch.Tensor, voxel: float) -> torch.Tensor:
    """Index of the point closest to its voxel centre, one per occupied voxel."""
    quant = 1000
    origin = torch.floor(points.min(dim=0)[0] / voxel).long()
    cell_f = torch.floor(points / voxel)
    centre = (cell_f + 0.5) * voxel
    dist = ((points - centre) ** 2).sum(dim=1) ** 0.5
    rank = (dist / dist.max() * (quant - 1)).long()
    cell = cell_f.long() - origin
    span = cell.max().ceil()
    key = cell[:, 0] + cell[:, 1] * span + cell[:, 2] * span * span
    uniq, inverse = torch.unique(key, return_inverse=True)
    order = torch.arange(inverse.size(0), dtype=inverse.dtype)
    base = 10 ** len(str(order.max().item()))
    packed = order + rank * base
    best = torch.empty(uniq.shape, dtype=inverse.dtype).scatter_reduce_(
        0, inverse, packed, reduce="amin", include_self=False
    )
    return best % base


def map_insert(m: OracleMap, points: torch.Tensor, sensor_position: torch.Tensor, cur_ts: int,
               generator: Optional[torch.Generator] = None) -> float:
    cfg = m.cfg
    res = cfg.voxel_size_m
    pick = voxel_downsample_indices(points, res)
    cand = points[pick]
    slot = voxel_hash((cand / res).floor().to(torch.int64), cfg.buffer_size)
    owner = m.table[slot]
    if m.points.shape[0] > 0 and cur_ts != m.reboot_ts:
        d2 = ((m.points[owner] - cand) ** 2).sum(-1)
        fresh = (owner == -1) | (d2 > 3 * res**2)
        if cfg.temporal_local_map_on:
            gap = m.travel_dist[cur_ts] - m.travel_dist[m.ts_update[owner]]
            fresh = fresh | (gap > cfg.diff_travel_dist_local)
    else:
        fresh = torch.ones(owner.shape, dtype=torch.bool)
    added = cand[fresh]
    n_new = added.shape[0]
    ratio = n_new / cand.shape[0]
    owner = m.table[slot]
    owner[fresh] = torch.arange(n_new, dtype=torch.int64) + m.points.shape[0]
    m.table[slot] = owner
    m.points = torch.cat((m.points, added), 0)
    stamp = torch.ones(n_new, dtype=torch.int32) * cur_ts
    m.ts_create = torch.cat((m.ts_create, stamp), 0)
    m.ts_update = torch.cat((m.ts_update, stamp), 0)
    new_feat = cfg.feature_std * torch.randn(n_new + 1, cfg.feature_dim, generator=generator)
    m.features = torch.cat((m.features[:-1], new_feat), 0)
    m.certainties = torch.cat((m.certainties, torch.zeros(n_new)), 0)
    reset_local_window(m, sensor_position, cur_ts, reboot_map=True)
    return ratio


def reset_local_window(m: OracleMap, sensor_position: torch.Tensor, cur_ts: int,
                       reboot_map: bool = False) -> None:
    cfg = m.cfg
    m.cur_ts = cur_ts
    count = m.points.shape[0]
    if cfg.temporal_local_map_on:
        used = ((m.ts_create + m.ts_update) / 2).int() if cfg.use_mid_ts else m.ts_create
        gap = (m.travel_dist[cur_ts] - m.travel_dist[used]).abs()
        in_time = gap < cfg.diff_travel_dist_local
        if reboot_map:
            in_time = in_time & (used >= m.reboot_ts)
        if in_time.sum() < 100:
            in_time = torch.ones(count, dtype=torch.bool)
    else:
        in_time = torch.ones(count, dtype=torch.bool)
    d2 = ((m.points[in_time] - sensor_position) ** 2).sum(-1)
    near = d2 < cfg.local_map_radius**2
    chosen = torch.nonzero(in_time).squeeze()[near]
    mask = torch.zeros(count, dtype=torch.bool)
    mask[chosen] = True
    m.local_points = m.points[mask]
    m.local_certainties = m.certainties[mask]
    m.local_ts_update = m.ts_update[mask]
    mask = torch.cat((mask, torch.tensor([True])))
    m.local_mask = mask
    g2l = torch.full_like(mask, -1, dtype=torch.int64)
    rows = torch.nonzero(mask).flatten()
    g2l[rows] = torch.arange(rows.numel())
    g2l[-1] = -1
    m.global2local = g2l
    m.local_features = m.features[mask].clone().requires_grad_(True)


def write_back_local(m: OracleMap) -> None:
    m.features[m.local_mask] = m.local_features.detach()
    m.certainties[m.local_mask[:-1]] = m.local_certainties
    m.ts_update[m.local_mask[:-1]] = m.local_ts_update


# --------------------------------------------------------------------------------------
# synthetic world / batches -- shared by tests and bench baselines
# --------------------------------------------------------------------------------------
def wavy_sheets(n_side: int, n_sheets: int, pitch: float, generator: torch.Generator) -> torch.Tensor:
    """Points on `n_sheets` wavy sheets z = 3 s + 0.8 sin(x/5) cos(y/7) + N(0, 0.05^2)."""
    half = n_side * pitch / 2
    ax = torch.arange(n_side, dtype=torch.float32) * pitch - half + pitch / 2
    gx, gy = torch.meshgrid(ax, ax, indexing="ij")
    sheets = []
    for s in range(n_sheets):
        z = 3.0 * s + 0.8 * torch.sin(gx / 5) * torch.cos(gy / 7)
        z = z + 0.05 * torch.randn(gx.shape, generator=generator)
        sheets.append(torch.stack((gx, gy, z), -1).reshape(-1, 3))
    return torch.cat(sheets, 0)


def sample_batch(anchor_points: torch.Tensor, n: int, generator: torch.Generator, max_range: float = 60.0):
    """N samples anchored at random neural points, displaced along z with the sampler's
    5:2:1 surface / front / behind mixture.  Returns (x, label, weight, ts)."""
    pick = torch.randint(0, anchor_points.shape[0], (n,), generator=generator)
    x = anchor_points[pick].clone()
    x[:, :2] += 0.1 * torch.randn(n, 2, generator=generator)
    kind = torch.rand(n, generator=generator)
    disp = 0.25 * torch.randn(n, generator=generator)
    front = -(0.5 + 8.0 * torch.rand(n, generator=generator))
    behind = 0.5 + 0.7 * torch.rand(n, generator=generator)
    is_front = (kind >= 0.625) & (kind < 0.875)
    is_behind = kind >= 0.875
    disp = torch.where(is_front, front, disp)
    disp = torch.where(is_behind, behind, disp)
    x[:, 2] += disp
    label = -disp
    weight = 1.0 + 0.4 - 0.8 * x.norm(dim=-1) / max_range
    weight = torch.where(is_front | is_behind, -weight, weight)
    ts = torch.zeros(n, dtype=torch.int32)
    return x, label, weight, ts
