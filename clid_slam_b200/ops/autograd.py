"""torch.autograd glue for NeuralPoints.query_feature.

The reference builds z = sum_k w_k [f_k ; x - p_k] out of differentiable torch ops, so callers
may (and do) differentiate it: the tracker and the dynamic filter take d sdf / d x with
``get_gradient`` (utils/error_state_iekf.py:227, utils/mapper.py:116) and the analytic-eikonal
training mode back-propagates *through* that gradient (utils/mapper.py:695-696, 780-835).
Two custom Functions reproduce exactly that connectivity on top of the CUDA kernels:

  QueryFeature          forward  -> clid_query_forward            (z, weights, counts, certainty)
                        backward -> QueryFeatureGrad.apply
  QueryFeatureGrad      forward  -> clid_query_backward           (gx = J^T gz, gfeat)
                        backward -> clid_query_backward_backward  (g_gz = J ggx, gfeat)

Second derivatives with respect to the query coordinates themselves are not produced (no
caller of the reference needs them; they would only matter for d/dx of the eikonal loss).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from .. import _lib
from . import query as _q


def _state_stamp(npm, query_locally, feats):
    """What the backward kernels re-read from the LIVE map by pointer: the gather points / features.  If any of it
    changed since the forward (optimiser step, reset_local_map, update ...) the saved neighbour rows are stale and
    the gradients would silently belong to another map: torch's own version check cannot see it, so we do."""
    pts = npm.local_neural_points if query_locally else npm.neural_points
    return (getattr(npm, "_map_version", 0), feats.data_ptr(), feats._version, tuple(feats.shape), pts.data_ptr())


def _check_stamp(ctx, what):
    feats = ctx.npm.local_geo_features if ctx.query_locally else ctx.npm.geo_features
    if _state_stamp(ctx.npm, ctx.query_locally, feats) != ctx.stamp:
        raise RuntimeError(
            f"{what}: the neural-point map or its feature table was modified between the forward of query_feature "
            "and this backward (optimizer.step / reset_local_map / update ...); the saved neighbour rows are stale")


def _launch_backward(npm, query_locally, x, idx, gz, need_gx, need_gfeat, feat_rows):
    lib = _lib.load()
    m, flags = _q.map_struct(npm, query_locally)
    n = x.shape[0]
    dev = x.device
    gx = torch.empty(n, 3, dtype=torch.float32, device=dev) if need_gx else None
    gfeat = torch.zeros(feat_rows, npm.geo_feature_dim, dtype=torch.float32, device=dev) if need_gfeat else None
    gz = gz.contiguous().float()
    with torch.cuda.device(dev):
        rc = lib.clid_query_backward(
            C.byref(m), x.data_ptr(), idx.data_ptr(), gz.data_ptr(), n, flags,
            None if gx is None else gx.data_ptr(), None if gfeat is None else gfeat.data_ptr(),
            _lib.current_stream(dev))
    _lib.check(rc, "clid_query_backward")
    return gx, gfeat


def _launch_backward_backward(npm, query_locally, x, idx, gz, ggx, need_gfeat, feat_rows):
    lib = _lib.load()
    m, flags = _q.map_struct(npm, query_locally)
    n = x.shape[0]
    dev = x.device
    g_gz = torch.empty(n, npm.geo_feature_dim + 3, dtype=torch.float32, device=dev)
    gfeat = torch.zeros(feat_rows, npm.geo_feature_dim, dtype=torch.float32, device=dev) if need_gfeat else None
    gz = gz.contiguous().float()
    ggx = ggx.contiguous().float()
    with torch.cuda.device(dev):
        rc = lib.clid_query_backward_backward(
            C.byref(m), x.data_ptr(), idx.data_ptr(), gz.data_ptr(), ggx.data_ptr(), n, flags,
            g_gz.data_ptr(), None if gfeat is None else gfeat.data_ptr(), _lib.current_stream(dev))
    _lib.check(rc, "clid_query_backward_backward")
    return g_gz, gfeat


class QueryFeatureGrad(torch.autograd.Function):
    """(gz, x, feats) -> (gx, gfeat); differentiable with respect to gz and feats."""

    @staticmethod
    def forward(ctx, gz, x, feats, idx, npm, query_locally, need_gx, need_gfeat):
        ctx.npm, ctx.query_locally = npm, query_locally
        ctx.feat_rows = feats.shape[0]
        ctx.stamp = _state_stamp(npm, query_locally, feats)
        ctx.save_for_backward(gz, x, idx)
        ctx.set_materialize_grads(False)
        gx, gfeat = _launch_backward(npm, query_locally, x, idx, gz, need_gx, need_gfeat, feats.shape[0])
        return gx, gfeat

    @staticmethod
    def backward(ctx, ggx, ggfeat):
        if ggfeat is not None:
            raise NotImplementedError("differentiating the feature gradient of query_feature is not supported")
        if ggx is None:
            return (None,) * 8
        gz, x, idx = ctx.saved_tensors
        _check_stamp(ctx, "double backward of query_feature")
        g_gz, g_feats = _launch_backward_backward(
            ctx.npm, ctx.query_locally, x, idx, gz, ggx, ctx.needs_input_grad[2], ctx.feat_rows)
        return g_gz if ctx.needs_input_grad[0] else None, None, g_feats, None, None, None, None, None


class QueryFeature(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, feats, npm, ts, training_mode, query_locally):
        xd = _q._prep_points(x)
        accum = None
        table = None
        if training_mode:
            # accumulate the certainty increments apart and add them afterwards, so the queried
            # certainty is gathered before the scatter like the reference (neural_points.py:654 vs 714)
            table = npm.local_point_certainties if query_locally else npm.point_certainties
            accum = torch.zeros_like(table)
        res = _q.forward(npm, None, xd, ts, training_mode, query_locally, want_z=True, want_weights=True,
                         want_idx=True, want_count=True, want_certainty=True, certainty_accum=accum)
        if training_mode:
            table.add_(accum)
        ctx.npm, ctx.query_locally = npm, query_locally
        ctx.stamp = _state_stamp(npm, query_locally, feats)
        ctx.save_for_backward(xd, feats, res["knn_idx"])
        ctx.set_materialize_grads(False)
        nn = res["nn_count"].long()
        ctx.mark_non_differentiable(res["weights"], nn, res["certainty"])
        return res["z"], res["weights"], nn, res["certainty"]

    @staticmethod
    def backward(ctx, gz, _gw, _gn, _gc):
        if gz is None:
            return (None,) * 6
        xd, feats, idx = ctx.saved_tensors
        _check_stamp(ctx, "backward of query_feature")
        need_gx, need_gfeat = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gx, gfeat = QueryFeatureGrad.apply(gz, xd, feats, idx, ctx.npm, ctx.query_locally, need_gx, need_gfeat)
        return gx, gfeat, None, None, None, None


def query_feature(npm, x: torch.Tensor, ts: Optional[torch.Tensor], training_mode: bool, query_locally: bool):
    feats = npm.local_geo_features if query_locally else npm.geo_features
    return QueryFeature.apply(x, feats, npm, ts, training_mode, query_locally)
