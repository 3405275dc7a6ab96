"""Losses of the mapping step: mirror of the reference's ``utils.loss`` (same names/signatures).

``sdf_bce_loss`` is the one every shipped configuration uses (utils/loss.py:44-62); inside
``Mapper.mapping`` its value and gradient come from the fused ``clid_sdf_loss`` kernel, and the
torch expression below serves callers that build their own graph (and the unfused training
path).  The selectable alternatives (l1 / l2 / zhong, colour) are mirrored; the ray-rendering
losses of the reference file are dead code upstream and not carried over.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def sdf_bce_loss(pred, label, sigma, weight, weighted=False, bce_reduction="mean"):
    """BCE between sigmoid-squashed prediction and label, both scaled by 1/sigma."""
    target = torch.sigmoid(label / sigma)
    return F.binary_cross_entropy_with_logits(
        pred / sigma, target, weight=weight if weighted else None, reduction=bce_reduction)


def sdf_diff_loss(pred, label, weight, scale=1.0, l2_loss=True):
    diff = (pred - label) / scale
    per_sample = diff**2 if l2_loss else diff.abs()
    return (weight * per_sample).sum() / pred.shape[0]


def sdf_l1_loss(pred, label):
    return (pred - label).abs().mean()


def sdf_l2_loss(pred, label):
    return ((pred - label) ** 2).mean()


def color_diff_loss(pred, label, weight, weighted=False, l2_loss=False):
    diff = pred - label
    w = weight.unsqueeze(1) if weighted else 1.0
    return (w * (diff**2 if l2_loss else diff.abs())).mean()


def sdf_zhong_loss(pred, label, trunc_dist=None, weight=None, weighted=False):
    """Zero inside the band between 0 and the label, L1 outside it (utils/loss.py:66-85)."""
    half = label / 2.0
    excess = (pred - half).abs() - half.abs()
    loss = torch.where(excess > 0, excess, torch.zeros_like(excess))
    if trunc_dist is not None:
        near = label.abs() < trunc_dist
        loss = torch.where(near, (pred - label).abs(), loss)
    if weighted:
        loss = loss * weight
    return loss.mean()
