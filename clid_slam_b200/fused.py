"""Fused entry points: one kernel launch per call, nothing materialised in between.

These are what ``Mapper`` uses internally and what the tracker / mesher can call instead of
the three-step ``query_feature`` -> ``Decoder.sdf`` -> ``get_gradient`` sequence
(utils/error_state_iekf.py:203-231, utils/mesher.py:92-135, utils/mapper.py:99-136).
"""
from __future__ import annotations

from typing import Optional

import torch

from .ops import query as _q


def sdf_and_gradient(neural_points, decoder, query_points: torch.Tensor, query_ts: Optional[torch.Tensor] = None,
                     training_mode: bool = False, query_locally: bool = True, with_gradient: bool = True,
                     with_certainty: bool = True, use_bricks: Optional[bool] = None):
    """SDF, its spatial gradient, the candidate count and the queried certainty of every point.

    Returns (sdf [N], grad [N,3] | None, nn_counts [N] int32, certainty [N] | None).  Values equal
    ``decoder.sdf(query_feature(x)[0])`` and ``get_gradient(x, sdf)`` of the reference."""
    res = _q.forward(
        neural_points, decoder, query_points, query_ts, training_mode, query_locally,
        want_sdf=True, want_grad=with_gradient, want_count=True, want_certainty=with_certainty,
        use_bricks=use_bricks,
    )
    return res["sdf"], res.get("grad"), res["nn_count"], res.get("certainty")
