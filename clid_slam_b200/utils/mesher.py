"""Dense-grid SDF inference of the mesher: mirror of ``Mesher.query_points`` (utils/mesher.py:38-163 of the
reference), the inference caller behind marching cubes and the SDF slices.

The reference chunks the grid by ``bs`` and runs query_feature (81 probes, sort, gathers) + Decoder.sdf under
no_grad per chunk; here every chunk is one fused launch on the GLOBAL map (query_locally=False is the mesher's
default).  Marching cubes, mesh clean-up and export (skimage / open3d) are visualisation and stay with the
reference; the semantic / colour heads are outside the neural-SDF hot path.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import fused


class Mesher:
    def __init__(self, config, neural_points, decoders: dict):
        self.config = config
        self.silence = config.silence
        self.neural_points = neural_points
        self.sdf_mlp = decoders["sdf"]
        self.sem_mlp = decoders.get("semantic")
        self.color_mlp = decoders.get("color")
        self.device = config.device
        self.cur_device = self.device
        self.dtype = config.dtype
        self.global_transform = np.eye(4)

    def query_points(self, coord, bs, query_sdf=True, query_sem=False, query_color=False, query_mask=True,
                     query_locally=False, mask_min_nn_count: int = 4, out_torch: bool = False):
        """(sdf_pred [N], sem_pred None, color_pred None, mc_mask [N]); numpy arrays unless out_torch.  sdf_pred is 0
        where no neural point is in reach (nn_count == 0), mc_mask marks nn_count >= mask_min_nn_count."""
        if query_sem or query_color:
            raise NotImplementedError("semantic / colour heads are outside the neural-SDF hot path")
        if not self.config.weighted_first:
            raise NotImplementedError("weighted_first=False is not used by any shipped configuration")
        n = coord.shape[0]
        sdf_pred = torch.zeros(n) if query_sdf else None
        mc_mask = torch.zeros(n) if query_mask else None
        for it in range(math.ceil(n / bs)):
            head, tail = it * bs, min((it + 1) * bs, n)
            chunk = coord[head:tail, :]
            if not chunk.is_cuda:
                chunk = chunk.to(self.device)
            sdf, _, nn, _ = fused.sdf_and_gradient(self.neural_points, self.sdf_mlp, chunk, training_mode=False,
                                                   query_locally=query_locally, with_gradient=False, with_certainty=False)
            if query_sdf:
                sdf_pred[head:tail] = torch.where(nn >= 1, sdf, torch.zeros_like(sdf)).cpu()
            if query_mask:
                mc_mask[head:tail] = (nn >= mask_min_nn_count).float().cpu()
        if not out_torch:
            sdf_pred = None if sdf_pred is None else sdf_pred.numpy().astype(np.float64)
            mc_mask = None if mc_mask is None else mc_mask.numpy().astype(np.float64)
        return sdf_pred, None, None, mc_mask
