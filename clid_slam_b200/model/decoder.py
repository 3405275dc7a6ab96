"""SDF decoder: mirror of the reference's ``model.decoder.Decoder`` (model/decoder.py:12-111).

Same constructor, attribute names (``layers``, ``lout``, ``sdf_scale``, ``use_leaky_relu``,
``out_dim``) and therefore the same ``state_dict`` keys / checkpoint format.  Called on a
tensor, ``mlp`` / ``sdf`` run through torch (cuBLAS) so any caller that differentiates
through them keeps working; the fused CUDA kernels read the very same parameters in place
through ``abi_struct()`` (no copy), which is how ``Mapper.mapping`` and
``fused.sdf_and_gradient`` evaluate the decoder.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib


class Decoder(nn.Module):
    def __init__(self, config, hidden_dim, hidden_level, out_dim, is_time_conditioned=False):
        super().__init__()
        self.out_dim = out_dim
        self.use_leaky_relu = config.mlp_leaky_relu
        bias_on = config.mlp_bias_on

        if config.use_gaussian_pe:
            position_dim = config.pos_input_dim + 2 * config.pos_encoding_band
        else:
            position_dim = config.pos_input_dim * (2 * config.pos_encoding_band + 1)
        input_dim = config.feature_dim + position_dim + (1 if is_time_conditioned else 0)

        widths = [input_dim] + [hidden_dim] * hidden_level
        self.layers = nn.ModuleList(
            nn.Linear(n_in, n_out, bias_on) for n_in, n_out in zip(widths[:-1], widths[1:])
        )
        self.lout = nn.Linear(hidden_dim, out_dim, bias_on)

        self.sdf_scale = 1.0
        if config.main_loss_type == "bce":
            self.sdf_scale = config.logistic_gaussian_ratio * config.sigma_sigmoid_m
        self.to(config.device)

    # ------------------------------------------------------------------ torch path
    def _act(self, t):
        return F.leaky_relu(t) if self.use_leaky_relu else F.relu(t)

    def mlp(self, features):
        h = features
        for layer in self.layers:
            h = self._act(layer(h))
        return self.lout(h)

    def sdf(self, features):
        """Scaled SDF, [N] for [N,D] input ([N,K,1] stays 3-D like the reference)."""
        return self.mlp(features).squeeze(1) * self.sdf_scale

    def time_conditionded_sdf(self, features, ts):
        k = features.shape[1]
        stamp = ts.repeat(k).view(-1, k, 1)
        return self.sdf(torch.cat((features, stamp), dim=-1))

    def occupancy(self, features):
        return torch.sigmoid(self.sdf(features) / -self.sdf_scale)

    def sem_label_prob(self, features):
        return F.log_softmax(self.mlp(features), dim=-1)

    def sem_label(self, features):
        return torch.argmax(self.sem_label_prob(features), dim=1)

    def regress_color(self, features):
        return torch.sigmoid(self.mlp(features))

    # ------------------------------------------------------------------ fused-kernel view
    def flat_parameters(self):
        """[W0, b0, (W1, b1, ...), Wout, bout]; absent biases are None."""
        out = []
        for layer in self.layers:
            out += [layer.weight, layer.bias]
        out += [self.lout.weight, self.lout.bias]
        return out

    def abi_struct(self) -> "_lib.ClidDecoder":
        """ClidDecoder pointing at the live parameters (borrowed, valid while they are)."""
        if self.out_dim != 1:
            raise ValueError("the fused kernels evaluate SDF decoders (out_dim == 1) only")
        if len(self.layers) > _lib.MAX_LEVELS:
            raise ValueError(f"at most {_lib.MAX_LEVELS} hidden levels")
        d = _lib.ClidDecoder()
        for i, layer in enumerate(self.layers):
            d.weight[i] = _lib.ptr(layer.weight.data, torch.float32, f"layers.{i}.weight")
            d.bias[i] = _lib.ptr(None if layer.bias is None else layer.bias.data, torch.float32, f"layers.{i}.bias")
        d.out_weight = _lib.ptr(self.lout.weight.data, torch.float32, "lout.weight")
        d.out_bias = _lib.ptr(None if self.lout.bias is None else self.lout.bias.data, torch.float32, "lout.bias")
        d.in_dim = self.layers[0].in_features
        d.hidden_dim = self.layers[0].out_features
        d.levels = len(self.layers)
        d.sdf_scale = float(self.sdf_scale)
        return d
