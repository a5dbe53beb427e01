"""Density-ratio estimator, inference side (reference: conformalbo/ratio_estimation.py:60-107).

`forward` -- the part on the acquisition hot path (ratio_estimation.py:97-99) -- runs in the CUDA
kernel `k_ratio_mlp_mma` (fp64 tensor path) when the input is on a CUDA device; the acquisition kernels themselves consume
the target-network weights directly (no per-candidate Python call).  Online training of the
classifier (`optimize_callback`, `_Dataset`; ratio_estimation.py:110-144) is out of scope
(SURVEY.md §2 #5): the EMA update of the target network is provided so that externally trained
classifiers can be plugged in.
"""
from __future__ import annotations

import copy

import torch
from torch import nn

from . import ops


class RatioEstimator(nn.Module):
    def __init__(self, in_size=1, device=None, dtype=None, ema_weight=1e-2, lr=1e-3, weight_decay=1e-4):
        super().__init__()
        self.device, self.dtype, self.in_size = device, dtype, in_size
        self._neg_samples = None
        self._pos_samples = []
        self.classifier = nn.Sequential(
            nn.Linear(in_size, 64), nn.ReLU(), nn.Linear(64, 64), nn.ReLU(), nn.Linear(64, 1),
        ).to(device=device, dtype=dtype)
        # density ratio estimates are exactly 1 when untrained (ratio_estimation.py:81-86)
        self._target_network = copy.deepcopy(self.classifier)
        self._target_network.requires_grad_(False)
        for tgt_p in self._target_network.parameters():
            tgt_p.data.fill_(0.0)
        self._ema_weight = ema_weight
        self._weights_cache = None

    def set_neg_samples(self, neg_samples):
        self._neg_samples = neg_samples

    def set_pos_samples(self, pos_samples):
        self._pos_samples  # no-op, as in the reference (ratio_estimation.py:94-95)

    def device_weights(self, device) -> "ops.RatioWeights":
        """fp64 device copies of the target-network weights for the kernels (refreshed after every update)."""
        if not hasattr(self, "_target_params"):
            self._target_params = list(self._target_network.parameters())
        key = (device, tuple(p._version for p in self._target_params))
        if self._weights_cache is None or self._weights_cache[0] != key:
            self._weights_cache = (key, ops.RatioWeights.from_sequential(self._target_network, device))
        return self._weights_cache[1]

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        if inputs.is_cuda:
            return ops.ratio_forward(self.device_weights(inputs.device), inputs).to(inputs.dtype)
        _p = self._target_network(inputs).squeeze(-1).sigmoid()     # CPU tensors: plain torch, not on the GPU hot path
        return _p.clamp_max(1 - 1e-6) / (1 - _p).clamp_min(1e-6)

    def update_target_network(self):
        with torch.no_grad():
            for src_p, tgt_p in zip(self.classifier.parameters(), self._target_network.parameters()):
                tgt_p.mul_(1.0 - self._ema_weight)
                tgt_p.add_(self._ema_weight * src_p)

    def optimize_callback(self, xk):
        raise NotImplementedError(
            "online ratio-estimator training (ratio_estimation.py:110-144) is outside the hot path (SURVEY.md §8f-4)")
