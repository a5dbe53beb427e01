"""Density-ratio estimator (reference: conformalbo/ratio_estimation.py:60-144).

`forward` -- the part on the acquisition hot path (ratio_estimation.py:97-99) -- runs in the CUDA
kernel `k_ratio_mlp_mma` (fp64 tensor path) when the input is on a CUDA device; the acquisition kernels themselves consume
the target-network weights directly (no per-candidate Python call).

`optimize_callback` (ratio_estimation.py:110-144) trains on the device: the SGLD iterates are appended to
a device-resident pool, and one call of `cbo_ratio_train_step_f64` (two launches) does the BCE forward/reverse pass over
[negatives; sampled positives], the Adam step and the EMA of the target network, in place on flat fp64 parameter vectors
that the acquisition kernels read directly -- no `.cpu()` round trip per SGLD step (optim/optimize.py:222).  The unused
`_Dataset` bookkeeping of the reference (train/val/test splits that `optimize_callback` never reads) is not rebuilt.
"""
from __future__ import annotations

import copy

import numpy as np
import torch
from torch import nn

from . import _lib, ops


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
        self._lr, self._weight_decay = lr, weight_decay       # torch.optim.Adam(lr, weight_decay) of the reference (:91)
        self._train = None                # device training state, created by the first training step
        self._pos_pool, self._num_pos = None, 0
        self._neg_dev = None
        # True (default): the positive minibatch is drawn on the device (cbo_sample_subset_i64, one launch, seeded from
        # torch's seed at construction + the step count); False: the reference's own host draw, np.random.permutation
        # (ratio_estimation.py:129-131) -- the parity setting the reference-executed goldens are checked with
        self.device_permutation = True
        self._perm_seed = int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF
        self._perm_idx = None
        self.last_loss = None             # device scalar, minibatch loss of the latest step

    def set_neg_samples(self, neg_samples):
        self._neg_samples = neg_samples

    def set_pos_samples(self, pos_samples):
        self._pos_samples  # no-op, as in the reference (ratio_estimation.py:94-95)

    @property
    def num_positive(self) -> int:
        return self._num_pos

    def _token(self, net):
        # (storage, version) of every parameter; the parameter list of a module is walked once (nn.Module.parameters() costs
        # ~10 us per call, and a training step asks half a dozen times) and again only if the module object is replaced
        cache = self.__dict__.setdefault("_param_lists", {})
        ent = cache.get(id(net))
        if ent is None or ent[0] is not net:
            ent = cache[id(net)] = (net, list(net.parameters()))
        return tuple((p.data_ptr(), p._version) for p in ent[1])

    def _train_state(self, device):
        """Flat fp64 device vectors (theta, Adam moments, target) shared with the kernels.  The nn modules stay the
        public view: cuda fp64 modules are re-pointed at slices of the flat vectors (zero copy), others are refreshed
        after every step; edits made to the modules from outside are picked up (parameter version counters)."""
        t = self._train
        if t is None or t["device"] != device:
            P = ops.ratio_num_params(self.in_size)
            z = lambda: torch.zeros(P, dtype=torch.float64, device=device)
            t = self._train = dict(device=device, theta=z(), m=z(), v=z(), target=z(), step=0, tok_c=None, tok_t=None,
                                   loss=torch.zeros((), dtype=torch.float64, device=device), ws=None, views=False)
            t["weights"] = ops.RatioWeights.from_flat(t["target"], self.in_size)
        flatten = lambda net: torch.cat([p.detach().reshape(-1) for p in net.parameters()]).to(device, torch.float64)
        if t["tok_c"] != self._token(self.classifier):
            t["theta"].copy_(flatten(self.classifier))
        if t["tok_t"] != self._token(self._target_network):
            t["target"].copy_(flatten(self._target_network))
        if t["tok_c"] is None:
            p0 = next(self.classifier.parameters())
            if p0.device == device and p0.dtype == torch.float64:
                for net, flat in ((self.classifier, t["theta"]), (self._target_network, t["target"])):
                    torch.nn.utils.vector_to_parameters(flat, net.parameters())
                t["views"] = True
        t["tok_c"], t["tok_t"] = self._token(self.classifier), self._token(self._target_network)
        return t

    def prepare_device_training(self, device):
        """Create the flat device vectors now: from here on `device_weights(device)` hands out views of the SAME target
        vector, which training updates in place (needed by CUDA graphs that captured the pointers)."""
        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self._publish(self._train_state(device))

    def _publish(self, t):
        if not t["views"]:
            with torch.no_grad():
                for net, flat in ((self.classifier, t["theta"]), (self._target_network, t["target"])):
                    p0 = next(net.parameters())
                    torch.nn.utils.vector_to_parameters(flat.to(device=p0.device, dtype=p0.dtype), net.parameters())
            self._weights_cache = None
        t["tok_c"], t["tok_t"] = self._token(self.classifier), self._token(self._target_network)

    def device_weights(self, device) -> "ops.RatioWeights":
        """fp64 device copies of the target-network weights for the kernels (refreshed after every update)."""
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        t = self._train
        if t is not None and t["device"] == device:
            return self._train_state(device)["weights"]       # the flat target vector itself: always current
        if not hasattr(self, "_target_params"):
            self._target_params = list(self._target_network.parameters())
        key = (device, tuple(p._version for p in self._target_params))
        if self._weights_cache is None or self._weights_cache[0] != key:
            self._weights_cache = (key, ops.RatioWeights.from_sequential(self._target_network, device))
        return self._weights_cache[1]

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        if not inputs.is_cuda:        # like every other entry point: no CPU path, never a silent torch fallback
            raise _lib.CboError("libcbo has no CPU path: RatioEstimator.forward needs CUDA inputs")
        return ops.ratio_forward(self.device_weights(inputs.device), inputs).to(inputs.dtype)

    def update_target_network(self):
        with torch.no_grad():
            for src_p, tgt_p in zip(self.classifier.parameters(), self._target_network.parameters()):
                tgt_p.mul_(1.0 - self._ema_weight)
                tgt_p.add_(self._ema_weight * src_p)

    def _training_device(self, xk) -> torch.device:
        for cand in (self.device, getattr(self._neg_samples, "device", None), xk.device):
            if cand is not None and torch.device(cand).type == "cuda":
                dev = torch.device(cand)
                return dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())
        raise _lib.CboError("RatioEstimator.optimize_callback trains on the device (libcbo has no CPU path): "
                            "construct the estimator with a CUDA device or pass CUDA samples")

    def _append_positives(self, xk, device):
        n = xk.shape[0]
        if self._pos_pool is None or self._pos_pool.device != device:
            old = self._pos_pool
            self._pos_pool = torch.empty(max(1024, 2 * n), self.in_size, dtype=torch.float64, device=device)
            if old is not None:
                self._pos_pool[:self._num_pos].copy_(old[:self._num_pos])
        if self._num_pos + n > self._pos_pool.shape[0]:
            grown = torch.empty(2 * (self._num_pos + n), self.in_size, dtype=torch.float64, device=device)
            grown[:self._num_pos].copy_(self._pos_pool[:self._num_pos])
            self._pos_pool = grown
        self._pos_pool[self._num_pos:self._num_pos + n].copy_(xk, non_blocking=True)
        self._num_pos += n

    def optimize_callback(self, xk):
        """ratio_estimation.py:110-144: collect the SGLD iterates as positives; once there are at least as many as
        negatives, one Adam step on BCE-with-logits over [negatives; an equally large random subset of the positives],
        then the EMA update of the target network.  Everything after the host-side bookkeeping runs in two kernels."""
        if isinstance(xk, np.ndarray):
            xk = torch.from_numpy(xk)
        xk = xk.detach().reshape(-1, self.in_size)
        device = self._training_device(xk)
        self._append_positives(xk, device)
        if self._neg_samples is None:
            return None
        num_negative, num_positive = self._neg_samples.size(0), self._num_pos
        if num_positive < num_negative:
            return None
        if self._neg_dev is None or self._neg_dev[0] is not self._neg_samples or self._neg_dev[1].device != device:
            self._neg_dev = (self._neg_samples, self._neg_samples.detach().to(device, torch.float64).contiguous())
        t = self._train_state(device)
        if self.device_permutation:
            if self._perm_idx is None or self._perm_idx.numel() < num_negative or self._perm_idx.device != device:
                self._perm_idx = torch.empty(num_negative, dtype=torch.int64, device=device)
            idx = ops.sample_subset(num_positive, num_negative, self._perm_seed + 0x9E3779B97F4A7C15 * (t["step"] + 1), device,
                                    out=self._perm_idx)
        else:       # the reference's generator and draw (:129-131), 8 bytes per row to the device
            from .helpers import _PinnedRing
            idx = _PinnedRing.upload(torch.from_numpy(np.random.permutation(num_positive)[:num_negative].copy()), device)
        t["step"] += 1
        t["ws"] = ops.ratio_train_step(t["theta"], t["m"], t["v"], t["target"], x_neg=self._neg_dev[1],
                                       x_pos=self._pos_pool, pos_idx=idx, n_pos=num_negative, step=t["step"],
                                       lr=self._lr, weight_decay=self._weight_decay, ema_weight=self._ema_weight,
                                       loss=t["loss"], workspace=t["ws"])
        self.last_loss = t["loss"]
        self._publish(t)
        return None
