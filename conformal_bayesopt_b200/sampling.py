"""Base-sample holder with the semantics of botorch 0.7 `IIDNormalSampler` that the path relies on
(SURVEY.md App. B-7): base samples are drawn once under a private seed and reused while the
requested shape is unchanged (`resample=False`); with `collapse_batch_dims=True` the batch
dimensions inside `batch_range` are collapsed to 1, i.e. one sample vector is shared by all candidates.

botorch is not installed in this image; this class only produces the base samples -- the
posterior arithmetic (mean + chol(cov) * base) happens inside the kernels.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


class IIDNormalSampler:
    def __init__(self, num_samples: int, resample: bool = False, seed: Optional[int] = None,
                 collapse_batch_dims: bool = True, batch_range: Tuple[int, int] = (0, -2)):
        self._sample_shape = torch.Size([num_samples])
        self.resample = resample
        self.collapse_batch_dims = collapse_batch_dims
        self.batch_range = batch_range
        self.seed = seed if seed is not None else int(torch.randint(0, 1000000, (1,)).item())
        self.base_samples: Optional[torch.Tensor] = None
        self._device_copy = None              # (device, tensor): the cached base samples stay resident between calls

    @property
    def sample_shape(self) -> torch.Size:
        return self._sample_shape

    def draw(self, shape: Tuple[int, ...], device=None) -> torch.Tensor:
        """Base samples of `shape` (already collapsed by the caller), fp64, cached like botorch's `_construct_base_samples`."""
        shape = tuple(int(s) for s in shape)
        if self.resample or self.base_samples is None or tuple(self.base_samples.shape) != shape:
            g = torch.Generator().manual_seed(self.seed)
            self.base_samples = torch.randn(shape, generator=g, dtype=torch.float64)
            self._device_copy = None
            if self.resample:
                self.seed += 1
        if device is None:
            return self.base_samples
        device = torch.device(device)
        if self._device_copy is None or self._device_copy[0] != device or self._device_copy[2] is not self.base_samples:
            self._device_copy = (device, self.base_samples.to(device), self.base_samples)
        return self._device_copy[1]

    def set_base_samples(self, base: torch.Tensor):
        """Inject base samples (parity tests)."""
        self.base_samples = base.detach().to("cpu", torch.float64).clone()
        self._device_copy = None
