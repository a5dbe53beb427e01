"""GPyTorchPosterior (botorch 0.7): wraps an MVN; single-output posteriors carry a trailing output dim of size 1."""
import torch
from gpytorch.distributions import MultitaskMultivariateNormal


class GPyTorchPosterior:
    def __init__(self, mvn):
        self.mvn = mvn
        self._is_mt = isinstance(mvn, MultitaskMultivariateNormal)

    @property
    def dtype(self):
        return self.mvn.loc.dtype

    @property
    def device(self):
        return self.mvn.loc.device

    @property
    def event_shape(self):
        shape = self.mvn.batch_shape + self.mvn.event_shape
        return shape if self._is_mt else shape + torch.Size([1])

    @property
    def base_sample_shape(self):
        shape = self.mvn.batch_shape + self.mvn.base_sample_shape
        return shape if self._is_mt else shape + torch.Size([1])

    @property
    def mean(self):
        return self.mvn.mean if self._is_mt else self.mvn.mean.unsqueeze(-1)

    @property
    def variance(self):
        return self.mvn.variance if self._is_mt else self.mvn.variance.unsqueeze(-1)

    def rsample(self, sample_shape=None, base_samples=None):
        sample_shape = torch.Size() if sample_shape is None else sample_shape
        if base_samples is not None:
            if base_samples.shape[:len(sample_shape)] != sample_shape:
                raise RuntimeError("sample_shape disagrees with shape of base_samples.")
            base_samples = base_samples.expand(sample_shape + self.event_shape)
            if not self._is_mt:
                base_samples = base_samples.squeeze(-1)
        samples = self.mvn.rsample(sample_shape=sample_shape, base_samples=base_samples)
        return samples if self._is_mt else samples.unsqueeze(-1)
