"""MultivariateNormal / MultitaskMultivariateNormal with the calls the reference makes (gpytorch 1.9.0 semantics):
mean, lazy_covariance_matrix, variance, log_prob (joint, Cholesky), rsample(base_samples) = mean + chol(cov) @ eps."""
import math

import torch

from .lazy import DenseLazy, lazify


class MultivariateNormal:
    def __init__(self, mean, covariance_matrix):
        self.loc = mean
        self._covar = lazify(covariance_matrix)

    @property
    def mean(self):
        return self.loc

    @property
    def lazy_covariance_matrix(self):
        return self._covar

    @property
    def covariance_matrix(self):
        return self._covar.evaluate()

    @property
    def batch_shape(self):
        return self.loc.shape[:-1]

    @property
    def event_shape(self):
        return self.loc.shape[-1:]

    @property
    def base_sample_shape(self):
        return self.event_shape

    @property
    def variance(self):
        return torch.diagonal(self._covar.evaluate(), dim1=-2, dim2=-1).expand(self.loc.shape)

    def _flat_log_prob(self, value):
        diff = value - self.loc                                        # (..., n)
        L = self._covar.cholesky()
        Lb = L.expand(*diff.shape[:-1], *L.shape[-2:])
        sol = torch.linalg.solve_triangular(Lb, diff.unsqueeze(-1), upper=False).squeeze(-1)
        logdet = 2.0 * torch.diagonal(L, dim1=-2, dim2=-1).log().sum(-1)
        return -0.5 * (sol.pow(2).sum(-1) + logdet + diff.shape[-1] * math.log(2 * math.pi))

    def log_prob(self, value):
        return self._flat_log_prob(value)

    def _flat_rsample(self, sample_shape, base_samples):
        """base_samples: sample_shape + loc.shape."""
        L = self._covar.cholesky()
        return self.loc + (L @ base_samples.unsqueeze(-1)).squeeze(-1)

    def rsample(self, sample_shape=torch.Size(), base_samples=None):
        if base_samples is None:
            base_samples = torch.randn(*sample_shape, *self.loc.shape, dtype=self.loc.dtype)
        return self._flat_rsample(sample_shape, base_samples)


class MultitaskMultivariateNormal(MultivariateNormal):
    """mean (..., q, t); covariance over the flattened event.  interleaved=False: flat index = task * q + point
    (block diagonal over tasks, as MultitaskMultivariateNormal.from_independent_mvns builds it)."""

    def __init__(self, mean, covariance_matrix, interleaved=True):
        self._output_shape = mean.shape
        self._interleaved = interleaved
        flat = mean.reshape(*mean.shape[:-2], -1) if interleaved else mean.transpose(-1, -2).reshape(*mean.shape[:-2], -1)
        super().__init__(flat, covariance_matrix)

    @classmethod
    def from_independent_mvns(cls, mvns):
        mean = torch.stack([m.loc for m in mvns], dim=-1)              # (..., q, t)
        covs = [m.lazy_covariance_matrix for m in mvns]
        q, t = mean.shape[-2], mean.shape[-1]

        def thunk():
            blocks = [c.evaluate() for c in covs]
            out = torch.zeros(*blocks[0].shape[:-2], q * t, q * t, dtype=mean.dtype)
            for k, B in enumerate(blocks):
                out[..., k * q:(k + 1) * q, k * q:(k + 1) * q] = B
            return out

        return cls(mean, DenseLazy(thunk=thunk, shape=(*mean.shape[:-2], q * t, q * t)), interleaved=False)

    @property
    def mean(self):
        if self._interleaved:
            return self.loc.view(self._output_shape)
        q, t = self._output_shape[-2:]
        return self.loc.view(*self._output_shape[:-2], t, q).transpose(-1, -2)

    @property
    def batch_shape(self):
        return self._output_shape[:-2]

    @property
    def event_shape(self):
        return self._output_shape[-2:]

    @property
    def base_sample_shape(self):
        return self._output_shape[-2:]

    @property
    def variance(self):
        v = torch.diagonal(self._covar.evaluate(), dim1=-2, dim2=-1)
        if self._interleaved:
            return v.reshape(self._output_shape)
        q, t = self._output_shape[-2:]
        return v.reshape(*self._output_shape[:-2], t, q).transpose(-1, -2)

    def log_prob(self, value):
        # gpytorch: interleaved -> flatten (q, t) row-major; else transpose first, then flatten
        if self._interleaved:
            flat = value.reshape(*value.shape[:-2], -1)
        else:
            flat = value.transpose(-1, -2).reshape(*value.shape[:-2], -1)
        return self._flat_log_prob(flat)

    def rsample(self, sample_shape=torch.Size(), base_samples=None):
        if base_samples is None:
            base_samples = torch.randn(*sample_shape, *self._output_shape, dtype=self.loc.dtype)
        # gpytorch views the (…, q, t) base samples as the flat event WITHOUT transposing, also when the distribution is
        # not interleaved; only the drawn samples are un-flattened with the task-major layout.
        sample_shape = base_samples.shape[: base_samples.ndim - len(self._output_shape)]
        flat = base_samples.reshape(*sample_shape, *self.loc.shape)
        s = self._flat_rsample(sample_shape, flat)
        if self._interleaved:
            return s.view(*sample_shape, *self._output_shape)
        q, t = self._output_shape[-2:]
        return s.view(*sample_shape, *self._output_shape[:-2], t, q).transpose(-1, -2).contiguous()
