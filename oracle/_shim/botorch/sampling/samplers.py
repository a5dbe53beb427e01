"""IIDNormalSampler (botorch 0.7 MCSampler semantics).

Base samples have shape sample_shape + posterior.base_sample_shape with the dims in `batch_range` collapsed to 1 when
`collapse_batch_dims`; they are drawn once (resample=False) and reused while the shape is unchanged; the posterior
expands them and returns mean + chol(cov) @ base.  DRAW_HOOK(shape, dtype) -> tensor replaces the private-seeded
`torch.randn` draw so that the golden generator can inject (and record) the base samples."""
import torch

DRAW_HOOK = None


class IIDNormalSampler(torch.nn.Module):
    def __init__(self, num_samples, resample=False, seed=None, collapse_batch_dims=True, batch_range=(0, -2)):
        super().__init__()
        self._sample_shape = torch.Size([num_samples])
        self.resample = resample
        self.collapse_batch_dims = collapse_batch_dims
        self.batch_range = tuple(batch_range)
        self.seed = seed if seed is not None else int(torch.randint(0, 1000000, (1,)).item())
        self.base_samples = None

    @property
    def sample_shape(self):
        return self._sample_shape

    def _get_base_sample_shape(self, posterior):
        shape = posterior.base_sample_shape
        if self.collapse_batch_dims:
            s, e = self.batch_range
            shape = shape[:s] + torch.Size([1 for _ in shape[s:e]]) + shape[e:]
        return self.sample_shape + shape

    def _shape_changed(self, shape):
        if self.base_samples is None:
            return True
        if not self.collapse_batch_dims:
            return self.base_samples.shape != shape
        s, e = self.batch_range
        old = self.base_samples.shape
        n_s = len(self.sample_shape)
        # compare everything outside the (collapsed) batch range
        return (old[:n_s + s] + old[e:]) != (shape[:n_s + s] + shape[e:]) or len(old) != len(shape)

    def forward(self, posterior):
        shape = self._get_base_sample_shape(posterior)
        if self.resample or self._shape_changed(shape):
            if DRAW_HOOK is not None:
                self.base_samples = DRAW_HOOK(shape, posterior.dtype)
            else:
                g = torch.Generator().manual_seed(self.seed)
                self.base_samples = torch.randn(shape, generator=g, dtype=posterior.dtype)
                self.seed += 1
        return posterior.rsample(sample_shape=self.sample_shape, base_samples=self.base_samples)


class SobolQMCNormalSampler(IIDNormalSampler):
    """Only referenced as a default argument on the reference path; the conformal configs always pass an IIDNormalSampler."""
