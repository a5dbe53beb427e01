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


def _posterior_samples(state, X: torch.Tensor, num_samples: int, base_samples: Optional[torch.Tensor], seed: Optional[int]):
    """Samples of the noise-free posterior f(X) (num_samples, n) on the host: mean and covariance from the library's own
    kernels (ops.NeiBaseline.build), factor with gpytorch's escalating jitter, standard-normal base samples from a scrambled
    Sobol sequence through the inverse normal CDF like botorch's SobolQMCNormalSampler (IID beyond the Sobol dimension limit)."""
    from . import ops
    bl = ops.NeiBaseline.build(state, X)
    mean, cov = bl.mub.cpu(), bl.Sbb.cpu()
    n = mean.shape[0]
    cov = 0.5 * (cov + cov.T)
    L, info = torch.linalg.cholesky_ex(cov)
    jitter = 1e-8
    while int(info) != 0 and jitter <= 1e-2:
        L, info = torch.linalg.cholesky_ex(cov + jitter * torch.eye(n, dtype=cov.dtype))
        jitter *= 10.0
    if int(info) != 0:
        raise RuntimeError("posterior covariance of the baseline is not positive definite")
    if base_samples is None:
        if n <= torch.quasirandom.SobolEngine.MAXDIM:
            u = torch.quasirandom.SobolEngine(n, scramble=True, seed=seed).draw(num_samples, dtype=torch.float64)
            v = 0.5 + (1 - 1e-10) * (u - 0.5)
            base_samples = torch.erfinv(2 * v - 1) * 2 ** 0.5
        else:
            g = torch.Generator().manual_seed(0 if seed is None else seed)
            base_samples = torch.randn(num_samples, n, generator=g, dtype=torch.float64)
    base_samples = base_samples.detach().to("cpu", torch.float64).reshape(-1, n)
    return mean.unsqueeze(0) + base_samples @ L.T


def prune_inferior_points(model, X: torch.Tensor, num_samples: int = 2048, max_frac: float = 1.0,
                          base_samples: Optional[torch.Tensor] = None, seed: Optional[int] = None) -> torch.Tensor:
    """botorch.acquisition.utils.prune_inferior_points (what `prune_baseline=True` runs in qNoisyExpectedImprovement.__init__,
    hydra_config/acq_fn/cnei.yaml:3): keep the points of X (n, d) that are the best one in at least one joint posterior sample;
    if more than ceil(max_frac n) survive, the most frequent winners.  A construction-time host step (like botorch's)."""
    import math
    from .helpers import as_state
    state = as_state(model)
    X2 = X.reshape(-1, X.shape[-1])
    samples = _posterior_samples(state, X2, num_samples, base_samples, seed)          # (S, n)
    is_best = samples.argmax(-1)
    idcs, counts = torch.unique(is_best, return_counts=True)
    max_points = math.ceil(max_frac * X2.shape[0])
    if idcs.numel() > max_points:
        order = torch.argsort(counts, descending=True, stable=True)
        idcs = idcs[order[:max_points]]
    return X2[idcs.to(X2.device)]


def is_non_dominated(Y: torch.Tensor) -> torch.Tensor:
    """(..., n, m) -> (..., n) bool: rows not dominated by another row (maximisation), as botorch.utils.multi_objective.pareto."""
    ge = (Y.unsqueeze(-3) >= Y.unsqueeze(-2)).all(-1)       # [.., i, j]: Y_j >= Y_i in every objective
    gt = (Y.unsqueeze(-3) > Y.unsqueeze(-2)).any(-1)
    return ~(ge & gt).any(-1)


def prune_inferior_points_multi_objective(models, X: torch.Tensor, ref_point: torch.Tensor, num_samples: int = 2048,
                                          max_frac: float = 1.0, base_samples: Optional[torch.Tensor] = None,
                                          seed: Optional[int] = None) -> torch.Tensor:
    """botorch prune_inferior_points_multi_objective (`prune_baseline=True` in qNoisyExpectedHypervolumeImprovement.__init__,
    scripts/bbo_multi_obj.py:253-258): keep the points that are Pareto-optimal and better than the reference point in at least
    one joint posterior sample of the m independent outputs.  base_samples: (S, n, m)."""
    import math
    from .helpers import as_state
    X2 = X.reshape(-1, X.shape[-1])
    n = X2.shape[0]
    cols = []
    for k, mdl in enumerate(models):
        bs = None if base_samples is None else base_samples[..., k]
        cols.append(_posterior_samples(as_state(mdl), X2, num_samples, bs, None if seed is None else seed + k))
    Y = torch.stack(cols, -1)                                                         # (S, n, m)
    ref = torch.as_tensor(ref_point, dtype=torch.float64).reshape(1, 1, -1)
    pareto = is_non_dominated(Y) & (Y > ref).all(-1)                                  # (S, n)
    probs = pareto.double().mean(0)
    idcs = probs.nonzero().view(-1)
    max_points = math.ceil(max_frac * n)
    if idcs.numel() > max_points:
        order = torch.argsort(probs[idcs], descending=True, stable=True)
        idcs = idcs[order[:max_points]]
    return X2[idcs.to(X2.device)]
