"""MC acquisition base classes as attribute bags (botorch 0.7 constructor signatures).  Their own
`forward`s are never reached on the reference path (the conformal subclasses override them)."""
import math

import torch

from ..sampling.samplers import SobolQMCNormalSampler
from .objective import IdentityMCObjective


class MCAcquisitionFunction(torch.nn.Module):
    def __init__(self, model, sampler=None, objective=None, posterior_transform=None, X_pending=None):
        super().__init__()
        self.model = model
        self.sampler = SobolQMCNormalSampler(512, collapse_batch_dims=True) if sampler is None else sampler
        self.objective = IdentityMCObjective() if objective is None else objective
        self.posterior_transform = posterior_transform
        self.X_pending = None
        self.set_X_pending(X_pending)

    def set_X_pending(self, X_pending=None):
        self.X_pending = X_pending.detach().clone() if X_pending is not None else None


class qExpectedImprovement(MCAcquisitionFunction):
    def __init__(self, model, best_f, sampler=None, objective=None, posterior_transform=None, X_pending=None, **kwargs):
        super().__init__(model, sampler, objective, posterior_transform, X_pending)
        self.register_buffer("best_f", torch.as_tensor(best_f, dtype=float))


class qNoisyExpectedImprovement(MCAcquisitionFunction):
    def __init__(self, model, X_baseline, sampler=None, objective=None, posterior_transform=None, X_pending=None,
                 prune_baseline=False, cache_root=True, **kwargs):
        super().__init__(model, sampler, objective, posterior_transform, X_pending)
        if prune_baseline:
            raise NotImplementedError("prune_baseline (prune_inferior_points, an MC heuristic) is outside the stand-in")
        self.register_buffer("X_baseline", X_baseline)
        self._cache_root = False


class qUpperConfidenceBound(MCAcquisitionFunction):
    def __init__(self, model, beta, sampler=None, objective=None, posterior_transform=None, X_pending=None):
        super().__init__(model, sampler, objective, posterior_transform, X_pending)
        self.beta_prime = math.sqrt(beta * math.pi / 2)
