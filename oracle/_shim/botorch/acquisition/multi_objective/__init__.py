"""qExpectedHypervolumeImprovement / qNoisyExpectedHypervolumeImprovement: attribute bags + `_compute_qehvi`
(botorch 0.7): per box-decomposition cell prod_m clamp(min(obj, u_cell) - l_cell, 0), inclusion-exclusion
over the q-subsets, sum over cells, mean over the MC dimension."""
from itertools import combinations

import torch

from ..monte_carlo import MCAcquisitionFunction
from ..objective import IdentityMCMultiOutputObjective


class qExpectedHypervolumeImprovement(MCAcquisitionFunction):
    def __init__(self, model, ref_point, partitioning, sampler=None, objective=None, constraints=None, X_pending=None,
                 eta=1e-3, **kwargs):
        super().__init__(model, sampler, IdentityMCMultiOutputObjective() if objective is None else objective, None,
                         X_pending)
        self.constraints = constraints
        self.eta = eta
        self.register_buffer("ref_point", torch.as_tensor(ref_point, dtype=torch.float64))
        cell_bounds = partitioning.get_hypercell_bounds()
        self.register_buffer("cell_lower_bounds", cell_bounds[0])
        self.register_buffer("cell_upper_bounds", cell_bounds[1])
        self.q_out = -1
        self.q_subset_indices = {}

    def _cache_q_subset_indices(self, q_out):
        if q_out != self.q_out:
            idx = list(range(q_out))
            self.q_subset_indices = {f"q_choose_{i}": torch.tensor(list(combinations(idx, i)), dtype=torch.long)
                                     for i in range(1, q_out + 1)}
            self.q_out = q_out

    def _compute_qehvi(self, samples, X=None):
        obj = self.objective(samples, X=X)                     # (S, *batch, q, m)
        q = obj.shape[-2]
        self._cache_q_subset_indices(q)
        batch_shape = obj.shape[:-2]
        n_cells = self.cell_lower_bounds.shape[-2]
        areas = torch.zeros(*batch_shape, n_cells, dtype=obj.dtype)
        view = (*[1] * len(batch_shape), n_cells, 1, self.cell_upper_bounds.shape[-1])
        upper, lower = self.cell_upper_bounds.view(view), self.cell_lower_bounds.view(view)
        for i in range(1, q + 1):
            sub = self.q_subset_indices[f"q_choose_{i}"]                                  # (C(q,i), i)
            objs = obj.index_select(-2, sub.view(-1)).view(*obj.shape[:-2], *sub.shape, obj.shape[-1])
            vertices = objs.min(dim=-2).values                                            # (..., C, m)
            vertices = torch.min(vertices.unsqueeze(-3), upper)                           # (..., cells, C, m)
            lengths = (vertices - lower).clamp_min(0.0)
            areas = areas + (-1) ** (i + 1) * lengths.prod(-1).sum(-1)
        return areas.sum(-1).mean(0)


class qNoisyExpectedHypervolumeImprovement(qExpectedHypervolumeImprovement):
    """Constructor state only; the stand-in does not restate botorch's cached box decompositions over baseline samples."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError
