"""Conformal acquisition classes with the reference's interface (conformalbo/acquisitions.py).

`forward(X: b x q x d) -> Tensor(b)` is one call into libcbo: the fused kernels take over
`ConformalAcquisition._conformal_fwd` as a whole (construct_conformal_bands + the refit-model
`_nonconformal_fwd` + `_conformal_integration`).  Constructor keywords follow the
reference (acquisitions.py:80-86, 107-113, 143-149) and the BoTorch bases it mixes in
(`model`, `beta` / `best_f`, `sampler`, `objective`, `posterior_transform`, `X_pending`).

Joint batches q <= 4 (incl. X_pending) are supported.  `forward` is differentiable in X for cUCB / cEI / cNEI
(q = 1 and joint batches up to 4, the reference's default SGLD loop uses q_batch_size = 2): when X
requires grad, ONE call of cbo_conformal_acq_grad_f64 returns the values and d acq / d X, and a custom
autograd.Function hands the stored Jacobian blocks to `backward` (what optimize_acqf_sgld / L-BFGS-B need:
conformalbo/optim/optimize.py:224-235; cEHVI / cNEHVI: mobo_acquisitions.py).  Not built (raise, never silently fall
back): non-identity objectives / posterior transforms.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
from torch import nn

from . import ops
from .helpers import _ratio_weights, as_state, draw_probes
from .sampling import IIDNormalSampler


class _ConformalAcqFn(torch.autograd.Function):
    """acq = f(X) with the per-candidate Jacobian rows computed by the kernels in the same call (acq[c] depends on X[c]
    only, so the vector-Jacobian product is a row scaling)."""

    @staticmethod
    def forward(ctx, X, acq, dacq_dX):
        ctx.save_for_backward(dacq_dX)
        return acq.clone()

    @staticmethod
    def backward(ctx, grad_out):
        (dacq_dX,) = ctx.saved_tensors
        return (grad_out.reshape(-1, 1, 1) * dacq_dX).to(dacq_dX.dtype), None, None


class ConformalAcquisition(nn.Module):
    """Mixin state of acquisitions.py:29-39 plus the BoTorch MC-acquisition attributes the callers probe
    (`model`, `X_pending`, `set_X_pending`, `ratio_estimator`; conformalbo/utils.py:243, optim/optimize.py:214)."""

    _kind = None

    def __init__(self, alpha, temp, grid_res, max_grid_refinements, ratio_estimator, optimistic=False,
                 grid_sampler=None, randomized=False, *, model=None, sampler=None, objective=None,
                 posterior_transform=None, X_pending=None):
        super().__init__()
        if objective is not None or posterior_transform is not None:
            raise NotImplementedError("only the identity objective / no posterior transform are built")
        self.alpha = alpha
        self.temp = temp
        self.grid_res = grid_res
        self.max_grid_refinements = max_grid_refinements
        self.ratio_estimator = ratio_estimator
        self.optimistic = optimistic
        self.randomized = randomized
        self.grid_sampler = IIDNormalSampler(grid_res) if grid_sampler is None else grid_sampler
        self.model = model
        # every conformal hydra config uses IIDNormalSampler(num_samples=1, batch_range=(0,-3)) (acq_fn/c*.yaml)
        self.sampler = IIDNormalSampler(1, batch_range=(0, -3)) if sampler is None else sampler
        self.X_pending = None
        self.set_X_pending(X_pending)
        self.fixed_probes: Optional[torch.Tensor] = None      # inject z; None: fresh CPU draw per call
        self.fixed_uniforms: Optional[torch.Tensor] = None    # inject the Bernoulli uniforms (App. C-3)
        self.device_probes = False        # True: draw the probes from the device generator (no host RNG / upload per call)
        # candidate-sharded scoring with max_grid_refinements > 0: distributed.refine_reducer() makes the refinement
        # decisions batch-global again; None: per-call statistics
        self.refine_reduce = None
        # the plain cUCB / cEI call goes through the dispatcher-registered operator torch.ops.cbo.conformal_acq (csrc/torch_ext.cpp,
        # Autograd kernel included); False (or CBO_NO_TORCH_EXT=1): the ctypes binding of the same C entry points
        self.use_torch_ops = os.environ.get("CBO_NO_TORCH_EXT", "0") != "1"
        # d acq / d X of joint batches (q + pending >= 2, no grid refinement; every acquisition class): False = the target grid is a constant
        # (the intent of the reference's no_grad block); True = also the gradient the reference's autograd leaks through the
        # lazily evaluated prior block of the grid covariance (cbo_params.grid_grad_leak)
        self.grid_grad_leak = False
        self.last_outputs = {}

    def set_X_pending(self, X_pending=None):
        self.X_pending = None if X_pending is None else X_pending.detach().clone()

    # --- the reference's hooks, kept for API parity -------------------------------------------------------
    def _conformalize_model(self, X):
        if self.model is None:
            raise AttributeError("ConformalAcquisition must be subclassed")     # acquisitions.py:42-43
        from .helpers import construct_conformal_bands
        return construct_conformal_bands(self.model, X, self.alpha, self.temp, self.grid_res, self.max_grid_refinements,
                                         self.grid_sampler, self.ratio_estimator, mask_ood=True,
                                         randomized=self.randomized, probes=self.fixed_probes, u_rand=self.fixed_uniforms)

    def _acq_params(self):
        raise NotImplementedError("ConformalAcquisition must be subclassed")     # acquisitions.py:50-51

    def _mc_width(self, q):
        """Trailing dimension of the acquisition sampler's base samples (q; nb + q for the noisy-EI joint sample)."""
        return q

    def value_and_grad(self, X: torch.Tensor):
        """(acq (b,), d acq / d X (b, q, d)) from ONE library call, without an autograd graph: what a device-side
        optimiser loop needs per step (optim.optimize_acqf_sgld)."""
        return self._conformal_fwd(X.detach(), raw_grad=True)

    def graphed(self, b: int, q: int, want_grad: bool = True):
        """This acquisition function for ONE fixed call shape (b, q) as a CUDA graph (`graphs.GraphedConformalAcq`): same
        base samples, parameters and ratio weights as `forward`; raises ValueError where a capture is impossible."""
        from .graphs import GraphedConformalAcq
        if self.model is None:
            raise AttributeError("ConformalAcquisition must be subclassed")
        if self.max_grid_refinements > 0 or self.X_pending is not None:
            raise ValueError("grid refinement / pending points are not captured")
        if not self.grid_sampler.collapse_batch_dims or self.grid_sampler.resample or self.sampler.resample:
            raise ValueError("base samples that change per call are not captured")
        if self._kind not in ("ucb", "ei", "nei"):
            raise ValueError("one library call per step is captured; the multi-objective acquisitions take several")
        state = as_state(self.model)
        dev, G = state.device, self.grid_res
        self.grid_sampler._sample_shape = torch.Size([G])
        eps = self.grid_sampler.draw((G, q), dev)
        eps_mc = self.sampler.draw((int(self.sampler.sample_shape[0]), G, self._mc_width(q)), dev)
        if self.ratio_estimator is not None and hasattr(self.ratio_estimator, "prepare_device_training"):
            self.ratio_estimator.prepare_device_training(dev)      # weights at a stable address while training goes on
        return GraphedConformalAcq(state, b, q, eps_grid=eps, eps_mc=eps_mc, ratio=_ratio_weights(self.ratio_estimator, dev),
                                   want_grad=want_grad, alpha=self.alpha, temp=self.temp, grid_res=G,
                                   max_grid_refinements=0, randomized=self.randomized, optimistic=self.optimistic,
                                   device_probes=self.device_probes,
                                   grid_grad_leak=bool(self.grid_grad_leak) and want_grad and self._kind in ("ucb", "ei", "nei"),
                                   **self._acq_params())

    def _conformal_fwd(self, X: torch.Tensor, outputs=(), raw_grad=False) -> torch.Tensor:
        if self.model is None:
            raise AttributeError("ConformalAcquisition must be subclassed")
        want_grad = raw_grad or (X.requires_grad and torch.is_grad_enabled())
        X_in = X
        state = as_state(self.model)
        if X.ndim == 2:                                          # t_batch_mode_transform: add the t-batch dim
            X = X.unsqueeze(0)
        if self.X_pending is not None:                           # concatenate_pending_points
            X = torch.cat([X, self.X_pending.to(X).expand(*X.shape[:-2], -1, -1)], dim=-2)
        lead, (q, d) = X.shape[:-2], X.shape[-2:]
        dev = X.device
        Xf = X.detach().reshape(-1, q, d)
        b, G = Xf.shape[0], self.grid_res
        self.grid_sampler._sample_shape = torch.Size([G])
        if self.grid_sampler.collapse_batch_dims:
            eps = self.grid_sampler.draw((G, q), dev)                              # shared by all candidates (App. B-7)
        else:
            eps = self.grid_sampler.draw((G, b, q), dev).movedim(0, 1).contiguous()
        S = int(self.sampler.sample_shape[0])
        eps_mc = self.sampler.draw((S, G, self._mc_width(q)), dev)
        z = self.fixed_probes.to(dev) if self.fixed_probes is not None else draw_probes(state.N + q, device=dev,
                                                                                        on_device=self.device_probes)
        u = self.fixed_uniforms
        if self.randomized and u is None:
            u = torch.rand(b, G, q, dtype=torch.float64, device=dev)
        rr = self.refine_reduce if self.max_grid_refinements > 0 else None
        leak = bool(self.grid_grad_leak) and want_grad and self._kind in ("ucb", "ei", "nei")
        if self.use_torch_ops and self._kind in ("ucb", "ei") and not outputs and not raw_grad and rr is None and not leak:
            from . import torch_ops
            acq = torch_ops.conformal_acq(state, X.reshape(-1, q, d).to(torch.float64), eps_grid=eps, z=z, eps_mc=eps_mc, u_rand=u,
                                          ratio=_ratio_weights(self.ratio_estimator, dev), alpha=self.alpha, temp=self.temp,
                                          grid_res=G, max_grid_refinements=self.max_grid_refinements,
                                          randomized=self.randomized, optimistic=self.optimistic, **self._acq_params())
            self.last_outputs = {"acq": acq.detach()}
            return acq.reshape(lead).to(X.dtype)
        out = ops.conformal_acq(state, Xf, eps_grid=eps, z=z, eps_mc=eps_mc, u_rand=u,
                                ratio=_ratio_weights(self.ratio_estimator, dev), outputs=outputs, want_grad=want_grad,
                                refine_reduce=self.refine_reduce if self.max_grid_refinements > 0 else None,
                                alpha=self.alpha, temp=self.temp, grid_res=G,
                                max_grid_refinements=self.max_grid_refinements, randomized=self.randomized,
                                optimistic=self.optimistic, grid_grad_leak=leak, **self._acq_params())
        self.last_outputs = out
        if raw_grad:
            q_in = (X_in.unsqueeze(0) if X_in.ndim == 2 else X_in).shape[-2]
            return out["acq"].reshape(lead), out["dacq_dX"][:, :q_in, :].reshape(*lead, q_in, d)
        if want_grad:
            Xg = X_in.unsqueeze(0) if X_in.ndim == 2 else X_in
            q_in = Xg.shape[-2]                                          # pending points (appended last) carry no gradient
            acq = _ConformalAcqFn.apply(Xg.reshape(-1, q_in, d), out["acq"], out["dacq_dX"][:, :q_in, :].to(X.dtype))
            return acq.reshape(lead).to(X.dtype)
        return out["acq"].reshape(lead).to(X.dtype)

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        return self._conformal_fwd(X)                            # max over q / mean over S happen in the kernel epilogue


class qConformalUpperConfidenceBound(ConformalAcquisition):
    """acquisitions.py:142-164 (+ botorch qUpperConfidenceBound(model, beta, sampler, ...))."""

    def __init__(self, alpha, temp, grid_res, max_grid_refinements, ratio_estimator, optimistic=True,
                 grid_sampler=None, randomized=False, *, model=None, beta=None, sampler=None, objective=None,
                 posterior_transform=None, X_pending=None):
        super().__init__(alpha, temp, grid_res, max_grid_refinements, ratio_estimator, optimistic, grid_sampler,
                         randomized, model=model, sampler=sampler, objective=objective,
                         posterior_transform=posterior_transform, X_pending=X_pending)
        if beta is None:
            raise TypeError("qConformalUpperConfidenceBound needs beta")
        self.beta_prime = math.sqrt(beta * math.pi / 2)          # botorch qUpperConfidenceBound (App. B-10)

    _kind = "ucb"

    def _acq_params(self):
        return dict(kind="ucb", beta_prime=self.beta_prime)


class qConformalExpectedImprovement(ConformalAcquisition):
    """acquisitions.py:79-103 (+ botorch qExpectedImprovement(model, best_f, sampler, ...))."""

    def __init__(self, alpha, temp, grid_res, max_grid_refinements, ratio_estimator, optimistic=False,
                 grid_sampler=None, randomized=False, *, model=None, best_f=None, sampler=None, objective=None,
                 posterior_transform=None, X_pending=None):
        super().__init__(alpha, temp, grid_res, max_grid_refinements, ratio_estimator, optimistic, grid_sampler,
                         randomized, model=model, sampler=sampler, objective=objective,
                         posterior_transform=posterior_transform, X_pending=X_pending)
        if best_f is None:
            raise TypeError("qConformalExpectedImprovement needs best_f")
        self.register_buffer("best_f", torch.as_tensor(best_f, dtype=torch.float64).reshape(()))

    _kind = "ei"

    def _acq_params(self):
        return dict(kind="ei", best_f=float(self.best_f))


class qConformalNoisyExpectedImprovement(ConformalAcquisition):
    """acquisitions.py:106-139 (+ botorch qNoisyExpectedImprovement(model, X_baseline, sampler, ..., prune_baseline)).

    Joint sampling over [X_baseline; X] under every refit model.  Built for q <= 4 (incl. pending points) and nb + q <= 64
    baseline + candidate points (one Cholesky of the joint (nb+q)^2 covariance per joint batch in shared memory).
    `prune_baseline=True` (every reference config: hydra_config/acq_fn/cnei.yaml:3) runs botorch's `prune_inferior_points`
    at construction (`sampling.prune_inferior_points`; base samples `prune_base_samples` (S, nb) or Sobol with `prune_seed`)."""

    def __init__(self, alpha, temp, grid_res, max_grid_refinements, ratio_estimator, optimistic=False,
                 grid_sampler=None, randomized=False, *, model=None, X_baseline=None, sampler=None, objective=None,
                 posterior_transform=None, X_pending=None, prune_baseline=False, cache_root=False, prune_base_samples=None,
                 prune_seed=None):
        super().__init__(alpha, temp, grid_res, max_grid_refinements, ratio_estimator, optimistic, grid_sampler,
                         randomized, model=model, sampler=sampler, objective=objective,
                         posterior_transform=posterior_transform, X_pending=X_pending)
        if X_baseline is None:
            raise TypeError("qConformalNoisyExpectedImprovement needs X_baseline")
        if prune_baseline:
            from .sampling import prune_inferior_points
            X_baseline = prune_inferior_points(model, X_baseline, base_samples=prune_base_samples, seed=prune_seed)
        if X_baseline.reshape(-1, X_baseline.shape[-1]).shape[0] > 60:
            raise NotImplementedError("cNEI is built for nb + q <= 64 joint points: prune the baseline (prune_baseline=True)")
        self.X_baseline = X_baseline
        self._baseline = None

    def _mc_width(self, q):
        return self.X_baseline.reshape(-1, self.X_baseline.shape[-1]).shape[0] + q

    def _acq_params(self):
        state = as_state(self.model)
        if self._baseline is None or self._baseline[0] is not state:        # rebuilt whenever the model state changes
            self._baseline = (state, ops.NeiBaseline.build(state, self.X_baseline))
        return dict(kind="nei", nei_baseline=self._baseline[1])
