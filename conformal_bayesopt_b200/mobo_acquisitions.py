"""Multi-objective conformal acquisition (reference: conformalbo/mobo_acquisitions.py:24-45).

`qConformalExpectedHypervolumeImprovement.forward(X: b x q x d) -> Tensor(b)`: one cbo_conformal_acq_f64 call per
objective (independent GPs) and one cbo_cehvi_integrate_q_f64.  The reference's EHVI integration is shape-broken
(raises unless b in {1, G}) and scrambles multi-output labels (App. C-6); this implements the
intent -- see oracle/mobo.py.  Joint batches q <= 4 incl. pending points; `qConformalNoisyExpectedHypervolumeImprovement`
(mobo_acquisitions.py:184-219) samples f(X) jointly with the baseline under every refit model and carries one box
decomposition per MC sample.
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import ops
from .acquisitions import ConformalAcquisition
from .helpers import _ratio_weights, as_state, draw_probes
from .sampling import IIDNormalSampler


def box_decomposition_2d(ref_point: torch.Tensor, Y: torch.Tensor):
    """Cells of the non-dominated region for m = 2 (maximisation): (lower, upper), each (n_front + 1, 2).
    Host-side stand-in for botorch's FastNondominatedPartitioning(ref_point, Y).get_hypercell_bounds()."""
    ref_point = ref_point.detach().double().cpu()
    Y = Y.detach().double().cpu()
    Y = Y[(Y > ref_point).all(-1)]
    Y = Y[torch.argsort(Y[:, 0])]
    keep, best2 = [], -float("inf")
    for i in range(Y.shape[0] - 1, -1, -1):
        if Y[i, 1] > best2:
            keep.append(i)
            best2 = float(Y[i, 1])
    front = Y[keep[::-1]] if keep else Y.new_zeros(0, 2)
    n, inf = front.shape[0], float("inf")
    lower = [[float(ref_point[0]) if i == 0 else float(front[i - 1, 0]),
              float(ref_point[1]) if i == n else float(front[i, 1])] for i in range(n + 1)]
    upper = [[inf if i == n else float(front[i, 0]), inf] for i in range(n + 1)]
    return torch.tensor(lower, dtype=torch.float64), torch.tensor(upper, dtype=torch.float64)


class qConformalExpectedHypervolumeImprovement(ConformalAcquisition):
    """mobo_acquisitions.py:24-45.  `model`: a list of m independent single-output models / GPModelStates (the reference's
    batched multi-output SingleTaskGP is m independent GPs).  Joint batches q <= 4 including pending
    points; with max_grid_refinements > 0 every output refines its own grid (the reference's loop takes its decisions over
    all outputs at once, helpers.py:150-182 -- unpinned either way, App. C-5/C-6)."""

    _kind = "ehvi"

    def __init__(self, alpha, temp, grid_res, max_grid_refinements, ratio_estimator, optimistic=False,
                 grid_sampler=None, randomized=False, *, model=None, ref_point=None, partitioning=None, sampler=None,
                 objective=None, constraints=None, X_pending=None, eta=1e-3):
        super().__init__(alpha, temp, grid_res, max_grid_refinements, ratio_estimator, optimistic, grid_sampler,
                         randomized, model=model, sampler=sampler, objective=objective, X_pending=X_pending)
        assert constraints is None   # for simplicity of code (mobo_acquisitions.py:33)
        if partitioning is None:
            raise TypeError("qConformalExpectedHypervolumeImprovement needs a partitioning")
        if hasattr(partitioning, "get_hypercell_bounds"):          # botorch partitioning object
            bounds = partitioning.get_hypercell_bounds()
            lower, upper = bounds[0], bounds[1]
        else:
            lower, upper = partitioning
        self.register_buffer("cell_lower", lower.detach().double().reshape(-1, lower.shape[-1]).contiguous())
        self.register_buffer("cell_upper", upper.detach().double().reshape(-1, upper.shape[-1]).contiguous())
        self.ref_point = ref_point

    def _states(self):
        models: Sequence = self.model if isinstance(self.model, (list, tuple)) else [self.model]
        return [as_state(m) for m in models]

    _PER_OUTPUT = ("conf_pred_mask", "grid_logp", "target_grid", "post_mean", "ood_mask", "post_var")

    def _prepare(self, X):
        """Shared front end of cEHVI / cNEHVI: shapes, pending points, base samples, probes, uniforms."""
        if self.model is None:
            raise AttributeError("ConformalAcquisition must be subclassed")
        states = self._states()
        m = len(states)
        if X.ndim == 2:
            X = X.unsqueeze(0)
        if self.X_pending is not None:                           # concatenate_pending_points
            X = torch.cat([X, self.X_pending.to(X).expand(*X.shape[:-2], -1, -1)], dim=-2)
        lead, (q, d) = X.shape[:-2], X.shape[-2:]
        if q > 4:
            raise NotImplementedError("joint batches are built for q <= 4 (incl. pending points)")
        dev = X.device
        Xf = X.detach().reshape(-1, q, d)
        b, G = Xf.shape[0], self.grid_res
        self.grid_sampler._sample_shape = torch.Size([G])
        eps = self.grid_sampler.draw((G, q, m), dev)                                # base samples (G, 1, q, t)
        z = (self.fixed_probes if self.fixed_probes is not None else draw_probes(states[0].N + q)).to(dev)
        u = self.fixed_uniforms                                                     # (b, G, q, m)
        if self.randomized and u is None:
            u = torch.rand(b, G, q, m, dtype=torch.float64, device=dev)
        return states, m, X, Xf, lead, q, b, G, dev, eps, z, u

    def _per_output(self, st, k, Xf, q, G, dev, eps, z, u, **extra):
        kw = dict(kind="ucb", eps_mc=torch.zeros(1, G, q, dtype=torch.float64, device=dev), outputs=self._PER_OUTPUT)
        kw.update(extra)
        return ops.conformal_acq(
            st, Xf, eps_grid=eps[:, :, k].contiguous(), z=z, u_rand=None if u is None else u.reshape(-1, G, q, u.shape[-1])[..., k].contiguous(),
            ratio=_ratio_weights(self.ratio_estimator, dev), alpha=self.alpha, temp=self.temp, grid_res=G,
            max_grid_refinements=self.max_grid_refinements, randomized=self.randomized, optimistic=self.optimistic, **kw)

    def _per_output_grad(self, st, k, Xf, q, G, dev, eps, z, u, bars, **extra):
        """d (this output's masks, ood masks, samples) / d X contracted with the integration's adjoints `bars`."""
        kw = dict(kind="ucb")
        kw.update(extra)
        if self.max_grid_refinements > 0:
            raise NotImplementedError("d acq / d X of the multi-objective acquisitions is built without grid refinement")
        out = ops.conformal_acq(
            st, Xf, eps_grid=eps[:, :, k].contiguous(), z=z, u_rand=None if u is None else u.reshape(-1, G, q, u.shape[-1])[..., k].contiguous(),
            ratio=_ratio_weights(self.ratio_estimator, dev), alpha=self.alpha, temp=self.temp, grid_res=G,
            max_grid_refinements=0, randomized=self.randomized, optimistic=self.optimistic, want_grad=True,
            ext_adjoints=bars, grid_grad_leak=bool(self.grid_grad_leak), **kw)
        return out["dacq_dX"]

    def _value_and_grad(self, X):
        states, m, X, Xf, lead, q, b, G, dev, eps, z, u = self._prepare(X)
        S = int(self.sampler.sample_shape[0])
        eps_mc = self.sampler.draw((S, G, q, m), dev).contiguous()                  # base samples (S, 1, G, q, t)
        outs = [self._per_output(st, k, Xf, q, G, dev, eps, z, u) for k, st in enumerate(states)]
        acq, values, bars = ops.cehvi_integrate_q_grad(outs, [st.noise for st in states], eps_mc, self.cell_lower.to(dev),
                                                       self.cell_upper.to(dev), q=q, alpha=self.alpha,
                                                       optimistic=self.optimistic)
        dX = None
        for k, st in enumerate(states):
            g = self._per_output_grad(st, k, Xf, q, G, dev, eps, z, u, bars[k], eps_mc=eps_mc[..., k].contiguous())
            dX = g if dX is None else dX + g
        self.last_outputs = {"per_output": outs, "acq": acq, "values": values}
        return acq, dX, lead, q

    def value_and_grad(self, X: torch.Tensor):
        """(acq (b,), d acq / d X (b, q_in, d)): 2 m + 1 library calls (m forward, the integration with its reverse pass, m
        vector-Jacobian products); pending points carry no gradient."""
        Xd = X.detach()
        q_in = (Xd.unsqueeze(0) if Xd.ndim == 2 else Xd).shape[-2]
        acq, dX, lead, q = self._value_and_grad(Xd)
        return acq.reshape(lead), dX[:, :q_in, :].reshape(*lead, q_in, dX.shape[-1])

    def graphed(self, b, q, want_grad=True):
        raise ValueError("one library call per step is captured; the multi-objective acquisitions take several")

    def _forward_with_grad(self, X):
        from .acquisitions import _ConformalAcqFn
        Xg = X.unsqueeze(0) if X.ndim == 2 else X
        q_in, d = Xg.shape[-2:]
        acq, dX, lead, q = self._value_and_grad(X.detach())
        out = _ConformalAcqFn.apply(Xg.reshape(-1, q_in, d), acq, dX[:, :q_in, :].to(X.dtype))
        return out.reshape(lead).to(X.dtype)

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        if X.requires_grad and torch.is_grad_enabled():
            return self._forward_with_grad(X)
        states, m, X, Xf, lead, q, b, G, dev, eps, z, u = self._prepare(X)
        S = int(self.sampler.sample_shape[0])
        eps_mc = self.sampler.draw((S, G, q, m), dev).contiguous()                  # base samples (S, 1, G, q, t)
        outs = [self._per_output(st, k, Xf, q, G, dev, eps, z, u) for k, st in enumerate(states)]
        acq, values = ops.cehvi_integrate_q(outs, [st.noise for st in states], eps_mc, self.cell_lower.to(dev),
                                            self.cell_upper.to(dev), q=q, alpha=self.alpha, optimistic=self.optimistic,
                                            want_values=True)
        self.last_outputs = {"per_output": outs, "acq": acq, "values": values}
        return acq.reshape(lead).to(X.dtype)


def psd_safe_cholesky(A: torch.Tensor) -> torch.Tensor:
    """Lower Cholesky factor with gpytorch's escalating diagonal jitter (1e-8 ... 1e-3 of nothing: absolute), used for the
    small baseline covariances at construction time (host side, like the box decompositions)."""
    L, info = torch.linalg.cholesky_ex(A)
    jitter = 1e-8
    while int(info.max()) != 0 and jitter <= 1e-2:
        L, info = torch.linalg.cholesky_ex(A + jitter * torch.eye(A.shape[-1], dtype=A.dtype))
        jitter *= 10.0
    if int(info.max()) != 0:
        raise RuntimeError("baseline posterior covariance is not positive definite")
    return L


def per_sample_cells(ref_point: torch.Tensor, fb: torch.Tensor):
    """One box decomposition per MC sample of the baseline values fb (S, nb, 2) -> (lower, upper), each (S, n_cells, 2),
    padded with zero-volume cells (lower = upper = ref_point) to the largest cell count."""
    cells = [box_decomposition_2d(ref_point, fb[s]) for s in range(fb.shape[0])]
    n = max(c[0].shape[0] for c in cells)
    ref = ref_point.detach().double().cpu().reshape(1, -1)
    lo = torch.stack([torch.cat([c[0], ref.expand(n - c[0].shape[0], -1)]) for c in cells])
    up = torch.stack([torch.cat([c[1], ref.expand(n - c[1].shape[0], -1)]) for c in cells])
    return lo.contiguous(), up.contiguous()


class qConformalNoisyExpectedHypervolumeImprovement(qConformalExpectedHypervolumeImprovement):
    """mobo_acquisitions.py:184-219 (+ botorch qNoisyExpectedHypervolumeImprovement(model, ref_point, X_baseline, sampler, ...,
    cache_root=False)).  At construction the baseline values f(X_baseline) are sampled under the model (base samples
    `baseline_base_samples`, (S, nb, m)) and every MC sample gets its own box decomposition of the non-dominated region; per
    call f(X) is sampled JOINTLY with the baseline under every refit model -- the baseline part of the joint base samples is
    the stored one, shared by all grid points (botorch `_set_sampler`), the new points draw (S, G, q, m) -- and the
    hyper-volume improvement of sample s is taken against the cells of sample s.  Built for two objectives (the 2-D box
    decomposition), a (pruned: `prune_baseline=True` runs sampling.prune_inferior_points_multi_objective) baseline of
    nb + q <= 64 points, q <= 4 incl. pending points."""

    _kind = "nehvi"

    def __init__(self, alpha, temp, grid_res, max_grid_refinements, ratio_estimator, optimistic=False,
                 grid_sampler=None, randomized=False, *, model=None, ref_point=None, X_baseline=None, sampler=None,
                 objective=None, constraints=None, X_pending=None, eta=1e-3, prune_baseline=False, cache_pending=True,
                 max_iep=0, incremental_nehvi=True, cache_root=False, baseline_base_samples=None, **unused):
        ConformalAcquisition.__init__(self, alpha, temp, grid_res, max_grid_refinements, ratio_estimator, optimistic,
                                      grid_sampler, randomized, model=model, sampler=sampler, objective=objective,
                                      X_pending=X_pending)
        assert constraints is None
        if X_baseline is None or ref_point is None:
            raise TypeError("qConformalNoisyExpectedHypervolumeImprovement needs X_baseline and ref_point")
        states = self._states()
        m = len(states)
        if m != 2:
            raise NotImplementedError("the box decomposition is built for two objectives")
        Xb = X_baseline.reshape(-1, X_baseline.shape[-1])
        if prune_baseline:
            from .sampling import prune_inferior_points_multi_objective
            Xb = prune_inferior_points_multi_objective(states, Xb, ref_point, base_samples=unused.get("prune_base_samples"),
                                                       seed=unused.get("prune_seed"))
            X_baseline = Xb
        if Xb.shape[0] > 60:
            raise NotImplementedError("cNEHVI is built for nb + q <= 64 joint points: prune the baseline (prune_baseline=True)")
        self.X_baseline = X_baseline
        self.ref_point = torch.as_tensor(ref_point, dtype=torch.float64).reshape(-1)
        self._baselines = [ops.NeiBaseline.build(st, Xb) for st in states]
        S, nb = int(self.sampler.sample_shape[0]), Xb.shape[0]
        if baseline_base_samples is None:          # botorch: base_sampler = deepcopy(sampler), drawn once for the cell bounds
            baseline_base_samples = IIDNormalSampler(S, seed=self.sampler.seed).draw((S, nb, m))
        self.baseline_base_samples = baseline_base_samples.detach().to("cpu", torch.float64).reshape(S, nb, m).contiguous()
        fb = torch.stack([bl.mub.cpu().unsqueeze(0) + self.baseline_base_samples[:, :, k] @ psd_safe_cholesky(bl.Sbb.cpu()).T
                          for k, bl in enumerate(self._baselines)], dim=-1)                       # (S, nb, m)
        self.baseline_samples = fb
        lo, up = per_sample_cells(self.ref_point, fb)
        self.register_buffer("cell_lower", lo)
        self.register_buffer("cell_upper", up)

    def _joint_outputs(self, X):
        states, m, X, Xf, lead, q, b, G, dev, eps, z, u = self._prepare(X)
        S, nb = self.baseline_base_samples.shape[0], self.baseline_base_samples.shape[1]
        if nb + q > 64:
            raise NotImplementedError("cNEHVI: nb + q <= 64")
        eps_new = self.sampler.draw((S, G, q, m), dev)                                           # base samples (S, 1, G, q, t)
        eps_joint = torch.cat([self.baseline_base_samples.to(dev).unsqueeze(1).expand(S, G, nb, m), eps_new], dim=2)
        outs = [self._per_output(st, k, Xf, q, G, dev, eps, z, u, kind="joint", nei_baseline=self._baselines[k],
                                 eps_mc=eps_joint[..., k].contiguous(), outputs=self._PER_OUTPUT + ("mc_values",))
                for k, st in enumerate(states)]
        return states, Xf, lead, q, G, dev, eps, z, u, eps_joint, outs

    def _value_and_grad(self, X):
        """The box decompositions belong to the baseline samples drawn at construction: constants of the reverse pass."""
        states, Xf, lead, q, G, dev, eps, z, u, eps_joint, outs = self._joint_outputs(X)
        acq, values, bars = ops.cehvi_integrate_q_grad(outs, [st.noise for st in states], None, self.cell_lower.to(dev),
                                                       self.cell_upper.to(dev), q=q, alpha=self.alpha,
                                                       optimistic=self.optimistic, cells_per_sample=True)
        dX = None
        for k, st in enumerate(states):
            g = self._per_output_grad(st, k, Xf, q, G, dev, eps, z, u, bars[k], kind="joint",
                                      nei_baseline=self._baselines[k], eps_mc=eps_joint[..., k].contiguous())
            dX = g if dX is None else dX + g
        self.last_outputs = {"per_output": outs, "acq": acq, "values": values}
        return acq, dX, lead, q

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        if X.requires_grad and torch.is_grad_enabled():
            return self._forward_with_grad(X)
        states, Xf, lead, q, G, dev, eps, z, u, eps_joint, outs = self._joint_outputs(X)
        dev = Xf.device
        acq, values = ops.cehvi_integrate_q(outs, [st.noise for st in states], None, self.cell_lower.to(dev),
                                            self.cell_upper.to(dev), q=q, alpha=self.alpha, optimistic=self.optimistic,
                                            cells_per_sample=True, want_values=True)
        self.last_outputs = {"per_output": outs, "acq": acq, "values": values}
        return acq.reshape(lead).to(X.dtype)               # _prev_nehvi = 0: pending points are scored jointly instead
