"""`optimize_acqf_sgld` with the reference's signature and loop (conformalbo/optim/optimize.py:23-283), run on the device.

Per SGLD step the reference does: `callback(iterates.cpu())` (the ratio-estimator training step), `acq(X)`, autograd
`backward()` through the conformal path and the box penalty, an SGHMC update (optim/sgld.py) and a cosine learning-rate
step.  Here a step is: `RatioEstimator.optimize_callback` on the device iterates (two kernels, no `.cpu()`), ONE
value+gradient library call (`acq.value_and_grad`), one normal draw and ONE `k_sgld_update` launch (penalty gradient +
momentum + noise + parameter update); the learning rate follows the closed form of CosineAnnealingLR.  The value+gradient
call of each optimiser batch has a fixed shape and is replayed as a CUDA graph (`options["cuda_graph"]`, default on).

The initial conditions come from botorch's `gen_batch_initial_conditions` in the reference [3P, not installed here]; its
heuristic is restated from the published algorithm (Sobol raw samples, Boltzmann selection with eta = 2 on the
standardised acquisition values, the best raw sample always kept).  `sequential=True` runs the reference's greedy loop (q
single-point problems with the earlier points pending; the SGLD settings are passed on, where the reference falls back to
its defaults).  Constraints, fixed features and one-shot acquisition functions are not built (the reference's scripts never
pass them, scripts/black_box_opt.py:131-143).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Tuple, Union

import torch

from . import ops


def gen_batch_initial_conditions(acq_function, bounds: torch.Tensor, q: int, num_restarts: int, raw_samples: int,
                                 options: Optional[Dict] = None) -> torch.Tensor:
    """botorch.optim.initializers.gen_batch_initial_conditions + initialize_q_batch [3P], restated."""
    options = options or {}
    d = bounds.shape[-1]
    seed = options.get("seed")
    batch_limit = options.get("init_batch_limit", options.get("batch_limit"))
    eta = options.get("eta", 2.0)
    engine = torch.quasirandom.SobolEngine(dimension=q * d, scramble=True, seed=seed)
    X = engine.draw(raw_samples, dtype=torch.float64).view(raw_samples, q, d).to(bounds)
    X = bounds[0] + (bounds[1] - bounds[0]) * X
    with torch.no_grad():
        Y = torch.cat([acq_function(x) for x in X.split(batch_limit or raw_samples)])
    if num_restarts > raw_samples:
        raise RuntimeError(f"num_restarts ({num_restarts}) exceeds raw_samples ({raw_samples})")
    if num_restarts == raw_samples:
        return X
    Ystd = Y.std()
    if not bool(Ystd > 0):
        return X[torch.randperm(raw_samples, device=X.device)[:num_restarts]]
    max_idx = Y.argmax()
    etaZ = eta * (Y - Y.mean()) / Ystd
    weights = torch.exp(etaZ)
    while torch.isinf(weights).any():
        etaZ = etaZ * 0.5
        weights = torch.exp(etaZ)
    idcs = torch.multinomial(weights, num_restarts)
    if not bool((idcs == max_idx).any()):
        idcs[-1] = max_idx
    return X[idcs]


def cosine_lr(step: int, lr: float, T_max: int, eta_min: float) -> float:
    """Closed form of torch.optim.lr_scheduler.CosineAnnealingLR after `step` scheduler steps (cycles past T_max)."""
    return eta_min + (lr - eta_min) * (1.0 + math.cos(math.pi * step / T_max)) / 2.0


def _stepper(acq_function, ic_batch, use_graph):
    """X -> (acq, d acq / d X) for the fixed shape of one optimiser batch: a CUDA-graph replay where the acquisition
    function can be captured (`ConformalAcquisition.graphed`), else one value+gradient call per step."""
    if use_graph and hasattr(acq_function, "graphed"):
        try:
            graph = acq_function.graphed(ic_batch.shape[0], ic_batch.shape[1])
        except ValueError:
            graph = None
        if graph is not None:
            def step(X):
                out = graph(X, z=acq_function.fixed_probes, u_rand=acq_function.fixed_uniforms)
                return out["acq"], out["dacq_dX"]
            return step
    return lambda X: _value_and_grad(acq_function, X)


def _value_and_grad(acq_function, X):
    if hasattr(acq_function, "value_and_grad"):
        return acq_function.value_and_grad(X)
    Xg = X.detach().clone().requires_grad_(True)           # any other differentiable acquisition function
    vals = acq_function(Xg)
    (g,) = torch.autograd.grad(vals.sum(), Xg)
    return vals.detach(), g


def optimize_acqf_sgld(
    acq_function,
    bounds: torch.Tensor,
    q: int,
    num_restarts: int,
    raw_samples: Optional[int] = None,
    options: Optional[Dict[str, Union[bool, float, int, str, Callable]]] = None,
    inequality_constraints=None,
    equality_constraints=None,
    nonlinear_inequality_constraints=None,
    fixed_features=None,
    post_processing_func: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
    batch_initial_conditions: Optional[torch.Tensor] = None,
    return_best_only: bool = True,
    sequential: bool = False,
    warmup_steps: int = 32,
    sgld_steps: int = 256,
    lr: float = 1e-3,
    temperature: float = 5e-2,
    return_full_traj: bool = False,
    **kwargs,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Multi-start SGLD on the acquisition surface; returns (candidates, acquisition values) as the reference does:
    per restart the best in-bounds iterate of its trajectory, then the best restart if `return_best_only`."""
    if inequality_constraints or equality_constraints or nonlinear_inequality_constraints or fixed_features:
        raise NotImplementedError("constraints / fixed features are not built (unused by the reference's scripts)")
    if sequential and q > 1:
        # sequential greedy: q single-point problems, each with the points found so far as pending (optimize.py:113-149)
        if not return_best_only:
            raise NotImplementedError("`return_best_only=False` only supported for joint optimization.")
        candidate_list, acq_value_list = [], []
        base_X_pending = acq_function.X_pending
        for _ in range(q):
            candidate, acq_value = optimize_acqf_sgld(
                acq_function, bounds, 1, num_restarts, raw_samples=raw_samples, options=options,
                post_processing_func=post_processing_func, batch_initial_conditions=None, return_best_only=True,
                sequential=False, warmup_steps=warmup_steps, sgld_steps=sgld_steps, lr=lr, temperature=temperature)
            candidate_list.append(candidate)
            acq_value_list.append(acq_value)
            candidates = torch.cat(candidate_list, dim=-2)
            acq_function.set_X_pending(torch.cat([base_X_pending, candidates], dim=-2) if base_X_pending is not None
                                       else candidates)
        acq_function.set_X_pending(base_X_pending)             # reset to the previous X_pending state
        return candidates, torch.stack(acq_value_list, dim=-1)
    if not bounds.is_cuda:
        raise RuntimeError("optimize_acqf_sgld runs on the device: `bounds` must be a CUDA tensor")
    options = options or {}
    bounds = bounds.to(torch.float64)
    d = bounds.shape[-1]
    if batch_initial_conditions is None:
        if raw_samples is None:
            raise ValueError("Must specify `raw_samples` when `batch_initial_conditions` is `None`.")
        batch_initial_conditions = gen_batch_initial_conditions(acq_function, bounds, q, num_restarts, raw_samples, options)
    batch_limit = options.get("batch_limit", num_restarts)
    # the trajectory is scored in calls of up to `score_batch_limit` joint batches: the reference scores it `batch_limit` at a
    # time (optimize.py:243-246, a memory limit of its dense path) -- here that would be one ~0.1 ms latency-path call per step
    score_limit = max(int(options.get("score_batch_limit", 4096)), 1)
    # one contiguous (restarts, q, d) fp64 buffer per optimiser batch, updated in place
    batched_ics = [b.contiguous() for b in
                   batch_initial_conditions.detach().to(bounds.device, torch.float64).clone().split(batch_limit)]
    momentum = 0.1
    scaled_temp = temperature / math.sqrt(d)               # rescale temp so noise has consistent norm (optimize.py:197)
    bufs = [torch.zeros_like(b) for b in batched_ics]
    T_max, eta_min = sgld_steps // 4, lr / 16.0            # cyclic learning rate (optimize.py:202-207)
    lb, ub = bounds[0].contiguous(), bounds[1].contiguous()
    callback = options.get("callback")
    sgld_iterates = [[b.clone()] for b in batched_ics]
    timing = options.get("timing")                          # dict or None: wall seconds of the stages (synchronises; diagnostics only)
    if timing is not None:
        import time as _t
        torch.cuda.synchronize(); _t0 = _t.perf_counter()
    steppers = [_stepper(acq_function, b, options.get("cuda_graph", True)) for b in batched_ics]
    if timing is not None:
        torch.cuda.synchronize(); _t1 = _t.perf_counter(); timing["setup_s"] = _t1 - _t0

    for _s in range(sgld_steps):
        lr_s = cosine_lr(_s, lr, T_max, eta_min)
        for batch_idx, ic_batch in enumerate(batched_ics):
            if _s >= warmup_steps:                         # once the chain has warmed up, train the ratio estimator
                sgld_iterates[batch_idx].append(ic_batch.clone())
                if callable(callback):
                    callback(ic_batch.clone())             # stays on the device (the reference hands over `.cpu()`)
            if _s < sgld_steps - 1:
                _, dacq = steppers[batch_idx](ic_batch)
                noise = torch.randn_like(ic_batch)
                ops.sgld_update(ic_batch, bufs[batch_idx], dacq.to(torch.float64).contiguous(), noise, lb, ub, lr=lr_s,
                                momentum=momentum, temperature=scaled_temp)

    if timing is not None:
        torch.cuda.synchronize(); _t2 = _t.perf_counter(); timing["steps_s"] = _t2 - _t1
    batch_candidates = torch.cat([torch.stack(it, dim=1) for it in sgld_iterates])      # (restarts, T, q, d)
    if return_full_traj:
        with torch.no_grad():
            batch_acq_values = torch.cat([acq_function(c) for c in batch_candidates.flatten(0, -3).split(score_limit)
                                          ]).view(batch_candidates.shape[:-2])
    else:
        select_cands, select_vals = [], []
        with torch.no_grad():
            for cand_traj in batch_candidates:
                in_bounds = ((cand_traj >= lb).all(-1) & (cand_traj <= ub).all(-1)).all(-1)
                feas = cand_traj[in_bounds]
                if feas.shape[0] == 0:
                    raise RuntimeError("no SGLD iterate of a restart lies inside the bounds")
                vals = torch.cat([acq_function(c) for c in feas.split(score_limit)])
                best = vals.argmax(0)
                select_cands.append(feas[best])
                select_vals.append(vals[best])
        batch_candidates, batch_acq_values = torch.stack(select_cands), torch.stack(select_vals)
        if return_best_only:
            best = batch_acq_values.argmax(0)
            batch_candidates, batch_acq_values = batch_candidates[best], batch_acq_values[best]
    if timing is not None:
        torch.cuda.synchronize(); timing["scoring_s"] = _t.perf_counter() - _t2
    if post_processing_func is not None:
        batch_candidates = post_processing_func(batch_candidates)
    return batch_candidates, batch_acq_values


def optimize_acqf_sgld_list(acq_function_list, bounds: torch.Tensor, num_restarts: int, raw_samples: Optional[int] = None,
                            options: Optional[Dict] = None, inequality_constraints=None, equality_constraints=None,
                            fixed_features=None, post_processing_func=None, return_best_only: bool = True, **sgld_kwargs):
    """One candidate per acquisition function, optimised in sequence with the earlier candidates pending (sequential greedy;
    conformalbo/optim/optimize.py:286-357).  The ratio estimator's `optimize_callback` is installed as the SGLD callback when
    the acquisition function has one (optimize.py:337-338).  `sgld_kwargs` (sgld_steps, lr, temperature, warmup_steps) are
    passed on; the reference uses its defaults here."""
    if not acq_function_list:
        raise ValueError("acq_function_list must be non-empty.")
    candidate_list, acq_value_list = [], []
    candidates = torch.tensor([], device=bounds.device, dtype=bounds.dtype)
    base_X_pending = acq_function_list[0].X_pending
    for acq_function in acq_function_list:
        if candidate_list:
            acq_function.set_X_pending(torch.cat([base_X_pending, candidates], dim=-2) if base_X_pending is not None
                                       else candidates)
        opts = dict(options or {})
        if hasattr(getattr(acq_function, "ratio_estimator", None), "optimize_callback"):
            opts["callback"] = acq_function.ratio_estimator.optimize_callback
        candidate, acq_value = optimize_acqf_sgld(
            acq_function, bounds, 1, num_restarts, raw_samples=raw_samples, options=opts,
            inequality_constraints=inequality_constraints, equality_constraints=equality_constraints,
            fixed_features=fixed_features, post_processing_func=post_processing_func, return_best_only=True,
            sequential=False, **sgld_kwargs)
        candidate_list.append(candidate)
        acq_value_list.append(acq_value)
        candidates = torch.cat(candidate_list, dim=-2)
    return candidates, torch.stack(acq_value_list, dim=-1)
