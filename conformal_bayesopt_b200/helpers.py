"""Drop-in replacements for the lower boundary of the reference (conformalbo/helpers.py).

Same names, argument meaning and return layout as the reference; the arithmetic runs in libcbo
(no torch fallback).  `model` is a `GPModelState` (or a duck-typed BoTorch `SingleTaskGP`).
Differences that cannot be avoided: the 4th return value of
`construct_conformal_bands` (a GPyTorch model object conditioned on every grid point) is `None` --
the fused kernels never materialise it; the acquisition classes take over `_conformal_fwd` instead.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .model_state import GPModelState
from .sampling import IIDNormalSampler


def _model_fingerprint(model):
    """(data_ptr, version, shape) of every tensor the extracted state depends on: training data and hyper-parameters.
    An in-place refit, `load_state_dict` (reference utils.py:149) or `set_train_data` changes it."""
    tensors = [model.train_inputs[0], model.train_targets, model.covar_module.base_kernel.lengthscale,
               model.covar_module.outputscale, model.likelihood.noise, model.mean_module.constant]
    fp = []
    for t in tensors:
        t = torch.as_tensor(t)
        fp.append((t.data_ptr(), t._version, tuple(t.shape)))
    return tuple(fp)


def as_state(model) -> GPModelState:
    """The kernels' plain-tensor view of `model`.  A BoTorch model is extracted once and the state cached on it together
    with a fingerprint of its data / hyper-parameters; the cache is rebuilt when the fingerprint changes."""
    if isinstance(model, GPModelState):
        return model
    fp = _model_fingerprint(model)
    cached = getattr(model, "_cbo_state", None)
    if cached is not None and cached[0] == fp:
        return cached[1]
    state = GPModelState.from_botorch(model)
    try:
        model._cbo_state = (fp, state)
    except Exception:
        pass
    return state


def as_states(model):
    """[GPModelState per output]: a list / tuple of single-output models or states, a batched multi-output BoTorch model
    (train_inputs (t, N, d): t independent GPs), or a single-output model (a list of one)."""
    if isinstance(model, (list, tuple)):
        return [as_state(m) for m in model]
    if isinstance(model, GPModelState):
        return [model]
    X = model.train_inputs[0]
    if X.ndim == 3:
        fp = _model_fingerprint(model)
        cached = getattr(model, "_cbo_states", None)
        if cached is not None and cached[0] == fp:
            return cached[1]
        states = [GPModelState.from_botorch(model, output_index=k) for k in range(X.shape[0])]
        try:
            model._cbo_states = (fp, states)
        except Exception:
            pass
        return states
    return [as_state(model)]


class _PinnedRing:
    """A small ring of pinned staging buffers per (shape, dtype): an async upload that does not synchronise the host with the
    stream and never allocates in steady state (torch's caching host allocator would `cudaHostAlloc` a new block whenever the
    host runs ahead of the device).  Before a slot is reused its previous copy is waited for, which bounds the run-ahead."""
    _rings = {}

    @classmethod
    def upload(cls, src: torch.Tensor, device, slots: int = 8) -> torch.Tensor:
        key = (tuple(src.shape), src.dtype, torch.device(device))
        ring = cls._rings.get(key)
        if ring is None:
            ring = cls._rings[key] = dict(i=0, buf=[torch.empty(src.shape, dtype=src.dtype).pin_memory() for _ in range(slots)],
                                          ev=[None] * slots)
        i = ring["i"]
        ring["i"] = (i + 1) % slots
        if ring["ev"][i] is not None:
            ring["ev"][i].synchronize()
        ring["buf"][i].copy_(src)
        out = ring["buf"][i].to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        ring["ev"][i] = ev
        return out


def draw_probes(n_rows: int, num_samples: int = 8, device=None, on_device: bool = False) -> torch.Tensor:
    """Rademacher probes drawn from the CPU default generator, as est_train_post_var does (helpers.py:399-400): same draws
    and values as `2 * ((rand > 0.5).double() - 0.5)`.  With `device`, only the sign bits cross the bus (1 byte per probe
    entry) and the +-1 doubles are formed on the device.  `on_device=True` draws from the device generator instead (no host
    work per call; a different random stream than the reference's)."""
    if on_device and device is not None:
        one = torch.ones((), dtype=torch.float64, device=device)
        return torch.where(torch.rand(n_rows, num_samples, device=device) > 0.5, one, -one)
    bits = torch.rand(n_rows, num_samples) > 0.5
    if device is None:
        return torch.where(bits, 1.0, -1.0).to(torch.float64)
    one = torch.ones((), dtype=torch.float64, device=device)
    # pinned staging: the upload does not synchronise the host with the stream, so the next step's host work overlaps this
    # step's kernels
    return torch.where(_PinnedRing.upload(bits, device), one, -one)


def _ratio_weights(ratio_estimator, device):
    if ratio_estimator is None:
        return None
    if isinstance(ratio_estimator, ops.RatioWeights):
        return ratio_estimator
    return ratio_estimator.device_weights(device)


def conf_mask_to_bounds(target_grid: torch.Tensor, conf_pred_mask: torch.Tensor):
    """helpers.py:15-91.  (*, G, 1, t) x 2 -> conf_lb, conf_ub (*, 1, t)."""
    lead = target_grid.shape[:-3]
    G, q, t = target_grid.shape[-3:]
    lb, ub = ops.conf_mask_to_bounds(target_grid.reshape(-1, G, q, t), conf_pred_mask.reshape(-1, G, q, t))
    return lb.reshape(*lead, 1, t).to(target_grid.dtype), ub.reshape(*lead, 1, t).to(target_grid.dtype)


def _bands_multi(states, X, alpha, temp, grid_res, max_grid_refinements, eps, z, u_rand, ratio, randomized,
                 extra_outputs=()):
    """t > 1 (label_layout='intent'): one library call per output and refinement step; the loop of helpers.py:129-189 runs
    HERE because its decisions are taken over all outputs at once (min / max / mean of the (b, q, t) accept counts in
    float32, one kept pass) -- every step calls the kernels with max_grid_refinements = 0 and the covariance row scales
    accumulated so far (cbo_params.grid_cov_scale).  eps (G,q,t) or (b,G,q,t); u_rand (b,G,t,q) (layout of `ub_prob`)."""
    t = len(states)
    b, q, _ = X.shape
    G = grid_res
    dev = X.device
    scale = [None] * t
    best = None
    outs_needed = ("target_grid", "grid_logp", "conf_pred_mask", "ood_mask", "num_accepted") + tuple(extra_outputs)
    for _ in range(max_grid_refinements + 1):
        outs = []
        for k, st in enumerate(states):
            outs.append(ops.conformal_acq(
                st, X, eps_grid=eps[..., k].contiguous(), z=z, eps_mc=torch.zeros(1, G, q, dtype=torch.float64, device=dev),
                u_rand=None if u_rand is None else u_rand[:, :, k, :].contiguous(), ratio=ratio, outputs=outs_needed,
                kind="ucb", alpha=alpha, temp=temp, grid_res=G, max_grid_refinements=0, randomized=randomized,
                optimistic=False, grid_cov_scale=scale[k]))
        num_acc = torch.stack([o["num_accepted"].reshape(b, q) for o in outs], -1).float()       # (b,q,t) float32
        diff = num_acc / G - 0.5
        score = float(diff.pow(2).mean())                                                       # one host sync per step, like the reference
        if best is None or score < best[0]:
            best = (score, outs)
        if float(num_acc.min()) < 2 or float(num_acc.max()) > G - 2:
            cs = (1.0 + diff.sign() * diff.abs().pow(2.0 / q)).double()                          # (b,q,t)
            for k in range(t):
                cur = cs[..., k].contiguous()
                scale[k] = cur if scale[k] is None else (scale[k] * cur).contiguous()
        else:
            break
    return best[1]


def construct_conformal_bands(model, inputs, alpha, temp, grid_res, max_grid_refinements, sampler,
                              ratio_estimator=None, mask_ood=True, randomized=False, *, probes=None, u_rand=None,
                              label_layout="intent"):
    """helpers.py:94-202.  inputs (b, q, d) -> target_grid (b,G,q,t), weights (b,G,1,1), conf_pred_mask (b,G,q,t),
    None, ood_mask (b,G,q,t).  `probes` / `u_rand` inject the reference's internal randomness (App. C-2/3).
    Multi-output models (a list of single-output models / states, or a batched BoTorch model): t independent GPs sharing the
    probes, joint refinement decisions, joint grid log-density.  `label_layout`: "intent" -- every
    output is scored at its own labels; "reference" -- the `.view` scramble of helpers.py:257-258, restated in
    oracle/multi.py and pinned there against the reference-source goldens, is NOT built on the device (it replaces a train
    label by another output's label or grid value, which breaks the per-output closed form the kernels rest on)."""
    shape_msg = "inputs should have shape (*q_batch_shape, q_batch_size, input_dim)" \
                f"instead has shape {inputs.shape}"
    assert inputs.ndim >= 3, shape_msg
    assert grid_res >= 2, "grid resolution must be at least 2"
    if label_layout not in ("intent", "reference"):
        raise ValueError("label_layout must be 'intent' or 'reference'")
    states = as_states(model)
    if len(states) > 1:
        if label_layout == "reference":
            raise NotImplementedError("label_layout='reference' (the multi-output .view scramble, SURVEY.md App. C-6) exists "
                                      "only in oracle/multi.py; the kernels implement label_layout='intent'")
        dev = inputs.device
        lead = inputs.shape[:-2]
        q, d = inputs.shape[-2:]
        X = inputs.reshape(-1, q, d)
        b, t = X.shape[0], len(states)
        sampler._sample_shape = torch.Size([grid_res])
        if sampler.collapse_batch_dims:
            eps = sampler.draw((grid_res, q, t), dev)
        else:
            eps = sampler.draw((grid_res, b, q, t), dev).movedim(0, 1).contiguous()
        z = (probes if probes is not None else draw_probes(states[0].N + q)).to(dev)
        if randomized and u_rand is None:
            u_rand = torch.rand(b, grid_res, t, q, dtype=torch.float64, device=dev)
        outs = _bands_multi(states, X, alpha, temp, grid_res, max_grid_refinements, eps, z,
                            None if u_rand is None else u_rand.reshape(b, grid_res, t, q), _ratio_weights(ratio_estimator, dev),
                            randomized)
        dt = inputs.dtype
        tg = torch.stack([o["target_grid"].reshape(b, grid_res, q) for o in outs], -1).reshape(*lead, grid_res, q, t).to(dt)
        w = sum(o["grid_logp"] for o in outs).reshape(*lead, grid_res, 1, 1).to(dt)
        mk = torch.stack([o["conf_pred_mask"].reshape(b, grid_res, q) for o in outs], -1).reshape(*lead, grid_res, q, t).to(dt)
        ood = torch.stack([o["ood_mask"].reshape(b, 1, q).expand(b, grid_res, q) for o in outs], -1)
        return tg, w, mk, None, ood.reshape(*lead, grid_res, q, t).to(dt)
    state = states[0]
    dev = inputs.device
    lead = inputs.shape[:-2]
    q, d = inputs.shape[-2:]
    X = inputs.reshape(-1, q, d)
    b = X.shape[0]
    sampler._sample_shape = torch.Size([grid_res])                      # helpers.py:132
    if sampler.collapse_batch_dims:
        eps = sampler.draw((grid_res, q), dev)
    else:
        eps = sampler.draw((grid_res, b, q), dev).movedim(0, 1).contiguous()
    z = (probes if probes is not None else draw_probes(state.N + q)).to(dev)
    if randomized and u_rand is None:
        u_rand = torch.rand(b, grid_res, q, dtype=torch.float64, device=dev)
    out = ops.conformal_acq(
        state, X, eps_grid=eps, z=z, eps_mc=torch.zeros(1, grid_res, q, dtype=torch.float64, device=dev), u_rand=u_rand,
        ratio=_ratio_weights(ratio_estimator, dev), outputs=("target_grid", "grid_logp", "conf_pred_mask", "ood_mask"),
        kind="ucb", alpha=alpha, temp=temp, grid_res=grid_res, max_grid_refinements=max_grid_refinements,
        randomized=randomized, optimistic=False)
    dt = inputs.dtype
    tg = out["target_grid"].reshape(*lead, grid_res, q, 1).to(dt)
    w = out["grid_logp"].reshape(*lead, grid_res, 1, 1).to(dt)
    mk = out["conf_pred_mask"].reshape(*lead, grid_res, q, 1).to(dt)
    ood = out["ood_mask"].reshape(*lead, 1, q, 1).expand(*lead, grid_res, q, 1).to(dt)
    return tg, w, mk, None, ood


def assess_coverage(model, inputs, targets, alpha, temp, grid_res, max_grid_refinements, ratio_estimator=None,
                    randomized=False, *, probes=None, u_rand=None, grid_sampler: Optional[IIDNormalSampler] = None,
                    label_layout="intent"):
    """helpers.py:336-387.  inputs (n, d), targets (n, t) -> (std_coverage, conformal_coverage) floats.  t > 1: a
    multi-output model (see construct_conformal_bands); u_rand is then (n, G, t, 1)."""
    from scipy import stats
    states = as_states(model)
    dev = inputs.device
    n = inputs.shape[0]
    if len(states) > 1:
        t = len(states)
        if grid_sampler is None:
            grid_sampler = IIDNormalSampler(grid_res, resample=False, collapse_batch_dims=False)   # helpers.py:364
        tg, _, mk, _, _ = construct_conformal_bands(states, inputs[:, None], alpha, temp, grid_res, max_grid_refinements,
                                                    grid_sampler, ratio_estimator, mask_ood=False, randomized=randomized,
                                                    probes=probes, u_rand=u_rand, label_layout=label_layout)
        X1 = inputs[:, None].to(torch.float64).contiguous()
        stats_k = [ops.conformal_acq(st, X1, eps_grid=torch.zeros(2, dtype=torch.float64, device=dev),
                                     z=(probes if probes is not None else draw_probes(st.N + 1)).to(dev),
                                     eps_mc=torch.zeros(1, 2, dtype=torch.float64, device=dev), outputs=("post_mean", "post_var"),
                                     kind="ucb", alpha=alpha, temp=temp, grid_res=2) for st in states]
        y_mean = torch.stack([o["post_mean"] for o in stats_k], -1)
        y_std = torch.stack([o["post_var"] for o in stats_k], -1).sqrt()
        tgt = targets.reshape(n, t).to(torch.float64)
        std_scale = stats.norm.ppf(1 - alpha / 2.0)
        std_coverage = ((tgt >= y_mean - std_scale * y_std) * (tgt <= y_mean + std_scale * y_std)).float().mean()
        lb, ub = conf_mask_to_bounds(tg.to(torch.float64), mk.to(torch.float64))
        conformal_coverage = ((tgt >= lb.squeeze(-2)) * (tgt <= ub.squeeze(-2))).float().nanmean()
        return std_coverage.item(), conformal_coverage.item()
    state = states[0]
    if grid_sampler is None:
        grid_sampler = IIDNormalSampler(grid_res, resample=False, collapse_batch_dims=False)   # helpers.py:364
    grid_sampler._sample_shape = torch.Size([grid_res])
    eps = grid_sampler.draw((grid_res, n, 1), dev).movedim(0, 1)[:, :, 0].contiguous()
    z = (probes if probes is not None else draw_probes(state.N + 1)).to(dev)
    if randomized and u_rand is None:
        u_rand = torch.rand(n, grid_res, dtype=torch.float64, device=dev)
    out = ops.conformal_acq(
        state, inputs[:, None], eps_grid=eps, z=z, eps_mc=torch.zeros(1, grid_res, dtype=torch.float64, device=dev),
        u_rand=u_rand, ratio=_ratio_weights(ratio_estimator, dev),
        outputs=("target_grid", "conf_pred_mask", "post_mean", "post_var"), kind="ucb", alpha=alpha, temp=temp,
        grid_res=grid_res, max_grid_refinements=max_grid_refinements, randomized=randomized, optimistic=False)
    tg = targets.reshape(n, -1).to(torch.float64)
    y_mean, y_std = out["post_mean"].unsqueeze(-1), out["post_var"].sqrt().unsqueeze(-1)
    std_scale = stats.norm.ppf(1 - alpha / 2.0)
    std_coverage = ((tg >= y_mean - std_scale * y_std) * (tg <= y_mean + std_scale * y_std)).float().mean()
    lb, ub = ops.conf_mask_to_bounds(out["target_grid"][:, :, None, None], out["conf_pred_mask"][:, :, None, None])
    conformal_coverage = ((tg >= lb.squeeze(-2)) * (tg <= ub.squeeze(-2))).float().nanmean()
    return std_coverage.item(), conformal_coverage.item()
