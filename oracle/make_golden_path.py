"""Generate tests/golden/ref_path*.pt by EXECUTING the reference's own hot-path source, unmodified.

Runs only in the build container (needs /root/reference).  `conformalbo/helpers.py`, `conformalbo/acquisitions.py`,
`conformalbo/mobo_acquisitions.py` and `conformalbo/ratio_estimation.py` are imported from /root/reference as they are;
their `gpytorch` / `botorch` imports resolve to the stand-ins under `oracle/_shim/` (exact-GP algebra + sampler shape
rules only, see oracle/_shim/README.md), `conformalbo.utils` (hydra, wandb, ...) is stubbed as in
oracle/make_golden_training.py.  Every line of the conformal logic that runs is therefore the reference's:
  construct_conformal_bands / refine_grid_dist / conformal_gp_regression / est_train_post_var / assess_coverage /
  conf_mask_to_bounds (helpers.py), ConformalAcquisition._conformal_fwd / forward, the cEI / cNEI / cUCB
  _nonconformal_fwd and _conformal_integration (acquisitions.py), cEHVI's _nonconformal_fwd (mobo_acquisitions.py),
  RatioEstimator.forward (ratio_estimation.py).
Randomness is injected, never re-seeded: sampler base samples through the stand-in's DRAW_HOOK, the Hutchinson probes by
patching `torch.rand` for the one (n, 8) draw of helpers.py:399, the randomised step by patching `Bernoulli.sample` to
`u < probs` (helpers.py:311).

Usage:  python oracle/make_golden_path.py [--only NAME] [--ehvi] [--grad] [--more]   (writes tests/golden/ref_path.pt,
        ref_path_n1024.pt, ref_path_ehvi.pt, ref_path_grad.pt, ref_path_more.pt)
"""
import argparse
import contextlib
import importlib
import math
import os
import sys
import time
import types
from collections import namedtuple

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def load_reference():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_shim"))
    sys.path.insert(0, REF)
    utils = types.ModuleType("conformalbo.utils")
    utils.DataSplit = namedtuple("DataSplit", ["inputs", "targets"], defaults=[np.array([]), np.array([])])
    utils.update_splits = lambda train_split, val_split, test_split, new_split, holdout_ratio: (train_split, val_split, test_split)
    utils.safe_np_cat = lambda arrays, **kw: np.concatenate([a for a in arrays if len(a)], **kw)
    import conformalbo                                          # the reference's (empty) package __init__
    assert conformalbo.__file__.startswith(REF)
    sys.modules["conformalbo.utils"] = utils
    mods = {n: importlib.import_module(f"conformalbo.{n}") for n in
            ("helpers", "acquisitions", "mobo_acquisitions", "ratio_estimation")}
    for m in mods.values():
        assert m.__file__.startswith(REF), m.__file__
    return types.SimpleNamespace(**mods)


@contextlib.contextmanager
def injected(z=None, u_rand=None, draws=None):
    """z (n, Sp) of +-1 -> the torch.rand draw of est_train_post_var; u_rand -> Bernoulli.sample; draws: dict
    shape-tuple -> tensor for sampler base samples (each must be consumed with exactly that shape)."""
    from botorch.sampling import samplers
    real_rand, real_sample = torch.rand, torch.distributions.Bernoulli.sample
    used = dict(z=0, u=0, draws=[])

    def rand(*size, **kw):
        if z is not None and tuple(size) == tuple(z.shape):
            used["z"] += 1
            return torch.where(z > 0, 0.75, 0.25).to(torch.get_default_dtype())
        return real_rand(*size, **kw)

    def sample(self, sample_shape=torch.Size()):
        used["u"] += 1
        u = u_rand
        while u.ndim < self.probs.ndim:
            u = u.unsqueeze(2)                                   # (b, G, q) -> (b, G, t=1, q)
        assert u.shape == self.probs.shape, (u.shape, self.probs.shape)
        return (u < self.probs).to(self.probs)

    def hook(shape, dtype):
        key = tuple(shape)
        assert draws is not None and key in draws, f"unexpected base-sample draw of shape {key}"
        used["draws"].append(key)
        return draws[key].to(dtype)

    torch.rand = rand
    if u_rand is not None:
        torch.distributions.Bernoulli.sample = sample
    samplers.DRAW_HOOK = hook
    try:
        yield used
    finally:
        torch.rand = real_rand
        torch.distributions.Bernoulli.sample = real_sample
        samplers.DRAW_HOOK = None


def make_model(p, t=1):
    from botorch.models import SingleTaskGP
    if t == 1:
        return SingleTaskGP(p["X"], p["y"].reshape(-1, 1), mean_const=p["mean"], lengthscale=p["lengthscale"],
                            outputscale=p["outputscale"], noise=p["noise"])
    return SingleTaskGP(p["X"], p["Y"], mean_const=p["mean"], lengthscale=p["lengthscale"],
                        outputscale=p["outputscale"], noise=p["noise"])


def make_ratio(ref, d, seed=1):
    """The reference's RatioEstimator with the seeded stand-in weights copied into its target network."""
    from conformal_bayesopt_b200.synthetic import make_ratio_weights
    est = ref.ratio_estimation.RatioEstimator(in_size=d, device=None, dtype=torch.float64)
    W = make_ratio_weights(d, seed)
    lin = [m for m in est._target_network if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for k, layer in enumerate(lin):
            layer.weight.copy_(W[2 * k])
            layer.bias.copy_(W[2 * k + 1])
    return est, W


def run_acq_case(ref, name, *, N, d, b, q, G, kind, temp, alpha, refine=0, randomized=False, ratio=False, S=1, seed=0,
                 noise=1e-2, optimistic=None, beta=0.2, best_f=0.3, nb=0, chunk=None, bounds=True, Sp=8, x_pending=0,
                 want_grad=False):
    """One acquisition-forward case through the reference classes.  chunk: candidates per forward call (the reference's
    callers chunk by batch_limit; only valid with refine == 0, where candidates are independent)."""
    from botorch.sampling import IIDNormalSampler
    from conformal_bayesopt_b200.synthetic import make_inputs, make_problem
    t0 = time.time()
    p = make_problem(N, d, seed=seed, noise=noise, ard_spread=0.1)
    qq = q + x_pending
    inp = make_inputs(N, d, b, qq, G, S=S, Sp=Sp, seed=seed)
    g = torch.Generator().manual_seed(77 + seed)
    Xb = torch.rand(nb, d, generator=g, dtype=torch.float64) if nb else None
    if nb:
        inp["eps_mc"] = torch.randn(S, G, nb + qq, generator=g, dtype=torch.float64)
    X_pending = inp["Xq"][0, q:].clone() if x_pending else None
    Xq = inp["Xq"][:, :q].contiguous()
    model = make_model(p)
    est, W = make_ratio(ref, d) if ratio else (None, None)
    A = ref.acquisitions
    common = dict(alpha=alpha, temp=temp, grid_res=G, max_grid_refinements=refine, ratio_estimator=est,
                  randomized=randomized)
    if optimistic is not None:
        common["optimistic"] = optimistic

    def build():
        grid_sampler = IIDNormalSampler(G)
        sampler = IIDNormalSampler(S, batch_range=(0, -3))
        kw = dict(common, grid_sampler=grid_sampler, model=model, sampler=sampler, X_pending=X_pending)
        if kind == "ucb":
            return A.qConformalUpperConfidenceBound(beta=beta, **kw)
        if kind == "ei":
            return A.qConformalExpectedImprovement(best_f=best_f, **kw)
        if kind == "nei":
            return A.qConformalNoisyExpectedImprovement(X_baseline=Xb, prune_baseline=False, **kw)
        raise ValueError(kind)

    outs = dict(acq=[], target_grid=[], grid_logp=[], conf_pred_mask=[], ood_mask=[], lb=[], ub=[], dacq_dX=[])
    step = chunk or b
    assert step == b or refine == 0
    for lo in range(0, b, step):
        Xc = Xq[lo:lo + step]
        bc = Xc.shape[0]
        grid_draw = inp["eps_grid"].reshape(G, 1, qq, 1)                  # shared by all candidates (collapse_batch_dims)
        mc = inp["eps_mc"]
        draws = {tuple(grid_draw.shape): grid_draw, (S, 1, G, mc.shape[-1], 1): mc.reshape(S, 1, G, mc.shape[-1], 1)}
        u = inp["u_rand"][lo:lo + step] if randomized else None
        acq_fn = build()
        with injected(z=inp["z"], u_rand=u, draws=draws) as used:
            Xfull = Xc if X_pending is None else torch.cat([Xc, X_pending.expand(bc, -1, -1)], dim=-2)
            grid, logp, mask, _, ood = acq_fn._conformalize_model(Xfull)
        n_pass = used["z"]
        acq_fn = build()                                                  # fresh samplers: same injected draws again
        with injected(z=inp["z"], u_rand=u, draws=draws) as used:
            if want_grad:     # d acq / d X by autograd through the reference's own code (optimize.py:224-235 does exactly this)
                Xg = Xc.clone().requires_grad_(True)
                val = acq_fn(Xg)
                (gX,) = torch.autograd.grad(val.sum(), Xg)
                outs["dacq_dX"].append(gX.detach())
            else:
                val = acq_fn(Xc)
        assert used["z"] == n_pass and n_pass >= 1
        outs["acq"].append(val.detach())
        outs["target_grid"].append(grid)
        outs["grid_logp"].append(logp.reshape(bc, G))
        outs["conf_pred_mask"].append(mask)
        outs["ood_mask"].append(ood.to(torch.float64))
        if bounds and qq == 1:
            lb, ub = ref.helpers.conf_mask_to_bounds(grid, mask)
            outs["lb"].append(lb.reshape(bc))
            outs["ub"].append(ub.reshape(bc))
    res = {k: torch.cat(v) for k, v in outs.items() if v}
    res["argmax"] = int(res["acq"].argmax())
    case = dict(name=name, kind=kind, N=N, d=d, b=b, q=q, G=G, S=S, temp=temp, alpha=alpha, refine=refine,
                randomized=randomized, ratio=ratio, noise=noise, beta=beta, best_f=best_f, seed=seed, nb=nb,
                optimistic=(kind == "ucb") if optimistic is None else optimistic, x_pending=x_pending,
                n_passes=n_pass,
                problem={k: (v.clone() if torch.is_tensor(v) else v) for k, v in p.items()},
                Xq=Xq, X_pending=X_pending, eps_grid=inp["eps_grid"], z=inp["z"], eps_mc=inp["eps_mc"],
                u_rand=inp["u_rand"] if randomized else None, X_baseline=Xb,
                ratio_weights=[w.clone() for w in W] if W is not None else None, out=res)
    print(f"  {name}: acq[{b}] in {time.time() - t0:.1f}s, passes={n_pass}, argmax={res['argmax']}, "
          f"acq range [{float(res['acq'].min()):.4g}, {float(res['acq'].max()):.4g}]", flush=True)
    return case


def run_coverage_case(ref, name, *, N, d, n, G, temp, alpha, refine, randomized, ratio, seed, t=1, noise=1e-2):
    """assess_coverage (helpers.py:336-387) on a labelled hold-out set: per-point base samples
    (collapse_batch_dims=False), bands from the reference's conf_mask_to_bounds."""
    from conformal_bayesopt_b200.synthetic import make_problem
    t0 = time.time()
    p = make_problem(N + n, d, seed=seed, noise=noise, ard_spread=0.1)
    g = torch.Generator().manual_seed(500 + seed)
    if t == 1:
        Y = p["y"].reshape(-1, 1)
        ls, os_, nz, mc = p["lengthscale"], p["outputscale"], p["noise"], p["mean"]
    else:
        extra = [torch.sin((2.0 + k) * p["X"].sum(-1)) + 0.1 * torch.randn(N + n, generator=g, dtype=torch.float64)
                 for k in range(1, t)]
        Y = torch.stack([p["y"]] + [(e - e.mean()) / e.std() for e in extra], dim=-1)
        ls = torch.stack([p["lengthscale"] * (1.0 + 0.3 * k) for k in range(t)])
        os_ = torch.tensor([1.0 + 0.5 * k for k in range(t)], dtype=torch.float64)
        nz = torch.tensor([noise * (1.0 + k) for k in range(t)], dtype=torch.float64)
        mc = torch.tensor([0.1 * k for k in range(t)], dtype=torch.float64)
    Xtr, Ytr, Xho, Yho = p["X"][:N], Y[:N], p["X"][N:], Y[N:]
    prob = dict(X=Xtr, Y=Ytr, y=Ytr[:, 0], mean=mc, lengthscale=ls, outputscale=os_, noise=nz)
    model = make_model(prob, t=t)
    est, W = make_ratio(ref, d) if ratio else (None, None)
    eps = torch.randn(G, n, 1, t, generator=g, dtype=torch.float64)
    z = 2.0 * ((torch.rand(N + 1, 8, generator=g) > 0.5).double() - 0.5)
    u = torch.rand(n, G, t, 1, generator=g, dtype=torch.float64)          # (b, G, t, q) = layout of ub_prob
    H = ref.helpers
    captured = {}
    real_c2b = H.conf_mask_to_bounds

    def spy(grid, mask):
        captured["grid"], captured["mask"] = grid.clone(), mask.clone()
        lb, ub = real_c2b(grid, mask)
        captured["lb"], captured["ub"] = lb.clone(), ub.clone()
        return lb, ub

    H.conf_mask_to_bounds = spy
    try:
        with injected(z=z, u_rand=u if randomized else None, draws={(G, n, 1, t): eps}) as used:
            std_cov, conf_cov = H.assess_coverage(model, Xho, Yho, alpha, temp, G, refine, est, randomized)
    finally:
        H.conf_mask_to_bounds = real_c2b
    assert "lb" in captured, "assess_coverage swallowed an exception"
    print(f"  {name}: cred {std_cov:.4f} conf {conf_cov:.4f} in {time.time() - t0:.1f}s passes={used['z']}", flush=True)
    return dict(name=name, N=N, d=d, n=n, G=G, t=t, temp=temp, alpha=alpha, refine=refine, randomized=randomized,
                ratio=ratio, problem=prob, X_holdout=Xho, Y_holdout=Yho, eps_grid=eps, z=z,
                u_rand=u if randomized else None, ratio_weights=[w.clone() for w in W] if W is not None else None,
                out=dict(std_coverage=std_cov, conformal_coverage=conf_cov, target_grid=captured["grid"],
                         conf_pred_mask=captured["mask"], lb=captured["lb"], ub=captured["ub"]))


def run_ehvi_case(ref, name, *, N, d, b, q, G, temp, alpha, S=2, randomized=False, ratio=False, optimistic=False, seed=0,
                  noise=1e-2):
    """cEHVI (mobo_acquisitions.py:24-45) on a two-output model: the reference's `_conformal_fwd` runs up to the call of
    `_conformal_integration`, whose ARGUMENTS are recorded -- `values` = the reference's `_nonconformal_fwd` (`_compute_qehvi`
    under every refit model: the cEHVI inner expectation, (b, G)), the masks, the grid log-density, the optimistic and ood
    masks.  The integration itself cannot be pinned: for these shapes the reference's own function raises unless b in
    {1, G} and then returns a (G,) tensor."""
    from botorch.sampling import IIDNormalSampler
    from conformal_bayesopt_b200.synthetic import make_problem
    from oracle.mobo import box_decomposition_2d
    t0 = time.time()
    t = 2
    p = make_problem(N, d, seed=seed, noise=noise, ard_spread=0.1)
    g = torch.Generator().manual_seed(900 + seed)
    extra = torch.sin(3.0 * p["X"].sum(-1)) + 0.1 * torch.randn(N, generator=g, dtype=torch.float64)
    Y = torch.stack([p["y"], (extra - extra.mean()) / extra.std()], dim=-1)
    ls = torch.stack([p["lengthscale"] * (1.0 + 0.3 * k) for k in range(t)])
    os_ = torch.tensor([1.0 + 0.5 * k for k in range(t)], dtype=torch.float64)
    nz = torch.tensor([noise * (1.0 + k) for k in range(t)], dtype=torch.float64)
    mc_ = torch.tensor([0.1 * k for k in range(t)], dtype=torch.float64)
    prob = dict(X=p["X"], Y=Y, y=Y[:, 0], mean=mc_, lengthscale=ls, outputscale=os_, noise=nz)
    model = make_model(prob, t=t)
    est, W = make_ratio(ref, d) if ratio else (None, None)
    Xq = torch.rand(b, q, d, generator=g, dtype=torch.float64)
    eps_grid = torch.randn(G, 1, q, t, generator=g, dtype=torch.float64)
    eps_mc = torch.randn(S, 1, G, q, t, generator=g, dtype=torch.float64)
    z = 2.0 * ((torch.rand(N + q, 8, generator=g) > 0.5).double() - 0.5)
    u = torch.rand(b, G, t, q, generator=g, dtype=torch.float64)          # layout of ub_prob for a batched model
    ref_point = Y.min(0).values - 0.1
    lower, upper = box_decomposition_2d(ref_point, Y)
    part = types.SimpleNamespace(get_hypercell_bounds=lambda: torch.stack([lower, upper]))
    acq = ref.mobo_acquisitions.qConformalExpectedHypervolumeImprovement(
        alpha, temp, G, 0, est, optimistic, IIDNormalSampler(G), randomized, model=model, ref_point=ref_point.tolist(),
        partitioning=part, sampler=IIDNormalSampler(S, batch_range=(0, -3)))
    captured = {}
    real_int = ref.acquisitions._conformal_integration

    def spy(values, conf_pred_mask, grid_logp, alpha_, opt_mask, ood_mask):
        captured.update(values=values.detach().clone(), conf_pred_mask=conf_pred_mask.detach().clone(),
                        grid_logp=grid_logp.detach().clone(), opt_mask=opt_mask.detach().clone(),
                        ood_mask=ood_mask.detach().clone().to(torch.float64))
        return values.new_zeros(1, values.shape[0], 1)            # the literal integration raises for these shapes

    ref.acquisitions._conformal_integration = spy
    try:
        draws = {tuple(eps_grid.shape): eps_grid, tuple(eps_mc.shape): eps_mc}
        with injected(z=z, u_rand=u if randomized else None, draws=draws) as used:
            acq._conformal_fwd(Xq)
    finally:
        ref.acquisitions._conformal_integration = real_int
    assert "values" in captured and captured["values"].shape == (b, G), captured.get("values", torch.zeros(0)).shape
    print(f"  {name}: cEHVI values {tuple(captured['values'].shape)} in {time.time() - t0:.1f}s, draws={used['draws']}, "
          f"range [{float(captured['values'].min()):.4g}, {float(captured['values'].max()):.4g}]", flush=True)
    return dict(name=name, N=N, d=d, b=b, q=q, G=G, S=S, t=t, temp=temp, alpha=alpha, randomized=randomized, ratio=ratio,
                optimistic=optimistic, problem=prob, Xq=Xq, eps_grid=eps_grid, eps_mc=eps_mc, z=z,
                u_rand=u if randomized else None, ref_point=ref_point, cell_lower=lower, cell_upper=upper,
                ratio_weights=[w.clone() for w in W] if W is not None else None, out=captured)


EHVI_CASES = dict(
    ehvi_q1=dict(N=40, d=2, b=12, q=1, G=32, temp=3e-2, alpha=0.1, S=2, seed=51),
    ehvi_q2_rand=dict(N=36, d=3, b=7, q=2, G=16, temp=3e-2, alpha=0.15, S=3, randomized=True, ratio=True, optimistic=True, seed=52),
)


BP = math.sqrt(0.2 * math.pi / 2)

ACQ_CASES = dict(
    # BASELINE configs[0]: branin-like 2-D, 8 initial points, 128 raw samples, the hydra defaults (temp 3e-2, randomized)
    cfg1_ucb_q1=dict(N=8, d=2, b=128, q=1, G=64, kind="ucb", temp=3e-2, alpha=8 ** -0.35, randomized=True, ratio=True, seed=11),
    cfg1_ucb_q2=dict(N=8, d=2, b=128, q=2, G=64, kind="ucb", temp=3e-2, alpha=8 ** -0.35, randomized=True, ratio=True, seed=12),
    cfg1_ei_q2_refine=dict(N=8, d=2, b=64, q=2, G=64, kind="ei", temp=3e-2, alpha=8 ** -0.35, randomized=True, refine=2, seed=13),
    # BASELINE configs[3]: tabular pool b=224, d=12, N=32, tab_bandits.yaml: temp 1e-6, one refinement, randomized
    cfg4_pool=dict(N=32, d=12, b=224, q=1, G=64, kind="ucb", temp=1e-6, alpha=32 ** -0.35, randomized=True, ratio=True, refine=1, seed=14),
    # shape / flag coverage at small sizes
    small_soft=dict(N=40, d=3, b=24, q=1, G=16, kind="ucb", temp=3e-2, alpha=0.1, seed=21),
    small_hard=dict(N=40, d=3, b=24, q=1, G=32, kind="ucb", temp=0.0, alpha=0.1, randomized=True, ratio=True, S=2, seed=22),
    small_ei_refine3=dict(N=64, d=2, b=20, q=1, G=64, kind="ei", temp=1e-2, alpha=0.1, randomized=True, refine=3, seed=23),
    small_q3=dict(N=50, d=4, b=12, q=3, G=32, kind="ucb", temp=3e-2, alpha=0.15, randomized=True, ratio=True, refine=1, seed=24),
    small_nei=dict(N=48, d=3, b=16, q=1, G=32, kind="nei", temp=3e-2, alpha=0.1, randomized=True, nb=6, S=2, seed=25),
    small_nei_q2=dict(N=48, d=3, b=10, q=2, G=32, kind="nei", temp=3e-2, alpha=0.1, ratio=True, nb=5, seed=26),
    small_pending=dict(N=30, d=2, b=9, q=1, x_pending=1, G=32, kind="ucb", temp=3e-2, alpha=0.1, randomized=True, seed=27),
    small_nonopt_lownoise=dict(N=120, d=5, b=16, q=1, G=64, kind="ei", temp=3e-2, alpha=0.05, optimistic=False, noise=1e-4, ratio=True, seed=28),
)

# BASELINE configs[1] / [4] sizes: N=1024, d=8, G=64, alpha = 1/sqrt(N); a 48-candidate subset scored in chunks of 2
# (dense (2, 64, 1025, 1025) temporaries; the reference's own callers chunk by batch_limit)
BIG_CASES = dict(
    cfg2_ucb=dict(N=1024, d=8, b=48, q=1, G=64, kind="ucb", temp=3e-2, alpha=1 / 32, chunk=2, seed=31),
    cfg2_ei_ratio_rand=dict(N=1024, d=8, b=48, q=1, G=64, kind="ei", temp=3e-2, alpha=1 / 32, randomized=True, ratio=True, chunk=2, seed=32),
    cfg2_ucb_q2=dict(N=1024, d=8, b=16, q=2, G=64, kind="ucb", temp=3e-2, alpha=1 / 32, randomized=True, ratio=True, chunk=2, seed=33),
)

# more shapes / flag combinations through the same reference classes (tests/golden/ref_path_more.pt)
MORE_CASES = dict(
    more_ei_q2_S2=dict(N=36, d=3, b=14, q=2, G=32, kind="ei", temp=3e-2, alpha=0.12, randomized=True, ratio=True, S=2, seed=71),
    more_ucb_q4=dict(N=28, d=2, b=6, q=4, G=16, kind="ucb", temp=3e-2, alpha=0.2, randomized=True, seed=72),
    more_nei_pending=dict(N=40, d=3, b=8, q=1, x_pending=1, G=32, kind="nei", temp=2e-1, alpha=0.1, ratio=True, nb=7, S=2, seed=73),
    more_hard_refine=dict(N=48, d=4, b=18, q=1, G=64, kind="ucb", temp=0.0, alpha=0.15, randomized=True, ratio=True, refine=2, seed=74),
    more_ei_q3_nonopt=dict(N=32, d=5, b=7, q=3, G=24, kind="ei", temp=1e-2, alpha=0.1, optimistic=False, refine=1, seed=75),
    more_g128=dict(N=64, d=6, b=10, q=1, G=128, kind="ucb", temp=3e-2, alpha=0.08, randomized=True, ratio=True, seed=76),
    more_ucb_optoff_S3=dict(N=44, d=2, b=12, q=2, G=32, kind="ucb", temp=5e-2, alpha=0.1, optimistic=False, S=3, seed=77),
    more_tiny_temp_q2=dict(N=30, d=3, b=9, q=2, G=32, kind="ucb", temp=1e-6, alpha=0.2, randomized=True, ratio=True, refine=1, seed=78),
)

# gradients: the same reference classes under torch.autograd (small shapes; refine = 0: the SGLD / L-BFGS callers' setting)
GRAD_CASES = dict(
    grad_ucb_q1=dict(N=40, d=3, b=10, q=1, G=32, kind="ucb", temp=3e-2, alpha=0.1, randomized=True, ratio=True, seed=61, want_grad=True),
    grad_ucb_q2=dict(N=30, d=2, b=8, q=2, G=32, kind="ucb", temp=3e-2, alpha=0.15, randomized=True, ratio=True, seed=62, want_grad=True),
    grad_ei_q1_S2=dict(N=36, d=4, b=9, q=1, G=16, kind="ei", temp=1e-2, alpha=0.1, S=2, seed=63, want_grad=True),
    grad_nei_q2=dict(N=32, d=3, b=6, q=2, G=16, kind="nei", temp=3e-2, alpha=0.1, ratio=True, nb=5, seed=64, want_grad=True),
    grad_pending=dict(N=30, d=2, b=7, q=1, x_pending=1, G=32, kind="ucb", temp=3e-2, alpha=0.1, randomized=True, seed=65, want_grad=True),
    # the reference's default SGLD shape at the BASELINE size: N = 1024, d = 8, G = 64, q = 2, ratio estimator, randomised
    grad_cfg2_ucb_q2=dict(N=1024, d=8, b=4, q=2, G=64, kind="ucb", temp=3e-2, alpha=1 / 32, randomized=True, ratio=True, chunk=2, seed=66, want_grad=True),
)

COV_CASES = dict(
    cov_soft=dict(N=60, d=3, n=40, G=64, temp=1e-2, alpha=0.1, refine=4, randomized=False, ratio=False, seed=41),
    cov_hard_rand=dict(N=60, d=3, n=40, G=64, temp=1e-6, alpha=0.2, refine=1, randomized=True, ratio=True, seed=42),
    cov_t2=dict(N=40, d=2, n=24, G=32, temp=1e-2, alpha=0.1, refine=2, randomized=False, ratio=False, seed=43, t=2),
    cov_t2_rand=dict(N=40, d=2, n=24, G=32, temp=1e-2, alpha=0.2, refine=0, randomized=True, ratio=True, seed=44, t=2),
)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--skip-big", action="store_true")
    ap.add_argument("--ehvi", action="store_true", help="only (re)generate tests/golden/ref_path_ehvi.pt")
    ap.add_argument("--grad", action="store_true", help="only (re)generate tests/golden/ref_path_grad.pt")
    ap.add_argument("--more", action="store_true", help="only (re)generate tests/golden/ref_path_more.pt")
    args = ap.parse_args()
    torch.set_default_dtype(torch.float64)       # the north star fixes fp64 (the reference's scripts default to fp32)
    ref = load_reference()
    os.makedirs(OUT, exist_ok=True)
    sel = lambda cases: {k: v for k, v in cases.items() if args.only in (None, k)}
    if args.more:
        torch.save(dict(acq=[run_acq_case(ref, k, **v) for k, v in MORE_CASES.items()], coverage=[]), os.path.join(OUT, "ref_path_more.pt"))
        return
    if args.grad:
        torch.save(dict(acq=[run_acq_case(ref, k, **v) for k, v in GRAD_CASES.items()]), os.path.join(OUT, "ref_path_grad.pt"))
        return
    if args.ehvi:
        torch.save(dict(ehvi=[run_ehvi_case(ref, k, **v) for k, v in EHVI_CASES.items()]), os.path.join(OUT, "ref_path_ehvi.pt"))
        return
    small = dict(acq=[run_acq_case(ref, k, **v) for k, v in sel(ACQ_CASES).items()],
                 coverage=[run_coverage_case(ref, k, **v) for k, v in sel(COV_CASES).items()])
    if args.only is None:
        torch.save(small, os.path.join(OUT, "ref_path.pt"))
    if args.only is None:
        torch.save(dict(ehvi=[run_ehvi_case(ref, k, **v) for k, v in EHVI_CASES.items()]), os.path.join(OUT, "ref_path_ehvi.pt"))
    if not args.skip_big:
        big = dict(acq=[run_acq_case(ref, k, **v) for k, v in sel(BIG_CASES).items()], coverage=[])
        if args.only is None:
            torch.save(big, os.path.join(OUT, "ref_path_n1024.pt"))


if __name__ == "__main__":
    main()
