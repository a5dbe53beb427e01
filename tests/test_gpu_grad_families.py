"""GPU parity of d acq / d X for the rest of the acquisition family: cNEI, cEHVI, cNEHVI.

The reference reaches `acq_function(X).sum().backward()` for every acquisition that carries a ratio estimator
(conformalbo/utils.py:243-246 -> conformalbo/optim/optimize.py:224-235).  Oracle: torch.autograd through the closed-form
restatements (oracle/closed_form.py, oracle/mobo.py), which the dense literal restatement and the reference-source goldens
pin.  Tolerance: 1e-6 relative on the gradient rows with an absolute floor of 1e-9 times the largest entry.
"""
import pytest
import torch

from oracle import closed_form
from oracle import gp as ogp
from tests.problems import make_inputs, make_problem, oracle_state

pytestmark = pytest.mark.gpu


def _dev():
    assert torch.cuda.is_available(), "run with -m gpu on a CUDA box"
    return torch.device("cuda:0")


def _state(p, dev):
    from conformal_bayesopt_b200 import GPModelState
    return GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)


def _estimator(d, dev):
    from tests.test_gpu_api import _estimator as mk
    return mk(d, dev)


def _grad_close(got, want):
    scale = want.abs().max().clamp_min(1e-300)
    return torch.allclose(got, want, rtol=1e-6, atol=float(1e-9 * scale))


@pytest.mark.parametrize("temp,q,ratio_on,S", [(3e-2, 2, True, 2), (3e-2, 1, False, 1), (1e-2, 3, True, 1), (3e-2, 4, False, 2),
                                               (0.0, 2, True, 2)])
def test_cnei_grad_matches_autograd_of_oracle(temp, q, ratio_on, S):
    """qConformalNoisyExpectedImprovement under autograd: k_score_q_bwd (values are constants, their adjoints exported) +
    k_nei_values_q_bwd (Cholesky reverse of the (nb+q)^2 joint factor)."""
    from conformal_bayesopt_b200.acquisitions import qConformalNoisyExpectedImprovement
    from conformal_bayesopt_b200.sampling import IIDNormalSampler
    dev = _dev()
    N, d, b, G, nb = 90, 4, 11, 32, 7
    p = make_problem(N, d, seed=13)
    inp = make_inputs(N, d, b, q, G, S=S, seed=13)
    g = torch.Generator().manual_seed(4)
    Xb = torch.cat([p["X"][torch.randperm(N, generator=g)[:nb - 1]], torch.rand(1, d, generator=g, dtype=torch.float64)])
    eps_mc = torch.randn(S, G, nb + q, generator=g, dtype=torch.float64)
    ratio, est = _estimator(d, dev) if ratio_on else (None, None)
    gs = IIDNormalSampler(G)
    gs.set_base_samples(inp["eps_grid"])
    mc = IIDNormalSampler(S, batch_range=(0, -3))
    mc.set_base_samples(eps_mc)
    acq = qConformalNoisyExpectedImprovement(alpha=0.15, temp=temp, grid_res=G, max_grid_refinements=0, ratio_estimator=est,
                                             grid_sampler=gs, randomized=True, model=_state(p, dev), X_baseline=Xb.to(dev),
                                             sampler=mc, prune_baseline=False)
    acq.fixed_probes, acq.fixed_uniforms = inp["z"], inp["u_rand"].to(dev)
    X = inp["Xq"].to(dev).requires_grad_(True)
    got = acq(X)
    got.sum().backward()
    Xo = inp["Xq"].clone().requires_grad_(True)
    want = closed_form.conformal_acquisition(oracle_state(p), Xo, kind="nei", alpha=0.15, temp=temp, grid_res=G,
                                             eps_grid=inp["eps_grid"], z=inp["z"], eps_mc=eps_mc, ratio=ratio, randomized=True,
                                             u_rand=inp["u_rand"], optimistic=False, X_baseline=Xb)
    (gw,) = torch.autograd.grad(want["acq"].sum(), Xo)
    assert torch.allclose(got.detach().cpu(), want["acq"].detach(), rtol=1e-6, atol=1e-10)
    assert gw.abs().max() > 0
    assert _grad_close(X.grad.cpu(), gw)
    # the device-side optimiser's entry point returns the same rows
    v, dX = acq.value_and_grad(inp["Xq"].to(dev))
    assert torch.equal(dX.cpu(), X.grad.cpu()) and torch.equal(v.cpu(), got.detach().cpu())


def _mobo_problem(q, pending, b, G, S, nb=0, seed=3):
    from conformal_bayesopt_b200 import GPModelState
    dev = _dev()
    N, d, m = 16, 2, 2
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(N, d, generator=g, dtype=torch.float64)
    Y = torch.stack([torch.sin(3 * X.sum(-1)), torch.cos(2 * X[:, 0]) - X[:, 1]], -1)
    hyp = [(0.0, [0.4 + 0.1 * k] * d, 1.0 + 0.2 * k, 1e-2 * (k + 1)) for k in range(m)]
    gps = [ogp.build_state(X, Y[:, k], *hyp[k]) for k in range(m)]
    states = [GPModelState.fit_fixed(X, Y[:, k], *hyp[k], device=dev) for k in range(m)]
    Xb = torch.cat([X[:nb - 1], torch.rand(1, d, generator=g, dtype=torch.float64)]) if nb else None
    Xq = torch.rand(b, q, d, generator=g, dtype=torch.float64)
    if pending:
        Xq[:, q - pending:] = Xq[0, q - pending:]
    eps_grid = torch.randn(G, q, m, generator=g, dtype=torch.float64)
    eps_new = torch.randn(S, G, q, m, generator=g, dtype=torch.float64)
    eps_b = torch.randn(S, nb, m, generator=g, dtype=torch.float64) if nb else None
    z = 2.0 * ((torch.rand(N + q, 8, generator=g) > 0.5).double() - 0.5)
    u = torch.rand(b, G, q, m, generator=g, dtype=torch.float64)
    ref = torch.tensor([-1.5, -1.5], dtype=torch.float64)
    return dict(dev=dev, N=N, d=d, m=m, X=X, Y=Y, gps=gps, states=states, Xb=Xb, Xq=Xq, eps_grid=eps_grid, eps_new=eps_new,
                eps_b=eps_b, z=z, u=u, ref=ref)


@pytest.mark.parametrize("q,pending,optimistic", [(1, 0, False), (2, 0, True), (3, 1, False)])
def test_cehvi_grad_matches_autograd_of_oracle(q, pending, optimistic):
    """qConformalExpectedHypervolumeImprovement under autograd: k_ehvi_q<Q, true> hands mask / ood / sample adjoints to one
    vector-Jacobian call per objective (cbo_ext_adjoints_f64)."""
    from conformal_bayesopt_b200.mobo_acquisitions import box_decomposition_2d, qConformalExpectedHypervolumeImprovement
    from conformal_bayesopt_b200.sampling import IIDNormalSampler
    from oracle import mobo
    b, G, S = 9, 32, 2
    P = _mobo_problem(q, pending, b, G, S)
    dev = P["dev"]
    lo, up = box_decomposition_2d(P["ref"], P["Y"])
    gs = IIDNormalSampler(G)
    gs.set_base_samples(P["eps_grid"])
    mc = IIDNormalSampler(S, batch_range=(0, -3))
    mc.set_base_samples(P["eps_new"])
    acq = qConformalExpectedHypervolumeImprovement(alpha=0.2, temp=3e-2, grid_res=G, max_grid_refinements=0,
                                                   ratio_estimator=None, optimistic=optimistic, grid_sampler=gs, randomized=True,
                                                   model=P["states"], ref_point=P["ref"], partitioning=(lo, up), sampler=mc,
                                                   X_pending=P["Xq"][0, q - pending:].to(dev) if pending else None)
    acq.fixed_probes, acq.fixed_uniforms = P["z"], P["u"].to(dev)
    X = P["Xq"][:, :q - pending].to(dev).requires_grad_(True)
    got = acq(X)
    got.sum().backward()
    Xo = P["Xq"].clone().requires_grad_(True)
    want = mobo.cehvi(P["gps"], Xo, alpha=0.2, temp=3e-2, grid_res=G, eps_grid=P["eps_grid"], z=P["z"], eps_mc=P["eps_new"],
                      cell_lower=lo, cell_upper=up, randomized=True, u_rand=P["u"], optimistic=optimistic)
    (gw,) = torch.autograd.grad(want["acq"].sum(), Xo)
    gw = gw[:, :q - pending]                               # pending points carry no gradient
    assert torch.allclose(got.detach().cpu(), want["acq"].detach(), rtol=1e-6, atol=1e-10)
    assert gw.abs().max() > 0
    assert _grad_close(X.grad.cpu(), gw)
    with torch.no_grad():
        plain = acq(P["Xq"][:, :q - pending].to(dev))
    assert torch.allclose(plain, got.detach(), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("q,pending", [(1, 0), (2, 0), (2, 1)])
def test_cnehvi_grad_matches_autograd_of_oracle(q, pending):
    """qConformalNoisyExpectedHypervolumeImprovement under autograd: joint samples with the baseline under every refit model
    (CBO_ACQ_JOINT + k_nei_values_q_bwd in raw-sample mode); the per-sample box decompositions are constants."""
    from conformal_bayesopt_b200.mobo_acquisitions import qConformalNoisyExpectedHypervolumeImprovement
    from conformal_bayesopt_b200.sampling import IIDNormalSampler
    from oracle import mobo
    b, G, S, nb = 8, 32, 3, 6
    P = _mobo_problem(q, pending, b, G, S, nb=nb, seed=9)
    dev = P["dev"]
    eps_mc = torch.cat([P["eps_b"].unsqueeze(1).expand(S, G, nb, P["m"]), P["eps_new"]], dim=2)
    lo, up, fb = mobo.nehvi_cells(P["gps"], P["Xb"], P["eps_b"], P["ref"])
    gs = IIDNormalSampler(G)
    gs.set_base_samples(P["eps_grid"])
    mc = IIDNormalSampler(S, batch_range=(0, -3))
    mc.set_base_samples(P["eps_new"])
    acq = qConformalNoisyExpectedHypervolumeImprovement(alpha=0.2, temp=3e-2, grid_res=G, max_grid_refinements=0,
                                                        ratio_estimator=None, optimistic=True, grid_sampler=gs, randomized=True,
                                                        model=P["states"], ref_point=P["ref"], X_baseline=P["Xb"].to(dev),
                                                        sampler=mc, prune_baseline=False, baseline_base_samples=P["eps_b"],
                                                        X_pending=P["Xq"][0, q - pending:].to(dev) if pending else None)
    acq.fixed_probes, acq.fixed_uniforms = P["z"], P["u"].to(dev)
    X = P["Xq"][:, :q - pending].to(dev).requires_grad_(True)
    got = acq(X)
    got.sum().backward()
    Xo = P["Xq"].clone().requires_grad_(True)
    want = mobo.cehvi(P["gps"], Xo, alpha=0.2, temp=3e-2, grid_res=G, eps_grid=P["eps_grid"], z=P["z"], eps_mc=eps_mc,
                      cell_lower=lo, cell_upper=up, randomized=True, u_rand=P["u"], optimistic=True, X_baseline=P["Xb"])
    (gw,) = torch.autograd.grad(want["acq"].sum(), Xo)
    gw = gw[:, :q - pending]
    assert torch.allclose(got.detach().cpu(), want["acq"].detach(), rtol=1e-6, atol=1e-10)
    assert gw.abs().max() > 0
    assert _grad_close(X.grad.cpu(), gw)


@pytest.mark.parametrize("nb,q", [(40, 2), (60, 4), (33, 1)])
def test_cnei_wide_baseline_value_and_grad(nb, q):
    """More than 32 joint points (each lane owns two rows of the (nb+q)^2 factor): value and gradient against the oracle."""
    from conformal_bayesopt_b200.acquisitions import qConformalNoisyExpectedImprovement
    from conformal_bayesopt_b200.sampling import IIDNormalSampler
    dev = _dev()
    N, d, b, G, S = 120, 3, 6, 16, 2
    p = make_problem(N, d, seed=21, noise=1e-2)
    inp = make_inputs(N, d, b, q, G, S=S, seed=21)
    g = torch.Generator().manual_seed(5)
    Xb = torch.rand(nb, d, generator=g, dtype=torch.float64)       # spread-out baseline: a well-conditioned joint covariance
    eps_mc = torch.randn(S, G, nb + q, generator=g, dtype=torch.float64)
    gs = IIDNormalSampler(G)
    gs.set_base_samples(inp["eps_grid"])
    mc = IIDNormalSampler(S, batch_range=(0, -3))
    mc.set_base_samples(eps_mc)
    acq = qConformalNoisyExpectedImprovement(alpha=0.15, temp=3e-2, grid_res=G, max_grid_refinements=0, ratio_estimator=None,
                                             grid_sampler=gs, randomized=True, model=_state(p, dev), X_baseline=Xb.to(dev),
                                             sampler=mc)
    acq.fixed_probes, acq.fixed_uniforms = inp["z"], inp["u_rand"].to(dev)
    X = inp["Xq"].to(dev).requires_grad_(True)
    got = acq(X)
    got.sum().backward()
    Xo = inp["Xq"].clone().requires_grad_(True)
    want = closed_form.conformal_acquisition(oracle_state(p), Xo, kind="nei", alpha=0.15, temp=3e-2, grid_res=G,
                                             eps_grid=inp["eps_grid"], z=inp["z"], eps_mc=eps_mc, randomized=True,
                                             u_rand=inp["u_rand"], optimistic=False, X_baseline=Xb)
    (gw,) = torch.autograd.grad(want["acq"].sum(), Xo)
    assert torch.allclose(got.detach().cpu(), want["acq"].detach(), rtol=1e-6, atol=1e-10)
    scale = gw.abs().max()
    assert scale > 0 and torch.allclose(X.grad.cpu(), gw, rtol=1e-5, atol=float(1e-7 * scale))   # 60-wide factor: cond(Sigma) enters
    with torch.no_grad():
        assert torch.allclose(acq(inp["Xq"].to(dev)), got.detach(), rtol=1e-12, atol=1e-14)


def test_prune_baseline_matches_botorch_semantics():
    """`prune_baseline=True` (hydra_config/acq_fn/cnei.yaml:3): botorch's prune_inferior_points restated on the CPU with the
    same injected base samples -- keep the points that win at least one joint posterior sample."""
    from conformal_bayesopt_b200.acquisitions import qConformalNoisyExpectedImprovement
    from conformal_bayesopt_b200.sampling import prune_inferior_points, prune_inferior_points_multi_objective
    dev = _dev()
    N, d, S = 90, 4, 256
    p = make_problem(N, d, seed=13, noise=1e-2)
    st = _state(p, dev)
    gp = oracle_state(p)
    g = torch.Generator().manual_seed(8)
    base = torch.randn(S, N, generator=g, dtype=torch.float64)
    kept = prune_inferior_points(st, p["X"].to(dev), base_samples=base)
    # oracle: noise-free posterior at the training inputs
    K = gp.k(gp.X, gp.X)
    Vb = torch.linalg.solve_triangular(gp.L, K, upper=False)
    mean = gp.mean + K @ gp.alpha
    cov = K - Vb.T @ Vb
    cov = 0.5 * (cov + cov.T)
    L, info = torch.linalg.cholesky_ex(cov)
    jitter = 1e-8
    while int(info) != 0:
        L, info = torch.linalg.cholesky_ex(cov + jitter * torch.eye(N, dtype=cov.dtype))
        jitter *= 10.0
    want_idx = torch.unique((mean.unsqueeze(0) + base @ L.T).argmax(-1))
    assert 0 < kept.shape[0] < N
    assert kept.shape[0] == want_idx.numel() and torch.allclose(kept.cpu(), p["X"][want_idx], atol=0)
    acq = qConformalNoisyExpectedImprovement(alpha=0.15, temp=3e-2, grid_res=16, max_grid_refinements=0, ratio_estimator=None,
                                             model=st, X_baseline=p["X"].to(dev), prune_baseline=True, prune_base_samples=base)
    assert torch.equal(acq.X_baseline.cpu(), kept.cpu())
    # Sobol default draws: a deterministic, non-empty subset
    a = prune_inferior_points(st, p["X"].to(dev), num_samples=128, seed=3)
    b = prune_inferior_points(st, p["X"].to(dev), num_samples=128, seed=3)
    assert torch.equal(a, b) and 0 < a.shape[0] <= N
    # two objectives: every kept point is Pareto-optimal in at least one sample
    P = _mobo_problem(1, 0, 2, 8, 1)
    base2 = torch.randn(64, P["N"], 2, generator=g, dtype=torch.float64)
    kept2 = prune_inferior_points_multi_objective(P["states"], P["X"].to(dev), P["ref"], base_samples=base2)
    assert 0 < kept2.shape[0] <= P["N"]


@pytest.mark.parametrize("cls_kind", ["cehvi", "cnehvi"])
def test_multi_objective_grad_with_the_grid_leak_switch(cls_kind):
    """`acq.grid_grad_leak = True` on the multi-objective classes (per-output vector-Jacobian calls with
    cbo_params.grid_grad_leak): the gradient then includes what flows through each output's target grid via the prior block
    k(x_j, x_k) -- compared with torch.autograd through the oracle under the same convention
    (`mobo.cehvi(lazy_prior_grad=True)`); without the switch the rows differ."""
    from conformal_bayesopt_b200.mobo_acquisitions import (box_decomposition_2d, qConformalExpectedHypervolumeImprovement,
                                                         qConformalNoisyExpectedHypervolumeImprovement)
    from conformal_bayesopt_b200.sampling import IIDNormalSampler
    from oracle import mobo
    q, b, G, S = 2, 7, 32, 2
    nb = 5 if cls_kind == "cnehvi" else 0
    P = _mobo_problem(q, 0, b, G, S, nb=nb)
    dev = P["dev"]
    gs = IIDNormalSampler(G)
    gs.set_base_samples(P["eps_grid"])
    mc = IIDNormalSampler(S, batch_range=(0, -3))
    mc.set_base_samples(P["eps_new"])
    common = dict(alpha=0.2, temp=3e-2, grid_res=G, max_grid_refinements=0, ratio_estimator=None, optimistic=False,
                  grid_sampler=gs, randomized=True, model=P["states"], ref_point=P["ref"], sampler=mc)
    if cls_kind == "cehvi":
        lo, up = box_decomposition_2d(P["ref"], P["Y"])
        acq = qConformalExpectedHypervolumeImprovement(partitioning=(lo, up), **common)
        okw = dict(eps_mc=P["eps_new"], cell_lower=lo, cell_upper=up)
    else:
        acq = qConformalNoisyExpectedHypervolumeImprovement(X_baseline=P["Xb"].to(dev), prune_baseline=False,
                                                            baseline_base_samples=P["eps_b"], **common)
        lo, up, _ = mobo.nehvi_cells(P["gps"], P["Xb"], P["eps_b"], P["ref"])
        eps_joint = torch.cat([P["eps_b"].unsqueeze(1).expand(S, G, nb, P["m"]), P["eps_new"]], dim=2)
        okw = dict(eps_mc=eps_joint, cell_lower=lo, cell_upper=up, X_baseline=P["Xb"])
    acq.fixed_probes, acq.fixed_uniforms = P["z"], P["u"].to(dev)
    grads = {}
    for leak in (False, True):
        acq.grid_grad_leak = leak
        X = P["Xq"].to(dev).requires_grad_(True)
        acq(X).sum().backward()
        grads[leak] = X.grad.cpu()
        Xo = P["Xq"].clone().requires_grad_(True)
        want = mobo.cehvi(P["gps"], Xo, alpha=0.2, temp=3e-2, grid_res=G, eps_grid=P["eps_grid"], z=P["z"], randomized=True,
                          u_rand=P["u"], optimistic=False, lazy_prior_grad=leak, **okw)
        (gw,) = torch.autograd.grad(want["acq"].sum(), Xo)
        assert gw.abs().max() > 0
        assert _grad_close(grads[leak], gw), (leak, float((grads[leak] - gw).abs().max() / gw.abs().max()))
    assert float((grads[True] - grads[False]).abs().max() / grads[False].abs().max()) > 1e-4
