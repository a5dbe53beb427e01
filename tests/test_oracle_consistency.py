"""Dense-literal oracle == closed-form oracle (SURVEY.md §8c known-answer test 2)."""
import math

import pytest
import torch

from oracle import closed_form, dense
from oracle import gp as ogp
from tests.problems import make_inputs, make_problem, oracle_state

CASES = [
    # N, d, q, b, G, temp, randomized, ratio, refine, kind
    (40, 3, 1, 6, 16, 3e-2, False, False, 0, "ucb"),
    (40, 3, 1, 6, 16, 3e-2, True, True, 2, "ei"),
    (33, 2, 2, 5, 12, 3e-2, True, True, 0, "ucb"),
    (33, 2, 2, 5, 12, 0.0, False, False, 1, "ucb"),
    (20, 4, 3, 4, 8, 1e-2, True, True, 3, "ei"),
    (40, 3, 1, 6, 16, 1e-6, True, True, 0, "ucb"),
    (8, 2, 1, 16, 64, 3e-2, True, False, 0, "ucb"),      # BASELINE config 1 shape (branin-like, N=8)
    (257, 8, 1, 2, 8, 3e-2, False, True, 0, "ucb"),
]


@pytest.mark.parametrize("N,d,q,b,G,temp,rand,ratio_on,refine,kind", CASES)
def test_dense_equals_closed_form(N, d, q, b, G, temp, rand, ratio_on, refine, kind):
    p = make_problem(N, d, seed=N, fn="sin")
    gp = oracle_state(p)
    inp = make_inputs(N, d, b, q, G, S=2, seed=N + q)
    ratio = ogp.make_ratio_mlp(d) if ratio_on else None
    kw = dict(kind=kind, alpha=0.2, temp=temp, grid_res=G, max_grid_refinements=refine,
              eps_grid=inp["eps_grid"], z=inp["z"], eps_mc=inp["eps_mc"], ratio=ratio, randomized=rand,
              u_rand=inp["u_rand"], optimistic=True, beta_prime=math.sqrt(0.2 * math.pi / 2), best_f=0.3)
    a = dense.conformal_acquisition(gp, inp["Xq"], **kw)
    c = closed_form.conformal_acquisition(gp, inp["Xq"], **kw)
    for key in ("acq", "target_grid", "logp", "mask", "ood", "values"):
        assert torch.allclose(a[key], c[key], rtol=1e-9, atol=1e-11), key


def test_untrained_ratio_estimator_is_identity():
    """ratio_estimation.py:81-86: a zero target network gives w == 1 (up to the 1e-6 clamps) -> same as None."""
    N, d, b, G = 30, 3, 4, 16
    gp = oracle_state(make_problem(N, d, seed=3, fn="sin"))
    inp = make_inputs(N, d, b, 1, G, seed=5)
    zero = ogp.RatioMLP(*(torch.zeros(s, dtype=torch.float64) for s in [(64, d), (64,), (64, 64), (64,), (1, 64), (1,)]))
    kw = dict(kind="ucb", alpha=0.1, temp=3e-2, grid_res=G, eps_grid=inp["eps_grid"], z=inp["z"],
              eps_mc=inp["eps_mc"], beta_prime=1.0)
    a = closed_form.conformal_acquisition(gp, inp["Xq"], ratio=None, **kw)
    c = closed_form.conformal_acquisition(gp, inp["Xq"], ratio=zero, **kw)
    assert torch.allclose(a["acq"], c["acq"], rtol=1e-12)


def test_hard_path_is_limit_of_soft_path():
    N, d, b, G = 30, 3, 6, 16
    gp = oracle_state(make_problem(N, d, seed=4, fn="sin"))
    inp = make_inputs(N, d, b, 1, G, seed=6)
    kw = dict(kind="ucb", alpha=0.15, grid_res=G, eps_grid=inp["eps_grid"], z=inp["z"], eps_mc=inp["eps_mc"], beta_prime=1.0)
    hard = closed_form.conformal_acquisition(gp, inp["Xq"], temp=0.0, **kw)
    soft = closed_form.conformal_acquisition(gp, inp["Xq"], temp=1e-9, **kw)
    assert torch.equal(hard["mask"], (soft["mask"] >= 0.5).double())


def test_finite_sample_coverage():
    """Unweighted, hard, non-randomised full conformal: marginal coverage >= 1 - alpha (the paper's guarantee)."""
    torch.manual_seed(0)
    N, d, alpha, G = 24, 2, 0.2, 128
    hits, trials = 0, 60
    from oracle.pure import conf_mask_to_bounds
    for trial in range(trials):
        Xall = torch.rand(N + 1, d, dtype=torch.float64)
        ls = torch.full((d,), 0.4, dtype=torch.float64)
        K = ogp.matern52(Xall, Xall, ls, 1.0) + 0.05 * torch.eye(N + 1, dtype=torch.float64)
        yall = torch.linalg.cholesky(K) @ torch.randn(N + 1, dtype=torch.float64)
        gp = ogp.build_state(Xall[:N], yall[:N], 0.0, ls, 1.0, 0.05)
        eps = torch.linspace(-4, 4, G, dtype=torch.float64).view(G, 1)
        z = 2.0 * ((torch.rand(N + 1, 8) > 0.5).double() - 0.5)
        out = closed_form.conformal_acquisition(
            gp, Xall[N:].view(1, 1, d), kind="ucb", alpha=alpha, temp=0.0, grid_res=G, eps_grid=eps, z=z,
            eps_mc=torch.zeros(1, G, 1, dtype=torch.float64), beta_prime=1.0)
        lb, ub = conf_mask_to_bounds(out["target_grid"].unsqueeze(-1), out["mask"].unsqueeze(-1))
        hits += int(lb.item() <= yall[N].item() <= ub.item())
    assert hits / trials >= 1 - alpha - 0.12   # 60 trials: ~2.3 sigma slack below the nominal 0.8


def _mobo_problem(N, d, m, seed):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(N, d, generator=g, dtype=torch.float64)
    Y = torch.stack([torch.sin(3 * X.sum(-1)), torch.cos(2 * X[:, 0]) - X[:, 1]], -1)
    gps = [ogp.build_state(X, Y[:, k], 0.0, [0.4 + 0.1 * k] * d, 1.0 + 0.2 * k, 1e-2 * (k + 1)) for k in range(m)]
    return g, X, Y, gps


@pytest.mark.parametrize("q,refine", [(1, 0), (2, 0), (2, 2), (3, 0)])
def test_cehvi_dense_equals_closed_form(q, refine):
    """Multi-objective oracle (intent semantics, App. C-5/C-6): dense refit pieces == closed-form pieces, joint batches."""
    from oracle import mobo
    N, d, b, G, m = 16, 2, 5, 12, 2          # BASELINE config 3 shape: BraninCurrin-like d=2, t=2, N=16
    g, X, Y, gps = _mobo_problem(N, d, m, 3)
    Xq = torch.rand(b, q, d, generator=g, dtype=torch.float64)
    eps_grid = torch.randn(G, q, m, generator=g, dtype=torch.float64)
    eps_mc = torch.randn(3, G, q, m, generator=g, dtype=torch.float64)
    z = 2.0 * ((torch.rand(N + q, 8, generator=g) > 0.5).double() - 0.5)
    u = torch.rand(b, G, q, m, generator=g, dtype=torch.float64)
    lo, up = mobo.box_decomposition_2d(torch.tensor([-1.5, -1.5], dtype=torch.float64), Y)
    assert lo.shape[0] >= 2 and torch.isinf(up[:, 1]).all()
    kw = dict(alpha=0.2, temp=3e-2, grid_res=G, eps_grid=eps_grid, z=z, eps_mc=eps_mc, cell_lower=lo, cell_upper=up,
              randomized=True, u_rand=u, optimistic=False, max_grid_refinements=refine)
    a = mobo.cehvi(gps, Xq, impl="dense", **kw)
    c = mobo.cehvi(gps, Xq, impl="closed", **kw)
    assert (a["acq"] > 0).any()
    assert torch.allclose(a["acq"], c["acq"], rtol=1e-9, atol=1e-12)
    assert torch.allclose(a["values"], c["values"], rtol=1e-9, atol=1e-12)
    if q == 1:      # q = 1 inclusion-exclusion has one term: the plain box overlap
        f = torch.tensor([[[[0.3, -0.2]]]], dtype=torch.float64)
        want = (torch.minimum(f[0, 0], up) - lo).clamp_min(0).prod(-1).sum()
        assert torch.allclose(mobo.qehvi(f, lo, up)[0, 0], want)


@pytest.mark.parametrize("q", [1, 2])
def test_cnehvi_dense_equals_closed_form(q):
    """qConformalNoisyExpectedHypervolumeImprovement (mobo_acquisitions.py:184-219): f(X) from the joint refit posterior over
    [X_baseline; X], one box decomposition per MC sample of the baseline values."""
    from oracle import mobo
    N, d, b, G, m, S, nb = 16, 2, 4, 10, 2, 3, 5
    g, X, Y, gps = _mobo_problem(N, d, m, 7)
    Xb = torch.cat([X[:nb - 1], torch.rand(1, d, generator=g, dtype=torch.float64)])
    Xq = torch.rand(b, q, d, generator=g, dtype=torch.float64)
    eps_grid = torch.randn(G, q, m, generator=g, dtype=torch.float64)
    eps_b = torch.randn(S, nb, m, generator=g, dtype=torch.float64)
    eps_new = torch.randn(S, G, q, m, generator=g, dtype=torch.float64)
    eps_mc = torch.cat([eps_b.unsqueeze(1).expand(S, G, nb, m), eps_new], dim=2)
    z = 2.0 * ((torch.rand(N + q, 8, generator=g) > 0.5).double() - 0.5)
    ref = torch.tensor([-1.5, -1.5], dtype=torch.float64)
    lo, up, fb = mobo.nehvi_cells(gps, Xb, eps_b, ref)
    assert lo.shape == up.shape and lo.shape[0] == S and lo.shape[2] == m
    kw = dict(alpha=0.2, temp=3e-2, grid_res=G, eps_grid=eps_grid, z=z, eps_mc=eps_mc, cell_lower=lo, cell_upper=up,
              optimistic=(q == 1), X_baseline=Xb)       # q = 2, optimistic: 4 masks must agree, G = 10 leaves no mass
    a = mobo.cehvi(gps, Xq, impl="dense", **kw)
    c = mobo.cehvi(gps, Xq, impl="closed", **kw)
    assert (a["acq"] > 0).any()
    assert torch.allclose(a["acq"], c["acq"], rtol=1e-8, atol=1e-12)
    assert torch.allclose(a["values"], c["values"], rtol=1e-8, atol=1e-12)


def test_qehvi_inclusion_exclusion_is_the_volume_of_the_union():
    """q = 2 against a direct computation: hv(a) + hv(b) - hv(min(a, b)) per cell; per-sample cells pick their own sample."""
    from oracle import mobo
    g = torch.Generator().manual_seed(5)
    lo = torch.tensor([[0.0, 0.5], [0.4, 0.0]], dtype=torch.float64)
    up = torch.tensor([[0.4, float("inf")], [float("inf"), float("inf")]], dtype=torch.float64)
    f = torch.rand(3, 4, 2, 2, generator=g, dtype=torch.float64) * 1.5
    hv = lambda y, l, u: (torch.minimum(y.unsqueeze(-2), u) - l).clamp_min(0).prod(-1).sum(-1)
    want = hv(f[..., 0, :], lo, up) + hv(f[..., 1, :], lo, up) - hv(f.min(-2).values, lo, up)
    assert torch.allclose(mobo.qehvi(f, lo, up), want)
    lo_s = torch.stack([lo, lo + 0.1, lo - 0.1])
    up_s = torch.stack([up, up, up])
    got = mobo.qehvi(f, lo_s, up_s)
    for s in range(3):
        assert torch.allclose(got[s], mobo.qehvi(f[s:s + 1], lo_s[s], up_s[s])[0])


@pytest.mark.parametrize("q", [1, 2, 3])
def test_cnei_dense_equals_closed_form(q):
    """qConformalNoisyExpectedImprovement (acquisitions.py:115-134): joint sampling over [X_baseline; X] under the refit model."""
    N, d, b, G, nb, S = 30, 3, 5, 12, 6, 2
    p = make_problem(N, d, seed=8, fn="sin")
    gp = oracle_state(p)
    inp = make_inputs(N, d, b, q, G, S=S, seed=8)
    g = torch.Generator().manual_seed(4)
    Xb = p["X"][torch.randperm(N, generator=g)[:nb]]                       # a pruned subset of the training inputs
    Xb = torch.cat([Xb[:-1], torch.rand(1, d, generator=g, dtype=torch.float64)])   # ... and one off-sample point
    eps_mc = torch.randn(S, G, nb + q, generator=g, dtype=torch.float64)
    kw = dict(kind="nei", alpha=0.2, temp=3e-2, grid_res=G, eps_grid=inp["eps_grid"], z=inp["z"], eps_mc=eps_mc,
              randomized=True, u_rand=inp["u_rand"], optimistic=False, X_baseline=Xb)
    a = dense.conformal_acquisition(gp, inp["Xq"], **kw)
    c = closed_form.conformal_acquisition(gp, inp["Xq"], **kw)
    assert (a["acq"] > 0).any()
    assert torch.allclose(a["values"], c["values"], rtol=1e-7, atol=1e-10)
    assert torch.allclose(a["acq"], c["acq"], rtol=1e-7, atol=1e-11)


@pytest.mark.parametrize("kind,ratio_on,rand", [("ucb", True, True), ("ei", False, False)])
def test_gradients_dense_equals_closed_form(kind, ratio_on, rand):
    """d acq / d X by autograd through both restatements (grid, its log-density and the optimistic mask are constants,
    as in the reference: helpers.py:118-119, acquisitions.py:58-59)."""
    N, d, b, G = 24, 3, 4, 10
    p = make_problem(N, d, seed=21, fn="sin")
    gp = oracle_state(p)
    inp = make_inputs(N, d, b, 1, G, S=2, seed=21)
    ratio = ogp.make_ratio_mlp(d) if ratio_on else None
    kw = dict(kind=kind, alpha=0.2, temp=5e-2, grid_res=G, eps_grid=inp["eps_grid"], z=inp["z"], eps_mc=inp["eps_mc"],
              ratio=ratio, randomized=rand, u_rand=inp["u_rand"], optimistic=True, beta_prime=1.2, best_f=0.1)
    grads = []
    for mod in (dense, closed_form):
        X = inp["Xq"].clone().requires_grad_(True)
        out = mod.conformal_acquisition(gp, X, **kw)
        (g,) = torch.autograd.grad(out["acq"].sum(), X)
        grads.append(g)
    assert grads[0].abs().max() > 1e-3
    assert torch.allclose(grads[0], grads[1], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("kind,ratio_on,rand,temp,S,opt", [
    ("ucb", True, True, 5e-2, 2, True), ("ei", False, False, 5e-2, 2, True), ("ucb", False, False, 3e-2, 1, True),
    ("ei", True, True, 3e-2, 1, False), ("ucb", True, True, 0.0, 2, True)])
def test_hand_derived_reverse_pass_equals_autograd(kind, ratio_on, rand, temp, S, opt):
    """oracle/analytic_grad.py (the algebra of the CUDA reverse pass) against autograd through the closed-form oracle."""
    from oracle import analytic_grad
    N, d, b, G = 24, 3, 5, 10
    gp = oracle_state(make_problem(N, d, seed=21, fn="sin"))
    inp = make_inputs(N, d, b, 1, G, S=S, seed=21)
    ratio = ogp.make_ratio_mlp(d) if ratio_on else None
    kw = dict(kind=kind, alpha=0.2, temp=temp, grid_res=G, eps_grid=inp["eps_grid"], z=inp["z"], eps_mc=inp["eps_mc"],
              ratio=ratio, randomized=rand, u_rand=inp["u_rand"], optimistic=opt, beta_prime=1.2, best_f=0.1)
    X = inp["Xq"].clone().requires_grad_(True)
    out = closed_form.conformal_acquisition(gp, X, **kw)
    (g,) = torch.autograd.grad(out["acq"].sum(), X)
    an = analytic_grad.conformal_acquisition_grad(gp, inp["Xq"], **kw)
    assert torch.allclose(an["acq"], out["acq"].detach(), rtol=1e-12, atol=1e-14)
    assert g.abs().max() > 1e-3
    assert torch.allclose(an["grad"], g, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("q,kind,ratio_on,rand,S", [(1, "ucb", True, True, 2), (2, "ucb", False, False, 1), (2, "ucb", True, True, 2),
                                                   (3, "ei", True, True, 2), (4, "ucb", True, True, 3)])
def test_hand_derived_joint_batch_reverse_pass_equals_autograd(q, kind, ratio_on, rand, S):
    """oracle/analytic_grad_q.py (the algebra of k_score_q_bwd / k_rows_bwd_q) against autograd through the closed-form oracle."""
    from oracle import analytic_grad_q
    N, d, b, G = 24, 3, 5, 10
    gp = oracle_state(make_problem(N, d, seed=21, fn="sin"))
    inp = make_inputs(N, d, b, q, G, S=S, seed=21 + q)
    ratio = ogp.make_ratio_mlp(d) if ratio_on else None
    kw = dict(kind=kind, alpha=0.2, temp=5e-2, grid_res=G, eps_grid=inp["eps_grid"], z=inp["z"], eps_mc=inp["eps_mc"],
              ratio=ratio, randomized=rand, u_rand=inp["u_rand"], optimistic=True, beta_prime=1.2, best_f=0.1)
    X = inp["Xq"].clone().requires_grad_(True)
    out = closed_form.conformal_acquisition(gp, X, **kw)
    (g,) = torch.autograd.grad(out["acq"].sum(), X)
    an = analytic_grad_q.conformal_acquisition_grad_q(gp, inp["Xq"], **kw)
    assert torch.allclose(an["acq"], out["acq"].detach(), rtol=1e-12, atol=1e-14)
    assert torch.allclose(an["grad"], g, rtol=1e-9, atol=1e-12)
