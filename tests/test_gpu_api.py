"""GPU tests of the reference-facing interface (acquisition classes / helpers) against the oracle."""
import math

import pytest
import torch

from oracle import closed_form, pure
from oracle import gp as ogp
from tests.problems import make_inputs, make_problem, oracle_state

pytestmark = pytest.mark.gpu


def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _state(p, dev):
    from conformal_bayesopt_b200 import GPModelState
    return GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)


def _estimator(d, dev):
    """RatioEstimator whose target network carries the oracle's seeded weights."""
    from conformal_bayesopt_b200.ratio_estimation import RatioEstimator
    r = ogp.make_ratio_mlp(d)
    est = RatioEstimator(in_size=d, device=dev, dtype=torch.float64)
    lin = [m for m in est._target_network if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for m, (W, b) in zip(lin, [(r.W1, r.b1), (r.W2, r.b2), (r.W3, r.b3)]):
            m.weight.copy_(W)
            m.bias.copy_(b)
    return r, est


@pytest.mark.parametrize("cls_name,extra,kind", [("qConformalUpperConfidenceBound", dict(beta=0.2), "ucb"),
                                                 ("qConformalExpectedImprovement", dict(best_f=0.3), "ei")])
def test_acquisition_forward_matches_oracle(cls_name, extra, kind):
    from conformal_bayesopt_b200 import acquisitions
    from conformal_bayesopt_b200.sampling import IIDNormalSampler
    dev = _dev()
    N, d, b, G = 120, 6, 33, 64
    p = make_problem(N, d, seed=11)
    inp = make_inputs(N, d, b, 1, G, S=1, seed=11)
    ratio, est = _estimator(d, dev)
    grid_sampler = IIDNormalSampler(G)
    grid_sampler.set_base_samples(inp["eps_grid"])
    mc = IIDNormalSampler(1, batch_range=(0, -3))
    mc.set_base_samples(inp["eps_mc"])
    acq = getattr(acquisitions, cls_name)(alpha=0.1, temp=3e-2, grid_res=G, max_grid_refinements=0, ratio_estimator=est,
                                          grid_sampler=grid_sampler, randomized=True, model=_state(p, dev), sampler=mc,
                                          **extra)
    acq.fixed_probes = inp["z"]
    acq.fixed_uniforms = inp["u_rand"][:, :, 0].to(dev)
    with torch.no_grad():
        got = acq(inp["Xq"].to(dev))
    assert got.shape == (b,)
    kw = dict(kind=kind, alpha=0.1, temp=3e-2, grid_res=G, randomized=True, optimistic=acq.optimistic,
              beta_prime=math.sqrt(0.2 * math.pi / 2), best_f=0.3)
    want = closed_form.conformal_acquisition(oracle_state(p), inp["Xq"], eps_grid=inp["eps_grid"], z=inp["z"],
                                             eps_mc=inp["eps_mc"], ratio=ratio, u_rand=inp["u_rand"], **kw)
    assert torch.allclose(got.cpu(), want["acq"], rtol=1e-6, atol=1e-9)
    # 2-D input gets a t-batch dimension (t_batch_mode_transform); output shape == X.shape[:-2]
    acq.fixed_uniforms = inp["u_rand"][:1, :, 0].to(dev)
    with torch.no_grad():
        one = acq(inp["Xq"][0].to(dev))
    assert one.shape == (1,) and torch.allclose(one.cpu(), want["acq"][:1], rtol=1e-6, atol=1e-9)
    # gradients: built for q = 1 (tests/test_gpu_grad.py); the value of a grad-enabled call is the same value
    acq.fixed_uniforms = inp["u_rand"][:, :, 0].to(dev)
    Xg = inp["Xq"].to(dev).requires_grad_(True)
    val = acq(Xg)
    val.sum().backward()
    assert torch.allclose(val.detach().cpu(), want["acq"], rtol=1e-6, atol=1e-9) and torch.isfinite(Xg.grad).all()
    # X_pending makes q = 2: injected inputs of the wrong shape (z has N+1 rows) are refused, not silently mis-read
    acq.set_X_pending(inp["Xq"][0].to(dev))
    from conformal_bayesopt_b200 import _lib
    with pytest.raises(_lib.CboError):
        with torch.no_grad():
            acq(inp["Xq"].to(dev))
    # ... and with consistent inputs the pending point is scored jointly (concatenate_pending_points)
    inp2 = make_inputs(N, d, 7, 2, G, S=1, seed=12)
    grid_sampler.set_base_samples(inp2["eps_grid"])
    mc.set_base_samples(inp2["eps_mc"])
    acq.fixed_probes = inp2["z"]
    acq.fixed_uniforms = inp2["u_rand"].to(dev)
    acq.set_X_pending(inp2["Xq"][0, 1:].to(dev))
    Xj = torch.cat([inp2["Xq"][:, :1], inp2["Xq"][0, 1:].expand(7, 1, d)], dim=1)
    with torch.no_grad():
        got2 = acq(inp2["Xq"][:, :1].to(dev))
    want2 = closed_form.conformal_acquisition(oracle_state(p), Xj, eps_grid=inp2["eps_grid"], z=inp2["z"],
                                              eps_mc=inp2["eps_mc"], ratio=ratio, u_rand=inp2["u_rand"], **kw)
    assert torch.allclose(got2.cpu(), want2["acq"], rtol=1e-6, atol=1e-9)


def test_fresh_probes_change_values_like_the_reference():
    """helpers.py:399: z is redrawn on every call, so values are stochastic call to call; seeding the CPU generator pins them."""
    from conformal_bayesopt_b200.acquisitions import qConformalUpperConfidenceBound
    dev = _dev()
    p = make_problem(60, 3, seed=2)
    X = torch.rand(16, 1, 3, dtype=torch.float64, device=dev)
    acq = qConformalUpperConfidenceBound(alpha=0.1, temp=3e-2, grid_res=32, max_grid_refinements=0, ratio_estimator=None,
                                         model=_state(p, dev), beta=0.2)
    with torch.no_grad():
        torch.manual_seed(5); a = acq(X)
        torch.manual_seed(5); b = acq(X)
        c = acq(X)
    assert torch.equal(a, b) and not torch.equal(a, c)


def test_construct_conformal_bands_and_coverage():
    from conformal_bayesopt_b200 import helpers
    from conformal_bayesopt_b200.sampling import IIDNormalSampler
    dev = _dev()
    N, d, n, G = 150, 4, 40, 64
    p = make_problem(N, d, seed=21)
    gp = oracle_state(p)
    g = torch.Generator().manual_seed(7)
    Xh = torch.rand(n, d, generator=g, dtype=torch.float64)
    mean_h = gp.mean + gp.k(Xh, gp.X) @ gp.alpha
    Yh = (mean_h + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)).unsqueeze(-1)
    z = 2.0 * ((torch.rand(N + 1, 8, generator=g) > 0.5).double() - 0.5)
    eps = torch.randn(G, n, 1, generator=g, dtype=torch.float64)           # collapse_batch_dims=False layout (G, b, q)
    sampler = IIDNormalSampler(G, collapse_batch_dims=False)
    sampler.set_base_samples(eps)
    alpha, temp = 0.1, 1e-2
    tg, w, mk, cond, ood = helpers.construct_conformal_bands(_state(p, dev), Xh[:, None].to(dev), alpha, temp, G, 0, sampler,
                                                             probes=z)
    assert tg.shape == (n, G, 1, 1) and w.shape == (n, G, 1, 1) and mk.shape == (n, G, 1, 1) and ood.shape == (n, G, 1, 1)
    assert cond is None
    want = closed_form.conformal_acquisition(gp, Xh[:, None], kind="ucb", alpha=alpha, temp=temp, grid_res=G,
                                             eps_grid=eps.movedim(0, 1), z=z, eps_mc=torch.zeros(1, G, 1, dtype=torch.float64),
                                             beta_prime=1.0, optimistic=False)
    assert torch.allclose(tg.cpu()[..., 0], want["target_grid"], rtol=1e-6, atol=1e-9)
    assert torch.allclose(mk.cpu()[..., 0], want["mask"], rtol=0, atol=1e-6)
    assert torch.allclose(w.cpu()[:, :, 0, 0], want["logp"], rtol=1e-6, atol=1e-9)
    lb, ub = helpers.conf_mask_to_bounds(tg, mk)
    lbo, ubo = pure.conf_mask_to_bounds(want["target_grid"].unsqueeze(-1), want["mask"].unsqueeze(-1))
    assert torch.allclose(lb.cpu(), lbo, rtol=1e-6, atol=1e-7) and torch.allclose(ub.cpu(), ubo, rtol=1e-6, atol=1e-7)
    # assess_coverage (helpers.py:336-387) against the same quantities computed from the oracle
    sampler2 = IIDNormalSampler(G, collapse_batch_dims=False)
    sampler2.set_base_samples(eps)
    std_c, conf_c = helpers.assess_coverage(_state(p, dev), Xh.to(dev), Yh.to(dev), alpha, temp, G, 0, probes=z,
                                            grid_sampler=sampler2)
    from scipy import stats
    sd = want["S"][:, 0, 0].sqrt().unsqueeze(-1)
    sc = stats.norm.ppf(1 - alpha / 2)
    std_o = ((Yh >= want["mu"] - sc * sd) * (Yh <= want["mu"] + sc * sd)).float().mean().item()
    conf_o = ((Yh >= lbo.squeeze(-2)) * (Yh <= ubo.squeeze(-2))).float().nanmean().item()
    assert abs(std_c - std_o) < 1e-6 and abs(conf_c - conf_o) < 1e-6
    assert conf_c >= 1 - alpha - 0.15                      # full conformal bands cover (n = 40: loose statistical check)


def test_reference_coverage_tests_on_cuda():
    """tests/test_coverage.py:38-87 ported verbatim onto the CUDA conf_mask_to_bounds (t = 1 and t = 2)."""
    from conformal_bayesopt_b200 import helpers
    dev = _dev()
    torch.manual_seed(0)
    for target_dim in (1, 2):
        b, G, cvrg = 256, 256, 0.8
        grid = torch.randn(b, G, 1, target_dim)
        idx_shape = (b, 1, 1, target_dim)
        s1 = torch.gather(grid, index=torch.randint(G // 2, G, idx_shape), dim=-3).squeeze(-3)
        s2 = torch.gather(grid, index=torch.randint(0, G // 2, idx_shape), dim=-3).squeeze(-3)
        real_lb, real_ub = torch.min(s1, s2) - 1e-6, torch.max(s1, s2) + 1e-6
        inb = real_lb + torch.rand_like(real_lb) * (real_ub - real_lb)
        rm = (torch.rand_like(real_lb) < 0.5).float()
        outb = rm * (real_lb - 1.0) + (1 - rm) * (real_ub + 1.0)
        rm = (torch.rand_like(real_lb) < cvrg).float()
        tgts = rm * inb + (1 - rm) * outb
        mask = ((grid >= real_lb.unsqueeze(-3)) * (grid <= real_ub.unsqueeze(-3))).float()
        est_lb, est_ub = helpers.conf_mask_to_bounds(grid.to(dev), mask.to(dev))
        est = ((tgts >= est_lb.cpu()) * (tgts <= est_ub.cpu())).float().mean(dim=[0, -2])
        act = ((tgts >= real_lb) * (tgts <= real_ub)).float().mean(dim=[0, -2])
        assert torch.allclose(act, est)


def test_ratio_estimator_forward_on_cuda():
    dev = _dev()
    ratio, est = _estimator(5, dev)
    x = torch.randn(3, 7, 5, dtype=torch.float64)
    out = est(x.to(dev))
    assert out.shape == (3, 7) and torch.allclose(out.cpu(), ratio(x), rtol=1e-10)


@pytest.mark.parametrize("q,pending,refine", [(1, 0, 0), (2, 0, 0), (1, 1, 0), (3, 1, 0), (2, 0, 2)])
def test_cehvi_matches_oracle(q, pending, refine):
    """BASELINE config 3 shape: two objectives, conformal bands over both (intent semantics, see oracle/mobo.py); joint
    batches of q points of which `pending` are X_pending (concatenate_pending_points), per-output grid refinement."""
    from conformal_bayesopt_b200 import GPModelState
    from conformal_bayesopt_b200.mobo_acquisitions import box_decomposition_2d, qConformalExpectedHypervolumeImprovement
    from conformal_bayesopt_b200.sampling import IIDNormalSampler
    from oracle import mobo
    dev = _dev()
    N, d, b, G, m, S = 16, 2, 40, 64, 2, 2
    g = torch.Generator().manual_seed(3)
    X = torch.rand(N, d, generator=g, dtype=torch.float64)
    Y = torch.stack([torch.sin(3 * X.sum(-1)), torch.cos(2 * X[:, 0]) - X[:, 1]], -1)
    hyp = [(0.0, [0.4 + 0.1 * k] * d, 1.0 + 0.2 * k, 1e-2 * (k + 1)) for k in range(m)]
    gps = [ogp.build_state(X, Y[:, k], *hyp[k]) for k in range(m)]
    states = [GPModelState.fit_fixed(X, Y[:, k], *hyp[k], device=dev) for k in range(m)]
    Xq = torch.rand(b, q, d, generator=g, dtype=torch.float64)
    if pending:
        Xq[:, q - pending:] = Xq[0, q - pending:]                  # the pending points are shared by every candidate batch
    eps_grid = torch.randn(G, q, m, generator=g, dtype=torch.float64)
    eps_mc = torch.randn(S, G, q, m, generator=g, dtype=torch.float64)
    z = 2.0 * ((torch.rand(N + q, 8, generator=g) > 0.5).double() - 0.5)
    u = torch.rand(b, G, q, m, generator=g, dtype=torch.float64)
    ref = torch.tensor([-1.5, -1.5], dtype=torch.float64)
    lo, up = box_decomposition_2d(ref, Y)
    lo_o, up_o = mobo.box_decomposition_2d(ref, Y)
    assert torch.equal(lo, lo_o) and torch.equal(up, up_o)
    gs = IIDNormalSampler(G)
    gs.set_base_samples(eps_grid)
    mc = IIDNormalSampler(S, batch_range=(0, -3))
    mc.set_base_samples(eps_mc)
    for optimistic in (False, True):
        acq = qConformalExpectedHypervolumeImprovement(alpha=0.2, temp=3e-2, grid_res=G, max_grid_refinements=refine,
                                                       ratio_estimator=None, optimistic=optimistic, grid_sampler=gs,
                                                       randomized=True, model=states, ref_point=ref, partitioning=(lo, up),
                                                       sampler=mc, X_pending=Xq[0, q - pending:].to(dev) if pending else None)
        acq.fixed_probes, acq.fixed_uniforms = z, u.to(dev)
        with torch.no_grad():
            got = acq(Xq[:, :q - pending].to(dev))
        want = mobo.cehvi(gps, Xq, alpha=0.2, temp=3e-2, grid_res=G, eps_grid=eps_grid, z=z, eps_mc=eps_mc, cell_lower=lo,
                          cell_upper=up, randomized=True, u_rand=u, optimistic=optimistic, max_grid_refinements=refine)
        assert got.shape == (b,) and (want["acq"] > 0).any()
        assert torch.allclose(got.cpu(), want["acq"], rtol=1e-6, atol=1e-10)
        assert torch.allclose(acq.last_outputs["values"].cpu(), want["values"], rtol=1e-6, atol=1e-10)


@pytest.mark.parametrize("q,pending", [(1, 0), (2, 0), (2, 1)])
def test_cnehvi_matches_oracle(q, pending):
    """qConformalNoisyExpectedHypervolumeImprovement (mobo_acquisitions.py:184-219): joint samples with a pruned baseline under
    every refit model, per-sample box decompositions; against oracle/mobo.py (dense == closed form on the CPU)."""
    from conformal_bayesopt_b200 import GPModelState
    from conformal_bayesopt_b200.mobo_acquisitions import qConformalNoisyExpectedHypervolumeImprovement
    from conformal_bayesopt_b200.sampling import IIDNormalSampler
    from oracle import mobo
    dev = _dev()
    N, d, b, G, m, S, nb = 16, 2, 30, 32, 2, 3, 6
    g = torch.Generator().manual_seed(9)
    X = torch.rand(N, d, generator=g, dtype=torch.float64)
    Y = torch.stack([torch.sin(3 * X.sum(-1)), torch.cos(2 * X[:, 0]) - X[:, 1]], -1)
    hyp = [(0.0, [0.4 + 0.1 * k] * d, 1.0 + 0.2 * k, 1e-2 * (k + 1)) for k in range(m)]
    gps = [ogp.build_state(X, Y[:, k], *hyp[k]) for k in range(m)]
    states = [GPModelState.fit_fixed(X, Y[:, k], *hyp[k], device=dev) for k in range(m)]
    Xb = torch.cat([X[:nb - 1], torch.rand(1, d, generator=g, dtype=torch.float64)])
    Xq = torch.rand(b, q, d, generator=g, dtype=torch.float64)
    if pending:
        Xq[:, q - pending:] = Xq[0, q - pending:]
    eps_grid = torch.randn(G, q, m, generator=g, dtype=torch.float64)
    eps_b = torch.randn(S, nb, m, generator=g, dtype=torch.float64)
    eps_new = torch.randn(S, G, q, m, generator=g, dtype=torch.float64)
    eps_mc = torch.cat([eps_b.unsqueeze(1).expand(S, G, nb, m), eps_new], dim=2)
    z = 2.0 * ((torch.rand(N + q, 8, generator=g) > 0.5).double() - 0.5)
    u = torch.rand(b, G, q, m, generator=g, dtype=torch.float64)
    ref = torch.tensor([-1.5, -1.5], dtype=torch.float64)
    lo, up, fb = mobo.nehvi_cells(gps, Xb, eps_b, ref)
    gs = IIDNormalSampler(G)
    gs.set_base_samples(eps_grid)
    mc = IIDNormalSampler(S, batch_range=(0, -3))
    mc.set_base_samples(eps_new)
    acq = qConformalNoisyExpectedHypervolumeImprovement(alpha=0.2, temp=3e-2, grid_res=G, max_grid_refinements=0,
                                                        ratio_estimator=None, optimistic=True, grid_sampler=gs, randomized=True,
                                                        model=states, ref_point=ref, X_baseline=Xb.to(dev), sampler=mc,
                                                        prune_baseline=False, baseline_base_samples=eps_b,
                                                        X_pending=Xq[0, q - pending:].to(dev) if pending else None)
    assert torch.allclose(acq.baseline_samples, fb, rtol=1e-9, atol=1e-11)
    assert torch.allclose(acq.cell_lower.cpu(), lo, rtol=1e-9, atol=1e-11) and torch.equal(torch.isinf(acq.cell_upper.cpu()), torch.isinf(up))
    acq.fixed_probes, acq.fixed_uniforms = z, u.to(dev)
    with torch.no_grad():
        got = acq(Xq[:, :q - pending].to(dev))
    want = mobo.cehvi(gps, Xq, alpha=0.2, temp=3e-2, grid_res=G, eps_grid=eps_grid, z=z, eps_mc=eps_mc, cell_lower=lo,
                      cell_upper=up, randomized=True, u_rand=u, optimistic=True, X_baseline=Xb)
    assert got.shape == (b,) and (want["acq"] > 0).any()
    assert torch.allclose(acq.last_outputs["values"].cpu(), want["values"], rtol=1e-6, atol=1e-10)
    assert torch.allclose(got.cpu(), want["acq"], rtol=1e-6, atol=1e-10)


@pytest.mark.parametrize("temp,refine,q", [(3e-2, 0, 1), (1e-2, 2, 1), (3e-2, 0, 2), (1e-2, 2, 2), (3e-2, 0, 3), (3e-2, 1, 4),
                                           (0.0, 0, 2)])
def test_cnei_matches_oracle(temp, refine, q):
    """qConformalNoisyExpectedImprovement (acquisitions.py:106-139) with a pruned baseline, q = 1 and joint batches, against
    the closed-form oracle (which the dense literal restatement pins, tests/test_oracle_consistency.py)."""
    from conformal_bayesopt_b200.acquisitions import qConformalNoisyExpectedImprovement
    from conformal_bayesopt_b200.sampling import IIDNormalSampler
    dev = _dev()
    N, d, b, G, nb, S = 90, 4, 21, 32, 7, 2
    p = make_problem(N, d, seed=13)
    inp = make_inputs(N, d, b, q, G, S=S, seed=13)
    g = torch.Generator().manual_seed(4)
    Xb = torch.cat([p["X"][torch.randperm(N, generator=g)[:nb - 1]], torch.rand(1, d, generator=g, dtype=torch.float64)])
    eps_mc = torch.randn(S, G, nb + q, generator=g, dtype=torch.float64)
    ratio, est = _estimator(d, dev)
    gs = IIDNormalSampler(G)
    gs.set_base_samples(inp["eps_grid"])
    mc = IIDNormalSampler(S, batch_range=(0, -3))
    mc.set_base_samples(eps_mc)
    acq = qConformalNoisyExpectedImprovement(alpha=0.15, temp=temp, grid_res=G, max_grid_refinements=refine,
                                             ratio_estimator=est, grid_sampler=gs, randomized=True, model=_state(p, dev),
                                             X_baseline=Xb.to(dev), sampler=mc, prune_baseline=False)
    acq.fixed_probes, acq.fixed_uniforms = inp["z"], inp["u_rand"].to(dev)
    with torch.no_grad():
        got = acq(inp["Xq"].to(dev))
    want = closed_form.conformal_acquisition(oracle_state(p), inp["Xq"], kind="nei", alpha=0.15, temp=temp, grid_res=G,
                                             max_grid_refinements=refine, eps_grid=inp["eps_grid"], z=inp["z"], eps_mc=eps_mc,
                                             ratio=ratio, randomized=True, u_rand=inp["u_rand"], optimistic=False, X_baseline=Xb)
    assert (want["acq"] > 0).any()
    assert torch.allclose(got.cpu(), want["acq"], rtol=1e-6, atol=1e-10)
    with pytest.raises(NotImplementedError):
        qConformalNoisyExpectedImprovement(alpha=0.15, temp=temp, grid_res=G, max_grid_refinements=0, ratio_estimator=None,
                                           model=_state(p, dev), X_baseline=p["X"].to(dev))      # 90 baseline points: prune first


@pytest.mark.parametrize("N,d,noise", [(5, 2, 1e-2), (128, 3, 1e-2), (200, 8, 1e-4), (1024, 8, 1e-2), (1100, 6, 1e-3)])
def test_model_build_matches_torch_linalg(N, d, noise):
    """cbo_model_build_f64 (blocked Cholesky + triangular inverse on the DMMA GEMM) against torch.linalg on the CPU."""
    from oracle import gp as ogp
    dev = _dev()
    p = make_problem(N, d, seed=N, noise=noise)
    st = _state(p, dev)
    gp = ogp.build_state(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"])
    linv_ref = torch.linalg.solve_triangular(gp.L, torch.eye(N, dtype=torch.float64), upper=False)
    linv = st.linv.cpu()
    assert st.Np % 128 == 0 and st.Np >= N
    scale = linv_ref.abs().max()
    assert torch.allclose(linv[:N, :N], linv_ref, rtol=1e-8, atol=float(1e-11 * scale))
    assert torch.equal(st.linv_t.cpu(), linv.T)
    assert torch.all(linv[N:, :] == 0) and torch.all(linv[:, N:] == 0) and torch.all(torch.triu(linv, 1) == 0)
    assert torch.allclose(st.alpha.cpu()[:N], gp.alpha, rtol=1e-7, atol=float(1e-10 * gp.alpha.abs().max()))
    assert torch.all(st.alpha.cpu()[N:] == 0)
    # A^-1 = L^-T L^-1 reproduces the identity against the kernel matrix
    A = gp.k(gp.X, gp.X) + noise * torch.eye(N, dtype=torch.float64)
    resid = (linv[:N, :N].T @ (linv[:N, :N] @ A) - torch.eye(N, dtype=torch.float64)).abs().max()
    assert resid < 1e-7


def test_model_build_reports_non_positive_definite():
    from conformal_bayesopt_b200 import GPModelState, _lib
    dev = _dev()
    X = torch.rand(40, 2, dtype=torch.float64)
    X[7] = X[3]                                            # duplicate point and (almost) no noise: singular to rounding
    with pytest.raises(_lib.CboError, match="positive definite|invalid"):
        GPModelState.fit_fixed(X, torch.zeros(40, dtype=torch.float64), 0.0, 0.5, 1.0, -1e-3, device=dev)


def test_bo_loop_demo_runs_on_the_drop_in_classes():
    """tools/bo_demo.py: model build -> raw scoring -> SGLD steps with grad (q = 2) -> re-scoring, two rounds."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bo_demo", os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools", "bo_demo.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    best = mod.run(rounds=2, sgld_steps=6, raw=256, verbose=False)
    assert len(best) == 3 and all(math.isfinite(v) for v in best) and best[-1] >= best[0]


@pytest.mark.parametrize("acq", ["cei", "cnei", "cehvi", "cnehvi"])
def test_bo_loop_demo_runs_every_acquisition_through_the_sgld_optimiser(acq):
    """conformalbo/utils.py:243-246 sends every acquisition with a ratio estimator to optimize_acqf_sgld: raw scoring, SGLD
    steps on d acq / d X (cNEI: graph replay; cEHVI / cNEHVI: 2 m + 1 library calls per step), estimator training, re-scoring."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bo_demo", os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools", "bo_demo.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    vals = mod.run_family(acq, rounds=2, sgld_steps=6, raw=64, restarts=3, verbose=False)
    assert len(vals) == 2 and all(math.isfinite(v) for v in vals)


def test_sharded_grid_refinement_matches_the_single_batch():
    """SURVEY.md §8e exception: with max_grid_refinements > 0 the reference's rescale / best-pass decisions are batch-global.
    Two 'ranks' (threads on their own streams, distinct workspaces) score the two halves of a batch; the refine_reduce hook
    combines their statistics per refinement pass -> same results as one call on the whole batch."""
    import threading
    from conformal_bayesopt_b200 import ops
    dev = _dev()
    N, d, b, G = 64, 2, 40, 32
    p = make_problem(N, d, seed=64)
    inp = make_inputs(N, d, b, 1, G, S=1, seed=64)
    st = _state(p, dev)
    kw = dict(kind="ei", alpha=0.1, temp=1e-6, grid_res=G, randomized=True, optimistic=True, best_f=0.3, max_grid_refinements=3)
    args = dict(eps_grid=inp["eps_grid"][:, 0].to(dev), z=inp["z"].to(dev), eps_mc=inp["eps_mc"][:, :, 0].to(dev))
    full = ops.conformal_acq(st, inp["Xq"].to(dev), u_rand=inp["u_rand"][:, :, 0].to(dev), outputs=("conf_pred_mask",), **args, **kw)
    torch.cuda.synchronize()
    half = b // 2
    barrier = threading.Barrier(2)
    slots, lock, res, calls = {}, threading.Lock(), {}, []

    def reducer_for(rank):
        def reduce(stats):
            with lock:
                slots[rank] = stats
            barrier.wait()
            a, c = slots[0], slots[1]
            out = [min(a[0], c[0]), max(a[1], c[1]), a[2] + c[2], a[3] + c[3]]
            barrier.wait()
            calls.append(rank)
            return out
        return reduce

    def worker(rank):
        sl = slice(0, half) if rank == 0 else slice(half, b)
        with torch.cuda.stream(torch.cuda.Stream(device=dev)):
            o = ops.conformal_acq(st, inp["Xq"][sl].to(dev), u_rand=inp["u_rand"][sl, :, 0].to(dev), outputs=("conf_pred_mask",),
                                  refine_reduce=reducer_for(rank), **args, **kw)
            torch.cuda.synchronize()
            res[rank] = {k: v.cpu() for k, v in o.items()}

    ts = [threading.Thread(target=worker, args=(r,)) for r in (0, 1)]
    [t.start() for t in ts]
    [t.join(timeout=120) for t in ts]
    assert len(res) == 2 and len(calls) >= 4          # at least two refinement passes on both ranks
    got_acq = torch.cat([res[0]["acq"], res[1]["acq"]])
    got_mask = torch.cat([res[0]["conf_pred_mask"], res[1]["conf_pred_mask"]])
    assert torch.allclose(got_acq, full["acq"].cpu(), rtol=1e-12, atol=1e-14)
    assert torch.allclose(got_mask, full["conf_pred_mask"].cpu(), rtol=0, atol=1e-12)
    # without the hook the two halves refine on their own statistics (documented per-shard semantics): generally different


@pytest.mark.gpu
@pytest.mark.parametrize("N,q", [(100, 2), (126, 4), (128, 1), (1000, 3), (1023, 2)])
def test_model_append_matches_a_rebuild(N, q):
    """cbo_model_append_f64 (rank-1 extensions, incl. growth across a 128-row padding boundary) against cbo_model_build_f64 on
    the concatenated data: L^-1, alpha, and an acquisition call on top."""
    from conformal_bayesopt_b200 import GPModelState, ops
    from conformal_bayesopt_b200.synthetic import make_inputs, make_problem
    d = 5
    p = make_problem(N + q, d, seed=N)
    dev = torch.device("cuda:0")
    kw = dict(mean=p["mean"], lengthscale=p["lengthscale"], outputscale=p["outputscale"], noise=p["noise"], device=dev)
    full = GPModelState.fit_fixed(p["X"], p["y"], **kw)
    base = GPModelState.fit_fixed(p["X"][:N], p["y"][:N], **kw)
    keep = base.linv.clone()
    app = base.append(p["X"][N:], p["y"][N:])
    assert torch.equal(base.linv, keep) and base.N == N                      # the old state is untouched
    assert app.N == N + q and app.Np == full.Np
    scale = full.linv.abs().max()
    assert float((app.linv - full.linv).abs().max() / scale) < 1e-10
    assert float((app.linv_t - full.linv_t).abs().max() / scale) < 1e-10
    assert torch.equal(app.linv_t, app.linv.T)
    assert float((app.alpha - full.alpha).abs().max() / full.alpha.abs().max()) < 1e-9
    assert torch.equal(app.X, full.X) and torch.equal(app.y, full.y)
    inp = make_inputs(N + q, d, 16, 1, 32, seed=1)
    call = dict(eps_grid=inp["eps_grid"][:, 0].to(dev), z=inp["z"].to(dev), eps_mc=inp["eps_mc"][:, :, 0].to(dev), kind="ucb",
                alpha=0.1, temp=3e-2, grid_res=32, optimistic=True, beta_prime=1.0)
    a = ops.conformal_acq(app, inp["Xq"].to(dev), **call)["acq"]
    b = ops.conformal_acq(full, inp["Xq"].to(dev), **call)["acq"]
    assert float((a - b).abs().max() / b.abs().max()) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("temp,ratio_on", [(3e-2, False), (3e-2, True), (1e-5, True)])
def test_pair_and_warp_specialised_score_kernels_are_bit_identical(monkeypatch, temp, ratio_on):
    """k_score (warp pairs) and k_score_ws (warp-specialised; picked by batch size) share ONE soft-rank sweep
    (score_rows_fx, also its clamped variant at tiny temp): forcing either kernel on the same call gives the same bits,
    so the batch-size switch can never move a result.  Both are checked against the closed-form oracle."""
    from conformal_bayesopt_b200 import ops
    dev = _dev()
    N, d, b, G = 200, 5, 150, 48
    p = make_problem(N, d, seed=21)
    inp = make_inputs(N, d, b, 1, G, S=1, seed=21)
    ratio = ogp.make_ratio_mlp(d) if ratio_on else None
    rw = ops.RatioWeights(*[t.to(dev) for t in (ratio.W1, ratio.b1, ratio.W2, ratio.b2, ratio.W3, ratio.b3)]) if ratio_on else None
    kw = dict(kind="ucb", alpha=0.1, temp=temp, grid_res=G, randomized=True, optimistic=True, beta_prime=1.0)
    outs = ("conf_pred_mask", "num_accepted", "ood_mask")
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("CBO_SCORE_WS", mode)
        res[mode] = ops.conformal_acq(_state(p, dev), inp["Xq"].to(dev), eps_grid=inp["eps_grid"][:, 0].to(dev),
                                      z=inp["z"].to(dev), eps_mc=inp["eps_mc"][:, :, 0].to(dev),
                                      u_rand=inp["u_rand"][:, :, 0].to(dev), ratio=rw, outputs=outs, **kw)
    for k in ("acq",) + outs:
        assert torch.equal(res["0"][k], res["1"][k]), k
    want = closed_form.conformal_acquisition(oracle_state(p), inp["Xq"], eps_grid=inp["eps_grid"], z=inp["z"],
                                             eps_mc=inp["eps_mc"], u_rand=inp["u_rand"], ratio=ratio, **kw)
    assert torch.allclose(res["1"]["acq"].cpu(), want["acq"], rtol=1e-6, atol=1e-9)
    assert torch.allclose(res["1"]["conf_pred_mask"].cpu(), want["mask"][:, :, 0], rtol=0, atol=1e-6)


@pytest.mark.parametrize("q,ratio_on,kind", [(1, True, "ucb"), (2, True, "ucb"), (2, False, "ei"), (1, True, "nei")])
def test_latency_path_branches_give_the_bits_of_the_single_stream_path(monkeypatch, q, ratio_on, kind):
    """A call with <= 32 candidate points forks into three stream branches (probe solves | kernel rows and candidate solves |
    ratio estimator; csrc/cbo_api.cu SideStream) and joins them through events.  The branches only reorder independent work:
    values and gradient rows must equal those of the same call issued on one stream (CBO_SIDE_STREAM=0), bit for bit, also
    when the call is repeated back to back (buffer reuse across calls) and replayed as a CUDA graph."""
    from conformal_bayesopt_b200 import ops
    dev = _dev()
    N, d, b, G = 150, 5, 8, 32
    p = make_problem(N, d, seed=31)
    inp = make_inputs(N, d, b, q, G, S=2, seed=31)
    ratio = ogp.make_ratio_mlp(d) if ratio_on else None
    rw = ops.RatioWeights(*[t.to(dev) for t in (ratio.W1, ratio.b1, ratio.W2, ratio.b2, ratio.W3, ratio.b3)]) if ratio_on else None
    kw = dict(kind=kind, alpha=0.1, temp=3e-2, grid_res=G, randomized=True, optimistic=True, beta_prime=1.0, best_f=0.3)
    st = _state(p, dev)
    if kind == "nei":
        kw["nei_baseline"] = ops.NeiBaseline.build(st, p["X"][:6].to(dev))
    eg = inp["eps_grid"].to(dev) if q > 1 else inp["eps_grid"][:, 0].to(dev)
    em = inp["eps_mc"].to(dev) if q > 1 else inp["eps_mc"][:, :, 0].to(dev)
    u = inp["u_rand"].to(dev) if q > 1 else inp["u_rand"][:, :, 0].to(dev)
    if kind == "nei":
        em = torch.randn(2, G, 6 + q, dtype=torch.float64, generator=torch.Generator().manual_seed(3)).to(dev)

    Xd, zd = inp["Xq"].to(dev), inp["z"].to(dev)

    def call():
        return ops.conformal_acq(st, Xd, eps_grid=eg, z=zd, eps_mc=em, u_rand=u, ratio=rw, want_grad=True, **kw)
    monkeypatch.setenv("CBO_SIDE_STREAM", "0")
    ref = call()
    monkeypatch.setenv("CBO_SIDE_STREAM", "1")
    for _ in range(3):
        got = call()
        torch.cuda.synchronize()
        assert torch.equal(got["acq"], ref["acq"]) and torch.equal(got["dacq_dX"], ref["dacq_dX"])
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):                 # every input is a device tensor made before the capture
        out = call()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out["acq"], ref["acq"]) and torch.equal(out["dacq_dX"], ref["dacq_dX"])
