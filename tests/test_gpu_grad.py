"""GPU parity of d acq / d X: cbo_conformal_acq_grad_f64 against torch.autograd through the closed-form
oracle and against the hand-derived reverse pass (oracle/analytic_grad.py), stage by stage.

Tolerance: 1e-6 relative (north_star) on the gradient rows, with an absolute floor of 1e-9 times the largest entry
(the soft-rank sigmoid is evaluated to ~3e-10 absolute in the kernels).
"""
import math

import pytest
import torch

from oracle import analytic_grad, closed_form
from oracle import gp as ogp
from tests.problems import make_inputs, make_problem, oracle_state

pytestmark = pytest.mark.gpu


def _dev():
    assert torch.cuda.is_available(), "run with -m gpu on a CUDA box"
    return torch.device("cuda:0")


def _state(p, dev):
    from conformal_bayesopt_b200 import GPModelState
    return GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)


def _ratio(d, dev):
    from conformal_bayesopt_b200 import ops
    r = ogp.make_ratio_mlp(d)
    return r, ops.RatioWeights(*[t.to(dev) for t in (r.W1, r.b1, r.W2, r.b2, r.W3, r.b3)])


def _oracle_grad(gp, inp, ratio, **kw):
    X = inp["Xq"].clone().requires_grad_(True)
    out = closed_form.conformal_acquisition(gp, X, eps_grid=inp["eps_grid"], z=inp["z"], eps_mc=inp["eps_mc"], ratio=ratio,
                                            u_rand=inp["u_rand"], **kw)
    (g,) = torch.autograd.grad(out["acq"].sum(), X)
    return out["acq"].detach(), g


def _gpu(p, inp, dev, ratio_w, chunk=0, outputs=(), **kw):
    from conformal_bayesopt_b200 import ops
    out = ops.conformal_acq(_state(p, dev), inp["Xq"].to(dev), eps_grid=inp["eps_grid"][:, 0].to(dev), z=inp["z"].to(dev),
                            eps_mc=inp["eps_mc"][:, :, 0].to(dev), u_rand=inp["u_rand"][:, :, 0].to(dev), ratio=ratio_w,
                            want_grad=True, chunk=chunk, outputs=outputs, **kw)
    torch.cuda.synchronize()
    return {k: v.cpu() for k, v in out.items()}


def _grad_close(got, want):
    scale = want.abs().max().clamp_min(1e-300)
    return torch.allclose(got, want, rtol=1e-6, atol=float(1e-9 * scale))


CASES = [
    # N, d, b, G, temp, randomized, ratio, kind, S, optimistic
    (40, 3, 6, 16, 3e-2, False, False, "ucb", 1, True),
    (200, 8, 37, 64, 3e-2, True, True, "ucb", 1, True),
    (130, 4, 9, 32, 5e-2, True, True, "ei", 2, False),
    (257, 8, 5, 7, 1e-2, False, True, "ei", 4, True),         # ragged N and G
    (300, 8, 50, 64, 3e-2, True, False, "ucb", 3, True),
    (8, 2, 128, 64, 3e-2, True, False, "ucb", 1, True),       # BASELINE config 1 shape
    (32, 12, 224, 64, 3e-2, False, True, "ucb", 1, True),     # BASELINE config 4 shape, d > 8 (two register chunks)
    (128, 3, 1, 2, 3e-2, False, False, "ucb", 1, True),       # minimum grid, single candidate
    (130, 4, 9, 32, 0.0, True, True, "ucb", 2, True),         # hard path: step masks, only the value path carries gradient
    (64, 5, 700, 200, 3e-2, False, True, "ucb", 1, True),     # G > threads of one 4-warp group
    (1500, 8, 24, 32, 3e-2, True, True, "ucb", 1, True),      # Np = 1536: 12 GEMM tiles, 2 score groups per SM
]


@pytest.mark.parametrize("N,d,b,G,temp,rand,ratio_on,kind,S,opt", CASES)
def test_grad_matches_autograd_of_oracle(N, d, b, G, temp, rand, ratio_on, kind, S, opt):
    dev = _dev()
    p = make_problem(N, d, seed=N)
    inp = make_inputs(N, d, b, 1, G, S=S, seed=N)
    ratio, ratio_w = _ratio(d, dev) if ratio_on else (None, None)
    kw = dict(kind=kind, alpha=0.1, temp=temp, grid_res=G, randomized=rand, optimistic=opt,
              beta_prime=math.sqrt(0.2 * math.pi / 2), best_f=0.3)
    acq, g = _oracle_grad(oracle_state(p), inp, ratio, **kw)
    got = _gpu(p, inp, dev, ratio_w, **kw)
    assert torch.allclose(got["acq"], acq, rtol=1e-6, atol=1e-9)
    assert G == 2 or g.abs().max() > 0          # (two grid points: the normalised weights do not move)
    assert _grad_close(got["dacq_dX"], g)


def test_grad_call_value_equals_forward_call_and_chunking():
    from conformal_bayesopt_b200 import ops
    dev = _dev()
    N, d, b, G = 150, 6, 100, 33
    p = make_problem(N, d, seed=5)
    inp = make_inputs(N, d, b, 1, G, S=2, seed=5)
    _, ratio_w = _ratio(d, dev)
    kw = dict(kind="ucb", alpha=0.1, temp=3e-2, grid_res=G, randomized=True, optimistic=True, beta_prime=1.1)
    full = _gpu(p, inp, dev, ratio_w, outputs=("conf_pred_mask",), **kw)
    chunked = _gpu(p, inp, dev, ratio_w, chunk=32, **kw)
    fwd = ops.conformal_acq(_state(p, dev), inp["Xq"].to(dev), eps_grid=inp["eps_grid"][:, 0].to(dev), z=inp["z"].to(dev),
                            eps_mc=inp["eps_mc"][:, :, 0].to(dev), u_rand=inp["u_rand"][:, :, 0].to(dev), ratio=ratio_w,
                            outputs=("conf_pred_mask",), **kw)
    assert torch.allclose(full["acq"], fwd["acq"].cpu(), rtol=1e-12, atol=1e-14)
    assert torch.equal(full["conf_pred_mask"], fwd["conf_pred_mask"].cpu())
    assert torch.equal(full["dacq_dX"], chunked["dacq_dX"]) and torch.equal(full["acq"], chunked["acq"])


def test_grad_stage_adjoints_match_analytic_reverse_pass():
    """Same call, checked against the hand-derived reverse pass (identical algebra) -- tight tolerance."""
    dev = _dev()
    N, d, b, G = 96, 4, 12, 24
    p = make_problem(N, d, seed=9)
    inp = make_inputs(N, d, b, 1, G, S=1, seed=9)
    ratio, ratio_w = _ratio(d, dev)
    kw = dict(kind="ucb", alpha=0.15, temp=3e-2, grid_res=G, randomized=True, optimistic=True, beta_prime=1.0)
    want = analytic_grad.conformal_acquisition_grad(oracle_state(p), inp["Xq"], eps_grid=inp["eps_grid"], z=inp["z"],
                                                    eps_mc=inp["eps_mc"], ratio=ratio, u_rand=inp["u_rand"], **kw)
    got = _gpu(p, inp, dev, ratio_w, **kw)
    assert _grad_close(got["dacq_dX"], want["grad"])


def test_grad_with_grid_refinement_uses_the_kept_pass():
    dev = _dev()
    N, d, b, G = 64, 2, 20, 32
    p = make_problem(N, d, seed=64)
    inp = make_inputs(N, d, b, 1, G, S=1, seed=64)
    kw = dict(kind="ei", alpha=0.1, temp=3e-2, grid_res=G, randomized=False, optimistic=True, best_f=0.3,
              max_grid_refinements=2)
    acq, g = _oracle_grad(oracle_state(p), inp, None, **kw)
    got = _gpu(p, inp, dev, None, **kw)
    assert torch.allclose(got["acq"], acq, rtol=1e-6, atol=1e-9)
    assert _grad_close(got["dacq_dX"], g)


def test_acquisition_class_backward():
    """forward(X) with X.requires_grad -> .sum().backward(), the call pattern of optimize_acqf_sgld (optim/optimize.py:224-235)."""
    from conformal_bayesopt_b200.acquisitions import qConformalExpectedImprovement, qConformalUpperConfidenceBound
    from conformal_bayesopt_b200.sampling import IIDNormalSampler
    dev = _dev()
    N, d, b, G = 80, 3, 10, 16
    p = make_problem(N, d, seed=80)
    inp = make_inputs(N, d, b, 1, G, S=1, seed=80)
    state = _state(p, dev)
    for cls, extra, okw in ((qConformalUpperConfidenceBound, dict(beta=0.2), dict(kind="ucb", beta_prime=math.sqrt(0.2 * math.pi / 2))),
                            (qConformalExpectedImprovement, dict(best_f=0.3), dict(kind="ei", best_f=0.3))):
        gs, ms = IIDNormalSampler(G), IIDNormalSampler(1, batch_range=(0, -3))
        gs.set_base_samples(inp["eps_grid"]); ms.set_base_samples(inp["eps_mc"])
        acqf = cls(0.1, 3e-2, G, 0, None, optimistic=True, grid_sampler=gs, model=state, sampler=ms, **extra)
        acqf.fixed_probes = inp["z"]
        X = inp["Xq"].to(dev).requires_grad_(True)
        val = acqf(X)
        (2.0 * val).sum().backward()
        acq, g = _oracle_grad(oracle_state(p), inp, None, alpha=0.1, temp=3e-2, grid_res=G, optimistic=True, **okw)
        assert torch.allclose(val.detach().cpu(), acq, rtol=1e-6, atol=1e-9)
        assert _grad_close(X.grad.cpu(), 2.0 * g)
        with torch.no_grad():
            assert torch.allclose(acqf(X), val.detach(), rtol=1e-12)


QCASES = [
    # N, d, q, b, G, temp, randomized, ratio, kind, S
    (40, 3, 2, 6, 16, 3e-2, False, False, "ucb", 1),
    (120, 8, 2, 8, 64, 3e-2, True, True, "ucb", 1),          # the reference's default SGLD shape: q_batch_size 2, 8 restarts
    (33, 2, 2, 5, 12, 5e-2, True, True, "ei", 2),
    (60, 4, 3, 4, 10, 3e-2, True, True, "ucb", 3),
    (200, 12, 4, 3, 24, 3e-2, True, False, "ei", 2),         # d > 8
    (64, 3, 2, 4, 16, 0.0, True, True, "ucb", 2),            # hard path
    (150, 5, 2, 24, 16, 3e-2, True, True, "ucb", 1),         # > 18 joint batches: one CTA per joint batch (no clusters)
    (700, 6, 3, 5, 32, 3e-2, True, True, "ei", 1),           # cluster windows of 96 train points, last ranks partly empty
]


def _gpu_q(p, inp, dev, ratio_w, **kw):
    from conformal_bayesopt_b200 import ops
    out = ops.conformal_acq(_state(p, dev), inp["Xq"].to(dev), eps_grid=inp["eps_grid"].to(dev), z=inp["z"].to(dev),
                            eps_mc=inp["eps_mc"].to(dev), u_rand=inp["u_rand"].to(dev), ratio=ratio_w, want_grad=True, **kw)
    torch.cuda.synchronize()
    return {k: v.cpu() for k, v in out.items()}


@pytest.mark.parametrize("N,d,q,b,G,temp,rand,ratio_on,kind,S", QCASES)
def test_joint_batch_grad_matches_autograd_of_oracle(N, d, q, b, G, temp, rand, ratio_on, kind, S):
    dev = _dev()
    p = make_problem(N, d, seed=N + q)
    inp = make_inputs(N, d, b, q, G, S=S, seed=N + q)
    ratio, ratio_w = _ratio(d, dev) if ratio_on else (None, None)
    kw = dict(kind=kind, alpha=0.15, temp=temp, grid_res=G, randomized=rand, optimistic=True,
              beta_prime=math.sqrt(0.2 * math.pi / 2), best_f=0.3)
    acq, g = _oracle_grad(oracle_state(p), inp, ratio, **kw)
    got = _gpu_q(p, inp, dev, ratio_w, **kw)
    assert torch.allclose(got["acq"], acq, rtol=1e-6, atol=1e-9)
    assert g.abs().max() > 0
    assert _grad_close(got["dacq_dX"], g)


def test_cluster_and_single_cta_reverse_kernels_agree():
    """The same joint batches scored in a small call (thread-block clusters, train points split over 8 CTAs) and inside a
    large call (one CTA per joint batch): identical algebra, different summation order."""
    from conformal_bayesopt_b200 import ops
    dev = _dev()
    N, d, q, G = 300, 4, 2, 32
    p = make_problem(N, d, seed=5)
    inp = make_inputs(N, d, 40, q, G, S=1, seed=5)
    _, ratio_w = _ratio(d, dev)
    st = _state(p, dev)
    kw = dict(eps_grid=inp["eps_grid"].to(dev), z=inp["z"].to(dev), eps_mc=inp["eps_mc"].to(dev), ratio=ratio_w, want_grad=True,
              kind="ucb", alpha=0.1, temp=3e-2, grid_res=G, randomized=True, optimistic=True, beta_prime=1.0)
    X, u = inp["Xq"].to(dev), inp["u_rand"].to(dev)
    big = ops.conformal_acq(st, X, u_rand=u, **kw)
    small = ops.conformal_acq(st, X[:8].contiguous(), u_rand=u[:8].contiguous(), **kw)
    assert torch.allclose(small["acq"], big["acq"][:8], rtol=1e-11, atol=1e-13)
    scale = big["dacq_dX"][:8].abs().max()
    assert float((small["dacq_dX"] - big["dacq_dX"][:8]).abs().max() / scale) < 1e-9


def test_joint_batch_grad_with_grid_refinement():
    dev = _dev()
    N, d, q, b, G = 64, 2, 2, 10, 32
    p = make_problem(N, d, seed=64)
    inp = make_inputs(N, d, b, q, G, S=1, seed=65)
    kw = dict(kind="ucb", alpha=0.1, temp=1e-2, grid_res=G, randomized=False, optimistic=True, beta_prime=1.0,
              max_grid_refinements=2)
    acq, g = _oracle_grad(oracle_state(p), inp, None, **kw)
    got = _gpu_q(p, inp, dev, None, **kw)
    assert torch.allclose(got["acq"], acq, rtol=1e-6, atol=1e-9)
    assert _grad_close(got["dacq_dX"], g)


def test_joint_batch_grad_matches_hand_derived_reverse_pass():
    from oracle import analytic_grad_q
    dev = _dev()
    N, d, q, b, G = 48, 3, 2, 6, 12
    p = make_problem(N, d, seed=3)
    inp = make_inputs(N, d, b, q, G, S=2, seed=3)
    ratio, ratio_w = _ratio(d, dev)
    kw = dict(kind="ucb", alpha=0.2, temp=5e-2, grid_res=G, randomized=True, optimistic=True, beta_prime=1.2)
    want = analytic_grad_q.conformal_acquisition_grad_q(oracle_state(p), inp["Xq"], eps_grid=inp["eps_grid"], z=inp["z"],
                                                        eps_mc=inp["eps_mc"], ratio=ratio, u_rand=inp["u_rand"], **kw)
    got = _gpu_q(p, inp, dev, ratio_w, **kw)
    assert _grad_close(got["dacq_dX"], want["grad"])


def test_acquisition_class_backward_joint_batch_with_pending_point():
    """q = 2 restarts with grad (optimize_acqf_sgld's call) and a pending point appended by concatenate_pending_points:
    the gradient w.r.t. the optimised points only."""
    from conformal_bayesopt_b200.acquisitions import qConformalUpperConfidenceBound
    from conformal_bayesopt_b200.sampling import IIDNormalSampler
    dev = _dev()
    N, d, b, G = 60, 3, 8, 16
    p = make_problem(N, d, seed=60)
    inp = make_inputs(N, d, b, 3, G, S=1, seed=60)
    gs, ms = IIDNormalSampler(G), IIDNormalSampler(1, batch_range=(0, -3))
    gs.set_base_samples(inp["eps_grid"]); ms.set_base_samples(inp["eps_mc"])
    acqf = qConformalUpperConfidenceBound(0.1, 3e-2, G, 0, None, optimistic=True, grid_sampler=gs, model=_state(p, dev),
                                          sampler=ms, beta=0.2)
    acqf.fixed_probes = inp["z"]
    pend = inp["Xq"][0, 2:].clone()
    acqf.set_X_pending(pend.to(dev))
    X = inp["Xq"][:, :2].to(dev).requires_grad_(True)
    val = acqf(X)
    val.sum().backward()
    Xo = inp["Xq"][:, :2].clone().requires_grad_(True)
    Xfull = torch.cat([Xo, pend.expand(b, 1, d)], dim=1)
    out = closed_form.conformal_acquisition(oracle_state(p), Xfull, kind="ucb", alpha=0.1, temp=3e-2, grid_res=G,
                                            eps_grid=inp["eps_grid"], z=inp["z"], eps_mc=inp["eps_mc"], optimistic=True,
                                            beta_prime=math.sqrt(0.2 * math.pi / 2))
    (g,) = torch.autograd.grad(out["acq"].sum(), Xo)
    assert torch.allclose(val.detach().cpu(), out["acq"].detach(), rtol=1e-6, atol=1e-9)
    assert X.grad.shape == (b, 2, d) and _grad_close(X.grad.cpu(), g)


def test_grad_refuses_what_is_not_built():
    from conformal_bayesopt_b200 import _lib
    dev = _dev()
    p = make_problem(40, 3, seed=1)
    inp = make_inputs(40, 3, 4, 5, 8, seed=1)
    from conformal_bayesopt_b200 import ops
    with pytest.raises(_lib.CboError, match="unsupported"):
        ops.conformal_acq(_state(p, dev), inp["Xq"].to(dev), eps_grid=inp["eps_grid"].to(dev), z=inp["z"].to(dev),
                          eps_mc=inp["eps_mc"].to(dev), want_grad=True, kind="ucb", alpha=0.1, temp=3e-2, grid_res=8,
                          beta_prime=1.0)


def test_ratio_mlp_tensor_path_edge_shapes():
    """k_ratio_mlp_mma / _bwd with hidden sizes and input dimensions that are not multiples of the DMMA tile (zero padding)."""
    from conformal_bayesopt_b200 import ops
    dev = _dev()
    for d, H, n in ((20, 16, 37), (3, 40, 5), (64, 64, 130), (1, 8, 1)):
        r = ogp.make_ratio_mlp(d, seed=d + H, hidden=H)
        rw = ops.RatioWeights(*[t.to(dev) for t in (r.W1, r.b1, r.W2, r.b2, r.W3, r.b3)])
        x = torch.randn(n, d, dtype=torch.float64, generator=torch.Generator().manual_seed(n))
        got = ops.ratio_forward(rw, x.to(dev)).cpu()
        assert torch.allclose(got, r(x), rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("N,d,b,G,Sp,H", [(20, 2, 3, 6, 16, 24), (2, 1, 2, 4, 8, 8), (31, 20, 40, 33, 5, 64)])
def test_grad_edge_shapes(N, d, b, G, Sp, H):
    """Few train points, 16 / odd numbers of Hutchinson probes, d not a multiple of 4 or 8, small ratio network."""
    from conformal_bayesopt_b200 import ops
    from conformal_bayesopt_b200.synthetic import make_inputs as mk
    dev = _dev()
    p = make_problem(N, d, seed=N + d)
    inp = mk(N, d, b, 1, G, S=2, Sp=Sp, seed=N)
    r = ogp.make_ratio_mlp(d, seed=3, hidden=H)
    rw = ops.RatioWeights(*[t.to(dev) for t in (r.W1, r.b1, r.W2, r.b2, r.W3, r.b3)])
    kw = dict(kind="ucb", alpha=0.2, temp=5e-2, grid_res=G, randomized=True, optimistic=True, beta_prime=1.1)
    acq, g = _oracle_grad(oracle_state(p), inp, r, **kw)
    got = _gpu(p, inp, dev, rw, **kw)
    assert torch.allclose(got["acq"], acq, rtol=1e-6, atol=1e-9)
    assert _grad_close(got["dacq_dX"], g)


@pytest.mark.parametrize("q,b", [(1, 8), (1, 3000), (2, 8), (2, 300), (4, 40)])
def test_repeated_calls_are_bitwise_identical(q, b):
    """No atomics, fixed reduction orders: the same call gives the same bits (values and gradient rows), also across the
    latency path (<= 32 candidate points: split-K GEMM, CTA-per-point kernels) and the throughput path."""
    from conformal_bayesopt_b200 import ops
    dev = _dev()
    N, d, G = 300, 5, 48
    p = make_problem(N, d, seed=1)
    inp = make_inputs(N, d, b, q, G, S=2, seed=2)
    _, rw = _ratio(d, dev)
    st = _state(p, dev)
    kw = dict(kind="ucb", alpha=0.1, temp=3e-2, grid_res=G, randomized=True, optimistic=True, beta_prime=1.0)
    outs = []
    for _ in range(3):
        o = ops.conformal_acq(st, inp["Xq"].to(dev), eps_grid=inp["eps_grid"].to(dev), z=inp["z"].to(dev),
                              eps_mc=inp["eps_mc"].to(dev), u_rand=inp["u_rand"].to(dev), ratio=rw, want_grad=True, **kw)
        torch.cuda.synchronize()
        outs.append((o["acq"].cpu().clone(), o["dacq_dX"].cpu().clone()))
    for a, g in outs[1:]:
        assert torch.equal(a, outs[0][0]) and torch.equal(g, outs[0][1])
    assert torch.isfinite(outs[0][1]).all()


def test_cuda_graph_replay_equals_direct_calls():
    """graphs.GraphedConformalAcq: the captured value+gradient call replays bit-identically to direct calls (SGLD shape)."""
    from conformal_bayesopt_b200 import ops
    from conformal_bayesopt_b200.graphs import GraphedConformalAcq
    dev = _dev()
    N, d, b, q, G = 200, 4, 8, 2, 32
    p = make_problem(N, d, seed=7)
    st = _state(p, dev)
    _, rw = _ratio(d, dev)
    inp = make_inputs(N, d, b, q, G, S=1, seed=7)
    kw = dict(kind="ucb", alpha=0.1, temp=3e-2, grid_res=G, randomized=True, optimistic=True, beta_prime=1.0)
    gr = GraphedConformalAcq(st, b, q, eps_grid=inp["eps_grid"].to(dev), eps_mc=inp["eps_mc"].to(dev), ratio=rw, **kw)
    for t in range(3):
        it = make_inputs(N, d, b, q, G, S=1, seed=20 + t)
        X, z, u = it["Xq"].to(dev), it["z"].to(dev), it["u_rand"].to(dev)
        got = gr(X, z=z, u_rand=u)
        want = ops.conformal_acq(st, X, eps_grid=inp["eps_grid"].to(dev), z=z, eps_mc=inp["eps_mc"].to(dev), u_rand=u, ratio=rw,
                                 want_grad=True, **kw)
        torch.cuda.synchronize()
        assert torch.equal(got["acq"], want["acq"]) and torch.equal(got["dacq_dX"], want["dacq_dX"])
    with pytest.raises(ValueError):
        GraphedConformalAcq(st, b, q, eps_grid=inp["eps_grid"].to(dev), eps_mc=inp["eps_mc"].to(dev), max_grid_refinements=1, **kw)
