"""GPU tests of the torch C++ extension (torch.ops.cbo.*, csrc/torch_ext.cpp): every operator against the ctypes binding of
the same C entry point (bit-identical: both are one call into libcbo) and, for the differentiable operator, against
autograd through the oracle."""
import math

import pytest
import torch

from oracle import closed_form
from tests.problems import make_inputs, make_problem, oracle_state

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _state(p):
    from conformal_bayesopt_b200 import GPModelState
    return GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=DEV)


@pytest.mark.parametrize("kind,q,use_ratio", [("ucb", 1, False), ("ei", 1, True), ("ucb", 2, True)])
def test_conformal_acq_operator_matches_ctypes_and_oracle_autograd(kind, q, use_ratio):
    from conformal_bayesopt_b200 import ops, torch_ops
    from conformal_bayesopt_b200.synthetic import make_ratio_weights
    N, d, b, G = 96, 5, 9, 32
    p = make_problem(N, d, seed=21)
    inp = make_inputs(N, d, b, q, G, seed=21)
    st = _state(p)
    ratio = ops.RatioWeights(*[t.to(DEV) for t in make_ratio_weights(d)]) if use_ratio else None
    kw = dict(kind=kind, alpha=0.1, temp=3e-2, grid_res=G, randomized=True, optimistic=True, beta_prime=1.1, best_f=0.2)
    dv = dict(eps_grid=(inp["eps_grid"] if q > 1 else inp["eps_grid"][:, 0]).to(DEV), z=inp["z"].to(DEV),
              eps_mc=(inp["eps_mc"] if q > 1 else inp["eps_mc"][:, :, 0]).to(DEV),
              u_rand=(inp["u_rand"] if q > 1 else inp["u_rand"][:, :, 0]).to(DEV))
    ref = ops.conformal_acq(st, inp["Xq"].to(DEV), ratio=ratio, want_grad=True, **dv, **kw)
    X = inp["Xq"].to(DEV).requires_grad_(True)
    acq = torch_ops.conformal_acq(st, X, ratio=ratio, **dv, **kw)
    assert acq.requires_grad and torch.equal(acq.detach(), ref["acq"])
    wts = torch.linspace(0.5, 1.5, b, dtype=torch.float64, device=DEV)
    (g,) = torch.autograd.grad((acq * wts).sum(), X)
    assert torch.equal(g, ref["dacq_dX"] * wts.reshape(-1, 1, 1))
    with torch.no_grad():        # value-only call: the forward kernels (the value + reverse kernels round differently in the last bits)
        plain = ops.conformal_acq(st, inp["Xq"].to(DEV), ratio=ratio, **dv, **kw)["acq"]
        assert torch.equal(torch_ops.conformal_acq(st, inp["Xq"].to(DEV), ratio=ratio, **dv, **kw), plain)
        assert torch.allclose(plain, ref["acq"], rtol=1e-10, atol=1e-12)
    # against autograd through the CPU oracle (tolerance 1e-6 relative, as everywhere)
    from oracle import gp as ogp
    orat = ogp.make_ratio_mlp(d) if use_ratio else None
    Xo = inp["Xq"].clone().requires_grad_(True)
    want = closed_form.conformal_acquisition(oracle_state(p), Xo, eps_grid=inp["eps_grid"], z=inp["z"], eps_mc=inp["eps_mc"],
                                             u_rand=inp["u_rand"], ratio=orat, **kw)
    (go,) = torch.autograd.grad((want["acq"] * wts.cpu()).sum(), Xo)
    assert torch.allclose(acq.detach().cpu(), want["acq"].detach(), rtol=1e-6, atol=1e-9)
    assert (g.cpu() - go).abs().max() <= 1e-6 * go.abs().max()


def test_acquisition_classes_use_the_operator_and_agree_with_the_ctypes_path():
    import conformal_bayesopt_b200 as cbo
    from conformal_bayesopt_b200.sampling import IIDNormalSampler
    N, d, b, G = 80, 4, 7, 32
    p = make_problem(N, d, seed=22)
    inp = make_inputs(N, d, b, 2, G, seed=22)
    st = _state(p)
    res = {}
    for use in (True, False):
        gs, s = IIDNormalSampler(G), IIDNormalSampler(1, batch_range=(0, -3))
        gs.set_base_samples(inp["eps_grid"])
        s.set_base_samples(inp["eps_mc"])
        fn = cbo.qConformalUpperConfidenceBound(alpha=0.1, temp=3e-2, grid_res=G, max_grid_refinements=0, ratio_estimator=None,
                                                grid_sampler=gs, randomized=True, model=st, beta=0.2, sampler=s,
                                                X_pending=inp["Xq"][0, 1:].to(DEV))
        fn.use_torch_ops = use
        fn.fixed_probes = inp["z"]
        fn.fixed_uniforms = inp["u_rand"].to(DEV)
        X = inp["Xq"][:, :1].to(DEV).requires_grad_(True)        # q = 1 + one pending point
        acq = fn(X)
        acq.sum().backward()
        res[use] = (acq.detach(), X.grad.clone())
    assert torch.equal(res[True][0], res[False][0]) and torch.equal(res[True][1], res[False][1])
    assert res[True][1].shape == (b, 1, d)


def test_model_build_bounds_ratio_and_training_operators_match_ctypes():
    from conformal_bayesopt_b200 import ops, torch_ops
    from conformal_bayesopt_b200.synthetic import make_ratio_weights
    t = torch_ops.load()
    N, d = 150, 6
    p = make_problem(N, d, seed=23)
    st = _state(p)
    alpha, linv, linv_t, info = t.model_build(p["X"].to(DEV), p["y"].to(DEV), (1.0 / st.lengthscale), st.mean, st.outputscale, st.noise)
    assert int(info) == 0 and torch.equal(alpha, st.alpha) and torch.equal(linv, st.linv) and torch.equal(linv_t, st.linv_t)
    g = torch.Generator().manual_seed(5)
    grid = torch.randn(11, 40, 1, 2, dtype=torch.float64, generator=g).to(DEV)
    mask = torch.rand(11, 40, 1, 2, dtype=torch.float64, generator=g).to(DEV)
    lb, ub = t.conf_mask_to_bounds(grid, mask)
    lb2, ub2 = ops.conf_mask_to_bounds(grid, mask)
    assert torch.equal(lb, lb2) and torch.equal(ub, ub2)
    rw = ops.RatioWeights(*[w.to(DEV) for w in make_ratio_weights(d)])
    x = torch.rand(33, d, dtype=torch.float64, generator=g).to(DEV)
    assert torch.equal(t.ratio_forward([rw.W1, rw.b1, rw.W2, rw.b2, rw.W3, rw.b3], x), ops.ratio_forward(rw, x))
    P = ops.ratio_num_params(d)
    th0 = (torch.randn(P, dtype=torch.float64, generator=g) * 0.2).to(DEV)
    neg, pos = p["X"][:64].to(DEV).contiguous(), torch.rand(64, d, dtype=torch.float64, generator=g).to(DEV)
    a = [th0.clone(), torch.zeros_like(th0), torch.zeros_like(th0), torch.zeros_like(th0)]
    c = [x_.clone() for x_ in a]
    loss = t.ratio_train_step(*a, neg, pos, None, 64, 1, 5e-3, 0.9, 0.999, 1e-8, 5e-6, 2e-2)
    loss2 = torch.empty(1, dtype=torch.float64, device=DEV)
    ops.ratio_train_step(*c, x_neg=neg, x_pos=pos, pos_idx=None, n_pos=64, step=1, lr=5e-3, weight_decay=5e-6, ema_weight=2e-2,
                         loss=loss2)
    assert all(torch.equal(u, v) for u, v in zip(a, c)) and not torch.equal(a[0], th0)
    assert math.isfinite(float(loss)) and torch.equal(loss, loss2)
    X = torch.rand(16, d, dtype=torch.float64, generator=g).to(DEV)
    X2, buf, buf2 = X.clone(), torch.zeros_like(X), torch.zeros_like(X)
    gr, nz = torch.randn(16, d, dtype=torch.float64, generator=g).to(DEV), torch.randn(16, d, dtype=torch.float64, generator=g).to(DEV)
    lo, hi = torch.zeros(d, dtype=torch.float64, device=DEV), torch.ones(d, dtype=torch.float64, device=DEV)
    t.sgld_update(X, buf, gr, nz, lo, hi, 1e-3, 0.1, 1e-3, 1e2, 1e-2)
    ops.sgld_update(X2, buf2, gr, nz, lo, hi, lr=1e-3, momentum=0.1, temperature=1e-3)
    assert torch.equal(X, X2) and torch.equal(buf, buf2)


def test_operator_errors_are_loud():
    from conformal_bayesopt_b200 import torch_ops
    p = make_problem(40, 3, seed=24)
    inp = make_inputs(40, 3, 4, 1, 16, seed=24)
    st = _state(p)
    kw = dict(kind="ucb", alpha=0.1, temp=3e-2, grid_res=16, beta_prime=1.0)
    with pytest.raises((RuntimeError, NotImplementedError)):          # CPU candidates: no CPU kernel is registered
        torch_ops.conformal_acq(st, inp["Xq"], eps_grid=inp["eps_grid"][:, 0], z=inp["z"], eps_mc=inp["eps_mc"][:, :, 0], **kw)
    with pytest.raises(RuntimeError, match="z must be"):
        torch_ops.conformal_acq(st, inp["Xq"].to(DEV), eps_grid=inp["eps_grid"][:, 0].to(DEV), z=inp["z"][:-1].to(DEV),
                                eps_mc=inp["eps_mc"][:, :, 0].to(DEV), **kw)
    with pytest.raises(RuntimeError, match="u_rand"):
        torch_ops.conformal_acq(st, inp["Xq"].to(DEV), eps_grid=inp["eps_grid"][:, 0].to(DEV), z=inp["z"].to(DEV),
                                eps_mc=inp["eps_mc"][:, :, 0].to(DEV), randomized=True, **kw)
