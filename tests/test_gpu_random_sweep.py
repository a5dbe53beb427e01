"""Seeded random sweep over shapes and flags: values AND gradient rows of the CUDA path against the closed-form oracle
(autograd).  Complements the hand-picked cases of test_gpu_parity.py / test_gpu_grad.py."""
import math
import os
import random

import pytest
import torch

from oracle import closed_form
from oracle import gp as ogp
from tests.problems import make_inputs, make_problem, oracle_state

pytestmark = pytest.mark.gpu


def _cases(n=28, seed=1234):
    rng = random.Random(seed)
    out = []
    for i in range(n):
        q = rng.choice([1, 1, 1, 2, 2, 3, 4])
        N = rng.choice([3, 17, 32, 33, 64, 100, 129, 200, 257])
        d = rng.choice([1, 2, 3, 5, 8, 9, 13])
        b = rng.choice([1, 2, 7, 33]) if q > 1 else rng.choice([1, 5, 40, 130])
        G = rng.choice([2, 5, 16, 31, 64, 70])
        out.append(dict(i=i, q=q, N=N, d=d, b=b, G=G, S=rng.choice([1, 1, 2, 3]), Sp=rng.choice([8, 8, 4, 16]),
                        temp=rng.choice([3e-2, 1e-2, 1e-1]), alpha=rng.choice([0.05, 0.1, 0.3]),
                        kind=rng.choice(["ucb", "ei"]), rand=rng.random() < 0.5, ratio=rng.random() < 0.5,
                        opt=rng.random() < 0.7, H=rng.choice([8, 24, 64])))
    return out


@pytest.mark.parametrize("c", _cases(int(os.environ.get("CBO_SWEEP_CASES", "28")), int(os.environ.get("CBO_SWEEP_SEED", "1234"))), ids=lambda c: f"{c['i']}-q{c['q']}-N{c['N']}-d{c['d']}-b{c['b']}-G{c['G']}")
def test_random_case_value_and_gradient(c):
    from conformal_bayesopt_b200 import GPModelState, ops
    from conformal_bayesopt_b200.synthetic import make_inputs as mk
    assert torch.cuda.is_available(), "run with -m gpu on a CUDA box"
    dev = torch.device("cuda:0")
    p = make_problem(c["N"], c["d"], seed=100 + c["i"])
    inp = mk(c["N"], c["d"], c["b"], c["q"], c["G"], S=c["S"], Sp=c["Sp"], seed=c["i"])
    ratio = rw = None
    if c["ratio"]:
        ratio = ogp.make_ratio_mlp(c["d"], seed=c["i"], hidden=c["H"])
        rw = ops.RatioWeights(*[t.to(dev) for t in (ratio.W1, ratio.b1, ratio.W2, ratio.b2, ratio.W3, ratio.b3)])
    kw = dict(kind=c["kind"], alpha=c["alpha"], temp=c["temp"], grid_res=c["G"], randomized=c["rand"], optimistic=c["opt"],
              beta_prime=math.sqrt(0.2 * math.pi / 2), best_f=0.2)
    X = inp["Xq"].clone().requires_grad_(True)
    want = closed_form.conformal_acquisition(oracle_state(p), X, eps_grid=inp["eps_grid"], z=inp["z"], eps_mc=inp["eps_mc"],
                                             ratio=ratio, u_rand=inp["u_rand"], **kw)
    (g,) = torch.autograd.grad(want["acq"].sum(), X)
    st = GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)
    got = ops.conformal_acq(st, inp["Xq"].to(dev), eps_grid=inp["eps_grid"].to(dev), z=inp["z"].to(dev),
                            eps_mc=inp["eps_mc"].to(dev), u_rand=inp["u_rand"].to(dev), ratio=rw, want_grad=True,
                            outputs=("conf_pred_mask",), **kw)
    torch.cuda.synchronize()
    assert torch.allclose(got["acq"].cpu(), want["acq"].detach(), rtol=1e-6, atol=1e-9)
    mask = got["conf_pred_mask"].cpu().reshape(want["mask"].shape)
    assert torch.allclose(mask, want["mask"], rtol=0, atol=1e-6)
    scale = float(g.abs().max().clamp_min(1e-300))
    assert torch.allclose(got["dacq_dX"].cpu(), g, rtol=1e-6, atol=1e-9 * scale + 1e-14)
