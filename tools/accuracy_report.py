"""GPU-vs-oracle error at the benchmark size (N=1024, d=8, G=64): prints max errors for profiles/README.md."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformal_bayesopt_b200 import GPModelState, ops
from conformal_bayesopt_b200.synthetic import make_inputs, make_problem
from oracle import closed_form, gp as ogp
N, d, G, b = 1024, 8, 64, 256
dev = torch.device("cuda:0")
for noise in (1e-2, 1e-4):
    p = make_problem(N, d, seed=0, noise=noise)
    inp = make_inputs(N, d, b, 1, G, seed=0)
    st = GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)
    gp = ogp.build_state(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"])
    for temp in (3e-2, 1e-6, 0.0):
        kw = dict(kind="ucb", alpha=1 / math.sqrt(N), temp=temp, grid_res=G, randomized=True, optimistic=True, beta_prime=1.0)
        want = closed_form.conformal_acquisition(gp, inp["Xq"], eps_grid=inp["eps_grid"], z=inp["z"], eps_mc=inp["eps_mc"],
                                                 u_rand=inp["u_rand"], **kw)
        got = ops.conformal_acq(st, inp["Xq"].to(dev), eps_grid=inp["eps_grid"].to(dev), z=inp["z"].to(dev),
                                eps_mc=inp["eps_mc"].to(dev), u_rand=inp["u_rand"].to(dev),
                                outputs=("conf_pred_mask", "post_var", "post_mean"), **kw)
        torch.cuda.synchronize()
        acq_rel = ((got["acq"].cpu() - want["acq"]).abs() / want["acq"].abs().clamp_min(1e-12)).max().item()
        mask_abs = (got["conf_pred_mask"].cpu() - want["mask"][:, :, 0]).abs().max().item()
        s_rel = ((got["post_var"].cpu() - want["S"][:, 0, 0]).abs() / want["S"][:, 0, 0]).max().item()
        flips = int(((got["conf_pred_mask"].cpu() >= 0.5) != (want["mask"][:, :, 0] >= 0.5)).sum())
        print(f"noise={noise:g} temp={temp:g}: acq max rel err {acq_rel:.2e}, mask max abs err {mask_abs:.2e}, "
              f"s max rel err {s_rel:.2e}, coverage-mask flips {flips}/{b * G}")

# ---- gradient rows (cbo_conformal_acq_grad_f64) against autograd through the closed-form oracle, same size ------------------
bg = 48
for q in (1, 2):
    p = make_problem(N, d, seed=0, noise=1e-2)
    inp = make_inputs(N, d, bg, q, G, seed=3)
    st = GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)
    gp = ogp.build_state(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"])
    kw = dict(kind="ucb", alpha=1 / math.sqrt(N), temp=3e-2, grid_res=G, randomized=True, optimistic=True, beta_prime=1.0)
    Xr = inp["Xq"].clone().requires_grad_(True)
    want = closed_form.conformal_acquisition(gp, Xr, eps_grid=inp["eps_grid"], z=inp["z"], eps_mc=inp["eps_mc"], u_rand=inp["u_rand"], **kw)
    (gref,) = torch.autograd.grad(want["acq"].sum(), Xr)
    got = ops.conformal_acq(st, inp["Xq"].to(dev), eps_grid=inp["eps_grid"].to(dev), z=inp["z"].to(dev), eps_mc=inp["eps_mc"].to(dev),
                            u_rand=inp["u_rand"].to(dev), want_grad=True, **kw)
    torch.cuda.synchronize()
    gerr = ((got["dacq_dX"].cpu() - gref).abs().max() / gref.abs().max()).item()
    aerr = ((got["acq"].cpu() - want["acq"].detach()).abs() / want["acq"].detach().abs().clamp_min(1e-12)).max().item()
    print(f"gradient q={q}: max |d acq/dX - autograd(oracle)| / max|grad| = {gerr:.2e} (max|grad| = {gref.abs().max().item():.3f}), acq rel err {aerr:.2e}")
