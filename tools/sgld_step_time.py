"""Latency of ONE SGLD step of the reference's default loop shape (optimize_acqf_sgld: num_restarts x q_batch_size x d with grad,
conformalbo/optim/optimize.py:224-235; hydra: num_opt_restarts 8, q_batch_size 2, ratio estimator, randomized, N grows to ~1000)."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformal_bayesopt_b200 import GPModelState, ops
from conformal_bayesopt_b200.synthetic import make_inputs, make_problem, make_ratio_weights
d, G = 8, 64
dev = torch.device("cuda:0")
for N in (64, 256, 1024):
    p = make_problem(N, d, seed=0)
    st = GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)
    ratio = ops.RatioWeights(*[t.to(dev) for t in make_ratio_weights(d)])
    kw = dict(kind="ucb", alpha=N ** -0.35, temp=3e-2, grid_res=G, randomized=True, optimistic=True, beta_prime=1.0, ratio=ratio)
    for (b, q) in ((8, 1), (8, 2), (64, 2), (8, 4)):
        inp = make_inputs(N, d, b, q, G, seed=0)
        Xq = inp["Xq"].to(dev); eg = inp["eps_grid"].to(dev); z = inp["z"].to(dev); em = inp["eps_mc"].to(dev); u = inp["u_rand"].to(dev)
        res = []
        for grad in (False, True):
            for _ in range(3):
                ops.conformal_acq(st, Xq, eps_grid=eg, z=z, eps_mc=em, u_rand=u, want_grad=grad, **kw)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                ops.conformal_acq(st, Xq, eps_grid=eg, z=z, eps_mc=em, u_rand=u, want_grad=grad, **kw)
            torch.cuda.synchronize()
            res.append((time.perf_counter() - t0) / 20 * 1e3)
        print(f"N={N:5d} restarts={b:3d} q={q}: value {res[0]:.3f} ms, value+grad {res[1]:.3f} ms per call")
