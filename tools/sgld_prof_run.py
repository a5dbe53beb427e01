"""One SGLD-shaped value+gradient call (8 restarts x q=2, N=1024, ratio estimator, randomised) for an ncu launch list."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformal_bayesopt_b200 import GPModelState, ops
from conformal_bayesopt_b200.synthetic import make_inputs, make_problem, make_ratio_weights
N, d, G, b, q = 1024, 8, 64, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
p = make_problem(N, d, seed=0)
st = GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)
ratio = ops.RatioWeights(*[t.to(dev) for t in make_ratio_weights(d)])
kw = dict(kind="ucb", alpha=N ** -0.35, temp=3e-2, grid_res=G, randomized=True, optimistic=True, beta_prime=1.0, ratio=ratio)
inp = make_inputs(N, d, b, q, G, seed=0)
Xq = inp["Xq"].to(dev); eg = inp["eps_grid"].to(dev); z = inp["z"].to(dev); em = inp["eps_mc"].to(dev); u = inp["u_rand"].to(dev)
for rep in range(3):
    if rep == 2:
        torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    out = ops.conformal_acq(st, Xq, eps_grid=eg, z=z, eps_mc=em, u_rand=u, want_grad=True, **kw)
torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
print("ok", out["acq"].mean().item())
