"""One small value+gradient call for ncu (b candidates, N=1024, d=8, G=64)."""
import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformal_bayesopt_b200 import GPModelState, ops
from conformal_bayesopt_b200.synthetic import make_inputs, make_problem
b = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
N, d, G = 1024, 8, 64
dev = torch.device("cuda:0")
p = make_problem(N, d, seed=0)
inp = make_inputs(N, d, b, 1, G, seed=0)
st = GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)
kw = dict(kind="ucb", alpha=1 / math.sqrt(N), temp=3e-2, grid_res=G, randomized=False, optimistic=True, beta_prime=1.0)
Xq = inp["Xq"].to(dev); eg = inp["eps_grid"][:, 0].to(dev); z = inp["z"].to(dev); em = inp["eps_mc"][:, :, 0].to(dev)
for rep in range(reps):
    if rep == reps - 1:                       # ncu --profile-from-start off: capture the last repetition only
        torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    out = ops.conformal_acq(st, Xq, eps_grid=eg, z=z, eps_mc=em, want_grad=True, **kw)
torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
print("acq mean", out["acq"].mean().item(), "grad abs mean", out["dacq_dX"].abs().mean().item())
