"""Forward / gradient timing for larger training sets (shared memory admits fewer candidates per SM)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformal_bayesopt_b200 import GPModelState, ops
from conformal_bayesopt_b200.synthetic import make_inputs, make_problem
d, G, b = 8, 64, 16384
dev = torch.device("cuda:0")
for N in (1024, 2048, 4096):
    p = make_problem(N, d, seed=0)
    st = GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)
    inp = make_inputs(N, d, b, 1, G, seed=0)
    kw = dict(kind="ucb", alpha=1 / math.sqrt(N), temp=3e-2, grid_res=G, randomized=False, optimistic=True, beta_prime=1.0)
    Xq = inp["Xq"].to(dev); eg = inp["eps_grid"].to(dev); z = inp["z"].to(dev); em = inp["eps_mc"].to(dev)
    for grad in (False, True):
        for _ in range(2):
            ops.conformal_acq(st, Xq, eps_grid=eg, z=z, eps_mc=em, want_grad=grad, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.conformal_acq(st, Xq, eps_grid=eg, z=z, eps_mc=em, want_grad=grad, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"N={N} grad={grad}: {ms:.2f} ms per {b} candidates, {b * G / ms * 1e3:.3e} evals/s", flush=True)
