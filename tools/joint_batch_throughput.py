"""Throughput of the joint-batch kernels (q = 2, 4) for larger batches (raw-sample scoring of gen_batch_initial_conditions)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformal_bayesopt_b200 import GPModelState, ops
from conformal_bayesopt_b200.synthetic import make_inputs, make_problem
N, d, G = 1024, 8, 64
dev = torch.device("cuda:0")
p = make_problem(N, d, seed=0)
st = GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)
kw = dict(kind="ucb", alpha=1 / math.sqrt(N), temp=3e-2, grid_res=G, randomized=False, optimistic=True, beta_prime=1.0)
for q in (1, 2, 4):
    for b in (512, 4096):
        inp = make_inputs(N, d, b, q, G, seed=0)
        Xq = inp["Xq"].to(dev); eg = inp["eps_grid"].to(dev); z = inp["z"].to(dev); em = inp["eps_mc"].to(dev)
        for grad in (False, True):
            for _ in range(2):
                ops.conformal_acq(st, Xq, eps_grid=eg, z=z, eps_mc=em, want_grad=grad, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.conformal_acq(st, Xq, eps_grid=eg, z=z, eps_mc=em, want_grad=grad, **kw)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"q={q} b={b:5d} grad={grad}: {ms:7.3f} ms  {b * q * G / ms * 1e3:.3e} point-evals/s", flush=True)
