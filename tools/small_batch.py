"""Throughput at BASELINE config 2 (4096 candidates per call) and other batch sizes, device-resident."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformal_bayesopt_b200 import GPModelState, ops
from conformal_bayesopt_b200.synthetic import make_inputs, make_problem
N, d, G = 1024, 8, 64
dev = torch.device("cuda:0")
p = make_problem(N, d, seed=0)
st = GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)
kw = dict(kind="ucb", alpha=1 / math.sqrt(N), temp=3e-2, grid_res=G, randomized=False, optimistic=True, beta_prime=1.0)
grad = len(sys.argv) > 1 and sys.argv[1] == "grad"
for b in (8, 128, 1024, 4096, 16384, 65536):
    inp = make_inputs(N, d, b, 1, G, seed=0)
    Xq = inp["Xq"].to(dev); eg = inp["eps_grid"].to(dev); z = inp["z"].to(dev); em = inp["eps_mc"].to(dev)
    for _ in range(3):
        ops.conformal_acq(st, Xq, eps_grid=eg, z=z, eps_mc=em, want_grad=grad, **kw)
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.conformal_acq(st, Xq, eps_grid=eg, z=z, eps_mc=em, want_grad=grad, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        ops.conformal_acq(st, Xq, eps_grid=eg, z=z, eps_mc=em, want_grad=grad, **kw)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    print(f"grad={grad} b={b:6d}: {ms:8.3f} ms/call (host wall {wall:.3f})  {b * G / ms * 1e3:.3e} evals/s")
