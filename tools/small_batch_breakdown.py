import math, os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformal_bayesopt_b200 import GPModelState, ops, _lib
from conformal_bayesopt_b200.synthetic import make_inputs, make_problem
N, d, G = 1024, 8, 64
dev = torch.device("cuda:0")
lib = _lib.load()
p = make_problem(N, d, seed=0)
st = GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)
kw = dict(kind="ucb", alpha=1 / math.sqrt(N), temp=3e-2, grid_res=G, randomized=False, optimistic=True, beta_prime=1.0)
q = int(sys.argv[1]) if len(sys.argv) > 1 else 1
grad = len(sys.argv) > 2 and sys.argv[2] == "grad"
kw["want_grad"] = grad
for b in (8, 128, 1024, 4096):
    inp = make_inputs(N, d, b, q, G, seed=0)
    Xq = inp["Xq"].to(dev); eg = inp["eps_grid"].to(dev); z = inp["z"].to(dev); em = inp["eps_mc"].to(dev)
    for _ in range(3):
        ops.conformal_acq(st, Xq, eps_grid=eg, z=z, eps_mc=em, **kw)
    torch.cuda.synchronize()
    lib.cbo_timing_enable(1)
    ms = (C.c_double * 8)(); nl = (C.c_int64 * 8)()
    lib.cbo_timing_collect(ms, nl)
    reps = 10
    for _ in range(reps):
        ops.conformal_acq(st, Xq, eps_grid=eg, z=z, eps_mc=em, **kw)
    torch.cuda.synchronize()
    lib.cbo_timing_collect(ms, nl)
    lib.cbo_timing_enable(0)
    print(f"q={q} grad={grad} b={b}:", " ".join(f"{n}={m / reps * 1e3:.0f}us" for n, m in zip(["prep", "rows", "gemm1", "gemm2", "score", "score_bwd", "gemm_bwd", "rows_bwd"], ms) if m > 0))
