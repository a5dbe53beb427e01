"""The graph-replayed SGLD step of bench.py's other_configs leg on its own (8 restarts x q = 2, N = 1024, d = 8, G = 64, cUCB,
randomised, ratio estimator; value + d acq / dX, fresh probes and uniforms per step)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformal_bayesopt_b200 import GPModelState, ops
from conformal_bayesopt_b200.graphs import GraphedConformalAcq
from conformal_bayesopt_b200.synthetic import make_inputs, make_problem, make_ratio_weights
dev = torch.device("cuda:0")
d, G = 8, 64
for N in (256, 1024):
    p = make_problem(N, d, seed=0)
    st = GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)
    rw = ops.RatioWeights(*[t.to(dev) for t in make_ratio_weights(d)])
    for (b, q) in ((8, 2), (8, 1), (8, 4)) + (((4, 2), (6, 2)) if os.environ.get("SGS_SMALL") else ()):
        inp = make_inputs(N, d, b, q, G, seed=3)
        kw = dict(kind="ucb", alpha=N ** -0.35, temp=3e-2, grid_res=G, randomized=True, optimistic=True, beta_prime=1.0, ratio=rw)
        gacq = GraphedConformalAcq(st, b, q, eps_grid=inp["eps_grid"].to(dev), eps_mc=inp["eps_mc"].to(dev), **kw)
        X = inp["Xq"].to(dev)
        best = 1e9
        for rep in range(3):
            for _ in range(5):
                gacq(X)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(100):
                gacq(X)
            torch.cuda.synchronize()
            best = min(best, (time.perf_counter() - t0) / 100 * 1e3)
        print(f"N={N:5d} restarts={b} q={q}: graph-replayed value+gradient step {best:.4f} ms")
