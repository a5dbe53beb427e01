"""Where k_score_q_bwd's time goes: clock64() of CTA 0 / thread 0 at the phase boundaries of one SGLD-shaped call
(8 restarts x q = 2, N = 1024, G = 64, ratio estimator, randomised).  Needs an experiment build of the library:
    nvcc <build.py flags> -DCBO_SQB_CLOCKS -o tools/libcbo_clk.so conformal_bayesopt_b200/csrc/cbo_api.cu
    CBO_LIB_PATH=tools/libcbo_clk.so python tools/sqb_phase_clocks.py
The product library carries none of this (the macro expands to nothing)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformal_bayesopt_b200 import GPModelState, ops, _lib
from conformal_bayesopt_b200.synthetic import make_inputs, make_problem, make_ratio_weights
N, d, G, b, q = 1024, 8, 64, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
p = make_problem(N, d, seed=0)
st = GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)
ratio = ops.RatioWeights(*[t.to(dev) for t in make_ratio_weights(d)])
kw = dict(kind="ucb", alpha=N ** -0.35, temp=3e-2, grid_res=G, randomized=True, optimistic=True, beta_prime=1.0, ratio=ratio)
inp = make_inputs(N, d, b, q, G, seed=0)
Xq = inp["Xq"].to(dev); eg = inp["eps_grid"].to(dev); z = inp["z"].to(dev); em = inp["eps_mc"].to(dev); u = inp["u_rand"].to(dev)
lib = _lib.load()
names = ["gram sums (block)", "gram cluster exchange", "thread-0 q x q forward algebra", "phase A tables", "phase B sweep",
         "cluster barrier 1", "window gather + barrier 2", "phase C + densities", "phase D integration", "phase E reverse (per grid point)",
         "commit + zero", "pass 2 (per train point)", "commit + C_bar", "cluster barrier 3 + gather", "final q x q algebra (thread 0)",
         "cluster barrier 4 + ut_bar", "fold probes", "last cluster barrier"]
for rep in range(4):
    out = ops.conformal_acq(st, Xq, eps_grid=eg, z=z, eps_mc=em, u_rand=u, want_grad=True, **kw)
    buf = (ctypes.c_longlong * 32)()
    rc = lib.cbo_debug_sqb_clocks(buf)
    assert rc == 0, rc
    t = list(buf)[:19]
    if rep < 2:
        continue
    print(f"rep {rep}: total {t[18] - t[0]} clk")
    for k, nm in enumerate(names):
        print(f"  {nm:40s} {t[k + 1] - t[k]:8d} clk  {100.0 * (t[k + 1] - t[k]) / (t[18] - t[0]):5.1f} %")
    print(f"  (phase E: the per-grid-point chain of thread 0 {buf[19] - t[9]} clk, the block reductions after it {t[10] - buf[19]} clk)")
