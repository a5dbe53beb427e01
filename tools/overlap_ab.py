"""A/B timing of the overlapped chunk pipeline (score(k) beside rows + GEMMs(k+1)) against the serial one, with the footprints
of the two co-resident kernels as parameters.  usage: python tools/overlap_ab.py [log2 candidates] [ratio 0/1]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformal_bayesopt_b200 import GPModelState, ops
from conformal_bayesopt_b200.synthetic import make_inputs, make_problem

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
with_ratio = len(sys.argv) > 2 and sys.argv[2] == "1"
b, N, d, G = 1 << lg, 1024, 8, 64
dev = torch.device("cuda:0")
p = make_problem(N, d, seed=0)
inp = make_inputs(N, d, 4096, 1, G, seed=0)
st = GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)
g = torch.Generator(device=dev).manual_seed(1)
Xq = torch.rand(b, 1, d, generator=g, dtype=torch.float64, device=dev)
eg = inp["eps_grid"][:, 0].to(dev); z = inp["z"].to(dev); em = inp["eps_mc"][:, :, 0].to(dev)
kw = dict(kind="ucb", alpha=1 / math.sqrt(N), temp=3e-2, grid_res=G, randomized=False, optimistic=True, beta_prime=1.0)
ratio = None
if with_ratio:
    from oracle import gp as ogp
    r = ogp.make_ratio_mlp(d)
    ratio = ops.RatioWeights(*[t.to(dev) for t in (r.W1, r.b1, r.W2, r.b2, r.W3, r.b3)])


def run(env, reps=3):
    for k in ("CBO_OVERLAP", "CBO_WS_GROUPS", "CBO_GEMM_STAGES"):
        os.environ.pop(k, None)
    os.environ.update(env)
    out = ops.conformal_acq(st, Xq, eps_grid=eg, z=z, eps_mc=em, ratio=ratio, **kw)["acq"]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = ops.conformal_acq(st, Xq, eps_grid=eg, z=z, eps_mc=em, ratio=ratio, **kw)["acq"]
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out.clone()


base_ms, base = run({"CBO_OVERLAP": "0"})
print(f"serial                         {base_ms:8.2f} ms per {b} candidates")
for groups in (1, 2, 3, 4):
    for stages in (2, 3, 4):
        ms, out = run({"CBO_OVERLAP": "1", "CBO_WS_GROUPS": str(groups), "CBO_GEMM_STAGES": str(stages)})
        same = torch.equal(out, base)
        print(f"overlap groups={groups} stages={stages}      {ms:8.2f} ms   x{base_ms / ms:5.3f}   bit-identical: {same}")
