"""Forward sweep timing at the bench shape with per-kernel-class CUDA-event times (honours CBO_LIB_PATH for experiment builds).
usage: python tools/fwd_time.py [candidates] [ratio 0|1]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import GRID, PARAMS, workload_cpu
from conformal_bayesopt_b200 import GPModelState, _lib, ops
from conformal_bayesopt_b200.synthetic import make_ratio_weights

b = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
use_ratio = len(sys.argv) > 2 and sys.argv[2] == "1"
dev = torch.device("cuda:0")
prob, inp, Xq = workload_cpu(0, b)
state = GPModelState.fit_fixed(prob["X"], prob["y"], prob["mean"], prob["lengthscale"], prob["outputscale"], prob["noise"], device=dev)
ratio = ops.RatioWeights(*[t.to(dev) for t in make_ratio_weights(8)]) if use_ratio else None
kw = dict(eps_grid=inp["eps_grid"][:, 0].to(dev), z=inp["z"].to(dev), eps_mc=inp["eps_mc"][:, :, 0].to(dev), temp=3e-2,
          ratio=ratio, **PARAMS)
Xq = Xq.to(dev)
lib = _lib.load()
names = ["prep", "rows", "gemm1", "gemm2", "score", "score_bwd", "gemm_bwd", "rows_bwd"]
for _ in range(2):
    ops.conformal_acq(state, Xq, **kw)
torch.cuda.synchronize()
lib.cbo_timing_enable(1)
ms = (C.c_double * 8)(); nl = (C.c_int64 * 8)()
lib.cbo_timing_collect(ms, nl)
steps = 4
for _ in range(steps):
    ops.conformal_acq(state, Xq, **kw)
torch.cuda.synchronize()
lib.cbo_timing_collect(ms, nl)
print(os.environ.get("CBO_LIB_PATH", "libcbo.so"), f"ratio={use_ratio} b={b}: " + "  ".join(f"{n}={ms[i] / steps:.2f}" for i, n in enumerate(names) if ms[i] > 0))
