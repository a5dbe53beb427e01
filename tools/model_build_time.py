"""cbo_model_build_f64 against the torch.linalg (cuSOLVER/cuBLAS) sequence it replaces: ms per model-state build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformal_bayesopt_b200 import GPModelState
from conformal_bayesopt_b200.model_state import matern52
from conformal_bayesopt_b200.synthetic import make_problem

dev = torch.device("cuda:0")
for N in (256, 1024, 2048, 4096):
    p = make_problem(N, 8, seed=0)
    X, y = p["X"].to(dev), p["y"].to(dev)
    ls = p["lengthscale"].to(dev)

    def ours():
        return GPModelState.fit_fixed(X, y, p["mean"], ls, p["outputscale"], p["noise"], device=dev, check_pd=False)

    def torch_ref():
        K = matern52(X, X, ls, p["outputscale"])
        A = K + p["noise"] * torch.eye(N, dtype=torch.float64, device=dev)
        L = torch.linalg.cholesky(A)
        alpha = torch.cholesky_solve((y - p["mean"]).unsqueeze(-1), L)
        linv = torch.linalg.solve_triangular(L, torch.eye(N, dtype=torch.float64, device=dev), upper=False)
        return linv, linv.t().contiguous(), alpha

    res = {}
    for name, fn in (("libcbo", ours), ("torch.linalg", torch_ref)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 10
    st = ours()
    linv_ref = torch_ref()[0]
    err = ((st.linv[:N, :N] - linv_ref).abs().max() / linv_ref.abs().max()).item()
    # appending q = 2 observations (one BO round of the default config) instead of rebuilding
    Xn, yn = torch.rand(2, 8, dtype=torch.float64, device=dev), torch.randn(2, dtype=torch.float64, device=dev) * 0.1
    for _ in range(3):
        st.append(Xn, yn, check_pd=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        st.append(Xn, yn, check_pd=False)
    e1.record()
    torch.cuda.synchronize()
    print(f"N={N}: libcbo {res['libcbo']:.2f} ms, torch.linalg {res['torch.linalg']:.2f} ms, max |dLinv| / max|Linv| = {err:.1e}; "
          f"append of 2 points {e0.elapsed_time(e1) / 10:.3f} ms (incl. the copy of the factors)")
