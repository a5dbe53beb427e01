"""cuBLAS DGEMM / fp64 elementwise peak on the box (library reference for the fp64 roofline)."""
import torch, time, json
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"
res = {}
for n in (4096, 8192):
    a = torch.randn(n, n, device=dev, dtype=torch.float64)
    b = torch.randn(n, n, device=dev, dtype=torch.float64)
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}_tflops"] = 2 * n**3 / best / 1e9
print(json.dumps(res))
