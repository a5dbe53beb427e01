"""DMMA GEMM (cbo_gemm_tn_f64) timing for full / lower / upper K-ranges and several K: how much of a tile's time is fixed
overhead (prologue, epilogue, CTA launch)?  usage: python tools/gemm_shape_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformal_bayesopt_b200 import ops

dev = torch.device("cuda:0")
M = 65536
for K in (256, 512, 1024, 2048):
    A = torch.randn(M, K, dtype=torch.float64, device=dev)
    B = torch.randn(1024, K, dtype=torch.float64, device=dev)
    for tri in ((0, 1, 2) if K == 1024 else (0,)):
        for _ in range(3):
            ops.gemm_tn(A, B, tri=tri)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm_tn(A, B, tri=tri)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        flop = 2.0 * M * 1024 * K * (1.0 if tri == 0 else 0.5)
        print(f"M={M} N=1024 K={K} tri={tri}: {ms:.3f} ms  {flop / ms / 1e9:.1f} TFLOP/s algorithmic ({flop / ms / 1e9 / 37.0:.3f} of 37.0)")
