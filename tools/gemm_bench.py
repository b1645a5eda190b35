#!/usr/bin/env python
"""Throughput of csrc/gemm.cu on the PPO update's shapes next to torch (cuBLAS TF32): tools/gemm_bench.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from rlgymppo_cpp_b200 import gemm as G

torch.backends.cuda.matmul.allow_tf32 = True
R = 32768


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


for name, M, N, K, split in [("fwd/dX 256", R, 256, 256, 1), ("fwd obs", R, 256, 92, 1), ("head 92", R, 92, 256, 1), ("dW 256", 256, 256, R, 74), ("dW obs", 256, 92, R, 74)]:
    a = torch.randn(M, K, device="cuda"); b = torch.randn(N, K, device="cuda")
    out = torch.zeros(M, N, device="cuda")
    t_mine = timeit(lambda: G.gemm(a, b, out=out, atomic=split > 1, split_k=split))
    t_torch = timeit(lambda: torch.matmul(a, b.t()))
    fl = 2.0 * M * N * K
    print(f"{name:12s} M={M:6d} N={N:4d} K={K:6d}  mine {t_mine*1e3:7.1f} us ({fl/t_mine/1e9:6.1f} TFLOP/s)   cuBLAS tf32 {t_torch*1e3:7.1f} us ({fl/t_torch/1e9:6.1f} TFLOP/s)")
