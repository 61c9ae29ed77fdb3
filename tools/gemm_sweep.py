#!/usr/bin/env python
"""Shape sweep of the FP64 DMMA GEMM family against cuBLAS DGEMM (torch.matmul) on the same shapes."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "optimal-py-learning_b200"))
import torch
from learning_b200 import _native as nat, _device as dev

def ours(layout, M, N, K, iters=5):
    a = torch.randn(M * K, dtype=torch.float64, device="cuda")
    b = torch.randn(N * K, dtype=torch.float64, device="cuda")
    c = torch.empty(M * N, dtype=torch.float64, device="cuda")
    scratch = torch.empty(320 * M * N if layout == 2 else 1, dtype=torch.float64, device="cuda")
    ms = ctypes.c_float(0)
    nat.check(nat.lib.opl_bench_gemm_f64(layout, M, N, K, dev.ptr(a), dev.ptr(b), dev.ptr(c), dev.ptr(scratch),
                                         scratch.numel(), iters, ctypes.byref(ms)))
    return ms.value

def cublas(layout, M, N, K, iters=5):
    if layout == 0: a, b = torch.randn(M, K, dtype=torch.float64, device="cuda"), torch.randn(K, N, dtype=torch.float64, device="cuda")
    elif layout == 1: a, b = torch.randn(M, K, dtype=torch.float64, device="cuda"), torch.randn(N, K, dtype=torch.float64, device="cuda").t()
    else: a, b = torch.randn(K, M, dtype=torch.float64, device="cuda").t(), torch.randn(K, N, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): torch.matmul(a, b)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

shapes = [("fwd0 NN 60000x256x784", 0, 60000, 256, 784), ("dW0 TN 784x256x60000", 2, 784, 256, 60000),
          ("dA1 NT 60000x256x10", 1, 60000, 256, 10), ("NN 8192^3", 0, 8192, 8192, 8192), ("NN 4096x4096x4096", 0, 4096, 4096, 4096),
          ("cfg3 fwd0 NN 65536x128x512", 0, 65536, 128, 512), ("cfg3 dW0 TN 512x128x65536", 2, 512, 128, 65536),
          ("cfg5 NN 65536x1024x1024", 0, 65536, 1024, 1024), ("cfg5 TN 1024x1024x65536", 2, 1024, 1024, 65536),
          ("cfg5 NT 65536x1024x1024", 1, 65536, 1024, 1024)]
for name, layout, M, N, K in shapes:
    t, tc = ours(layout, M, N, K), cublas(layout, M, N, K)
    fl = 2.0 * M * N * K
    print("%-30s ours %8.3f ms %6.2f TF | cuBLAS %8.3f ms %6.2f TF | ratio %.2f" % (name, t, fl / t / 1e9, tc, fl / tc / 1e9, tc / t), flush=True)
