#!/usr/bin/env python
"""Validate the fused tcgen05 INT8 Ozaki GEMM against torch float64 matmul; time it next to cuBLAS DGEMM and our DMMA GEMM."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "optimal-py-learning_b200"))
import torch
from learning_b200 import _native as nat, _device as dev

def run(M, N, K, heavy=False):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.rand(M, K, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    b = (torch.rand(K, N, dtype=torch.float64, device="cuda", generator=g) * 2 - 1) * 0.25
    if heavy:   # heavy-tailed columns / rows, as deltas are
        b = b * torch.exp(torch.randn(K, 1, dtype=torch.float64, device="cuda", generator=g) * 2) * 1e-5
    ref = a @ b
    c = torch.full((M, N), float("nan"), dtype=torch.float64, device="cuda")
    sm, gm = ctypes.c_float(0), ctypes.c_float(0)
    nat.check(nat.lib.opl_bench_gemm_ozaki(M, N, K, dev.ptr(a), dev.ptr(b), dev.ptr(c), 3, ctypes.byref(sm), ctypes.byref(gm)))
    torch.cuda.synchronize()
    err = float((c - ref).norm() / ref.norm())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a @ b; e0.record()
    for _ in range(3): a @ b
    e1.record(); torch.cuda.synchronize()
    tc = e0.elapsed_time(e1) / 3
    ok = err < 1e-11
    print("%s %6dx%5dx%6d heavy=%d rel err %.2e | ozaki gemm %7.3f ms (slicing %6.3f) | cuBLAS DGEMM %7.3f ms | speed-up %.2fx" %
          ("OK  " if ok else "FAIL", M, N, K, heavy, err, gm.value, sm.value, tc, tc / gm.value), flush=True)
    return ok

ok = True
for args in [(128, 64, 128), (300, 100, 500), (1000, 256, 784), (784, 256, 20000, True), (60000, 256, 784), (784, 256, 60000, True),
             (8192, 8192, 8192)]:
    ok = run(*args) and ok
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
