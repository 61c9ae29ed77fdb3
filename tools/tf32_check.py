#!/usr/bin/env python
"""Validate + time the tcgen05 TF32 GEMM family against torch (exact FP32 reference, cuBLAS TF32 timing)."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "optimal-py-learning_b200"))
import torch
from learning_b200 import _native as nat, _device as dev

torch.backends.cuda.matmul.allow_tf32 = False

def run(layout, M, N, K, splits=1, iters=5):
    g = torch.Generator(device="cuda").manual_seed(M + N + K + layout)
    if layout == 0: a, b = torch.randn(M, K, device="cuda", generator=g), torch.randn(K, N, device="cuda", generator=g); ref = a @ b
    elif layout == 1: a, b = torch.randn(M, K, device="cuda", generator=g), torch.randn(N, K, device="cuda", generator=g); ref = a @ b.t()
    else: a, b = torch.randn(K, M, device="cuda", generator=g), torch.randn(K, N, device="cuda", generator=g); ref = a.t() @ b
    c = torch.full((M, N), float("nan"), device="cuda")
    scratch = torch.empty(max(1, splits * M * N if splits > 1 else 1), device="cuda")
    ms = ctypes.c_float(0)
    nat.check(nat.lib.opl_bench_gemm_tf32(layout, M, N, K, dev.ptr(a), dev.ptr(b), dev.ptr(c), dev.ptr(scratch),
                                          scratch.numel(), splits, iters, ctypes.byref(ms)))
    torch.cuda.synchronize()
    err = float((c - ref).norm() / ref.norm())
    torch.backends.cuda.matmul.allow_tf32 = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f = (lambda: a @ b) if layout == 0 else ((lambda: a @ b.t()) if layout == 1 else (lambda: a.t() @ b))
    f(); e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    torch.backends.cuda.matmul.allow_tf32 = False
    tc = e0.elapsed_time(e1) / iters
    fl = 2.0 * M * N * K
    ok = err < 2e-3
    print("%s layout=%d %6dx%5dx%6d splits=%2d  rel err %.2e  ours %7.3f ms %7.1f TF | cuBLAS-tf32 %7.3f ms %7.1f TF" %
          ("OK  " if ok else "FAIL", layout, M, N, K, splits, err, ms.value, fl / ms.value / 1e9, tc, fl / tc / 1e9), flush=True)
    return ok

ok = True
for args in [(0, 128, 32, 32), (0, 128, 128, 64), (1, 128, 64, 96), (2, 128, 32, 64), (2, 256, 128, 512),
             (0, 300, 160, 800), (1, 1000, 800, 256), (2, 800, 256, 4096, 4),
             (0, 60000, 256, 800), (1, 60000, 256, 32), (2, 800, 256, 60000, 16), (0, 8192, 8192, 8192)]:
    ok = run(*args) and ok
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
