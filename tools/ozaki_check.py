#!/usr/bin/env python
"""Validate the fused tcgen05 INT8 Ozaki GEMM against torch float64 matmul; time its three A-operand variants next to
cuBLAS DGEMM.  usage: ozaki_check.py [variants e.g. 012] [quick]"""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "optimal-py-learning_b200"))
import torch
from learning_b200 import _native as nat, _device as dev

VARIANTS = [0]
NAMES = {0: "SS"}

def run(M, N, K, heavy=False, epi=0):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.rand(M, K, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    b = (torch.rand(K, N, dtype=torch.float64, device="cuda", generator=g) * 2 - 1) * 0.25
    if heavy:   # heavy-tailed columns / rows, as deltas are
        b = b * torch.exp(torch.randn(K, 1, dtype=torch.float64, device="cuda", generator=g) * 2) * 1e-5
    ref = a @ b
    if epi == 2:
        ref = torch.nn.functional.softplus(ref)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a @ b; e0.record()
    for _ in range(3): a @ b
    e1.record(); torch.cuda.synchronize()
    tc = e0.elapsed_time(e1) / 3
    ok_all = True
    for variant in VARIANTS:
        c = torch.full((M, N), float("nan"), dtype=torch.float64, device="cuda")
        sm, gm = ctypes.c_float(0), ctypes.c_float(0)
        rc = nat.lib.opl_bench_gemm_ozaki(M, N, K, dev.ptr(a), dev.ptr(b), dev.ptr(c), epi, 3, ctypes.byref(sm), ctypes.byref(gm))
        if rc != 0:
            print("ERR  variant %s: %s" % (NAMES[variant], nat.lib.opl_last_error().decode()), flush=True)
            return False
        torch.cuda.synchronize()
        err = float((c - ref).norm() / ref.norm())
        ok = err < 1e-11
        ok_all &= ok
        tops = 21 * 2.0 * M * N * K / (gm.value * 1e-3) / 1e12
        print("%s %-5s %6dx%5dx%6d heavy=%d epi=%d rel err %.2e | gemm %7.3f ms (%6.0f INT8 TOPS, %5.1f FP64-equiv TFLOP/s; digits %6.3f ms) | cuBLAS DGEMM %7.3f ms | x%.2f" %
              ("OK  " if ok else "FAIL", NAMES[variant], M, N, K, heavy, epi, err, gm.value, tops, 2.0 * M * N * K / (gm.value * 1e-3) / 1e12, sm.value, tc, tc / gm.value), flush=True)
    return ok_all

shapes = [(128, 64, 128), (300, 100, 500), (1000, 256, 784), (60000, 256, 784), (60000, 256, 784, False, 2), (785, 256, 60000, True),
          (8192, 8192, 8192)]
if "quick" in sys.argv:
    shapes = shapes[:3]
ok = True
for args in shapes:
    ok = run(*args) and ok
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
