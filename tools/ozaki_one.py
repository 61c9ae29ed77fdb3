#!/usr/bin/env python
"""One shape through the Ozaki GEMM hook (for ncu): ozaki_one.py M N K variant epi"""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "optimal-py-learning_b200"))
import torch
from learning_b200 import _native as nat, _device as dev
M, N, K, epi = [int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (60000, 256, 784, 0))]
a = torch.rand(M, K, dtype=torch.float64, device="cuda") * 2 - 1
b = (torch.rand(K, N, dtype=torch.float64, device="cuda") * 2 - 1) * 0.25
c = torch.empty(M, N, dtype=torch.float64, device="cuda")
sm, gm = ctypes.c_float(0), ctypes.c_float(0)
nat.check(nat.lib.opl_bench_gemm_ozaki(M, N, K, dev.ptr(a), dev.ptr(b), dev.ptr(c), epi, 2, ctypes.byref(sm), ctypes.byref(gm)))
print("gemm %.3f ms, digits %.3f ms" % (gm.value, sm.value))
