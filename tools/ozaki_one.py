import ctypes, os, sys
sys.path.insert(0, "/root/repo/optimal-py-learning_b200")
import torch
from learning_b200 import _native as nat, _device as dev
M,N,K=60000,256,784
a = torch.rand(M, K, dtype=torch.float64, device="cuda")*2-1
b = (torch.rand(K, N, dtype=torch.float64, device="cuda")*2-1)*0.25
c = torch.empty(M, N, dtype=torch.float64, device="cuda")
sm, gm = ctypes.c_float(0), ctypes.c_float(0)
nat.check(nat.lib.opl_bench_gemm_ozaki(M, N, K, dev.ptr(a), dev.ptr(b), dev.ptr(c), 2, ctypes.byref(sm), ctypes.byref(gm)))
print(gm.value)
