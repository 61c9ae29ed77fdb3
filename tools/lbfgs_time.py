"""Time of one L-BFGS direction (two-loop recursion kernel) at config 2's n, m = 5."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "optimal-py-learning_b200"))
import torch
from learning_b200 import _native as nat, _device as dev
n = int(sys.argv[1]) if len(sys.argv) > 1 else 203520
h = ctypes.c_void_p()
nat.check(nat.lib.opl_lbfgs_create(ctypes.byref(h), n, 5, dev.stream_ptr()))
g = torch.Generator(device="cuda").manual_seed(0)
vecs = [torch.randn(n, dtype=torch.float64, device="cuda", generator=g) for _ in range(12)]
d = torch.empty(n, dtype=torch.float64, device="cuda")
def step(i):
    nat.check(nat.lib.opl_lbfgs_direction_device(h, dev.ptr(vecs[i % 4]), dev.ptr(vecs[4 + i % 4]), dev.ptr(vecs[8 + i % 4]), 0, dev.ptr(d)))
for i in range(10): step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(100): step(i)
e1.record(); torch.cuda.synchronize()
print("L-BFGS direction n=%d m=5: %.1f us (%.0f GB/s on (4m+2) n 8 bytes), launches per call %.1f" % (
    n, e0.elapsed_time(e1) * 10, 22 * n * 8 / (e0.elapsed_time(e1) * 1e-5) / 1e9, nat.lib.opl_lbfgs_launch_count(h) / 110.0))
