"""Where the host-buffer evaluation (MLP._get_obj_jac with NumPy parameters) spends its time."""
import sys, os, time, ctypes, cProfile, pstats
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "optimal-py-learning_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy, torch
import learning_b200 as L
from learning_b200 import _native as nat
from learning_b200.architecture import mlp as M
rng = numpy.random.default_rng(0)
x = rng.random((60000, 784)) * 2 - 1
y = numpy.zeros((60000, 10)); y[numpy.arange(60000), rng.integers(0, 10, 60000)] = 1.0
m = L.MLP((784, 256, 10), transfers=L.SoftmaxTransfer(), error_func=L.CrossEntropyError(), optimizer=L.optimize.LBFGS())
n = 256 + 784 * 256 + 2560
w = torch.empty(n, dtype=torch.float64).pin_memory().numpy(); w[:] = (rng.random(n) * 2 - 1) * 0.25
scope = m.resident(x, y); scope.__enter__()          # what Model.train opens around its train_steps
for _ in range(5): m._get_obj_jac(w, x, y)
def timeit(f, k=200):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e6
print("full _get_obj_jac           %8.1f us" % timeit(lambda: m._get_obj_jac(w, x, y), 100))
print("pinned torch.empty().numpy()%8.1f us" % timeit(lambda: torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()))
ws = m._workspace
g = torch.empty(n, dtype=torch.float64, pin_memory=True).numpy(); obj = ctypes.c_double()
wp, gp = w.ctypes.data_as(ctypes.c_void_p), g.ctypes.data_as(ctypes.c_void_p)
print("C call opl_mlp_obj_jac_host %8.1f us" % timeit(lambda: nat.lib.opl_mlp_obj_jac_host(ws.pointer, wp, ctypes.byref(obj), gp), 100))
wd = torch.from_numpy(w).cuda(); out = torch.empty(n + 2, dtype=torch.float64, device="cuda")
print("device eval only            %8.1f us" % timeit(lambda: nat.lib.opl_mlp_eval_device(ws.pointer, ctypes.c_void_p(wd.data_ptr()), 0.0, None, 1, ctypes.c_void_p(out.data_ptr())), 100))
hp = torch.empty(n, dtype=torch.float64, pin_memory=True)
print("H2D 1.6 MB pinned + sync    %8.1f us" % timeit(lambda: (wd.copy_(hp, non_blocking=True), torch.cuda.synchronize())))
print("D2H 1.6 MB pinned + sync    %8.1f us" % timeit(lambda: (hp.copy_(wd, non_blocking=True), torch.cuda.synchronize())))
pr = cProfile.Profile(); pr.enable()
for _ in range(100): m._get_obj_jac(w, x, y)
pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(12)
