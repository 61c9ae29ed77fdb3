import sys, os, cProfile, pstats, time, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "optimal-py-learning_b200"))
import numpy, torch
import learning_b200 as L
from learning_b200 import _native as nat, _device as dev
rng = numpy.random.default_rng(0)
x = rng.random((90, 4)) * 2 - 1
y = numpy.zeros((90, 3)); y[numpy.arange(90), rng.integers(0, 3, 90)] = 1.0
m = L.MLP((4, 2, 3), transfers=L.SoftmaxTransfer(), error_func=L.CrossEntropyError(), optimizer=L.optimize.BFGS())
m.logging = False
numpy.random.seed(0); m.reset()
for _ in range(5): m.train_step(x, y)
ws = m._workspace
n = 16
wd = dev.to_device(rng.random(n)); out = torch.empty(n + 2, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
l0 = nat.lib.opl_mlp_launch_count(ws.pointer)
t0 = time.perf_counter()
for _ in range(500):
    nat.lib.opl_mlp_eval_device(ws.pointer, dev.ptr(wd), 0.0, None, 1, dev.ptr(out))
torch.cuda.synchronize()
print("device eval, async back-to-back: %.1f us, %.1f launches per eval" % ((time.perf_counter() - t0) / 500 * 1e6, (nat.lib.opl_mlp_launch_count(ws.pointer) - l0) / 500))
t0 = time.perf_counter()
for _ in range(500): m._probe(ws, wd, 0.0, None, True)
print("_probe (eval + scalars readback): %.1f us" % ((time.perf_counter() - t0) / 500 * 1e6))
e0 = m.obj_jac_evaluations
t0 = time.perf_counter()
for _ in range(200): m.train_step(x, y)
dt = time.perf_counter() - t0
print("train_step: %.1f us, %.2f evals per step" % (dt / 200 * 1e6, (m.obj_jac_evaluations - e0) / 200))
pr = cProfile.Profile(); pr.enable()
for _ in range(200): m.train_step(x, y)
pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(14)
