import sys, os, cProfile, pstats, time
sys.path.insert(0, "/root/repo/optimal-py-learning_b200"); sys.path.insert(0, "/root/repo/tools")
import numpy, torch
import learning_b200 as L
from run_configs import synthetic
x, y = synthetic(60000, 784, 10, 1)
m = L.MLP((784, 256, 10), transfers=L.SoftmaxTransfer(), error_func=L.CrossEntropyError(), optimizer=L.optimize.LBFGS())
m.logging = False
numpy.random.seed(0); m.reset()
for _ in range(3): m.train_step(x, y)
torch.cuda.synchronize()
e0 = m.obj_jac_evaluations
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
for _ in range(10): m.train_step(x, y)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
pr.disable()
print("10 steps %.4f s, evals %d" % (dt, m.obj_jac_evaluations - e0))
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
