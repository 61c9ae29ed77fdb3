#!/usr/bin/env python
"""Per-launch stage times of one obj+grad evaluation for an arbitrary MLP shape:
    stage_times.py ROWS D0,D1,...,DL [precision ...]      e.g.  stage_times.py 65536 512,128,128,10 f64 f64-dmma"""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "optimal-py-learning_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy, torch
import learning_b200 as L
from learning_b200 import _native as nat, _device as dev
from run_configs import synthetic

NAMES = {1: "forward GEMM", 2: "row kernel", 3: "dW GEMM", 4: "dW reduce", 5: "dA GEMM", 6: "bias colsum", 7: "loss", 8: "digits",
         10: "fused head", 0: "trial point"}
rows = int(sys.argv[1]); shape = tuple(int(v) for v in sys.argv[2].split(","))
precisions = sys.argv[3:] or ["f64", "f64-dmma"]
x, y = synthetic(rows, shape[0], shape[-1], 1)
n = shape[1] + sum(a * b for a, b in zip(shape[:-1], shape[1:]))
w = dev.to_device((numpy.random.default_rng(0).random(n) * 2 - 1) * 0.25)
for precision in precisions:
    m = L.MLP(shape, transfers=L.SoftmaxTransfer(), error_func=L.CrossEntropyError(), optimizer=L.optimize.LBFGS(), precision=precision)
    ws = m._make_resident(x, y)
    out = torch.empty(n + 2, dtype=torch.float64, device="cuda")
    steps = 10
    for _ in range(3): nat.check(nat.lib.opl_mlp_eval_device(ws.pointer, dev.ptr(w), 0.0, None, 1, dev.ptr(out)))
    torch.cuda.synchronize()
    nat.check(nat.lib.opl_mlp_profile(ws.pointer, 1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): nat.check(nat.lib.opl_mlp_eval_device(ws.pointer, dev.ptr(w), 0.0, None, 1, dev.ptr(out)))
    e1.record(); torch.cuda.synchronize()
    cap = 64 * steps
    kinds = (ctypes.c_int32 * cap)(); ms = (ctypes.c_float * cap)(); count = ctypes.c_int32(0)
    nat.check(nat.lib.opl_mlp_profile_read(ws.pointer, cap, kinds, ms, ctypes.byref(count)))
    nat.check(nat.lib.opl_mlp_profile(ws.pointer, 0))
    agg = {}
    for i in range(count.value): agg[int(kinds[i])] = agg.get(int(kinds[i]), 0.0) + float(ms[i]) / steps
    total = e0.elapsed_time(e1) / steps
    print("%-9s rows=%d shape=%s: %.3f ms/eval (%.0f evals/s)  " % (precision, rows, shape, total, 1e3 / total) +
          ", ".join("%s L%d %.3f" % (NAMES.get(k // 100, "?"), k % 100, v) for k, v in sorted(agg.items())), flush=True)
    del m, ws
