#!/usr/bin/env python
"""Diagnostics for the teacher-forced step comparison: per step of one golden trace, the relative error of OUR gradient at
the reference's x_k, of the direction, the accepted step sizes, and the step / iterate norms.

    python tools/diag_teacher.py tr_iris_lbfgs 36 50
"""
import json
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("optimal-py-learning_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, sub))
import learning_b200 as L  # noqa: E402
import mlp_oracle as mo  # noqa: E402
from learning_b200 import _device as dev  # noqa: E402
from oracle_helpers import ORACLE_OPTIMIZERS, install_state, oracle_trace_with_states, restore  # noqa: E402
from util import make_error, make_transfers, rel_err, specs  # noqa: E402


def main():
    key = sys.argv[1] if len(sys.argv) > 1 else "tr_iris_lbfgs"
    lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 1000)
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "meta.json")))
    t = numpy.load(os.path.join(ROOT, "tests", "golden", "traces.npz"))
    run = [r for r in meta["traces"] if r["key"] == key][0]
    shape, transfers, eid = specs(run)
    x, y = t[run["data"] + "_x"], t[run["data"] + "_y"]
    o = L.optimize
    makers = {"bfgs": o.BFGS, "lbfgs": o.LBFGS, "sd": o.SteepestDescent, "sdm": o.SteepestDescentMomentum}
    records = oracle_trace_with_states(run, t)
    print("k  |x|  |step|  e_obj  e_grad  alpha_ref  alpha_ours  e_alpha  e_dir  e_x")
    for k, (x_k, state, val, x_next, _) in enumerate(records):
        if not lo <= k < hi:
            continue
        model = L.MLP(shape, transfers=make_transfers(L, transfers), error_func=make_error(L, eid),
                      optimizer=makers[run["optimizer"]]())
        model.logging = False
        model.small_path = os.environ.get("DIAG_SMALL", "1") != "0"
        model._bias_vec = x_k[:shape[1]].copy()
        mats, at = [], shape[1]
        for a, b in zip(shape[:-1], shape[1:]):
            mats.append(x_k[at:at + a * b].reshape(a, b).copy())
            at += a * b
        model._weight_matrices = mats
        install_state(L, model._optimizer, state)
        err = model.train_step(x, y)
        got = numpy.hstack([model._bias_vec] + [m.ravel() for m in model._weight_matrices])
        g_ours = model._optimizer.jacobian
        g_ours = dev.to_host(g_ours) if dev.is_device_vector(g_ours) else g_ours
        _, g_ref = mo.objective_gradient(x_k, shape, transfers, eid, x, y)
        ref = restore(ORACLE_OPTIMIZERS[run["optimizer"]](), state)
        with numpy.errstate(all="ignore"):
            ref.step(lambda v: mo.objective_gradient(v, shape, transfers, eid, x, y),
                     lambda v: mo.objective(v, shape, transfers, eid, x, y), numpy.array(x_k, dtype=float))
        a_ref = ref.search.initial.prev_step
        a_ours = model._optimizer._step_size_getter._initial_step_getter._prev_step_size
        step_ref, step_ours = x_next - x_k, got - x_k
        d_ref = ref.prev_move / a_ref if hasattr(ref, "prev_move") else step_ref / a_ref
        prev = model._optimizer._prev_step
        d_ours = dev.to_host(prev) / a_ours if prev is not None else step_ours / a_ours
        print("%2d %.3g %.3g  %.1e %.1e  %.6g %.6g %.1e  %.1e  %.1e" % (
            k, numpy.linalg.norm(x_k), numpy.linalg.norm(step_ref), rel_err(err, val), rel_err(g_ours, g_ref),
            a_ref, a_ours, abs(a_ours - a_ref) / abs(a_ref), rel_err(d_ours, d_ref), rel_err(got, x_next)))


if __name__ == "__main__":
    main()
