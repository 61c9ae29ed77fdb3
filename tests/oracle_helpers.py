"""Oracle-side helpers for the parity tests (tests may use oracle/; the product never does)."""
import numpy

import mlp_oracle as mo
import optimize_oracle as oo


def first_step_from(x_k, shape, transfers, eid, x, y):
    """x_{k+1} of a FRESH BFGS + Wolfe + FOChange optimizer started at x_k (teacher forcing:
    with empty state the reference takes d = -g and a first trial step of 1)."""
    opt = oo.BFGS(oo.Wolfe(1e-4, 0.9, oo.FOChange()))
    with numpy.errstate(all='ignore'):
        _, nxt = opt.step(lambda v: mo.objective_gradient(v, shape, transfers, eid, x, y),
                          lambda v: mo.objective(v, shape, transfers, eid, x, y), numpy.array(x_k, dtype=float))
    return nxt
