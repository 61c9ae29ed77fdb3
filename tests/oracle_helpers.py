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


# ----------------------------------------------------------------------------------------
# stateful teacher forcing (hard part (i)): run the pinned oracle, snapshot its
# optimizer state BEFORE every step, and install that state into the product's optimizer
# ----------------------------------------------------------------------------------------
ORACLE_OPTIMIZERS = {
    "bfgs": lambda: oo.BFGS(), "bfgs_scaled": lambda: oo.BFGS(scaled_identity=True),
    "bfgs_reset5": lambda: oo.BFGS(reset_every=5),
    "lbfgs": lambda: oo.LBFGS(), "sd": lambda: oo.SteepestDescent(),
    "sdm": lambda: oo.SteepestDescentMomentum(),
    "sd_backtracking": lambda: oo.SteepestDescent(oo.Backtracking()),
    "bfgs_wolfe_quadratic": lambda: oo.BFGS(oo.Wolfe()),
}


def _copy(v):
    return None if v is None else numpy.array(v, dtype=float, copy=True)


def snapshot(opt):
    """Everything the oracle optimizer carries from one step to the next (the reference's pickled attributes:
    optimizer.py:117,212-221,397-403; linesearch / initialstep state: initialstep.py:77,131-132,203)."""
    snap = {"kind": type(opt).__name__}
    initial = getattr(opt.search, "initial", None)
    snap["initial"] = dict(initial.__dict__) if initial is not None else None
    if isinstance(opt, oo.BFGS):
        snap.update(count=opt.count, prev_move=_copy(opt.prev_move), prev_grad=_copy(opt.prev_grad), h=_copy(opt.h))
    elif isinstance(opt, oo.LBFGS):
        snap.update(prev_move=_copy(opt.prev_move), prev_grad=_copy(opt.prev_grad),
                    pairs=[(_copy(s), _copy(y)) for s, y in zip(opt.s_hist, opt.y_hist)])
    elif isinstance(opt, oo.SteepestDescentMomentum):
        snap.update(prev=_copy(opt.prev))
    return snap


def oracle_trace_with_states(run, t):
    """Free-running oracle on one golden trace: [(x_k, state before step k, objective at x_k, x_{k+1}, H after step k)]."""
    shape = tuple(run["shape"])
    transfers = [(int(a), float(b)) for a, b in run["transfers"]]
    eid = int(run["error"])
    x, y = t[run["data"] + "_x"], t[run["data"] + "_y"]
    opt = ORACLE_OPTIMIZERS[run["optimizer"]]()
    cur = t[run["key"] + "_w0"].copy()
    f = lambda v: mo.objective_gradient(v, shape, transfers, eid, x, y)
    fo = lambda v: mo.objective(v, shape, transfers, eid, x, y)
    out = []
    with numpy.errstate(all='ignore'):
        for _ in range(run["steps"]):
            before = snapshot(opt)
            val, nxt = opt.step(f, fo, cur)
            out.append((cur.copy(), before, val, nxt.copy(), _copy(getattr(opt, "h", None))))
            cur = nxt
    return out


def install_state(L, optimizer, snap):
    """Put an oracle snapshot into the product's optimizer (through its public state-loading calls)."""
    o = L.optimize
    search = optimizer._step_size_getter
    initial = getattr(search, "_initial_step_getter", None)
    if snap["initial"] is not None and initial is not None:
        src = snap["initial"]
        if isinstance(initial, o.FOChangeInitialStep):
            initial._prev_step_size, initial._prev_jac_dot_dir = src["prev_step"], src["prev_slope"]
        elif isinstance(initial, o.QuadraticInitialStep):
            initial._prev_obj_value = src["prev_obj"]
        elif isinstance(initial, o.IncrPrevStep):
            initial._prev_step_size = src["prev"]
        else:
            raise TypeError(type(initial))
    if snap["kind"] == "BFGS":
        optimizer.load_state(snap["count"], snap["prev_move"], snap["prev_grad"], snap["h"])
    elif snap["kind"] == "LBFGS":
        optimizer.load_state(snap["prev_move"], snap["prev_grad"], snap["pairs"])
    elif snap["kind"] == "SteepestDescentMomentum":
        from learning_b200 import _device as dev
        optimizer._prev_step = None if snap["prev"] is None else dev.to_device(snap["prev"])


def restore(opt, snap):
    """Put a snapshot back into an ORACLE optimizer (inverse of ``snapshot``)."""
    initial = getattr(opt.search, "initial", None)
    if snap["initial"] is not None and initial is not None:
        initial.__dict__.update(snap["initial"])
    if snap["kind"] == "BFGS":
        opt.count, opt.prev_move, opt.prev_grad, opt.h = snap["count"], _copy(snap["prev_move"]), _copy(snap["prev_grad"]), _copy(snap["h"])
    elif snap["kind"] == "LBFGS":
        opt.prev_move, opt.prev_grad = _copy(snap["prev_move"]), _copy(snap["prev_grad"])
        opt.s_hist = [_copy(s) for s, _ in snap["pairs"]]
        opt.y_hist = [_copy(y) for _, y in snap["pairs"]]
    elif snap["kind"] == "SteepestDescentMomentum":
        opt.prev = _copy(snap["prev"])
    return opt


def reference_step_sensitivity(run, t, record, trials=4, seed=0):
    """How far the REFERENCE's own x_{k+1} moves when its arithmetic is merely re-ordered: the same step, from the same
    state, on the same patterns presented in a different ROW ORDER (objective and gradient are sums over the patterns, so a
    permutation changes nothing but the order in which the host BLAS adds them up -- what a different BLAS build or thread
    count does as well).  Quasi-Newton steps late in an ill-conditioned run subtract nearly equal gradients
    (y = g_k - g_{k-1}) of a nearly cancelled sum, which amplifies that rounding by orders of magnitude: this amplification,
    not 1e-10, is the attainable agreement there.
    Returns the largest relative change of the step x_{k+1} - x_k over `trials` random row orders."""
    shape = tuple(run["shape"])
    transfers = [(int(a), float(b)) for a, b in run["transfers"]]
    eid = int(run["error"])
    x, y = t[run["data"] + "_x"], t[run["data"] + "_y"]
    x_k, state, _, x_next, _ = record
    rng = numpy.random.default_rng(seed)
    worst = 0.0
    base = x_next - x_k
    for _ in range(trials):
        perm = rng.permutation(x.shape[0])
        xp, yp = numpy.ascontiguousarray(x[perm]), numpy.ascontiguousarray(y[perm])
        opt = restore(ORACLE_OPTIMIZERS[run["optimizer"]](), state)
        with numpy.errstate(all='ignore'):
            _, nxt = opt.step(lambda v: mo.objective_gradient(v, shape, transfers, eid, xp, yp),
                              lambda v: mo.objective(v, shape, transfers, eid, xp, yp), numpy.array(x_k, dtype=float))
        den = numpy.linalg.norm(base)
        worst = max(worst, float(numpy.linalg.norm((nxt - x_k) - base) / (den if den > 0 else 1.0)))
    return worst
