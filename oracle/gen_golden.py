#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REAL reference and pin the oracle to it.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/gen_golden.py

For every case the reference (Py3-shimmed into a scratch dir by
``oracle/make_ref_shim.py``) is executed, the NumPy oracle is executed on the
same inputs, the two are asserted equal to round-off, and the reference's
inputs + outputs are written as small fixtures.  The GPU parity tests compare
the CUDA path with these fixtures (and with the oracle at larger sizes).
"""
import io
import json
import logging
import os
import random
import sys
import zlib

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_ref_shim                      # noqa: E402
import mlp_oracle as mo                   # noqa: E402
import optimize_oracle as oo              # noqa: E402

GOLDEN = os.path.join(HERE, os.pardir, "tests", "golden")


def _load_reference():
    out = make_ref_shim.build()
    sys.path.insert(0, out)
    import learning                       # noqa: F401
    return learning


def _ref_transfer(L, tid, param):
    return {mo.LINEAR: L.LinearTransfer, mo.TANH: L.TanhTransfer, mo.SOFTPLUS: L.ReluTransfer,
            mo.SOFTMAX: L.SoftmaxTransfer}[tid]() if tid != mo.GAUSSIAN else L.GaussianTransfer(param)


def _ref_error(L, eid):
    return L.MeanSquaredError() if eid == mo.MSE else L.CrossEntropyError()


def _rel(a, b):
    """Relative l2 difference; non-finite entries must match exactly (same inf sign / both nan)."""
    a, b = numpy.asarray(a, dtype=float).ravel(), numpy.asarray(b, dtype=float).ravel()
    bad = ~numpy.isfinite(b)
    if bad.any() or (~numpy.isfinite(a)).any():
        same = numpy.array_equal(a[bad], b[bad], equal_nan=True) and numpy.isfinite(a[~bad]).all()
        if not same:
            return float('inf')
        a, b = a[~bad], b[~bad]
    den = numpy.linalg.norm(b)
    return float(numpy.linalg.norm(a - b) / (den if den > 0 else 1.0))


# ----------------------------------------------------------------------------
# A. objective / gradient / activate cases
# ----------------------------------------------------------------------------
OBJGRAD_SPECS = [
    # name, shape, transfers [(id, param)], error, rows, classification, param scale
    ("lin_mse", (3, 4, 2), [(mo.SOFTPLUS, 0), (mo.LINEAR, 0)], mo.MSE, 10, False, 1.0),
    ("softplus_ce", (5, 7, 3), [(mo.SOFTPLUS, 0), (mo.SOFTPLUS, 0)], mo.CROSS_ENTROPY, 10, True, 1.0),
    ("softmax_mse", (4, 6, 5), [(mo.SOFTPLUS, 0), (mo.SOFTMAX, 0)], mo.MSE, 10, False, 1.0),
    ("softmax_ce", (6, 3, 4), [(mo.SOFTPLUS, 0), (mo.SOFTMAX, 0)], mo.CROSS_ENTROPY, 10, True, 1.0),
    ("iris_shape", (4, 2, 3), [(mo.SOFTPLUS, 0), (mo.SOFTMAX, 0)], mo.CROSS_ENTROPY, 90, True, 1.0),
    ("perceptron", (2, 1), [(mo.LINEAR, 0)], mo.MSE, 7, False, 1.0),
    ("one_row", (3, 2, 2), [(mo.TANH, 0), (mo.SOFTMAX, 0)], mo.CROSS_ENTROPY, 1, True, 1.0),
    ("tanh_gauss", (9, 8, 7, 3), [(mo.TANH, 0), (mo.GAUSSIAN, 1.7), (mo.LINEAR, 0)], mo.MSE, 33, False, 2.0),
    ("deep_mixed", (10, 9, 8, 7, 6), [(mo.SOFTPLUS, 0), (mo.TANH, 0), (mo.GAUSSIAN, 0.6), (mo.SOFTMAX, 0)],
     mo.CROSS_ENTROPY, 41, True, 3.0),
    ("softmax_hidden", (5, 6, 4), [(mo.SOFTMAX, 0), (mo.LINEAR, 0)], mo.MSE, 17, False, 4.0),
    ("wide", (70, 130, 37), [(mo.SOFTPLUS, 0), (mo.SOFTMAX, 0)], mo.CROSS_ENTROPY, 300, True, 1.0),
    ("wide3", (33, 65, 129, 10), [(mo.SOFTPLUS, 0), (mo.SOFTPLUS, 0), (mo.SOFTMAX, 0)],
     mo.CROSS_ENTROPY, 257, True, 1.0),
    # large weights: softplus overflow branch (exp -> inf -> x) and softmax underflow to exactly 0
    ("overflow", (4, 5, 3), [(mo.SOFTPLUS, 0), (mo.SOFTMAX, 0)], mo.CROSS_ENTROPY, 12, True, 4000.0),
    ("underflow_some_rows", (4, 5, 3), [(mo.SOFTPLUS, 0), (mo.SOFTMAX, 0)], mo.CROSS_ENTROPY, 12, True, 80.0),
    ("big_softplus_lin", (4, 5, 3), [(mo.SOFTPLUS, 0), (mo.LINEAR, 0)], mo.MSE, 12, False, 4000.0),
]


def gen_objgrad(L, store, meta):
    from learning.architecture import mlp as rmlp
    for k, (name, shape, tr, eid, rows, classification, scale) in enumerate(OBJGRAD_SPECS):
        rng = numpy.random.default_rng(1000 + k)
        if classification:
            x, y = mo.random_classification(rows, shape[0], shape[-1], rng)
        else:
            x, y = mo.random_regression(rows, shape[0], shape[-1], rng)
        w = mo.random_params(shape, rng) * scale
        model = rmlp.MLP(shape, transfers=[_ref_transfer(L, t, p) for t, p in tr],
                         error_func=_ref_error(L, eid), optimizer=L.optimize.SteepestDescent())
        with numpy.errstate(all='ignore'):
            ref_obj = float(model._get_obj(w.copy(), x, y))
            ref_obj2, ref_grad = model._get_obj_jac(w.copy(), x, y)
            bias, mats = rmlp._unflatten_weights(w.copy(), shape)
            model._bias_vec, model._weight_matrices = bias, mats
            ref_out = model.activate(x)
            ora_obj = float(mo.objective(w, shape, tr, eid, x, y))
            ora_obj2, ora_grad = mo.objective_gradient(w, shape, tr, eid, x, y)
            ora_cf = mo.objective_gradient(w, shape, tr, eid, x, y, literal_softmax=False)[1]
            ora_out = mo.activate(w, shape, tr, x)
        assert ref_obj == float(ref_obj2)
        assert _rel(ora_obj, ref_obj) <= 1e-14, (name, ora_obj, ref_obj)
        assert _rel(ora_grad, ref_grad) <= 1e-14, (name, _rel(ora_grad, ref_grad))
        assert _rel(ora_cf, ref_grad) <= 1e-12, (name, 'closed form', _rel(ora_cf, ref_grad))
        assert _rel(ora_out, ref_out) <= 1e-15, name
        store["og%d_x" % k], store["og%d_y" % k], store["og%d_w" % k] = x, y, w
        store["og%d_obj" % k] = numpy.array(ref_obj)
        store["og%d_grad" % k] = ref_grad
        store["og%d_out" % k] = ref_out
        meta["objgrad"].append(dict(key="og%d" % k, name=name, shape=list(shape),
                                    transfers=[[int(t), float(p)] for t, p in tr], error=int(eid)))
        print("objgrad %-16s obj=%.12g |g|=%.6g oracle-vs-ref grad rel=%.2e" %
              (name, ref_obj, numpy.linalg.norm(ref_grad), _rel(ora_grad, ref_grad)))


# ----------------------------------------------------------------------------
# B. known-answer vectors lifted from the reference's own tests
# ----------------------------------------------------------------------------
def gen_known_answers(L, meta):
    """Values asserted in the reference's tests (tests/architecture/test_mlp.py:41-66,116-126;
    tests/test_error.py:79-128; tests/test_calculate.py:162-246) -- re-evaluated on the reference."""
    from learning import calculate
    ka = {}
    ka["softmax_1000"] = calculate.softmax(numpy.array([-1000.0, 1000.0])).tolist()          # [0, 1]
    ka["softplus_0_1"] = calculate.relu(numpy.array([0.0, 1.0])).tolist()                    # [.6931,1.3133]
    ka["softplus_big"] = calculate.relu(numpy.array([1000.0, -1000.0])).tolist()
    ce = L.CrossEntropyError()
    ka["ce_exact"] = [float(ce(numpy.array([0.0, 1.0]), numpy.array([0.0, 1.0]))),
                      float(ce(numpy.array([[0.5, 0.5], [0.5, 0.5]]), numpy.array([[0.0, 1.0], [1.0, 0.0]])))]
    ka["ce_deriv_01"] = ce.derivative(numpy.array([0.0, 1.0]), numpy.array([0.0, 1.0]))[1].tolist()
    meta["known_answers"] = ka
    assert ka["softmax_1000"] == [0.0, 1.0]


# ----------------------------------------------------------------------------
# C. BFGS update / L-BFGS direction cases
# ----------------------------------------------------------------------------
def gen_quasi_newton(L, store, meta):
    from learning.optimize import optimizer as ropt
    cases = []
    for k, (n, symmetric, zero_ys) in enumerate([(2, True, False), (7, False, False), (33, True, False),
                                                 (64, False, False), (96, True, False), (5, True, True)]):
        rng = numpy.random.default_rng(2000 + k)
        h = rng.standard_normal((n, n))
        if symmetric:
            h = h.dot(h.T) / n + numpy.identity(n)
        s = rng.standard_normal(n)
        y = rng.standard_normal(n)
        if zero_ys:
            y = numpy.zeros(n)
        g = rng.standard_normal(n)
        ref = ropt._bfgs_eq(h.copy(), s.copy(), y.copy())
        lit = oo.bfgs_update_literal(h.copy(), s, y)
        rk2 = oo.bfgs_update_rank2(h.copy(), s, y)
        assert _rel(lit, ref) <= 1e-15
        assert _rel(rk2, ref) <= 1e-12, _rel(rk2, ref)
        key = "bf%d" % k
        store[key + "_h"], store[key + "_s"], store[key + "_y"], store[key + "_g"] = h, s, y, g
        store[key + "_hnext"] = ref
        store[key + "_dir"] = -(ref.dot(g))
        cases.append(dict(key=key, n=n))
        print("bfgs n=%d rank2-vs-ref rel=%.2e" % (n, _rel(rk2, ref)))
    # seeding from identity and scaled identity (optimizer.py:148-167,266-271)
    rng = numpy.random.default_rng(2100)
    n = 19
    s, y, g = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    gprev = g - y
    store["seed_s"], store["seed_y"], store["seed_g"] = s, y, g
    store["seed_ident"] = ropt._bfgs_eq(ropt.initial_hessian_identity(s, g, gprev), s, y)
    store["seed_scaled"] = ropt._bfgs_eq(ropt.initial_hessian_scaled_identity(s, g, gprev), s, y)
    store["seed_gamma"] = numpy.array(ropt.initial_hessian_gamma_scalar(s, y))
    assert abs(oo.gamma_scalar(s, y) - float(store["seed_gamma"])) < 1e-16
    meta["bfgs"] = cases

    lcases = []
    for k, (n, m, zero_pair) in enumerate([(11, 0, None), (11, 1, None), (50, 3, None), (50, 5, None),
                                           (50, 5, 2), (403, 5, None)]):
        rng = numpy.random.default_rng(2200 + k)
        S = rng.standard_normal((m, n)) * 0.1
        Y = S * rng.uniform(0.5, 2.0, size=(m, n)) + 0.01 * rng.standard_normal((m, n))
        if zero_pair is not None:
            Y[zero_pair] = 0.0
        g = rng.standard_normal(n)
        opt = ropt.LBFGS()
        opt._prev_param_diffs = [S[i].copy() for i in range(m)]
        opt._prev_jac_diffs = [Y[i].copy() for i in range(m)]
        ref = opt._lbfgs_step_dir(g.copy())
        ora = oo.lbfgs_direction([S[i] for i in range(m)], [Y[i] for i in range(m)], g)
        assert _rel(ora, ref) <= 1e-15, _rel(ora, ref)
        key = "lb%d" % k
        store[key + "_S"], store[key + "_Y"], store[key + "_g"], store[key + "_dir"] = S, Y, g, ref
        lcases.append(dict(key=key, n=n, m=m))
    meta["lbfgs"] = lcases


# ----------------------------------------------------------------------------
# D. optimizer iterate traces (free-running reference), incl. README config 1
# ----------------------------------------------------------------------------
def _trace_reference(L, make_opt, shape, tr, eid, x, y, w0, steps):
    from learning.architecture import mlp as rmlp
    model = rmlp.MLP(shape, transfers=[_ref_transfer(L, t, p) for t, p in tr],
                     error_func=_ref_error(L, eid), optimizer=L.optimize.SteepestDescent())
    evals = [0]

    def obj_jac(v):
        evals[0] += 1
        return model._get_obj_jac(v, x, y)
    problem = L.optimize.Problem(obj_func=lambda v: model._get_obj(v, x, y), obj_jac_func=obj_jac)
    opt = make_opt(L)
    xs, objs = [w0.copy()], []
    cur = w0.copy()
    for _ in range(steps):
        val, cur = opt.next(problem, cur)
        objs.append(float(val))
        xs.append(numpy.array(cur, dtype=float))
    return numpy.array(xs), numpy.array(objs), evals[0]


def _trace_oracle(make_opt, shape, tr, eid, x, y, w0, steps):
    opt = make_opt()
    xs, objs = [w0.copy()], []
    cur = w0.copy()
    f = lambda v: mo.objective_gradient(v, shape, tr, eid, x, y)
    fo = lambda v: mo.objective(v, shape, tr, eid, x, y)
    for _ in range(steps):
        val, cur = opt.step(f, fo, cur)
        objs.append(float(val))
        xs.append(numpy.array(cur, dtype=float))
    return numpy.array(xs), numpy.array(objs)


def gen_traces(L, store, meta):
    from learning import datasets, validation
    # README config 1 data: iris, 30 per class, seed 0 (examples/readme_example.py:28-68)
    random.seed(0)
    numpy.random.seed(0)
    iris = datasets.get_iris()
    train, test = validation.make_train_test_sets(*iris, train_per_class=30)
    iris_x, iris_y = numpy.array(train[0], dtype=float), numpy.array(train[1], dtype=float)
    store["iris_x"], store["iris_y"] = iris_x, iris_y
    store["iris_test_x"], store["iris_test_y"] = numpy.array(test[0], dtype=float), numpy.array(test[1], dtype=float)

    rng = numpy.random.default_rng(3000)
    mid_shape = (20, 16, 5)
    mid_tr = [(mo.SOFTPLUS, 0), (mo.SOFTMAX, 0)]
    mid_x, mid_y = mo.random_classification(500, 20, 5, rng)
    store["mid_x"], store["mid_y"] = mid_x, mid_y
    reg_shape = (6, 8, 2)
    reg_tr = [(mo.TANH, 0), (mo.LINEAR, 0)]
    reg_x, reg_y = mo.random_regression(64, 6, 2, rng)
    store["reg_x"], store["reg_y"] = reg_x, reg_y

    iris_tr = [(mo.SOFTPLUS, 0), (mo.SOFTMAX, 0)]
    runs = [
        # key, data, shape, transfers, error, steps, reference optimizer factory, oracle factory
        ("tr_iris_bfgs", "iris", (4, 2, 3), iris_tr, mo.CROSS_ENTROPY, 50,
         lambda L: L.optimize.BFGS(step_size_getter=L.optimize.WolfeLineSearch(
             initial_step_getter=L.optimize.FOChangeInitialStep())),
         lambda: oo.BFGS(oo.Wolfe(1e-4, 0.9, oo.FOChange())), "bfgs"),
        ("tr_mid_bfgs", "mid", mid_shape, mid_tr, mo.CROSS_ENTROPY, 50,
         lambda L: L.optimize.BFGS(), lambda: oo.BFGS(), "bfgs"),
        ("tr_mid_bfgs_scaled", "mid", mid_shape, mid_tr, mo.CROSS_ENTROPY, 12,
         lambda L: L.optimize.BFGS(initial_hessian_func=L.optimize.optimizer.initial_hessian_scaled_identity),
         lambda: oo.BFGS(scaled_identity=True), "bfgs_scaled"),
        # the hard reset of BFGS.next at call `iterations_per_reset` (optimizer.py:236-238), here every 5th call
        ("tr_mid_bfgs_reset", "mid", mid_shape, mid_tr, mo.CROSS_ENTROPY, 14,
         lambda L: L.optimize.BFGS(iterations_per_reset=5), lambda: oo.BFGS(reset_every=5), "bfgs_reset5"),
        ("tr_mid_lbfgs", "mid", mid_shape, mid_tr, mo.CROSS_ENTROPY, 50,
         lambda L: L.optimize.LBFGS(), lambda: oo.LBFGS(), "lbfgs"),
        ("tr_iris_lbfgs", "iris", (4, 2, 3), iris_tr, mo.CROSS_ENTROPY, 50,
         lambda L: L.optimize.LBFGS(), lambda: oo.LBFGS(), "lbfgs"),
        ("tr_reg_sd", "reg", reg_shape, reg_tr, mo.MSE, 30,
         lambda L: L.optimize.SteepestDescent(), lambda: oo.SteepestDescent(), "sd"),
        ("tr_reg_sdm", "reg", reg_shape, reg_tr, mo.MSE, 30,
         lambda L: L.optimize.SteepestDescentMomentum(), lambda: oo.SteepestDescentMomentum(), "sdm"),
        ("tr_reg_sd_backtrack", "reg", reg_shape, reg_tr, mo.MSE, 30,
         lambda L: L.optimize.SteepestDescent(step_size_getter=L.optimize.BacktrackingLineSearch()),
         lambda: oo.SteepestDescent(oo.Backtracking()), "sd_backtracking"),
        ("tr_reg_bfgs_quadratic", "reg", reg_shape, reg_tr, mo.MSE, 30,
         lambda L: L.optimize.BFGS(step_size_getter=L.optimize.WolfeLineSearch()),
         lambda: oo.BFGS(oo.Wolfe()), "bfgs_wolfe_quadratic"),
    ]
    data = {"iris": (iris_x, iris_y), "mid": (mid_x, mid_y), "reg": (reg_x, reg_y)}
    meta["traces"] = []
    for key, dname, shape, tr, eid, steps, ref_factory, ora_factory, optname in runs:
        x, y = data[dname]
        w0 = mo.random_params(shape, numpy.random.default_rng(zlib.crc32(key.encode()) % 1000 + 7))
        w0 = numpy.round(w0, 12)
        with numpy.errstate(all='ignore'):
            rxs, robjs, nevals = _trace_reference(L, ref_factory, shape, tr, eid, x, y, w0, steps)
            oxs, oobjs = _trace_oracle(ora_factory, shape, tr, eid, x, y, w0, steps)
        drift = [_rel(oxs[i], rxs[i]) for i in range(len(rxs))]
        first_bad = next((i for i, d in enumerate(drift) if d > 1e-9), None)
        print("trace %-22s steps=%d evals=%d oracle-vs-ref max drift=%.2e first>1e-9 at %s final obj=%.6g" %
              (key, steps, nevals, max(drift), first_bad, robjs[-1]))
        assert first_bad is None, (key, first_bad)
        store[key + "_xs"], store[key + "_objs"] = rxs, robjs
        store[key + "_w0"] = w0
        meta["traces"].append(dict(key=key, data=dname, shape=list(shape),
                                   transfers=[[int(t), float(p)] for t, p in tr], error=int(eid),
                                   steps=steps, optimizer=optname, reference_evals=nevals))


# ----------------------------------------------------------------------------
# E. SURVEY section 8(f) widening: DropoutMLP, Linear / Logistic regression (+ L1 / L2 penalties)
#    No separate oracle restatement: the golden vectors ARE the reference's outputs.
# ----------------------------------------------------------------------------
def gen_widening(L, store, meta):
    from learning.architecture import mlp as rmlp
    from learning.architecture import regression as rreg
    rng = numpy.random.default_rng(4000)
    meta["dropout"] = []
    for k, (shape, hidden_t, out_t, eid, rows, steps) in enumerate([
            ((5, 7, 3), "softplus", "softmax", mo.CROSS_ENTROPY, 40, 8),
            ((6, 8, 5, 2), "tanh", "linear", mo.MSE, 30, 8),
            ((4, 9, 2), "gaussian", "linear", mo.MSE, 25, 6)]):
        gen = mo.random_classification if eid == mo.CROSS_ENTROPY else mo.random_regression
        x, y = gen(rows, shape[0], shape[-1], rng)
        tmap = {"softplus": L.ReluTransfer, "tanh": L.TanhTransfer, "gaussian": L.GaussianTransfer,
                "linear": L.LinearTransfer, "softmax": L.SoftmaxTransfer}
        transfers = [tmap[hidden_t]() for _ in shape[1:-1]] + [tmap[out_t]()]
        random.seed(50 + k)
        numpy.random.seed(50 + k)
        model = rmlp.DropoutMLP(shape, transfers=transfers, error_func=_ref_error(L, eid),
                                input_active_probability=0.8, hidden_active_probability=0.6)
        model.logging = False
        weights, errors = [], []
        with numpy.errstate(all='ignore'):
            for _ in range(steps):
                errors.append(float(model.train_step(x, y)))
                weights.append(rmlp._flatten(model._bias_vec, model._weight_matrices).copy())
            out = model.activate(x)          # first activate after training: post-training weight scaling
        key = "do%d" % k
        store[key + "_x"], store[key + "_y"] = x, y
        store[key + "_weights"], store[key + "_errors"], store[key + "_out"] = numpy.array(weights), numpy.array(errors), out
        meta["dropout"].append(dict(key=key, shape=list(shape), hidden=hidden_t, out=out_t, error=int(eid), seed=50 + k,
                                    steps=steps))
        print("dropout %s steps=%d final err=%.6g" % (shape, steps, errors[-1]))

    meta["regression"] = []
    for k, (kind, attrs, outs, eid, pen, rows) in enumerate([
            ("linear", 5, 2, mo.MSE, None, 40), ("linear", 7, 3, mo.MSE, ("l1", 0.05), 33),
            ("linear", 4, 1, mo.MSE, ("l2", 0.3), 21), ("logistic", 6, 3, mo.MSE, None, 50),
            ("logistic", 5, 2, mo.CROSS_ENTROPY, ("l2", 0.01), 45), ("logistic", 3, 4, mo.MSE, ("l1", 0.02), 28)]):
        x, y = mo.random_regression(rows, attrs, outs, rng)
        if kind == "logistic":
            y = (y + 1.0) / 2.0
        penalty = None if pen is None else (L.L1Penalty if pen[0] == "l1" else L.L2Penalty)(penalty_weight=pen[1])
        cls = rreg.LinearRegressionModel if kind == "linear" else rreg.LogisticRegressionModel
        random.seed(70 + k)
        numpy.random.seed(70 + k)
        model = cls(attrs, outs, error_func=_ref_error(L, eid), penalty_func=penalty)
        model.logging = False
        w = (2.0 * rng.random((attrs + 1) * outs) - 1.0) * 0.7
        with numpy.errstate(all='ignore'):
            obj = float(model._get_obj(w.copy(), x, y))
            obj2, jac = model._get_obj_jac(w.copy(), x, y)
            model._weight_matrix = w.copy().reshape(attrs + 1, outs)
            out = model.activate(x)
            # short training trace from the seeded initial weights
            random.seed(70 + k)
            numpy.random.seed(70 + k)
            model = cls(attrs, outs, error_func=_ref_error(L, eid), penalty_func=penalty)
            model.logging = False
            w0 = model._weight_matrix.ravel().copy()
            weights, errors = [], []
            for _ in range(10):
                errors.append(float(model.train_step(x, y)))
                weights.append(model._weight_matrix.ravel().copy())
        assert abs(obj - float(obj2)) <= 1e-15 * max(1, abs(obj))
        key = "rg%d" % k
        store[key + "_x"], store[key + "_y"], store[key + "_w"] = x, y, w
        store[key + "_obj"], store[key + "_jac"], store[key + "_out"] = numpy.array(obj), jac, out
        store[key + "_w0"], store[key + "_weights"], store[key + "_errors"] = w0, numpy.array(weights), numpy.array(errors)
        meta["regression"].append(dict(key=key, kind=kind, attrs=attrs, outs=outs, error=int(eid),
                                       penalty=list(pen) if pen else None, seed=70 + k))
        print("regression %-8s pen=%s obj=%.8g |jac|=%.4g trained err=%.6g" % (kind, pen, obj, numpy.linalg.norm(jac), errors[-1]))

    # ---- SOM + RBF ----
    from learning.architecture import som as rsom
    from learning.architecture import rbf as rrbf
    random.seed(90)
    numpy.random.seed(90)
    xs = rng.random((30, 3)) * 2 - 1
    som = rsom.SOM(3, 5)
    som.logging = False
    store["som_w0"] = som._weights.copy()
    som.train(xs, xs, iterations=25)
    store["som_x"], store["som_w"] = xs, som._weights.copy()
    store["som_dist"] = som.activate(xs).copy()
    store["som_dist_vec"] = som.activate(xs[3]).copy()
    print("som: 25 passes over 30 patterns, |w| = %.6g" % numpy.linalg.norm(som._weights))

    meta["rbf"] = []
    for k, (attrs, clusters, outs, rows, scale, variance, incremental, steps) in enumerate([
            (3, 5, 2, 40, True, None, False, 8), (4, 6, 1, 35, False, 0.7, False, 8), (2, 4, 2, 30, True, None, True, 6)]):
        x, y = mo.random_regression(rows, attrs, outs, rng)
        random.seed(95 + k)
        numpy.random.seed(95 + k)
        model = rrbf.RBF(attrs, clusters, outs, variance=variance, scale_by_similarity=scale,
                         cluster_incrementally=incremental)
        model.logging = False
        w = (2.0 * rng.random(outs + clusters * outs) - 1.0) * 0.6
        with numpy.errstate(all='ignore'):
            model._pre_train(x, y)                      # SOM clustering: 1000 passes (Model.train defaults)
            centers = model._clustering_model._weights.copy()
            obj = float(model._get_obj(w.copy(), x, y))
            obj2, jac = model._get_obj_jac(w.copy(), x, y)
            out = model.activate(x)
            sim = model._similarity_tensor.copy()
            # training trace from the seeded initial weights
            random.seed(95 + k)
            numpy.random.seed(95 + k)
            model = rrbf.RBF(attrs, clusters, outs, variance=variance, scale_by_similarity=scale,
                             cluster_incrementally=incremental)
            model.logging = False
            model._pre_train(x, y)
            weights, errors = [], []
            for _ in range(steps):
                errors.append(float(model.train_step(x, y)))
                weights.append(rrbf._flatten_weights(model._weight_matrix, model._bias_vec).copy())
        assert abs(obj - float(obj2)) <= 1e-15 * max(1, abs(obj))
        key = "rbf%d" % k
        store[key + "_x"], store[key + "_y"], store[key + "_w"] = x, y, w
        store[key + "_centers"], store[key + "_obj"], store[key + "_jac"] = centers, numpy.array(obj), jac
        store[key + "_out"], store[key + "_sim"] = out, sim
        store[key + "_weights"], store[key + "_errors"] = numpy.array(weights), numpy.array(errors)
        store[key + "_centers_end"] = model._clustering_model._weights.copy()
        meta["rbf"].append(dict(key=key, attrs=attrs, clusters=clusters, outs=outs, scale=scale, variance=variance,
                                incremental=incremental, steps=steps, seed=95 + k))
        print("rbf %s obj=%.8g |jac|=%.4g trained err=%.6g" % ((attrs, clusters, outs), obj, numpy.linalg.norm(jac), errors[-1]))

    # ---- Model.stochastic_train / select_sample (base.py:39-135) ----
    tr = numpy.load(os.path.join(GOLDEN, "traces.npz"))
    random.seed(7)
    numpy.random.seed(7)
    model = rmlp.MLP((4, 6, 3), transfers=L.SoftmaxTransfer(), error_func=L.CrossEntropyError())
    model.logging = False
    with numpy.errstate(all='ignore'):
        err = model.stochastic_train(tr["iris_x"], tr["iris_y"], max_iterations=4, train_kwargs={'iterations': 6})
    store["st_error"] = numpy.array(float(err))
    store["st_weights"] = rmlp._flatten(model._bias_vec, model._weight_matrices).copy()
    store["st_iteration"] = numpy.array(model.iteration)
    print("stochastic_train: 4 x 6 steps on random iris subsets, err=%.6g" % float(err))

    # penalty known answers
    w = rng.standard_normal(9)
    store["pen_w"] = w
    store["pen_l1"] = numpy.array([L.L1Penalty(0.3)(w)])
    store["pen_l1_d"] = L.L1Penalty(0.3).derivative(w)
    store["pen_l2"] = numpy.array([L.L2Penalty(0.7)(w)])
    store["pen_l2_d"] = L.L2Penalty(0.7).derivative(w)


def main():
    logging.disable(logging.WARNING)
    L = _load_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    meta = {"objgrad": [], "generated_by": "oracle/gen_golden.py",
            "numpy": numpy.__version__, "reference": "justinlovinger/optimal-py-learning (learning 0.1.0)"}

    store = {}
    gen_objgrad(L, store, meta)
    gen_known_answers(L, meta)
    numpy.savez_compressed(os.path.join(GOLDEN, "objgrad.npz"), **store)

    store = {}
    gen_quasi_newton(L, store, meta)
    numpy.savez_compressed(os.path.join(GOLDEN, "quasi_newton.npz"), **store)

    store = {}
    gen_traces(L, store, meta)
    numpy.savez_compressed(os.path.join(GOLDEN, "traces.npz"), **store)

    store = {}
    gen_widening(L, store, meta)
    numpy.savez_compressed(os.path.join(GOLDEN, "widening.npz"), **store)

    with io.open(os.path.join(GOLDEN, "meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
    print("golden vectors written to", os.path.normpath(GOLDEN))


if __name__ == "__main__":
    main()
