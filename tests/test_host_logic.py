"""CPU: host-side control flow (line searches, initial steps, Problem, training loop).

The scalar searches are pure functions of phi(alpha); they are checked against the oracle's
restatement (itself pinned to the reference) on analytic 1-D problems, no GPU needed.
"""
import math
import random

import numpy
import pytest

import optimize_oracle as oo
from learning_b200.base import Model
from learning_b200.optimize import Problem, linesearch, initialstep


def _phis():
    numpy.seterr(all='ignore')
    quad = lambda a: ((a - 2.0) ** 2 + 1.0, 2.0 * (a - 2.0))
    steep = lambda a: (float(numpy.exp(numpy.float64(3 * a))) - 10 * a, 3 * float(numpy.exp(numpy.float64(3 * a))) - 10)
    wiggle = lambda a: (math.sin(5 * a) + 0.3 * a * a - a, 5 * math.cos(5 * a) + 0.6 * a - 1)
    flat = lambda a: (1.0 - 1e-9 * a, -1e-9)
    return [quad, steep, wiggle, flat]


@pytest.mark.parametrize("first", [1.0, 0.01, 7.3, 1e-6, 300.0])
def test_wolfe_search_takes_the_reference_path(first):
    for phi in _phis():
        f0, d0 = phi(0.0)
        probes_a, probes_b = [], []

        def pa(a):
            probes_a.append(a)
            return phi(a)

        def pb(a):
            probes_b.append(a)
            return phi(a)
        got = linesearch._wolfe_bracket(pa, f0, d0, 1e-4, 0.9, first)
        want = oo.wolfe_search(pb, f0, d0, 1e-4, 0.9, first)
        assert got == want
        assert probes_a == probes_b


def test_wolfe_nan_and_abort_failsafes():
    assert linesearch._wolfe_bracket(lambda a: (0.0, 0.0), float('nan'), -1.0, 1e-4, 0.9, 1.0) == 1e-10
    # never satisfiable: decreasing forever with steep slope -> aborts at the 100th probe
    calls = []
    step = linesearch._wolfe_bracket(lambda a: (calls.append(a) or -a, -1.0), 0.0, -1.0, 1e-4, 0.1, 1.0)
    assert len(calls) == 99 and step == pytest.approx(1.5 ** 99)
    assert linesearch._bisect_value(1.0, 3.0) == 2.0


def test_backtracking_matches_oracle():
    for phi in _phis():
        f0, d0 = phi(0.0)
        for first in (1.0, 3.0, 0.5):
            got = linesearch._backtrack(lambda a: phi(a)[0], f0, d0, 0.5, first, 0.5)
            want = oo.backtracking_search(lambda a: phi(a)[0], f0, d0, 0.5, first, 0.5)
            assert got == want
    assert linesearch._backtrack(lambda a: 0.0, float('nan'), -1.0, 0.5, 1.0, 0.5) == 1e-25
    # objective that never improves: shrinks below 1e-25 and gives up
    assert linesearch._backtrack(lambda a: 1.0, 0.0, -1.0, 0.5, 1.0, 0.5) < 1e-25


class _MemoProblem(object):
    """Carries a precomputed slope so the initial-step rules run without touching the device."""

    def __init__(self, jac, direction, slope):
        self._slope_memo = (jac, direction, slope)


def test_initial_step_rules_follow_reference():
    jac, direction = object(), object()
    ours, theirs = initialstep.FOChangeInitialStep(), oo.FOChange()
    rng = random.Random(3)
    for _ in range(40):
        slope = rng.choice([-1.0, -0.3, 0.0, 2.0, float('nan'), -1e-300, float('inf')]) * rng.random()
        try:
            want = theirs.first(1.0, slope)
        except AssertionError:      # a zero step trips the reference's own `assert initial_step > 0`
            with pytest.raises(AssertionError):
                ours(None, 1.0, jac, direction, _MemoProblem(jac, direction, slope))
            ours._prev_jac_dot_dir = theirs.prev_slope
            continue
        got = ours(None, 1.0, jac, direction, _MemoProblem(jac, direction, slope))
        assert got == want and got > 0
        accepted = rng.choice([1.0, 0.5, 1e-10, 3.0])
        ours.update(accepted)
        theirs.update(accepted)

    ours, theirs = initialstep.QuadraticInitialStep(), oo.Quadratic()
    for obj in [3.0, 2.5, 2.6, float("nan"), 1.0, 0.9]:
        slope = rng.choice([-1.0, 0.0, 0.5, -1e-8])
        assert ours(None, obj, jac, direction, _MemoProblem(jac, direction, slope)) == theirs.first(obj, slope)

    incr = initialstep.IncrPrevStep(incr_rate=3.0, lower_bound=0.0, upper_bound=None)
    assert incr(None, 0, None, None, None) == 3.0
    incr.update(0.25)
    assert incr(None, 0, None, None, None) == 0.75
    assert initialstep.IncrPrevStep()(None, 0, None, None, None) == 1.0   # clamped to upper_bound 1
    with pytest.raises(ValueError):
        initialstep.IncrPrevStep(incr_rate=0.5)


def test_problem_derivation_table():
    # reference tests/optimize/test_problem.py:32-116
    p = Problem(obj_func=lambda x: 'o', jac_func=lambda x: 'j', hess_func=lambda x: 'h')
    assert (p.get_obj(0), p.get_jac(0), p.get_hess(0)) == ('o', 'j', 'h')
    assert list(p.get_obj_jac(0)) == ['o', 'j'] and list(p.get_obj_jac_hess(0)) == ['o', 'j', 'h']
    p = Problem(obj_jac_func=lambda x: ('o', 'j'))
    assert p.get_obj(0) == 'o' and p.get_jac(0) == 'j' and p.get_hess(0) is None
    assert tuple(p.get_obj_jac(0)) == ('o', 'j')
    p = Problem(obj_jac_hess_func=lambda x: ('o', 'j', 'h'))
    assert p.get_obj(0) == 'o' and p.get_jac(0) == 'j' and p.get_hess(0) == 'h'
    assert list(p.get_obj_hess(0)) == ['o', 'h'] and list(p.get_jac_hess(0)) == ['j', 'h']
    p = Problem(obj_hess_func=lambda x: ('o', 'h'), jac_func=lambda x: 'j')
    assert tuple(p.get_obj_jac_hess(0)) == ('o', 'j', 'h')
    p = Problem(jac_hess_func=lambda x: ('j', 'h'), obj_func=lambda x: 'o')
    assert tuple(p.get_obj_jac_hess(0)) == ('o', 'j', 'h') and p.get_jac(0) == 'j'
    p = Problem(obj_jac_func=lambda x: ('o', 'j'), hess_func=lambda x: 'h')
    assert tuple(p.get_obj_jac_hess(0)) == ('o', 'j', 'h')
    assert Problem().get_obj(0) is None


class _Scripted(Model):
    """Returns a scripted error per train_step (role of the reference's testing/helpers.py doubles)."""

    def __init__(self, errors):
        super(_Scripted, self).__init__()
        self.errors = list(errors)
        self.logging = False

    def train_step(self, x, y):
        return self.errors[min(self.iteration - 1, len(self.errors) - 1)]


def test_train_loop_stop_rules():
    x = y = numpy.zeros((2, 1))
    m = _Scripted([1.0, 0.5, 0.001])                 # error_break
    assert m.train(x, y) == 0.001 and m.converged and m.iteration == 3
    m = _Scripted([1.0] * 50)                         # stagnation: 5 equal errors then stop
    m.train(x, y)
    assert not m.converged and m.iteration == 6
    m = _Scripted([1.0 + 0.1 * i for i in range(100)])   # no improvement for 20 iterations
    m.train(x, y)
    assert m.iteration == 21
    m = _Scripted([1.0 / (i + 1) for i in range(100)])   # iteration cap
    m.train(x, y, iterations=7, error_break=0.0)
    assert m.iteration == 7
    m = _Scripted([None])                              # models that report no error
    assert m.train(x, y, iterations=3) is None and m.iteration == 3


class _Retry(_Scripted):
    def __init__(self):
        super(_Retry, self).__init__([0.5])
        self.tag = 0
        self.resets = 0

    def reset(self):
        super(_Retry, self).reset()
        self.resets += 1
        self.tag = self.resets
        self.errors = [0.5 + 0.1 * self.resets]


def test_train_retries_keep_best_attempt():
    m = _Retry()
    err = m.train(numpy.zeros((1, 1)), numpy.zeros((1, 1)), iterations=2, retries=2)
    assert err == pytest.approx(0.7)     # last attempt's error is returned (base.py:213-221)
    assert m.tag == 0                    # ... but the best attempt (the first) was restored


def test_table_driven_softplus_logistic_header_is_accurate(tmp_path):
    """csrc/oz_transcend.cuh (the FP64 softplus / logistic of the forward GEMM epilogue: table-driven exp and log, ~26 FP64
    instructions per element) compiled for the HOST and checked against mpmath, the reference's literal form
    log(1.0 + exp(z)) (calculate.py:128-135) and 1 / (1 + exp(-z)) (:138-141), including the special cases."""
    import ctypes
    import os
    import shutil
    import subprocess
    import mpmath
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no host C++ compiler")
    here = os.path.dirname(os.path.abspath(__file__))
    csrc = os.path.join(here, "..", "optimal-py-learning_b200", "csrc")
    so = str(tmp_path / "libozt.so")
    subprocess.check_call([gxx, "-O2", "-ffp-contract=off", "-std=c++14", "-shared", "-fPIC", "-I", csrc,
                           os.path.join(here, "ozt_host_check.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    lib.ozt_softplus_logistic.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p]
    lib.ozt_softplus_logistic_dispatch.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p,
                                                   ctypes.c_void_p]

    def run(z):
        """Full form, and the epilogue's dispatch (regular-argument form per group of 8 unless the group holds a special
        value): the two must agree bit for bit wherever both are defined, so everything below holds for both."""
        z = numpy.ascontiguousarray(z, dtype=numpy.float64)
        sp, sg = numpy.empty_like(z), numpy.empty_like(z)
        lib.ozt_softplus_logistic(z.ctypes.data, z.size, sp.ctypes.data, sg.ctypes.data)
        sp2, sg2 = numpy.empty_like(z), numpy.empty_like(z)
        groups = ctypes.c_long(0)
        lib.ozt_softplus_logistic_dispatch(z.ctypes.data, z.size, sp2.ctypes.data, sg2.ctypes.data, ctypes.byref(groups))
        assert numpy.array_equal(sp, sp2, equal_nan=True) and numpy.array_equal(sg, sg2, equal_nan=True)
        run.regular_groups += groups.value
        return sp, sg
    run.regular_groups = 0

    rng = numpy.random.default_rng(0)
    z = numpy.concatenate([rng.uniform(-40, 40, 3000), rng.uniform(-2, 2, 3000), rng.uniform(-750, 750, 500),
                           rng.normal(0, 1e-3, 500), [0.0, -0.0, 1e-300, -1e-300, 707.9999, -707.9999, 708.0, -708.0,
                                                      36.7, -36.7, -745.2, 1e300, -1e300]])
    sp, sg = run(z)
    mpmath.mp.prec = 200
    for zi, a, b in zip(z, sp, sg):
        zm = mpmath.mpf(float(zi))
        want_sp = mpmath.log1p(mpmath.exp(zm)) if zi < 700 else zm
        want_sg = 1 / (1 + mpmath.exp(-zm))
        # 1 + exp(z) is rounded to a double in the reference as well: the allowance is absolute below 1
        assert abs(mpmath.mpf(float(a)) - want_sp) <= 4.5e-16 * max(1, abs(want_sp)), (zi, a)
        if zi > -708.0:
            assert abs(mpmath.mpf(float(b)) - want_sg) <= 6e-16 * want_sg, (zi, b)
        else:
            assert b == 0.0          # exp(-708) = 3e-308: flushed instead of going subnormal
    # the reference's literal NumPy form on the same inputs
    with numpy.errstate(over='ignore'):
        lit = numpy.log(1.0 + numpy.exp(z))
    lit = numpy.where(numpy.isinf(lit), z, lit)
    assert (numpy.abs(sp - lit) <= 4.5e-16 * numpy.maximum(1, numpy.abs(lit))).all()
    # special values: NaN propagates, +-inf and the overflow detour (calculate.py:132-135), softplus(0) = ln 2
    sp, sg = run([numpy.nan, numpy.inf, -numpy.inf, 0.0, 1000.0, -1000.0, 1.0])
    assert numpy.isnan(sp[0]) and numpy.isnan(sg[0])
    assert sp[1] == numpy.inf and sg[1] == 1.0 and sp[2] == 0.0 and sg[2] == 0.0
    assert sp[3] == math.log(2.0) and sg[3] == 0.5
    assert sp[4] == 1000.0 and sg[4] == 1.0 and sp[5] == 0.0 and sg[5] == 0.0
    assert abs(sp[6] - 1.3132616875182228) <= 3e-16      # test_calculate.py:223: relu([0, 1]) == [0.6931, 1.3133]
    assert run.regular_groups > 700                       # the regular-argument form really was exercised


def test_scalar_exp_log_rcp_forms_of_the_head(tmp_path):
    """The single-value forms used by the fused head's softmax + cross-entropy row stage (csrc/oz_transcend.cuh:
    ozt_exp_nonpos, ozt_log_pos, ozt_rcp), host build against mpmath: exp incl. the subnormal range and the underflow
    point, log on (0, 1] incl. subnormals and a -> 1 (relative accuracy), reciprocal."""
    import ctypes
    import os
    import shutil
    import subprocess
    import mpmath
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no host C++ compiler")
    here = os.path.dirname(os.path.abspath(__file__))
    csrc = os.path.join(here, "..", "optimal-py-learning_b200", "csrc")
    so = str(tmp_path / "libozt.so")
    subprocess.check_call([gxx, "-O2", "-ffp-contract=off", "-std=c++14", "-shared", "-fPIC", "-I", csrc,
                           os.path.join(here, "ozt_host_check.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    lib.ozt_scalar_forms.argtypes = [ctypes.c_void_p, ctypes.c_long] + [ctypes.c_void_p] * 3
    rng = numpy.random.default_rng(1)
    x = numpy.concatenate([rng.uniform(0, 50, 1500), rng.uniform(0, 1, 1500), 1 - 10.0 ** rng.uniform(-16, -1, 1000),
                           10.0 ** rng.uniform(-320, 0, 1500), rng.uniform(700, 750, 300),
                           [1.0, 0.5, 4.9e-324, 2.2250738585072014e-308, 1e-310, 745.13, 745.2, 746.0, 1000.0, numpy.inf]])
    e, l, r = numpy.empty_like(x), numpy.empty_like(x), numpy.empty_like(x)
    lib.ozt_scalar_forms(x.ctypes.data, x.size, e.ctypes.data, l.ctypes.data, r.ctypes.data)
    mpmath.mp.prec = 200
    for xi, ei, li, ri in zip(x, e, l, r):
        if numpy.isinf(xi):
            assert ei == 0.0
            continue
        xm = mpmath.mpf(float(xi))
        want = mpmath.exp(-xm)
        if want > mpmath.mpf(2) ** -1022:
            assert abs(mpmath.mpf(float(ei)) - want) <= 4e-16 * want, (xi, ei)
        else:                                   # subnormal results: within one spacing, 0 below the smallest one
            assert abs(mpmath.mpf(float(ei)) - want) <= mpmath.mpf(2) ** -1074, (xi, ei)
        if 0 < xi <= 1:
            lg = mpmath.log(xm)
            assert abs(mpmath.mpf(float(li)) - lg) <= 4e-16 * abs(lg), (xi, li)
        if 1e-300 < xi < 1e300:
            assert abs(mpmath.mpf(float(ri)) * xm - 1) <= 3e-16, (xi, ri)
    assert l[list(x).index(1.0)] == 0.0 and l[list(x).index(0.5)] == -math.log(2.0)


def test_make_train_test_sets_splits_like_the_reference():
    """validation.make_train_test_sets (validation.py:342-381): first k patterns of every label to the training set, order kept."""
    from learning_b200 import validation
    labels = numpy.eye(3)[[0, 1, 0, 2, 1, 0, 2, 2, 1, 0]]
    inputs = numpy.arange(10, dtype=float).reshape(10, 1)
    (tr_x, tr_y), (te_x, te_y) = validation.make_train_test_sets(inputs, labels, train_per_class=2)
    assert tr_x.ravel().tolist() == [0, 1, 2, 3, 4, 6] and te_x.ravel().tolist() == [5, 7, 8, 9]
    assert numpy.array_equal(tr_y, labels[[0, 1, 2, 3, 4, 6]]) and numpy.array_equal(te_y, labels[[5, 7, 8, 9]])
    with pytest.raises(ValueError):
        validation.make_train_test_sets(inputs, labels, train_per_class=5)


class _Table(Model):
    """activate() looks the pattern up: row r of `outputs` for row r of `inputs` (vector or matrix)."""

    def __init__(self, inputs, outputs, matrix_ok=True):
        super(_Table, self).__init__()
        self.inputs, self.outputs, self.matrix_ok = numpy.asarray(inputs, float), numpy.asarray(outputs, float), matrix_ok
        self.calls = 0

    def activate(self, v):
        self.calls += 1
        v = numpy.asarray(v, float)
        if v.ndim == 2:
            if not self.matrix_ok:
                raise ValueError("single patterns only")
            return numpy.array([self.outputs[numpy.where((self.inputs == r).all(axis=1))[0][0]] for r in v])
        return self.outputs[numpy.where((self.inputs == v).all(axis=1))[0][0]]


def test_get_accuracy_column_labels_and_onehot():
    """validation._get_classes (validation.py:257-271): a column matrix holds the labels themselves (NOT arg-maxed: that
    would report 1.0 for every model), a wider matrix is arg-maxed."""
    from learning_b200 import validation
    x = numpy.arange(8.0).reshape(4, 2)
    labels = numpy.array([[0.0], [1.0], [2.0], [1.0]])
    model = _Table(x, numpy.array([[0.0], [1.0], [1.0], [0.0]]))          # 2 of 4 right
    assert validation.get_accuracy(model, x, labels) == 0.5
    onehot = numpy.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 1, 0.0]])
    scores = numpy.array([[0.9, 0.1, 0], [0.2, 0.5, 0.3], [0.6, 0.3, 0.1], [0.1, 0.8, 0.1]])   # 3 of 4 right
    assert validation.get_accuracy(_Table(x, scores), x, onehot) == 0.75
    assert validation._get_classes(labels).tolist() == [0.0, 1.0, 2.0, 1.0]
    assert validation._get_classes(onehot).tolist() == [0, 1, 2, 1]


def test_print_results_is_batched_and_prints_the_reference_lines(capsys):
    """Model.print_results (base.py:378-381): one batched activate, the same lines as the per-pattern loop; models that
    take single patterns only keep the loop."""
    x = numpy.array([[0.0, 1.0], [2.0, 3.0], [4.0, 5.0]])
    y = numpy.array([[1.0], [0.0], [1.0]])
    out = numpy.array([[0.25], [0.5], [0.75]])
    want = "".join("%s -> %s (%s)\n" % (a, b, c) for a, b, c in zip(x, out, y))
    batched = _Table(x, out)
    batched.print_results(x, y)
    assert capsys.readouterr().out == want and batched.calls == 1
    single = _Table(x, out, matrix_ok=False)
    single.print_results(x, y)
    assert capsys.readouterr().out == want and single.calls == 1 + len(x)


def test_row_selectors_replay_the_reference_draws():
    """base.select_sample / select_random (base.py:39-64): the row numbers used by the device gather are the same
    ``random`` draws, in the same order, as the host selection."""
    from learning_b200 import base
    x, y = numpy.arange(40.0).reshape(20, 2), numpy.arange(20.0).reshape(20, 1)
    for select, rows_of in base.ROW_SELECTORS.items():
        for size in (None, 7):
            random.seed(11)
            xs, ys = select(x, y, size)
            random.seed(11)
            rows = rows_of(20, size)
            assert numpy.array_equal(xs, x[rows]) and numpy.array_equal(ys, y[rows])
    assert base._selection_size_heuristic(20) == min(20, int(20.0 * (math.log(30, 2) - math.log(10.0) + 1.0)))


def test_optimizer_state_pickles_without_a_device():
    """Optimizers pickle as host state only (base.py:350-373): a fresh or reset optimizer needs no GPU to round-trip, and
    restored vectors stay host arrays until the first step uses them."""
    import pickle
    from learning_b200 import optimize as o
    from learning_b200.optimize import optimizer as opt
    for make in (o.BFGS, o.LBFGS, o.SteepestDescent, o.SteepestDescentMomentum):
        clone = pickle.loads(pickle.dumps(make(), protocol=2))
        assert type(clone) is make
    b = o.BFGS()
    b.__setstate__(dict(b.__getstate__(), _prev_step=numpy.ones(3), _prev_jacobian=numpy.zeros(3), _iteration=4,
                        _restore_hessian=numpy.identity(3)))
    assert isinstance(b._prev_step, opt._LazyDeviceVector) and b._iteration == 4
    again = pickle.loads(pickle.dumps(b, protocol=2))
    assert numpy.array_equal(again._prev_step.host, numpy.ones(3)) and numpy.array_equal(again.inverse_hessian(), numpy.identity(3))
    lb = o.LBFGS()
    lb.__setstate__(dict(lb.__getstate__(), _restore_history=[(numpy.ones(2), numpy.ones(2) * 2)]))
    assert len(pickle.loads(pickle.dumps(lb, protocol=2)).history()) == 1
