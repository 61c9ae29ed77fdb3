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
