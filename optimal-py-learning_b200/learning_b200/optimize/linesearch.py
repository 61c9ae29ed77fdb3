"""Step-size plugins (reference: optimize/linesearch.py:37-447).

The searches are pure host control flow over a one-dimensional function
``phi(alpha) -> (objective, slope)``; every probe is one fused device evaluation at
``x + alpha*p`` returning two scalars (``Problem.ray_obj_jac``), or -- for a Problem without
fused probes -- ``get_obj_jac(x + alpha*p)`` followed by a device dot.

Reference quirks kept on purpose (they decide which iterate sequence is produced):
  * the Armijo test of the bracketing phase uses the slope at the NEW point (linesearch.py:281);
    zoom uses phi'(0) (:360);
  * growth factor 1.5, bisection zoom, both loops give up at their 100th probe;
  * NaN objective short-circuits to 1e-10 (Wolfe) / 1e-25 (backtracking).
"""
import logging

import numpy

from .. import _device as dev
from .initialstep import IncrPrevStep, QuadraticInitialStep, directional_slope, _as_vector

WOLFE_INCR_RATE = 1.5


class StepSizeGetter(object):
    """Returns a step size when called; used by Optimizer (linesearch.py:37-56)."""

    def reset(self):
        pass

    def __call__(self, xk, obj_xk, jac_xk, step_dir, problem):
        raise NotImplementedError()


class SetStepSize(StepSizeGetter):
    """Constant step (linesearch.py:62-82)."""

    def __init__(self, step_size):
        super(SetStepSize, self).__init__()
        self._step_size = step_size

    def __call__(self, xk, obj_xk, jac_xk, step_dir, problem):
        return self._step_size


# ----------------------------------------------------------------------------------------
# probes
# ----------------------------------------------------------------------------------------
def _ray_obj_jac(problem, xk, step_dir):
    """phi(alpha) -> (f(x + alpha p), grad f(x + alpha p) . p)   (linesearch.py:392-399)."""
    fused = getattr(problem, 'ray_obj_jac', None)
    if fused is not None:
        return lambda alpha: fused(xk, alpha, step_dir)
    x, p = _as_vector(xk), _as_vector(step_dir)

    def phi(alpha):
        value, grad = problem.get_obj_jac(dev.axpy(x, alpha, p))
        return value, dev.dot(_as_vector(grad), p)
    return phi


def _ray_obj(problem, xk, step_dir):
    fused = getattr(problem, 'ray_obj', None)
    if fused is not None:
        return lambda alpha: fused(xk, alpha, step_dir)
    x, p = _as_vector(xk), _as_vector(step_dir)
    return lambda alpha: problem.get_obj(dev.axpy(x, alpha, p))


# ----------------------------------------------------------------------------------------
# scalar searches
# ----------------------------------------------------------------------------------------
def _wolfe_bracket(phi, phi0, dphi0, c_1, c_2, first_step):
    """Bracketing phase (Numerical Optimization 2nd ed. alg. 3.5 as the reference codes it)."""
    if numpy.isnan(phi0):
        logging.warning('nan objective value in _line_search_wolfe, defaulting to 1e-10 step size')
        return 1e-10
    last_step, last_value = 0.0, phi0
    step = first_step
    probe = 0
    while True:
        probe += 1
        if probe >= 100:
            logging.info('Wolfe line search aborting after 100 iterations')
            return step
        value, slope = phi(step)
        worse_than_last = probe > 1 and value >= last_value
        if worse_than_last or value > phi0 + c_1 * step * slope:
            return _wolfe_zoom(phi, last_step, last_value, step, phi0, dphi0, c_1, c_2)
        if abs(slope) <= -c_2 * dphi0:
            return step
        if slope >= 0:
            return _wolfe_zoom(phi, step, value, last_step, phi0, dphi0, c_1, c_2)
        last_step, last_value = step, value
        step *= WOLFE_INCR_RATE


def _wolfe_zoom(phi, low, low_value, high, phi0, dphi0, c_1, c_2):
    """Bisection zoom between a good (low objective) and a bad end (linesearch.py:320-389)."""
    probe = 0
    while True:
        probe += 1
        step = _bisect_value(min(low, high), max(low, high))
        assert step >= 0
        if probe >= 100:
            logging.info('Wolfe line search (zoom) aborting after 100 iterations')
            return step
        value, slope = phi(step)
        if value > phi0 + c_1 * step * dphi0 or value >= low_value:
            high = step
            continue
        if abs(slope) <= -c_2 * dphi0:
            return step
        if slope * (high - low) >= 0:
            high = low
        low, low_value = step, value


def _bisect_value(min_, max_):
    return min_ + 0.5 * (max_ - min_)


def _backtrack(phi_obj, phi0, dphi0, c_1, first_step, decr_rate):
    """Shrink until the Armijo rule holds (linesearch.py:179-224)."""
    if numpy.isnan(phi0):
        logging.warning('nan objective value in _backtracking_line_search, defaulting to 1e-25 step size')
        return 1e-25
    step = first_step
    while True:
        if step < 1e-25:
            logging.info('_backtracking_line_search failed Armijo with step_size ~= 1e-25, returning')
            return step
        if phi_obj(step) <= phi0 + (c_1 * step) * dphi0:
            assert step > 0
            return step
        step *= decr_rate


# ----------------------------------------------------------------------------------------
# reference-named function forms (vector arguments), kept for callers / tests that use them
# ----------------------------------------------------------------------------------------
def _line_search_wolfe(parameters, obj_xk, jac_xk, step_dir, obj_jac_func, c_1, c_2, initial_step):
    x, p = _as_vector(parameters), _as_vector(step_dir)

    def phi(alpha):
        value, grad = obj_jac_func(dev.axpy(x, alpha, p))
        return value, dev.dot(_as_vector(grad), p)
    return _wolfe_bracket(phi, obj_xk, dev.dot(_as_vector(jac_xk), p), c_1, c_2, initial_step)


def _backtracking_line_search(parameters, obj_xk, jac_xk, step_dir, obj_func, c_1, initial_step, decr_rate=0.9):
    x, p = _as_vector(parameters), _as_vector(step_dir)
    return _backtrack(lambda alpha: obj_func(dev.axpy(x, alpha, p)), obj_xk, dev.dot(_as_vector(jac_xk), p),
                      c_1, initial_step, decr_rate)


def _armijo_rule(step_size, obj_xk, jac_xk, step_dir, obj_xk_plus_ap, c_1):
    """f(x + a p) <= f(x) + c_1 a p.grad_f(x)   (linesearch.py:430-447)."""
    return obj_xk_plus_ap <= obj_xk + (c_1 * step_size) * dev.dot(_as_vector(jac_xk), _as_vector(step_dir))


def _curvature_condition(jac_xk, step_dir, jac_xk_plus_ap, c_2):
    """grad_f(x + a p).p >= c_2 grad_f(x).p   (linesearch.py:450-465)."""
    p = _as_vector(step_dir)
    return dev.dot(_as_vector(jac_xk_plus_ap), p) >= c_2 * dev.dot(_as_vector(jac_xk), p)


def _wolfe_conditions(step_size, parameters, obj_xk, jac_xk, step_dir, obj_xk_plus_ap, jac_xk_plus_ap, c_1, c_2):
    if not (0 < c_1 < c_2 < 1):
        raise ValueError('0 < c_1 < c_2 < 1')
    return bool(_armijo_rule(step_size, obj_xk, jac_xk, step_dir, obj_xk_plus_ap, c_1)
                and _curvature_condition(jac_xk, step_dir, jac_xk_plus_ap, c_2))


# ----------------------------------------------------------------------------------------
# plugins
# ----------------------------------------------------------------------------------------
class BacktrackingLineSearch(StepSizeGetter):
    """Backtracking with the Armijo rule; uses get_obj only (linesearch.py:85-131)."""

    def __init__(self, c_1=0.5, decr_rate=0.5, initial_step_getter=None):
        super(BacktrackingLineSearch, self).__init__()
        self._c_1 = c_1
        self._decr_rate = decr_rate
        if initial_step_getter is None:
            initial_step_getter = IncrPrevStep(incr_rate=2.0 / self._decr_rate - 1.0, lower_bound=0.0,
                                               upper_bound=None)
        self._initial_step_getter = initial_step_getter

    def reset(self):
        super(BacktrackingLineSearch, self).reset()
        self._initial_step_getter.reset()

    def __call__(self, xk, obj_xk, jac_xk, step_dir, problem):
        first = self._initial_step_getter(xk, obj_xk, jac_xk, step_dir, problem)
        slope0 = directional_slope(jac_xk, step_dir, problem)
        step_size = _backtrack(_ray_obj(problem, xk, step_dir), obj_xk, slope0, self._c_1, first,
                               self._decr_rate)
        self._initial_step_getter.update(step_size)
        return step_size


class WolfeLineSearch(StepSizeGetter):
    """Strong-Wolfe line search: bracketing + bisection zoom (linesearch.py:134-176)."""

    def __init__(self, c_1=1e-4, c_2=0.9, initial_step_getter=None):
        super(WolfeLineSearch, self).__init__()
        self._c_1 = c_1
        self._c_2 = c_2
        if initial_step_getter is None:
            initial_step_getter = QuadraticInitialStep()
        self._initial_step_getter = initial_step_getter

    def reset(self):
        super(WolfeLineSearch, self).reset()
        self._initial_step_getter.reset()

    def __call__(self, xk, obj_xk, jac_xk, step_dir, problem):
        first = self._initial_step_getter(xk, obj_xk, jac_xk, step_dir, problem)
        slope0 = directional_slope(jac_xk, step_dir, problem)
        step_size = _wolfe_bracket(_ray_obj_jac(problem, xk, step_dir), obj_xk, slope0, self._c_1, self._c_2,
                                   first)
        self._initial_step_getter.update(step_size)
        return step_size
