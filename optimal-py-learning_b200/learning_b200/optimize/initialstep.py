"""Initial step-size rules for the line searches (reference: optimize/initialstep.py:34-256).

Host scalars only; the one n-vector operation (grad . direction) runs on the device and is
shared with the line search through ``directional_slope`` so it is computed once per step.
"""
import logging

import numpy

from .. import _device as dev

BIGGEST = 1.79769313e+308


def directional_slope(jac_xk, step_dir, problem=None):
    """grad_f(x_k) . p_k, memoised on the problem object for the current (jac, dir) pair."""
    memo = getattr(problem, '_slope_memo', None) if problem is not None else None
    if memo is not None and memo[0] is jac_xk and memo[1] is step_dir:
        return memo[2]
    value = dev.dot(_as_vector(jac_xk), _as_vector(step_dir))
    if problem is not None:
        try:
            problem._slope_memo = (jac_xk, step_dir, value)
        except AttributeError:
            pass
    return value


def _as_vector(value):
    return value if dev.is_device_vector(value) else dev.to_device(value)


def _guard(step, who):
    """nan -> 1e-10, negative -> 1, inf -> 1.797e308, then must be > 0 (initialstep.py:153-178)."""
    if numpy.isnan(step):
        logging.warning('nan in jacobian of objective, in %s call, returning 1e-10', who)
        step = 1e-10
    if step < 0:
        logging.warning('Negative initial step in %s call, defaulting to 1. objective value may have '
                        'increased or step direction is negative.', who)
        step = 1.0
    if numpy.isinf(step):
        logging.warning('inf step size, in %s call, returning 1.79769313e+308', who)
        step = BIGGEST
    assert step > 0
    return step


class InitialStepGetter(object):
    """Returns the first trial step of a line search (initialstep.py:34-60)."""

    def reset(self):
        pass

    def __call__(self, xk, obj_xk, jac_xk, step_dir, problem):
        raise NotImplementedError()

    def update(self, step_size):
        pass


class IncrPrevStep(InitialStepGetter):
    """Previous accepted step times ``incr_rate``, clamped (initialstep.py:66-113)."""

    def __init__(self, incr_rate=1.05, lower_bound=0, upper_bound=1.0):
        if incr_rate < 1.0:
            raise ValueError('incr_rate > 1 to increment')
        if upper_bound is not None and upper_bound < 0:
            raise ValueError('upper_bound must be positive')
        self._incr_rate = incr_rate
        self._lower_bound = lower_bound
        self._upper_bound = float('inf') if upper_bound is None else upper_bound
        self._prev_step_size = 1.0

    def reset(self):
        self._prev_step_size = 1.0

    def __call__(self, xk, obj_xk, jac_xk, step_dir, problem):
        trial = self._incr_rate * self._prev_step_size
        assert trial > 0
        return max(self._lower_bound, min(self._upper_bound, trial))

    def update(self, step_size):
        self._prev_step_size = step_size


class FOChangeInitialStep(InitialStepGetter):
    """alpha_0 = alpha_{k-1} (g_{k-1}.p_{k-1}) / (g_k.p_k)  (initialstep.py:116-187)."""

    def __init__(self):
        self._prev_step_size = None
        self._prev_jac_dot_dir = None

    def reset(self):
        self._prev_step_size = None
        self._prev_jac_dot_dir = None

    def __call__(self, xk, obj_xk, jac_xk, step_dir, problem):
        slope = directional_slope(jac_xk, step_dir, problem)
        if self._prev_step_size is None:
            step = 1.0
        elif slope == 0.0:
            logging.warning('jac_dot_dir == 0 in FOChangeInitialStep, defaulting to 1')
            step = 1.0
        else:
            with numpy.errstate(all='ignore'):
                step = self._prev_step_size * (numpy.float64(self._prev_jac_dot_dir) / slope)
        step = _guard(step, 'FOChangeInitialStep')
        self._prev_jac_dot_dir = slope
        return step

    def update(self, step_size):
        self._prev_step_size = step_size


class QuadraticInitialStep(InitialStepGetter):
    """alpha_0 = 2 (f_k - f_{k-1}) / (g_k.p_k)  (initialstep.py:190-256)."""

    def __init__(self):
        self._prev_obj_value = None

    def reset(self):
        self._prev_obj_value = None

    def __call__(self, xk, obj_xk, jac_xk, step_dir, problem):
        if self._prev_obj_value is None:
            step = 1.0
        else:
            slope = directional_slope(jac_xk, step_dir, problem)
            if slope == 0.0:
                logging.warning('jac_dot_dir == 0 in QuadraticInitialStep, defaulting to 1')
                step = 1.0
            else:
                with numpy.errstate(all='ignore'):
                    step = (2.0 * (obj_xk - self._prev_obj_value)) / numpy.float64(slope)
        step = _guard(step, 'QuadraticInitialStep')
        self._prev_obj_value = obj_xk
        return step
