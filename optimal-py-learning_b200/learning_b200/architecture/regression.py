"""Linear and logistic regression through the same device path as the MLP.

Reference: learning/architecture/regression.py:35-290.  A regression model is the single-layer case of
the MLP kernel schedule: its weight matrix ``(attributes + 1, outputs)`` with the bias in row 0 is, raveled,
exactly the MLP flat layout ``[bias ; W0]`` (regression.py:221-242 vs mlp.py:313-327).  LinearRegressionModel
uses the linear output transfer, LogisticRegressionModel the logit transfer (``calculate.logit`` /
``calculate.dlogit`` incl. its overflow quirk).  L1 / L2 weight penalties are added on the device.
"""
import ctypes
import functools
import operator

import numpy
import torch

from .. import _device as dev
from .. import _native as nat
from .. import error as error_mod
from ..error import MeanSquaredError
from ..optimize import Problem, make_optimizer
from .mlp import _KernelModel

INITIAL_WEIGHTS_RANGE = 0.25
LINEAR, LOGIT = 0, 5


class RegressionModel(_KernelModel):
    """A model that optimizes the weight matrix of an equation of a set form (regression.py:35-215)."""
    _transfer_id = None

    def __init__(self, attributes, num_outputs, optimizer=None, error_func=None, penalty_func=None,
                 jacobian_norm_break=1e-10):
        super(RegressionModel, self).__init__()
        self._init_kernel_state('f64')
        if self._transfer_id is None:
            raise NotImplementedError()
        self._shape = (attributes, num_outputs)
        self._lowered_transfers = [(self._transfer_id, 0.0)]
        self._weight_matrix = self._random_weight_matrix(self._weights_shape(attributes, num_outputs))
        if optimizer is None:
            optimizer = make_optimizer(functools.reduce(operator.mul, self._weight_matrix.shape))
        self._optimizer = optimizer
        if error_func is None:
            error_func = MeanSquaredError()
        self._error_func = error_func
        self._lowered_error = error_mod.lower(error_func)
        self._penalty_func = penalty_func
        self._lowered_penalty = error_mod.lower_penalty(penalty_func)
        self._jacobian_norm_break = jacobian_norm_break

    def reset(self):
        super(RegressionModel, self).reset()
        self._weight_matrix = self._random_weight_matrix(self._weight_matrix.shape)
        self._optimizer.reset()

    def _random_weight_matrix(self, shape):
        return (2 * numpy.random.random(shape) - 1) * INITIAL_WEIGHTS_RANGE

    def _weights_shape(self, attributes, num_outputs):
        return (attributes + 1, num_outputs)     # +1: bias row (regression.py:221-224)

    def activate(self, input_tensor):
        """Model outputs for one pattern or a matrix of patterns."""
        return self._forward_host(self._weight_matrix.ravel(), input_tensor)

    # ------------------------------------------------------------------ training
    def train_step(self, input_matrix, target_matrix):
        """One optimizer step (regression.py:101-124)."""
        with self.resident(input_matrix, target_matrix):
            return self._train_step_resident(input_matrix, target_matrix)

    def _train_step_resident(self, input_matrix, target_matrix):
        ws = self._make_resident(input_matrix, target_matrix)
        problem = Problem(obj_func=lambda xk: self._get_obj(xk, input_matrix, target_matrix),
                          obj_jac_func=lambda xk: self._get_obj_jac(xk, input_matrix, target_matrix))
        if self._lowered_penalty is None:
            problem.ray_obj_jac = lambda xk, alpha, direction: self._probe(ws, xk, alpha, direction, True)[:2]
            problem.ray_obj = lambda xk, alpha, direction: self._probe(ws, xk, alpha, direction, False)[0]
        error, flat = self._optimizer.next(problem, dev.to_device(self._weight_matrix.ravel()))
        self._weight_matrix = dev.to_host(flat).reshape(self._weight_matrix.shape)
        jacobian = self._optimizer.jacobian
        if jacobian is None:
            self.converged = False
        else:
            norm = dev.norm(jacobian) if dev.is_device_vector(jacobian) else float(numpy.linalg.norm(jacobian))
            self.converged = norm < self._jacobian_norm_break
        return error

    # ------------------------------------------------------------------ objective / gradient
    def _get_obj(self, parameter_vec, input_matrix, target_matrix):
        return self._with_penalty(parameter_vec, input_matrix, target_matrix, False)[0]

    def _get_obj_jac(self, parameter_vec, input_matrix, target_matrix):
        return self._with_penalty(parameter_vec, input_matrix, target_matrix, True)

    def _with_penalty(self, parameter_vec, input_matrix, target_matrix, want_grad):
        """error (+ penalty), jacobian (+ penalty jacobian)  (regression.py:152-180)."""
        if numpy.shape(target_matrix)[-1] != self._shape[-1]:
            raise ValueError('target_matrix.shape does not match output_matrix.shape')
        obj, grad = self._objective(parameter_vec, input_matrix, target_matrix, want_grad)
        if self._lowered_penalty is None:
            return obj, grad
        kind, weight = self._lowered_penalty
        on_device = dev.is_device_vector(parameter_vec)
        w = parameter_vec if on_device else dev.to_device(parameter_vec)
        g = None
        if want_grad:
            g = grad if on_device else dev.to_device(grad)
        value = ctypes.c_double(0.0)
        nat.check(nat.lib.opl_vec_penalty_device(w.numel(), dev.ptr(w), kind, weight, dev.ptr(g), ctypes.byref(value),
                                                 dev.stream_ptr()))
        if want_grad and not on_device:
            g = dev.to_host(g)
        return obj + value.value, g

    def _adopt(self, params):
        self._weight_matrix = params.reshape(self._weight_matrix.shape)


class LinearRegressionModel(RegressionModel):
    r"""f(x) = W x + b  (regression.py:218-242)."""
    _transfer_id = LINEAR


class LogisticRegressionModel(RegressionModel):
    r"""f(x) = 1 / (1 + e^{-(W x + b)})  (regression.py:255-290)."""
    _transfer_id = LOGIT
