"""Error functions of the MLP hot path (reference: learning/error.py:31-116).

Descriptors for the fused output-row kernel (``kernel_id`` = OPL_ERROR_*); direct calls on
host arrays go through the same kernel via the C-ABI.
"""
from . import _host_ops

MSE, CROSS_ENTROPY = 0, 1


class ErrorFunc(object):
    """An error function (error.py:31-43)."""
    kernel_id = None

    def __call__(self, tensor_a, tensor_b):
        """Return the error between two tensors (a: prediction, b: target)."""
        if self.kernel_id is None:
            raise NotImplementedError()
        return _host_ops.error_value(self.kernel_id, tensor_a, tensor_b)

    def derivative(self, tensor_a, tensor_b):
        """Return (error, derivative tensor)."""
        if self.kernel_id is None:
            raise NotImplementedError()
        return _host_ops.error_value_derivative(self.kernel_id, tensor_a, tensor_b)


class MeanSquaredError(ErrorFunc):
    """mean((a - b)^2) over all elements; derivative (a - b) * 2/size  (error.py:46-64)."""
    kernel_id = MSE


class CrossEntropyError(ErrorFunc):
    """-mean_rows(sum(log(a) * b)); log(0) -> -1.797e308, log(negative) raises
    FloatingPointError; derivative -(b/a)/rows with 0/0 -> 0, x/0 -> 1.797e308  (error.py:67-116)."""
    kernel_id = CROSS_ENTROPY


def lower(error_func):
    kid = getattr(error_func, 'kernel_id', None)
    if not isinstance(error_func, ErrorFunc) or kid is None:
        raise TypeError('%r cannot run on the CUDA path: only MeanSquaredError / CrossEntropyError '
                        'are supported (no CPU fallback)' % (error_func,))
    return int(kid)


#############################
# Penalty functions (error.py:122-193): used by the regression models
#############################
class PenaltyFunc(object):
    """A penalty on the flat weight vector: ``penalty_weight * norm(w)`` and its gradient.  The norms
    and the gradient update run on the device (opl_vec_penalty_device)."""
    kernel_kind = None

    def __init__(self, penalty_weight=1.0):
        super(PenaltyFunc, self).__init__()
        self._penalty_weight = penalty_weight

    def __call__(self, weight_tensor):
        """Return penalty of given weight tensor."""
        return _host_ops.penalty(self._kind(), self._penalty_weight, weight_tensor, False)[0]

    def derivative(self, weight_tensor, penalty_output=None):
        """Return jacobian of the penalty at the given weight tensor (penalty_output is accepted for
        signature compatibility; the device kernel recomputes the norm in the same pass)."""
        return _host_ops.penalty(self._kind(), self._penalty_weight, weight_tensor, True)[1]

    def _kind(self):
        if self.kernel_kind is None:
            raise NotImplementedError()
        return self.kernel_kind


class L1Penalty(PenaltyFunc):
    """Penalize weights by ||W||_1 (lasso); derivative sign(W)  (error.py:166-178)."""
    kernel_kind = 1


class L2Penalty(PenaltyFunc):
    """Penalize weights by ||W||_2; derivative W / ||W||_2  (error.py:181-193)."""
    kernel_kind = 2


def lower_penalty(penalty_func):
    """(kind, weight) of a built-in penalty, None for no penalty, TypeError otherwise."""
    if penalty_func is None:
        return None
    kind = getattr(penalty_func, 'kernel_kind', None)
    if not isinstance(penalty_func, PenaltyFunc) or kind is None:
        raise TypeError('%r cannot run on the CUDA path: only L1Penalty / L2Penalty are supported' % (penalty_func,))
    return int(kind), float(penalty_func._penalty_weight)
