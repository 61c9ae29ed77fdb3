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
