"""Transfer (activation) layers of the MLP hot path.

Same class names and call signatures as the reference (learning/transfer.py:33-130).  In the
MLP these objects are *descriptors*: the math runs as GEMM epilogues / row kernels selected by
``kernel_id`` (include/opl_b200.h OPL_TRANSFER_*).  Called directly on host arrays they run
the same device kernels through the C-ABI (no NumPy arithmetic).
"""
import numpy

from . import _host_ops

LINEAR, TANH, SOFTPLUS, GAUSSIAN, SOFTMAX = 0, 1, 2, 3, 4


class Transfer(object):
    """Base class (transfer.py:33-44).  Subclasses outside this module cannot be lowered to
    the CUDA path; MLP raises TypeError for them (there is no CPU fallback)."""
    kernel_id = None
    kernel_param = 0.0

    def __call__(self, input_vec):
        if self.kernel_id is None:
            raise NotImplementedError()
        return _host_ops.transfer_apply(self.kernel_id, self.kernel_param, input_vec)

    def derivative(self, input_vec, output_vec):
        """Diagonal of the Jacobian (same shape as the input), or for softmax the full per-row
        Jacobian (calculate.py:156-176)."""
        if self.kernel_id is None:
            raise NotImplementedError()
        return _host_ops.transfer_derivative(self.kernel_id, self.kernel_param, input_vec)


class LinearTransfer(Transfer):          # transfer.py:47-58
    kernel_id = LINEAR


class TanhTransfer(Transfer):            # transfer.py:61-72
    kernel_id = TANH


class ReluTransfer(Transfer):            # transfer.py:76-91 -- softplus, log(1 + e^x)
    """Smooth approximation of a rectified linear unit, also known as softplus."""
    kernel_id = SOFTPLUS


class GaussianTransfer(Transfer):        # transfer.py:100-116
    kernel_id = GAUSSIAN

    def __init__(self, variance=1.0):
        super(GaussianTransfer, self).__init__()
        self._variance = variance

    @property
    def kernel_param(self):
        return float(self._variance)


class SoftmaxTransfer(Transfer):         # transfer.py:119-130
    kernel_id = SOFTMAX

    def derivative(self, input_vec, output_vec):
        # the reference derives the softmax Jacobian from the OUTPUT (calculate.py:156)
        return _host_ops.softmax_jacobian(numpy.asarray(output_vec, dtype=float))


def lower(transfer):
    """(kernel id, parameter) of a Transfer instance, or TypeError."""
    kid = getattr(transfer, 'kernel_id', None)
    if not isinstance(transfer, Transfer) or kid is None:
        raise TypeError('%r cannot run on the CUDA path: only the built-in transfers are supported '
                        '(no CPU fallback)' % (transfer,))
    return int(kid), float(transfer.kernel_param)
