"""Host-array forms of the transfer / error plugin methods, executed by the device row kernel.

Used when a user calls ``SoftmaxTransfer()(x)`` or ``CrossEntropyError().derivative(a, b)``
directly on NumPy arrays (reference: learning/transfer.py, learning/error.py).  Inputs may be
1-D (one pattern) or 2-D (patterns in rows), like the reference's.
"""
import ctypes

import numpy

from . import _native as nat


def _as_matrix(array):
    host = numpy.ascontiguousarray(array, dtype=numpy.float64)
    if host.ndim == 1:
        return host.reshape(1, -1), True
    if host.ndim == 2:
        return host, False
    raise ValueError('expected a vector or a matrix, got shape %s' % (host.shape,))


def _p(array):
    return array.ctypes.data_as(ctypes.c_void_p)


def transfer_apply(kernel_id, param, input_vec):
    x, vector = _as_matrix(input_vec)
    out = numpy.empty_like(x)
    nat.check(nat.lib.opl_transfer_apply_host(kernel_id, float(param), _p(x), x.shape[0], x.shape[1], _p(out)))
    return out[0] if vector else out


def transfer_backprop(kernel_id, param, input_vec, upstream):
    x, vector = _as_matrix(input_vec)
    g, _ = _as_matrix(upstream)
    if g.shape != x.shape:
        raise ValueError('upstream shape %s != input shape %s' % (g.shape, x.shape))
    out = numpy.empty_like(x)
    nat.check(nat.lib.opl_transfer_backprop_host(kernel_id, float(param), _p(x), _p(g), x.shape[0], x.shape[1],
                                                 _p(out)))
    return out[0] if vector else out


def transfer_derivative(kernel_id, param, input_vec):
    """Diagonal Jacobian of an element-wise transfer = backprop of an all-ones upstream."""
    x = numpy.asarray(input_vec, dtype=numpy.float64)
    return transfer_backprop(kernel_id, param, x, numpy.ones(x.shape))


def softmax_jacobian(output):
    y, vector = _as_matrix(output)
    jac = numpy.empty((y.shape[0], y.shape[1], y.shape[1]))
    nat.check(nat.lib.opl_softmax_jacobian_host(_p(y), y.shape[0], y.shape[1], _p(jac)))
    return jac[0] if vector else jac


def _error(kernel_id, tensor_a, tensor_b, want_derivative):
    a, vector = _as_matrix(tensor_a)
    b, _ = _as_matrix(tensor_b)
    if a.shape != b.shape:
        raise ValueError('shape mismatch %s vs %s' % (a.shape, b.shape))
    err = ctypes.c_double(0.0)
    deriv = numpy.empty_like(a) if want_derivative else None
    nat.check(nat.lib.opl_error_host(kernel_id, _p(a), _p(b), a.shape[0], a.shape[1], ctypes.byref(err),
                                     _p(deriv) if want_derivative else None))
    if want_derivative:
        return err.value, (deriv[0] if vector else deriv)
    return err.value


def error_value(kernel_id, tensor_a, tensor_b):
    return _error(kernel_id, tensor_a, tensor_b, False)


def error_value_derivative(kernel_id, tensor_a, tensor_b):
    return _error(kernel_id, tensor_a, tensor_b, True)


def penalty(kind, weight, weight_tensor, want_derivative):
    """(penalty value, gradient or None) of an L1 / L2 weight penalty, computed on the device."""
    import torch
    from . import _device as dev
    w = dev.to_device(numpy.asarray(weight_tensor, dtype=numpy.float64).ravel())
    grad = torch.zeros_like(w) if want_derivative else None
    value = ctypes.c_double(0.0)
    nat.check(nat.lib.opl_vec_penalty_device(w.numel(), dev.ptr(w), int(kind), float(weight), dev.ptr(grad),
                                             ctypes.byref(value), dev.stream_ptr()))
    out = dev.to_host(grad).reshape(numpy.shape(weight_tensor)) if want_derivative else None
    return value.value, out
