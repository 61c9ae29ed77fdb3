"""Device n-vectors for the optimizer stack.

Vectors are ``torch.float64`` CUDA tensors (torch is the allocator / stream owner); the
arithmetic is done by the kernels behind ``opl_vec_*`` so that iterates round exactly like
the reference's NumPy expressions (``x + a*d`` as two roundings, deterministic dots).
"""
import ctypes

import numpy
import torch

from . import _native as nat


def is_device_vector(value):
    return isinstance(value, torch.Tensor)


def current_stream_handle():
    """cudaStream_t of the caller's current torch stream on the current device, as an int.  (torch.cuda.current_stream()
    costs ~15 us of Python per call -- four of them were a third of a small model's train_step; the raw query is ~0.3 us.)"""
    try:
        return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())
    except AttributeError:       # a torch build without the raw query
        return torch.cuda.current_stream().cuda_stream


def stream_ptr():
    return ctypes.c_void_p(current_stream_handle())


def ptr(tensor):
    return ctypes.c_void_p(tensor.data_ptr()) if tensor is not None else None


def to_device(array):
    """Host array-like -> contiguous float64 CUDA tensor (H2D copy)."""
    nat.require_device()
    if isinstance(array, torch.Tensor):
        return array.to(device="cuda", dtype=torch.float64).contiguous()
    host = numpy.ascontiguousarray(array, dtype=numpy.float64)
    return torch.from_numpy(host).cuda()


def to_host(tensor):
    """Device tensor -> fresh NumPy array.  The download lands in page-locked memory from torch's caching host allocator
    (full PCIe rate; a pageable destination is several times slower) and the array keeps that block alive."""
    tensor = tensor.detach()
    if not tensor.is_cuda:
        return tensor.numpy()
    out = torch.empty(tensor.shape, dtype=tensor.dtype, pin_memory=True)
    out.copy_(tensor)
    return out.numpy()


def empty(n):
    nat.require_device()
    return torch.empty(int(n), dtype=torch.float64, device="cuda")


def _check_vec(*tensors):
    for t in tensors:
        if t is None:
            continue
        if t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous():
            raise TypeError("expected contiguous float64 CUDA tensors")


def axpy(x, a, d, b=0.0, p=None, out=None):
    """x + fl(a*d) [+ fl(b*p)]  (optimizer.py:97,136-141,253; linesearch.py:394)."""
    _check_vec(x, d, p)
    if out is None:
        out = torch.empty_like(x)
    nat.check(nat.lib.opl_vec_axpy_device(x.numel(), ptr(x), float(a), ptr(d), float(b), ptr(p), ptr(out),
                                          stream_ptr()))
    return out


def step(x, a, d):
    """(fl(a*d), x + fl(a*d)): the accepted step and the next iterate in one launch (optimizer.py:251-253)."""
    _check_vec(x, d)
    moved, following = torch.empty_like(x), torch.empty_like(x)
    nat.check(nat.lib.opl_vec_step_device(x.numel(), ptr(x), float(a), ptr(d), ptr(moved), ptr(following), stream_ptr()))
    # provenance: `following` is x + fl(a*d), rounded exactly like the trial point of a line-search probe at (x, a, d); a model
    # that has just evaluated that very probe may hand its result back instead of evaluating again (see MLP._make_problem)
    following.opl_origin = (x, float(a), d)
    return moved, following


def scale(a, x, out=None):
    """fl(a*x)  (prev_step = step_size * step_dir, optimizer.py:251)."""
    _check_vec(x)
    if out is None:
        out = torch.empty_like(x)
    nat.check(nat.lib.opl_vec_scale_device(x.numel(), float(a), ptr(x), ptr(out), stream_ptr()))
    return out


def dot(x, y):
    """x . y as a Python float (deterministic two-stage reduction; one host sync)."""
    _check_vec(x, y)
    if x.numel() != y.numel():
        raise ValueError("dot: length mismatch %d vs %d" % (x.numel(), y.numel()))
    out = ctypes.c_double(0.0)
    nat.check(nat.lib.opl_vec_dot_device(x.numel(), ptr(x), ptr(y), ctypes.byref(out), stream_ptr()))
    return out.value


def norm(x):
    return float(numpy.sqrt(dot(x, x)))
