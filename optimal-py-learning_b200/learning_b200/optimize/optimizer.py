"""Optimizers (reference: optimize/optimizer.py:33-508) driving device-resident n-vectors.

Plugin API unchanged: ``Optimizer.next(problem, parameters) -> (objective, next_parameters)``,
``reset()``, attributes ``jacobian`` / ``hessian``.  ``parameters`` may be

  * a float64 CUDA ``torch.Tensor`` -- the native mode used by MLP.train_step: all vectors stay
    on the device, the problem's callables take and return device vectors;
  * a NumPy array -- a user-supplied Problem whose callables work on NumPy arrays: vectors are
    uploaded once per call, the optimizer state (H, histories) still lives on the device and
    NumPy comes back.

All n-vector and n x n arithmetic runs in the CUDA kernels behind include/opl_b200.h; there
is no NumPy arithmetic here.
"""
import ctypes
import pickle

import numpy
import torch

from .. import _device as dev
from .. import _native as nat
from .initialstep import FOChangeInitialStep
from .linesearch import WolfeLineSearch


def make_optimizer(num_parameters):
    """BFGS up to 500 parameters, L-BFGS beyond (optimizer.py:33-43)."""
    if num_parameters > 500:
        return LBFGS()
    return BFGS()


def _default_step_size_getter():
    # c_1 = 1e-4, c_2 = 0.9: Numerical Optimization 2nd ed. p. 161 (optimizer.py:204-210)
    return WolfeLineSearch(c_1=1e-4, c_2=0.9, initial_step_getter=FOChangeInitialStep())


class _HostProblem(object):
    """Adapts a Problem over NumPy arrays to device vectors (one D2H + one H2D per call)."""
    ray_obj_jac = None
    ray_obj = None

    def __init__(self, problem):
        self._problem = problem

    def get_obj(self, x):
        return self._problem.get_obj(dev.to_host(x))

    def get_obj_jac(self, x):
        value, grad = self._problem.get_obj_jac(dev.to_host(x))
        return value, dev.to_device(grad)


class Optimizer(object):
    """Optimizer for optimizing model parameters (optimizer.py:49-63)."""

    def __init__(self):
        self.jacobian = None   # last computed gradient (same kind as the parameters passed to next)
        self.hessian = None

    def reset(self):
        self.jacobian = None
        self.hessian = None

    def next(self, problem, parameters):
        """Return (objective at ``parameters``, next parameters)."""
        if dev.is_device_vector(parameters):
            return self._next(problem, parameters)
        value, following = self._next(_HostProblem(problem), dev.to_device(parameters))
        if self.jacobian is not None and dev.is_device_vector(self.jacobian):
            self.jacobian = dev.to_host(self.jacobian)
        return value, dev.to_host(following)

    def _next(self, problem, x):
        raise NotImplementedError()

    # device handles never travel through pickle / deepcopy (base.py:350-373)
    def __getstate__(self):
        state = dict(self.__dict__)
        for key, value in list(state.items()):
            if isinstance(value, torch.Tensor):
                state[key] = dev.to_host(value)
            elif isinstance(value, _LazyDeviceVector):
                state[key] = value.host
        return state


class SteepestDescent(Optimizer):
    """x - alpha * grad, alpha from the step-size getter (optimizer.py:73-97)."""

    def __init__(self, step_size_getter=None):
        super(SteepestDescent, self).__init__()
        if step_size_getter is None:
            step_size_getter = WolfeLineSearch(initial_step_getter=FOChangeInitialStep())
        self._step_size_getter = step_size_getter

    def reset(self):
        super(SteepestDescent, self).reset()
        self._step_size_getter.reset()

    def _next(self, problem, x):
        value, grad = problem.get_obj_jac(x)
        self.jacobian = grad
        downhill = dev.scale(-1.0, grad)
        step_size = self._step_size_getter(x, value, grad, downhill, problem)
        return value, dev.axpy(x, -step_size, grad)      # parameters - step_size * jacobian


class SteepestDescentMomentum(Optimizer):
    """Steepest descent plus ``momentum_rate`` times the previous step, added AFTER the line
    search (optimizer.py:100-145)."""

    def __init__(self, step_size_getter=None, momentum_rate=0.2):
        super(SteepestDescentMomentum, self).__init__()
        if step_size_getter is None:
            step_size_getter = WolfeLineSearch(initial_step_getter=FOChangeInitialStep())
        self._step_size_getter = step_size_getter
        self._momentum_rate = momentum_rate
        self._prev_step = None

    def reset(self):
        super(SteepestDescentMomentum, self).reset()
        self._step_size_getter.reset()
        self._prev_step = None

    def __setstate__(self, state):
        self.__dict__.update(state)
        if isinstance(self._prev_step, numpy.ndarray):       # the reference pickles _prev_step too (optimizer.py:117)
            self._prev_step = _LazyDeviceVector(self._prev_step)

    def _next(self, problem, x):
        self._prev_step = _materialise(self._prev_step)
        value, grad = problem.get_obj_jac(x)
        self.jacobian = grad
        downhill = dev.scale(-1.0, grad)
        step_size = self._step_size_getter(x, value, grad, downhill, problem)
        step = dev.scale(step_size, downhill)
        if self._prev_step is None:
            following = dev.axpy(x, 1.0, step)
        else:
            following = dev.axpy(x, 1.0, step, self._momentum_rate, self._prev_step)
        self._prev_step = step
        return value, following


class _LazyDeviceVector(object):
    """A host vector restored from pickle that becomes a device vector on first use (unpickling must not need a GPU)."""

    def __init__(self, host):
        self.host = host


def _materialise(value):
    return dev.to_device(value.host) if isinstance(value, _LazyDeviceVector) else value


# ----------------------------------------------------------------------------------------
# BFGS
# ----------------------------------------------------------------------------------------
def initial_hessian_identity(param_diff, jacobian, previous_jacobian):
    """Identity seed (optimizer.py:148-151).  As a BFGS argument it selects the seed kernel
    mode; the matrix itself is never formed on the training path."""
    return numpy.identity(jacobian.shape[0])


def initial_hessian_scaled_identity(param_diff, jacobian, previous_jacobian):
    """gamma * I with gamma = (s.y)/(y.y) (optimizer.py:154-167)."""
    s = dev.to_device(param_diff)
    y = dev.axpy(dev.to_device(jacobian), -1.0, dev.to_device(previous_jacobian))
    return numpy.diag(numpy.repeat(initial_hessian_gamma_scalar(s, y), jacobian.shape[0]))


_SEED_MODES = {initial_hessian_identity: 0, initial_hessian_scaled_identity: 1}


# H larger than this is not carried through pickle (config 3's H is 55 GB); see BFGS.__getstate__
MAX_PICKLED_HESSIAN_BYTES = 1 << 30


class _NativeHandle(object):
    """Owns one C-ABI object; destroyed with the Python object, never pickled."""

    def __init__(self, pointer, destroy, set_stream=None):
        self.pointer = pointer
        self._destroy = destroy
        self._set_stream = set_stream
        self._stream = dev.current_stream_handle()

    def bind_current_stream(self):
        """The handle's kernels run on the caller's CURRENT stream: re-bind when it changed since the last call."""
        stream = dev.current_stream_handle()
        if stream != self._stream and self._set_stream is not None:
            nat.check(self._set_stream(self.pointer, ctypes.c_void_p(stream)))
            self._stream = stream

    def __del__(self):
        try:
            if self.pointer:
                self._destroy(self.pointer)
                self.pointer = None
        except Exception:
            pass


class BFGS(Optimizer):
    """Quasi-Newton BFGS (optimizer.py:170-285) on a device-resident inverse Hessian.

    The reference rebuilds H with two dense n x n products per iteration (4 n^3 flop); here the
    identical rank-2 update is applied in place and lazily by one fused HBM-bound pass per
    iteration (csrc/bfgs.cu).  ``shard_rows=True`` splits the rows of H over the ranks of an
    initialised torch.distributed process group (one GPU per rank).
    """

    def __init__(self, step_size_getter=None, initial_hessian_func=initial_hessian_identity,
                 iterations_per_reset=100, shard_rows=False):
        super(BFGS, self).__init__()
        if step_size_getter is None:
            step_size_getter = _default_step_size_getter()
        self._step_size_getter = step_size_getter
        if initial_hessian_func not in _SEED_MODES:
            raise TypeError('initial_hessian_func must be initial_hessian_identity or '
                            'initial_hessian_scaled_identity on the CUDA path (no CPU fallback)')
        self._initial_hessian_func = initial_hessian_func
        self._iterations_per_reset = iterations_per_reset
        self._shard_rows = shard_rows
        self._iteration = 0
        self._prev_step = None
        self._prev_jacobian = None
        self._native = None
        self._restore_hessian = None     # host H waiting to be uploaded (after unpickling / load_state)

    def reset(self):
        super(BFGS, self).reset()
        self._step_size_getter.reset()
        self._iteration = 0
        self._prev_step = None
        self._prev_jacobian = None
        self._restore_hessian = None
        if self._native is not None:
            nat.check(nat.lib.opl_bfgs_reset(self._native.pointer))

    def __getstate__(self):
        """The reference pickles ``_prev_step``, ``_prev_jacobian`` and ``_prev_inv_hessian`` with the model
        (optimizer.py:212-221, base.py:350-360), so a model restored mid-training continues the same iterate sequence.
        Here they come down as host arrays (Optimizer.__getstate__ converts the vectors; H is materialised) and go back
        up on the first step after the restore.  An H above MAX_PICKLED_HESSIAN_BYTES (or a row-sharded one) is not
        carried: the restored optimizer then starts over like a reset one (iteration counter included), which the
        reference would too after its own ``reset()``."""
        state = super(BFGS, self).__getstate__()
        native, n = self._native, getattr(self, '_native_n', 0)
        state['_native'] = None
        state.pop('_peer_exchange', None)
        state['_restore_hessian'] = getattr(self, '_restore_hessian', None)
        if native is not None and nat.lib.opl_bfgs_state(native.pointer) != 0:
            if self._world()[1] == 1 and 8 * n * n <= MAX_PICKLED_HESSIAN_BYTES:
                state['_restore_hessian'] = self.inverse_hessian()
            else:
                state['_prev_step'] = state['_prev_jacobian'] = state['_restore_hessian'] = None
                state['_iteration'] = 0
                state['_step_size_getter'] = pickle.loads(pickle.dumps(self._step_size_getter))
                state['_step_size_getter'].reset()
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        for key in ('_prev_step', '_prev_jacobian'):          # back to device vectors on first use
            if isinstance(self.__dict__.get(key), numpy.ndarray):
                self.__dict__[key] = _LazyDeviceVector(self.__dict__[key])

    def load_state(self, iteration, prev_step, prev_jacobian, inverse_hessian):
        """Install an explicit optimizer state (host arrays; ``inverse_hessian`` None before the first update): what a
        restore from pickle does, and what the teacher-forced parity tests use to start a step from the reference's
        state (optimizer.py:212-221 attributes)."""
        self._iteration = int(iteration)
        self._prev_step = None if prev_step is None else dev.to_device(prev_step)
        self._prev_jacobian = None if prev_jacobian is None else dev.to_device(prev_jacobian)
        if self._native is not None:
            nat.check(nat.lib.opl_bfgs_reset(self._native.pointer))
        self._restore_hessian = None if inverse_hessian is None else numpy.ascontiguousarray(inverse_hessian, dtype=float)

    # ---- native state -------------------------------------------------------------------
    def _world(self):
        if self._shard_rows and torch.distributed.is_available() and torch.distributed.is_initialized():
            return torch.distributed.get_rank(), torch.distributed.get_world_size()
        return 0, 1

    def _row_range(self, n):
        rank, world = self._world()
        per = -(-n // world)
        return min(n, rank * per), min(n, (rank + 1) * per), per

    def _state(self, n):
        if self._native is None or self._native_n != n:
            begin, end, _ = self._row_range(n)
            out = ctypes.c_void_p()
            nat.check(nat.lib.opl_bfgs_create(ctypes.byref(out), n, begin, end, dev.stream_ptr()))
            self._native = _NativeHandle(out, nat.lib.opl_bfgs_destroy, nat.lib.opl_bfgs_set_stream)
            self._native_n = n
        self._native.bind_current_stream()
        pending = getattr(self, '_restore_hessian', None)
        if pending is not None:                    # a restored / explicitly loaded H goes up once
            self._restore_hessian = None
            begin, end, _ = self._row_range(n)
            if pending.shape != (n, n):
                raise ValueError('restored inverse Hessian has shape %s, expected (%d, %d)' % (pending.shape, n, n))
            h = dev.to_device(pending[begin:end])
            nat.check(nat.lib.opl_bfgs_load_device(self._native.pointer, dev.ptr(h)))
            torch.cuda.current_stream().synchronize()     # h may be freed after this call
        return self._native.pointer

    def _direction(self, grad):
        """-(H.dot(jac)) with H from _get_approx_inv_hessian (optimizer.py:242-285)."""
        n = grad.numel()
        self._prev_step = _materialise(self._prev_step)
        self._prev_jacobian = _materialise(self._prev_jacobian)
        handle = self._state(n)
        seed = _SEED_MODES[self._initial_hessian_func]
        direction = torch.empty_like(grad)
        rank, world = self._world()
        if world == 1:
            nat.check(nat.lib.opl_bfgs_direction_device(handle, dev.ptr(grad), dev.ptr(self._prev_step),
                                                        dev.ptr(self._prev_jacobian), seed, dev.ptr(direction)))
            return direction
        # row-sharded H: local pass, ONE exchange of the three n-vectors, redundant finish on every rank
        begin, end, per = self._row_range(n)
        width = per * world
        handle_state = nat.lib.opl_bfgs_state(handle)
        exchange = self._prev_step is not None and handle_state != 0          # same decision on every rank
        peer = self._exchange(2 * width + n) if exchange else None
        if peer is not None:
            # The pass writes THIS rank's slices of u = H y and w = H g and its partial v = y^T H straight into a
            # symmetric (peer-mapped) vector [u ; w ; v] whose other entries stay zero; one launch of the peer all-reduce
            # kernel (csrc/collective.cu: NVSwitch multimem.ld_reduce / multimem.st, or rank-ordered pulls) then IS the
            # all-gather of u and w (x + 0 + .. + 0 is exact) and the all-reduce of v -- instead of three NCCL collectives
            # and two staging copies per iteration.
            mine = peer.next_slot()
            u, w, v = mine[:width], mine[width:2 * width], mine[2 * width:]
        else:
            u = torch.zeros(width, dtype=torch.float64, device=grad.device)
            w = torch.zeros_like(u)
            v = torch.zeros(n, dtype=torch.float64, device=grad.device)
        nat.check(nat.lib.opl_bfgs_pass_device(handle, dev.ptr(grad), dev.ptr(self._prev_step),
                                               dev.ptr(self._prev_jacobian), seed, dev.ptr(u), dev.ptr(v),
                                               dev.ptr(w)))
        if exchange:
            if peer is not None:
                total = torch.empty(2 * width + n, dtype=torch.float64, device=grad.device)
                peer.reduce(total)
                u, w, v = total[:width], total[width:2 * width], total[2 * width:]
            else:
                torch.distributed.all_gather_into_tensor(u, u[begin:begin + per].clone())
                torch.distributed.all_gather_into_tensor(w, w[begin:begin + per].clone())
                torch.distributed.all_reduce(v)
        nat.check(nat.lib.opl_bfgs_finish_device(handle, dev.ptr(grad), dev.ptr(self._prev_step),
                                                 dev.ptr(self._prev_jacobian), seed, dev.ptr(u), dev.ptr(v),
                                                 dev.ptr(w), dev.ptr(direction)))
        return direction

    def _exchange(self, count):
        """Peer-memory all-reduce for the [u ; w ; v] exchange of the row-sharded pass (None: NCCL collectives)."""
        state = getattr(self, '_peer_exchange', False)
        if state is False or (state is not None and state.count != count):
            from .. import _collective
            state = _collective.make(count)
            self._peer_exchange = state
        return state

    def _offer_slope(self, problem, grad, direction):
        """The single-launch direction kernel of small problems also forms grad . direction: hand it to the line search
        (initialstep.directional_slope memo), which would otherwise launch a dot product for phi'(0)."""
        if self._native is None or self._world()[1] != 1:
            return
        value = ctypes.c_double(0.0)
        if nat.lib.opl_bfgs_last_slope(self._native.pointer, ctypes.byref(value)) == nat.OPL_OK:
            try:
                problem._slope_memo = (grad, direction, value.value)
            except AttributeError:
                pass

    def inverse_hessian(self):
        """The reference's ``_prev_inv_hessian`` as a NumPy array (local rows when sharded);
        None before the second iteration."""
        pending = getattr(self, '_restore_hessian', None)
        if pending is not None:
            return pending
        if self._native is None or nat.lib.opl_bfgs_state(self._native.pointer) == 0:
            return None
        self._native.bind_current_stream()
        n = self._native_n
        begin, end, _ = self._row_range(n)
        out = torch.empty((end - begin, n), dtype=torch.float64, device='cuda')
        nat.check(nat.lib.opl_bfgs_materialize_device(self._native.pointer, dev.ptr(out)))
        return dev.to_host(out)

    # ---- Optimizer API --------------------------------------------------------------------
    def _next(self, problem, x):
        self._iteration += 1
        if self._iteration == self._iterations_per_reset:     # optimizer.py:236-238
            self.reset()
        value, grad = problem.get_obj_jac(x)
        self.jacobian = grad
        direction = self._direction(grad)
        self._prev_jacobian = grad
        self._offer_slope(problem, grad, direction)
        step_size = self._step_size_getter(x, value, grad, direction, problem)
        self._prev_step, following = dev.step(x, step_size, direction)
        return value, following


def _bfgs_eq(H_k, s_k, y_k):
    """Next inverse Hessian from (H_k, s_k, y_k) on host arrays (optimizer.py:288-338), computed by
    the device rank-2 pass.  Returns H_k itself when y.s == 0."""
    h = dev.to_device(numpy.asarray(H_k, dtype=float))
    s, y = dev.to_device(s_k), dev.to_device(y_k)
    n = s.numel()
    out = ctypes.c_void_p()
    nat.check(nat.lib.opl_bfgs_create(ctypes.byref(out), n, 0, n, dev.stream_ptr()))
    handle = _NativeHandle(out, nat.lib.opl_bfgs_destroy)
    nat.check(nat.lib.opl_bfgs_load_device(handle.pointer, dev.ptr(h)))
    zero = torch.zeros_like(y)
    scratch = torch.empty_like(y)
    # grad := y, prev_grad := 0  =>  the state's y is y_k; the direction output is discarded
    nat.check(nat.lib.opl_bfgs_direction_device(handle.pointer, dev.ptr(y), dev.ptr(s), dev.ptr(zero), 0,
                                                dev.ptr(scratch)))
    result = torch.empty((n, n), dtype=torch.float64, device='cuda')
    nat.check(nat.lib.opl_bfgs_materialize_device(handle.pointer, dev.ptr(result)))
    return dev.to_host(result)


# ----------------------------------------------------------------------------------------
# L-BFGS
# ----------------------------------------------------------------------------------------
def initial_hessian_one_scalar(param_diff, jac_diff):
    return 1.0


def initial_hessian_gamma_scalar(param_diff, jac_diff):
    """(s.y)/(y.y); 1.0 when y.y == 0 (optimizer.py:346-362)."""
    s = param_diff if dev.is_device_vector(param_diff) else dev.to_device(param_diff)
    y = jac_diff if dev.is_device_vector(jac_diff) else dev.to_device(jac_diff)
    yy = dev.dot(y, y)
    if yy == 0:
        return 1.0
    with numpy.errstate(all='ignore'):
        return float(numpy.nan_to_num(numpy.float64(dev.dot(s, y)) / yy))


_GAMMA_MODES = {initial_hessian_one_scalar: 0, initial_hessian_gamma_scalar: 1}


class LBFGS(Optimizer):
    """Limited-memory BFGS (optimizer.py:365-508); the two-loop recursion is one persistent
    cooperative kernel over the device-resident (s, y) history (csrc/lbfgs.cu)."""

    def __init__(self, step_size_getter=None, num_remembered_iterations=5,
                 initial_hessian_scalar_func=initial_hessian_gamma_scalar):
        super(LBFGS, self).__init__()
        if step_size_getter is None:
            step_size_getter = _default_step_size_getter()
        self._step_size_getter = step_size_getter
        if initial_hessian_scalar_func not in _GAMMA_MODES:
            raise TypeError('initial_hessian_scalar_func must be initial_hessian_one_scalar or '
                            'initial_hessian_gamma_scalar on the CUDA path (no CPU fallback)')
        self._initial_hessian_scalar_func = initial_hessian_scalar_func
        self._num_remembered_iterations = num_remembered_iterations
        self._prev_step = None
        self._prev_jacobian = None
        self._native = None
        self._restore_history = None     # host (s, y) pairs waiting to be uploaded (after unpickling / load_state)

    def reset(self):
        super(LBFGS, self).reset()
        self._step_size_getter.reset()
        self._prev_step = None
        self._prev_jacobian = None
        self._restore_history = None
        if self._native is not None:
            nat.check(nat.lib.opl_lbfgs_reset(self._native.pointer))

    def __getstate__(self):
        """The (s, y) history, ``_prev_step`` and ``_prev_jacobian`` travel through pickle as host arrays, like the
        reference's ``_prev_param_diffs`` / ``_prev_jac_diffs`` (optimizer.py:397-403)."""
        state = super(LBFGS, self).__getstate__()
        state['_native'] = None
        state['_restore_history'] = self.history()
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        for key in ('_prev_step', '_prev_jacobian'):
            if isinstance(self.__dict__.get(key), numpy.ndarray):
                self.__dict__[key] = _LazyDeviceVector(self.__dict__[key])

    def history(self):
        """[(s_i, y_i)] oldest first, as host arrays: the reference's ``_prev_param_diffs`` / ``_prev_jac_diffs``."""
        pending = getattr(self, '_restore_history', None)
        if pending is not None:
            return list(pending)
        if self._native is None:
            return []
        self._native.bind_current_stream()
        pairs = []
        n = self._native_n
        for i in range(nat.lib.opl_lbfgs_history_len(self._native.pointer)):
            s, y = dev.empty(n), dev.empty(n)
            nat.check(nat.lib.opl_lbfgs_get_pair_device(self._native.pointer, i, dev.ptr(s), dev.ptr(y)))
            pairs.append((dev.to_host(s), dev.to_host(y)))
        return pairs

    def load_state(self, prev_step, prev_jacobian, pairs):
        """Install an explicit optimizer state (host arrays): restore from pickle / teacher-forced parity tests."""
        self._prev_step = None if prev_step is None else dev.to_device(prev_step)
        self._prev_jacobian = None if prev_jacobian is None else dev.to_device(prev_jacobian)
        if self._native is not None:
            nat.check(nat.lib.opl_lbfgs_reset(self._native.pointer))
        self._restore_history = [(numpy.asarray(s, dtype=float), numpy.asarray(y, dtype=float)) for s, y in pairs]

    def _state(self, n):
        if self._native is None or self._native_n != n:
            memory = self._num_remembered_iterations
            memory = 0 if memory == float('inf') else int(memory)
            out = ctypes.c_void_p()
            nat.check(nat.lib.opl_lbfgs_create(ctypes.byref(out), n, memory, dev.stream_ptr()))
            self._native = _NativeHandle(out, nat.lib.opl_lbfgs_destroy, nat.lib.opl_lbfgs_set_stream)
            self._native_n = n
        self._native.bind_current_stream()
        pending = getattr(self, '_restore_history', None)
        if pending:
            for s, y in pending:
                sd, yd = dev.to_device(s), dev.to_device(y)
                nat.check(nat.lib.opl_lbfgs_push_pair_device(self._native.pointer, dev.ptr(sd), dev.ptr(yd)))
            torch.cuda.current_stream().synchronize()        # the staging tensors may be freed after this call
        self._restore_history = None
        return self._native.pointer

    def _step_direction(self, grad):
        """_update_diffs + _lbfgs_step_dir (optimizer.py:433-493)."""
        self._prev_step = _materialise(self._prev_step)
        self._prev_jacobian = _materialise(self._prev_jacobian)
        handle = self._state(grad.numel())
        direction = torch.empty_like(grad)
        nat.check(nat.lib.opl_lbfgs_direction_device(handle, dev.ptr(grad), dev.ptr(self._prev_step),
                                                     dev.ptr(self._prev_jacobian),
                                                     _GAMMA_MODES[self._initial_hessian_scalar_func],
                                                     dev.ptr(direction)))
        return direction

    def _next(self, problem, x):
        value, grad = problem.get_obj_jac(x)
        self.jacobian = grad
        direction = self._step_direction(grad)
        self._prev_jacobian = grad
        step_size = self._step_size_getter(x, value, grad, direction, problem)
        self._prev_step, following = dev.step(x, step_size, direction)
        return value, following
