"""MultiLayer Perceptron whose objective / gradient evaluation runs on a B200.

Drop-in for learning/architecture/mlp.py:41-327: same constructor, same ``activate`` /
``train`` / ``train_step`` / ``reset`` / ``serialize`` behaviour, same flat parameter layout
``[bias(d1); W0.ravel(); W1.ravel(); ...]``, same weight-initialisation draw order.  The hot
functions -- ``activate``, ``_get_obj``, ``_get_obj_jac`` and the line-search probes that
call them -- are a fixed schedule of hand-written CUDA kernels behind the C-ABI
(include/opl_b200.h, csrc/mlp.cu); nothing here does NumPy arithmetic.

Host state: ``_bias_vec`` / ``_weight_matrices`` are NumPy arrays (authoritative between
training steps, user-editable as in the reference).  During ``train_step`` the flat
parameter vector lives on the device and only the objective, the slope and the gradient norm
come back per evaluation.
"""
import ctypes
import functools
import math
import operator
import os
import random

import numpy
import torch

from .. import _device as dev
from .. import _native as nat
from .. import error as error_mod
from .. import transfer as transfer_mod
from ..base import Model
from ..error import MeanSquaredError
from ..optimize import Problem, make_optimizer
from ..transfer import LinearTransfer, ReluTransfer, Transfer

INITIAL_WEIGHTS_RANGE = 0.25


def _flatten(bias_vec, weight_matrices):
    """Flat vector: bias, then each weight matrix row-major (mlp.py:313-315)."""
    return numpy.hstack([numpy.asarray(bias_vec, dtype=float)]
                        + [numpy.asarray(m, dtype=float).ravel() for m in weight_matrices])


def _unflatten_weights(flat_weights, shape):
    """Views of the flat vector as (bias, [W_l]) (mlp.py:318-327)."""
    bias_vec = flat_weights[:shape[1]]
    matrices, at = [], shape[1]
    for rows, cols in zip(shape[:-1], shape[1:]):
        matrices.append(flat_weights[at:at + rows * cols].reshape((rows, cols)))
        at += rows * cols
    return bias_vec, matrices


class _Workspace(object):
    """One opl_mlp handle plus the identity of the data currently resident in it."""

    def __init__(self, shape, transfers, error_id, max_rows, precision=0):
        widths = (ctypes.c_int64 * len(shape))(*[int(w) for w in shape])
        tids = (ctypes.c_int32 * len(transfers))(*[t for t, _ in transfers])
        tparams = (ctypes.c_double * len(transfers))(*[p for _, p in transfers])
        out = ctypes.c_void_p()
        nat.check(nat.lib.opl_mlp_create(ctypes.byref(out), len(shape), widths, tids, tparams, error_id,
                                         int(max_rows), int(precision), dev.stream_ptr()))
        self.pointer = out
        self.max_rows = int(max_rows)
        self.stream = dev.current_stream_handle()
        self.resident = None      # identity of the (X, Y) resident on the device (see _make_resident)
        self.keepalive = None     # device tensors borrowed by the workspace
        self.sharded = False
        self.scalars = (ctypes.c_double * 3)()     # probe scalars of the last evaluation: objective, slope, ||g||^2
        self.data_epoch = 0                        # bumped whenever the resident data or the dropout masks change

    def bind_current_stream(self):
        """Kernels are enqueued on the caller's CURRENT stream: re-bind when it changed since the last call."""
        stream = dev.current_stream_handle()
        if stream != self.stream:
            nat.check(nat.lib.opl_mlp_set_stream(self.pointer, ctypes.c_void_p(stream)))
            self.stream = stream

    def __del__(self):
        try:
            if self.pointer:
                nat.lib.opl_mlp_destroy(self.pointer)
                self.pointer = None
        except Exception:
            pass


class _Residency(object):
    """What a workspace holds: the very objects that were uploaded (strong references, compared with ``is`` -- an
    ``id()`` can be recycled by a new mini-batch, a reference cannot) and, for device tensors, their version counters."""

    def __init__(self, x, y, first, count):
        self.x, self.y, self.first, self.count = x, y, first, count
        self.versions = (getattr(x, '_version', None), getattr(y, '_version', None))

    def holds(self, x, y, first, count):
        return (self.x is x and self.y is y and self.first == first and self.count == count
                and self.versions == (getattr(x, '_version', None), getattr(y, '_version', None)))


class _KernelModel(Model):
    """Device plumbing shared by the models whose objective is the MLP kernel schedule (MLP,
    DropoutMLP, Linear/LogisticRegressionModel): workspace lifetime, dataset residency, row
    sharding over torch.distributed ranks and the fused evaluation probe.

    Subclasses provide ``_shape``, ``_lowered_transfers``, ``_lowered_error`` and ``_precision``.
    """

    PRECISIONS = {'f64': 0, 'f64-dmma': 0, 'tf32': 1}
    accepted_probe_memo = True   # hand the accepted line-search probe back as the next step's evaluation at x_{k+1} (same bits)
    small_path = True      # networks of <= 4096 weights evaluate as ONE kernel launch (csrc/mlp_small.inl); False: general schedule

    def _init_kernel_state(self, precision=None):
        if precision is None:
            precision = os.environ.get('OPL_PRECISION', 'f64')
        if precision not in self.PRECISIONS:
            raise ValueError("precision must be 'f64', 'f64-dmma' or 'tf32', got %r" % (precision,))
        self._precision = precision
        self._workspace = None
        self._residency_scope = None     # (X, Y) of the enclosing train() / train_step() / resident() scope
        self._last_grad_norm = None
        self.obj_jac_evaluations = 0     # evaluations actually run on the device (bench / diagnostics)
        self.obj_jac_memo_hits = 0       # evaluations at x_{k+1} answered by the accepted probe of the previous line search

    def _param_count(self):
        return self._shape[1] + sum(a * b for a, b in zip(self._shape[:-1], self._shape[1:]))

    # ------------------------------------------------------------------ device workspace
    def _workspace_for(self, rows):
        ws = self._workspace
        if ws is None or ws.max_rows < rows:
            ws = _Workspace(self._shape, self._lowered_transfers, self._lowered_error, rows,
                            self.PRECISIONS[self._precision])
            if self._precision == 'f64-dmma':
                nat.check(nat.lib.opl_mlp_set_engine(ws.pointer, 0))
            if not self.small_path:
                nat.check(nat.lib.opl_mlp_set_small_path(ws.pointer, 0))
            self._workspace = ws
        ws.bind_current_stream()
        return ws

    def _shard(self, rows):
        """(first row, row count) of this rank's shard of a batch of ``rows`` patterns."""
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world, rank = torch.distributed.get_world_size(), torch.distributed.get_rank()
            if world > 1 and rows >= world:
                per = -(-rows // world)
                first = min(rows, rank * per)
                return first, min(rows, first + per) - first, True
        return 0, rows, False

    def _make_resident(self, input_matrix, target_matrix):
        """Make (this rank's shard of) the dataset resident in the workspace.

        Residency is explicit, never guessed from the contents (the reference always reads the arrays it is given):

          * host arrays are uploaded on every call, EXCEPT inside a residency scope -- ``train()`` (between the
            ``_pre_train`` and ``_post_train`` hooks, base.py:166-182), one ``train_step`` or ``with model.resident(X, Y)``
            -- where the scope's own two objects (compared with ``is``) were uploaded when the scope was entered.
            An in-place edit of X between two scopes is therefore always seen;
          * device tensors are borrowed, keyed on the tensor objects and their ``_version`` counters: an in-place edit bumps
            the version and the digits of X are cut again.
        """
        ws = self._workspace
        scope = self._residency_scope
        if (ws is not None and scope is not None and ws.resident is not None and scope[0] is input_matrix
                and scope[1] is target_matrix and ws.resident.x is input_matrix and ws.resident.y is target_matrix
                and ws.resident.versions == (getattr(input_matrix, '_version', None), getattr(target_matrix, '_version', None))):
            ws.bind_current_stream()          # inside a scope, on the scope's own objects: validated when they were uploaded
            return ws
        if dev.is_device_vector(input_matrix):
            x, y = input_matrix, target_matrix
            if x.dim() != 2 or y.dim() != 2 or x.shape[0] != y.shape[0]:
                raise ValueError('input_matrix and target_matrix must be matrices with equal row counts')
            ws = self._workspace_for(x.shape[0])
            if ws.resident is None or not ws.resident.holds(x, y, 0, x.shape[0]):
                self._check_widths(x.shape[-1], y.shape[-1])
                token = _Residency(x, y, 0, x.shape[0])
                xc = x.to(torch.float64).contiguous()
                yc = y.to(torch.float64).contiguous()
                nat.check(nat.lib.opl_mlp_set_data_device(ws.pointer, dev.ptr(xc), dev.ptr(yc), xc.shape[0],
                                                          self._global_rows(xc.shape[0])))
                ws.keepalive = (xc, yc)
                ws.resident = token
                ws.data_epoch += 1
                ws.sharded = self._world_size() > 1
            return ws
        x = input_matrix if (isinstance(input_matrix, numpy.ndarray) and input_matrix.dtype == numpy.float64) \
            else numpy.asarray(input_matrix, dtype=numpy.float64)
        y = target_matrix if (isinstance(target_matrix, numpy.ndarray) and target_matrix.dtype == numpy.float64) \
            else numpy.asarray(target_matrix, dtype=numpy.float64)
        if x.ndim != 2 or y.ndim != 2:
            raise ValueError('the gradient path needs matrices (patterns in rows), got shapes %s and %s'
                             % (x.shape, y.shape))
        if x.shape[0] != y.shape[0]:
            raise ValueError('input_matrix has %d rows, target_matrix %d' % (x.shape[0], y.shape[0]))
        self._check_widths(x.shape[-1], y.shape[-1])
        first, count, sharded = self._shard(x.shape[0])
        ws = self._workspace_for(count)
        scope = self._residency_scope
        in_scope = scope is not None and scope[0] is input_matrix and scope[1] is target_matrix
        if in_scope and ws.resident is not None and ws.resident.holds(input_matrix, target_matrix, first, count):
            return ws
        xs = numpy.ascontiguousarray(x[first:first + count])
        ys = numpy.ascontiguousarray(y[first:first + count])
        nat.check(nat.lib.opl_mlp_set_data_host(ws.pointer, xs.ctypes.data_as(ctypes.c_void_p),
                                                ys.ctypes.data_as(ctypes.c_void_p), count, x.shape[0]))
        ws.resident = _Residency(input_matrix, target_matrix, first, count) if in_scope else None
        ws.data_epoch += 1
        ws.keepalive = None
        ws.sharded = sharded
        return ws

    def _select_patterns(self, pattern_selection_func, input_matrix, target_matrix):
        """Mini-batch of Model.stochastic_train (base.py:39-50,100-135) selected ON THE DEVICE: the whole dataset is
        uploaded once per stochastic_train call, the row numbers are drawn on the host with the reference's ``random``
        calls (same order, so seeded runs select the same rows) and ``opl_gather_rows_device`` copies the rows; a
        mini-batch never crosses PCIe.  Other selection functions run on the host as in the reference."""
        from ..base import ROW_SELECTORS
        rows_of = ROW_SELECTORS.get(pattern_selection_func)
        if (rows_of is None and isinstance(pattern_selection_func, functools.partial)
                and not pattern_selection_func.args and pattern_selection_func.func in ROW_SELECTORS):
            rows_of = functools.partial(ROW_SELECTORS[pattern_selection_func.func], **pattern_selection_func.keywords)
        if rows_of is None or dev.is_device_vector(input_matrix) or self._world_size() > 1:
            return pattern_selection_func(input_matrix, target_matrix)
        full = getattr(self, '_stochastic_full', None)
        if full is None or full[0] is not input_matrix or full[1] is not target_matrix:
            x = numpy.ascontiguousarray(input_matrix, dtype=numpy.float64)
            y = numpy.ascontiguousarray(target_matrix, dtype=numpy.float64)
            if x.ndim != 2 or y.ndim != 2 or x.shape[0] != y.shape[0]:
                return pattern_selection_func(input_matrix, target_matrix)
            full = (input_matrix, target_matrix, dev.to_device(x), dev.to_device(y))
            self._stochastic_full = full
        xd, yd = full[2], full[3]
        rows = rows_of(xd.shape[0])
        index = torch.tensor(rows, dtype=torch.int64).cuda()
        xb = torch.empty((len(rows), xd.shape[1]), dtype=torch.float64, device=xd.device)
        yb = torch.empty((len(rows), yd.shape[1]), dtype=torch.float64, device=yd.device)
        for src, dst in ((xd, xb), (yd, yb)):
            nat.check(nat.lib.opl_gather_rows_device(dev.ptr(src), src.shape[0], src.shape[1], dev.ptr(index), len(rows),
                                                     dev.ptr(dst), dev.stream_ptr()))
        return xb, yb

    def stochastic_train(self, input_matrix, target_matrix, *args, **kwargs):
        """Model.stochastic_train (base.py:100-135) with the dataset resident on the device for the whole call."""
        try:
            return super(_KernelModel, self).stochastic_train(input_matrix, target_matrix, *args, **kwargs)
        finally:
            self._stochastic_full = None

    def resident(self, input_matrix, target_matrix):
        """``with model.resident(X, Y):`` -- upload (X, Y) once; every evaluation inside the block that is handed these same
        two objects runs on the resident copy.  ``train()`` and ``train_step()`` open this scope themselves."""
        return _ResidencyScope(self, input_matrix, target_matrix)

    def _enter_scope(self, input_matrix, target_matrix):
        """A residency scope begins: upload now (edits made outside the scope are always seen)."""
        if self._workspace is not None and not dev.is_device_vector(input_matrix):
            self._workspace.resident = None
        self._make_resident(input_matrix, target_matrix)

    def _pre_train(self, input_matrix, target_matrix):
        """Model.train hook (base.py:166): the dataset becomes device-resident once per train call."""
        self._train_scope = self.resident(input_matrix, target_matrix)
        self._train_scope.__enter__()

    def _post_train(self, input_matrix, target_matrix):
        """Model.train hook (base.py:181): leave the scope, reset the optimizer (mlp.py:184-190, regression.py:126-132)."""
        scope = getattr(self, '_train_scope', None)
        if scope is not None:
            scope.__exit__(None, None, None)
            self._train_scope = None
        self._problem_cache = None
        self._optimizer.reset()

    def invalidate(self):
        """Forget what is resident: the next evaluation uploads (host arrays) or re-reads (device tensors) its data."""
        if self._workspace is not None:
            self._workspace.resident = None

    def _world_size(self):
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            return torch.distributed.get_world_size()
        return 1

    def _global_rows(self, local_rows):
        """Device-resident shards: every rank passes its own rows; N of the 1/N scale is their sum."""
        if self._world_size() == 1:
            return local_rows
        total = torch.tensor([local_rows], dtype=torch.int64, device='cuda')
        torch.distributed.all_reduce(total)
        return int(total.item())

    def _check_widths(self, attrs, outs):
        if attrs != self._shape[0]:
            raise ValueError('input_tensor attributes == %s, expected %s' % (attrs, self._shape[0]))
        if outs != self._shape[-1]:
            raise ValueError('target attributes == %s, expected %s' % (outs, self._shape[-1]))

    def _forward_host(self, flat_params, input_tensor):
        """Batched forward of host patterns with host parameters (one H2D + kernels + one D2H)."""
        if not isinstance(input_tensor, numpy.ndarray):
            input_tensor = numpy.array(input_tensor)
        if input_tensor.shape[-1] != self._shape[0]:
            raise ValueError('input_tensor attributes == %s, expected %s' % (input_tensor.shape[-1], self._shape[0]))
        x = numpy.ascontiguousarray(input_tensor, dtype=numpy.float64)
        vector = x.ndim == 1
        x2 = x.reshape(1, -1) if vector else x
        if x2.ndim != 2:
            raise ValueError('activate expects a vector or a matrix')
        ws = self._workspace_for(x2.shape[0])
        params = numpy.ascontiguousarray(flat_params, dtype=numpy.float64)
        out = numpy.empty((x2.shape[0], self._shape[-1]))
        nat.check(nat.lib.opl_mlp_activate_host(ws.pointer, params.ctypes.data_as(ctypes.c_void_p),
                                                x2.ctypes.data_as(ctypes.c_void_p), x2.shape[0],
                                                out.ctypes.data_as(ctypes.c_void_p)))
        return out[0] if vector else out

    def _probe(self, ws, x_dev, alpha, direction, want_grad):
        """Evaluate at x + alpha*direction on the device.  Returns (obj, slope, grad tensor | None)."""
        n = x_dev.numel()
        out = torch.empty(n + 2, dtype=torch.float64, device=x_dev.device)     # [gradient ; objective ; domain flag]
        scalars = ws.scalars
        d_ptr = direction.data_ptr() if direction is not None else None
        flag = 1 if want_grad else 0
        if not ws.sharded:
            # one C-ABI call per probe: evaluation + the three scalars (small networks: one kernel launch in all)
            nat.check(nat.lib.opl_mlp_probe_device(ws.pointer, x_dev.data_ptr(), alpha, d_ptr, flag, out.data_ptr(), scalars))
        else:
            peer = self._peer_allreduce(ws, n + 2) if want_grad else None
            mine = peer.next_slot() if peer is not None else out
            nat.check(nat.lib.opl_mlp_eval_device(ws.pointer, x_dev.data_ptr(), alpha, d_ptr, flag, mine.data_ptr()))
            if peer is not None:
                # batch-sharded ranks hold partial sums of [gradient ; objective ; flag]: the evaluation wrote this rank's
                # vector into peer-mapped memory, one kernel reduces every rank's vector over NVLink (rank order / in the switch)
                peer.reduce(out)
            else:
                torch.distributed.all_reduce(out if want_grad else out[n:])     # objective only / no peer memory: NCCL
            nat.check(nat.lib.opl_mlp_probe_scalars(ws.pointer, out.data_ptr(), d_ptr, flag, scalars))
        self.obj_jac_evaluations += 1
        if not want_grad:
            self._last_grad_norm = None
            return scalars[0], scalars[1], None
        self._last_grad_norm = math.sqrt(scalars[2]) if scalars[2] >= 0.0 else float('nan')
        grad = out[:n]
        grad.opl_norm = self._last_grad_norm      # ||g|| came back with the probe scalars: train_step needs no extra reduction
        # what was just evaluated, for the accepted-probe memo of _make_problem (ws.data_epoch: same resident data and masks)
        self._last_probe = (ws, ws.data_epoch, x_dev, float(alpha), direction, scalars[0], grad)
        return scalars[0], scalars[1], grad

    def _peer_allreduce(self, ws, count):
        """The workspace's peer-memory all-reduce for vectors of ``count`` doubles (None: use NCCL)."""
        state = getattr(ws, 'peer_allreduce', False)
        if state is False or (state is not None and state.count != count):
            from .. import _collective
            state = _collective.make(count)
            ws.peer_allreduce = state
        return state

    def _objective(self, parameter_vec, input_matrix, target_matrix, want_grad):
        """Shared body of _get_obj / _get_obj_jac: NumPy in -> host-buffer C-ABI call, device in -> device call."""
        ws = self._make_resident(input_matrix, target_matrix)
        if dev.is_device_vector(parameter_vec):
            obj, _, grad = self._probe(ws, parameter_vec, 0.0, None, want_grad)
            return obj, grad
        params = numpy.ascontiguousarray(parameter_vec, dtype=numpy.float64)
        if params.shape != (self._param_count(),):
            raise ValueError('parameter_vec has shape %s, expected (%d,)' % (params.shape, self._param_count()))
        self._adopt(params)
        if getattr(ws, 'sharded', False):
            obj, _, grad = self._probe(ws, dev.to_device(params), 0.0, None, want_grad)
            return obj, (dev.to_host(grad) if want_grad else None)
        obj = ctypes.c_double(0.0)
        if not want_grad:
            nat.check(nat.lib.opl_mlp_obj_host(ws.pointer, params.ctypes.data_as(ctypes.c_void_p), ctypes.byref(obj)))
            return obj.value, None
        # a FRESH array per call (the optimizers keep references to earlier gradients), page-locked so the
        # download runs at the PCIe rate straight into it; torch's caching host allocator recycles the blocks
        grad = torch.empty(params.shape[0], dtype=torch.float64, pin_memory=True).numpy()
        nat.check(nat.lib.opl_mlp_obj_jac_host(ws.pointer, params.ctypes.data_as(ctypes.c_void_p),
                                               ctypes.byref(obj), grad.ctypes.data_as(ctypes.c_void_p)))
        self.obj_jac_evaluations += 1
        return obj.value, grad

    def _adopt(self, params):
        raise NotImplementedError()

    # ------------------------------------------------------------------ pickling (base.py:350-373)
    def __getstate__(self):
        state = dict(self.__dict__)
        if state.get('_flat_dev') is not None:                  # weights travel as the reference's host arrays
            state['_bias_host'], state['_mats_host'] = _unflatten_weights(dev.to_host(state['_flat_dev']), self._shape)
        state['_flat_dev'] = None
        state['_workspace'] = None       # device buffers are rebuilt on demand
        state['_residency_scope'] = None
        state.pop('_train_scope', None)
        state.pop('_problem_cache', None)
        state.pop('_last_probe', None)
        state.pop('_stochastic_full', None)
        state.pop('_flat_stage', None)
        return state


class _ResidencyScope(object):
    """Context manager behind ``_KernelModel.resident`` (re-entrant: an inner scope on the same objects is a no-op)."""

    def __init__(self, model, x, y):
        self.model, self.x, self.y = model, x, y
        self.outer = None

    def __enter__(self):
        model = self.model
        self.outer = model._residency_scope
        if not (self.outer is not None and self.outer[0] is self.x and self.outer[1] is self.y):
            model._residency_scope = (self.x, self.y)
            try:
                model._enter_scope(self.x, self.y)
            except Exception:
                model._residency_scope = self.outer
                raise
        return model

    def __exit__(self, *exc):
        self.model._residency_scope = self.outer
        return False


class MLP(_KernelModel):
    """MultiLayer Perceptron.

    Args:
        shape: Number of inputs, followed by number of outputs of each layer.
        transfers: Optional list of transfer layers, or a single Transfer for the output layer.
            Defaults to softplus ("ReLU") hidden layers followed by a linear output.
        optimizer: Optimizer used to optimize the weight matrices
            (default: BFGS up to 500 weights, else L-BFGS; bias excluded from the count).
        error_func: ErrorFunc (default MeanSquaredError).
        jacobian_norm_break: Training ends when the gradient norm falls below this value.
        precision: the one knob this build adds.  'f64' (default): FP64 in, FP64 out, objective and
            gradient match the reference to 1e-10; the large layer GEMMs run as exact INT8 digit products
            on the tcgen05 tensor cores (Ozaki splitting), the small ones on FP64 DMMA.  'f64-dmma': every
            GEMM on FP64 DMMA.  'tf32': tcgen05 TF32 tensor-core GEMMs with FP32
            accumulation (1e-3).  Defaults to the OPL_PRECISION environment variable.
    """

    def __init__(self, shape, transfers=None, optimizer=None, error_func=None, jacobian_norm_break=1e-10,
                 precision=None):
        super(MLP, self).__init__()
        self._init_kernel_state(precision)
        hidden = len(shape) - 2
        if transfers is None:
            transfers = [ReluTransfer() for _ in range(hidden)] + [LinearTransfer()]
        elif isinstance(transfers, Transfer):
            transfers = [ReluTransfer() for _ in range(hidden)] + [transfers]
        if len(transfers) != len(shape) - 1:
            raise ValueError('Must have exactly 1 transfer between each pair of layers, and after the output')
        self._shape = shape

        # draw order matters for seeded reproducibility (mlp.py:79-82): bias first, then matrices
        self._bias_vec = self._random_weight_matrix(shape[1])
        self._weight_matrices = []
        self._setup_weight_matrices()
        self._transfers = transfers
        self._lowered_transfers = [transfer_mod.lower(t) for t in transfers]

        if optimizer is None:
            optimizer = make_optimizer(sum(functools.reduce(operator.mul, m.shape)
                                           for m in self._weight_matrices))
        self._optimizer = optimizer
        if error_func is None:
            error_func = MeanSquaredError()
        self._error_func = error_func
        self._lowered_error = error_mod.lower(error_func)
        self._jacobian_norm_break = jacobian_norm_break

        self.reset()

    # ------------------------------------------------------------------ weights
    # ``_bias_vec`` / ``_weight_matrices`` are the reference's attributes (NumPy arrays, user-editable).  Between the
    # train_steps of a training run the authoritative copy is the flat DEVICE vector the optimizer returned; the host
    # arrays are produced from it on first access (one D2H) and are authoritative from then on -- so a training loop that
    # never looks at the weights moves no parameters over PCIe, and anything that reads or assigns them sees / sets the
    # real values.
    def _host_weights(self):
        state = self.__dict__
        flat = state.get('_flat_dev')
        if flat is not None:
            state['_flat_dev'] = None
            state['_bias_host'], state['_mats_host'] = _unflatten_weights(dev.to_host(flat), self._shape)

    @property
    def _bias_vec(self):
        self._host_weights()
        return self.__dict__['_bias_host']

    @_bias_vec.setter
    def _bias_vec(self, value):
        self._host_weights()
        self.__dict__['_bias_host'] = value

    @property
    def _weight_matrices(self):
        self._host_weights()
        return self.__dict__['_mats_host']

    @_weight_matrices.setter
    def _weight_matrices(self, value):
        self._host_weights()
        self.__dict__['_mats_host'] = value

    def _setup_weight_matrices(self):
        self._weight_matrices = [self._random_weight_matrix((rows, cols))
                                 for rows, cols in zip(self._shape[:-1], self._shape[1:])]

    def _random_weight_matrix(self, shape):
        return (2 * numpy.random.random(shape) - 1) * INITIAL_WEIGHTS_RANGE

    def reset(self):
        """New random weights (matrices first, then bias: mlp.py:125-130) and a reset optimizer."""
        super(MLP, self).reset()
        self._setup_weight_matrices()
        self._bias_vec = self._random_weight_matrix(self._shape[1])
        self._optimizer.reset()

    # ------------------------------------------------------------------ forward
    def activate(self, input_tensor):
        """Model outputs for one pattern (1-D) or a matrix of patterns (mlp.py:134-163)."""
        return self._forward_host(_flatten(self._bias_vec, self._weight_matrices), input_tensor)

    # ------------------------------------------------------------------ training
    def train_step(self, input_matrix, target_matrix):
        """One optimizer step on the full batch (mlp.py:165-182); returns the objective at x_k."""
        with self.resident(input_matrix, target_matrix):
            return self._train_step_resident(input_matrix, target_matrix)

    def _train_step_resident(self, input_matrix, target_matrix):
        ws = self._make_resident(input_matrix, target_matrix)
        cached = self.__dict__.get('_problem_cache')
        if cached is not None and cached[0] is ws and cached[1] is input_matrix and cached[2] is target_matrix:
            problem = cached[3]          # the reference builds an identical Problem every step (mlp.py:171-175)
        else:
            problem = self._make_problem(ws, input_matrix, target_matrix)
            self._problem_cache = (ws, input_matrix, target_matrix, problem)
        flat = self.__dict__.get('_flat_dev')          # the previous step's result, still authoritative: no upload
        if flat is None:
            flat = self._flat_params_to_device()
        error, flat_next = self._optimizer.next(problem, flat)
        if dev.is_device_vector(flat_next):
            self.__dict__['_flat_dev'] = flat_next      # host arrays on demand (see _host_weights)
        else:
            self._bias_vec, self._weight_matrices = _unflatten_weights(flat_next, self._shape)

        jacobian = self._optimizer.jacobian
        if jacobian is None:
            self.converged = False
        else:
            norm = getattr(jacobian, 'opl_norm', None)
            if norm is None:
                norm = dev.norm(jacobian) if dev.is_device_vector(jacobian) else float(numpy.linalg.norm(jacobian))
            self.converged = norm < self._jacobian_norm_break
        return error

    def _make_problem(self, ws, input_matrix, target_matrix):
        """Problem(obj_func, obj_jac_func) over the resident workspace (mlp.py:171-175) plus the fused line-search probes."""
        probe = self._probe

        def obj_jac(xk):      # device vectors (the optimizer's native mode) go straight to the resident workspace
            if isinstance(xk, torch.Tensor):
                # The optimizers evaluate obj + grad at x_{k+1} although the line search has just evaluated that very point as
                # its accepted probe (the reference's own TODO, optimizer.py:69-72).  When xk is provably that point --
                # dev.step tagged it as x + fl(a*d) of the SAME x, a, d objects the last probe was evaluated at, on the same
                # resident data -- the probe's objective and gradient are handed back: bit-identical values, one evaluation
                # per step saved.
                origin = getattr(xk, 'opl_origin', None)
                last = self.__dict__.get('_last_probe')
                if (origin is not None and last is not None and self.accepted_probe_memo and last[0] is ws
                        and last[1] == ws.data_epoch and origin[0] is last[2] and origin[1] == last[3]
                        and origin[2] is last[4] and last[4] is not None):
                    self.obj_jac_memo_hits += 1
                    return last[5], last[6]
                value, _, grad = probe(ws, xk, 0.0, None, True)
                return value, grad
            return self._get_obj_jac(xk, input_matrix, target_matrix)

        def obj(xk):
            if isinstance(xk, torch.Tensor):
                return probe(ws, xk, 0.0, None, False)[0]
            return self._get_obj(xk, input_matrix, target_matrix)
        problem = Problem(obj_func=obj, obj_jac_func=obj_jac)
        problem.ray_obj_jac = lambda xk, alpha, direction: probe(ws, xk, alpha, direction, True)[:2]
        problem.ray_obj = lambda xk, alpha, direction: probe(ws, xk, alpha, direction, False)[0]
        return problem

    def _flat_params_to_device(self):
        """[bias ; W0 ; W1 ; ..] (mlp.py:313-315) assembled in a page-locked staging buffer and uploaded from there (the
        reference keeps the weights as host arrays, so every step starts with this upload)."""
        parts = [numpy.asarray(self._bias_vec, dtype=float).ravel()] + [numpy.asarray(m, dtype=float).ravel()
                                                                        for m in self._weight_matrices]
        n = sum(p.size for p in parts)
        stage = getattr(self, '_flat_stage', None)
        if stage is None or stage.numel() != n:
            stage = torch.empty(n, dtype=torch.float64, pin_memory=True)
            self._flat_stage = stage
        torch.cuda.current_stream().synchronize()        # an earlier upload from this buffer has been consumed
        numpy.concatenate(parts, out=stage.numpy())
        return stage.cuda(non_blocking=True)

    # ------------------------------------------------------------------ objective / gradient
    def _get_obj(self, parameter_vec, input_matrix, target_matrix):
        """Objective at ``parameter_vec`` (mlp.py:195-199)."""
        return self._objective(parameter_vec, input_matrix, target_matrix, False)[0]

    def _get_obj_jac(self, parameter_vec, input_matrix, target_matrix):
        """(objective, flat gradient) at ``parameter_vec`` (mlp.py:202-208)."""
        return self._objective(parameter_vec, input_matrix, target_matrix, True)

    def _adopt(self, params):
        """The reference leaves the model weights at the last evaluated point (mlp.py:197,204)."""
        self._bias_vec, self._weight_matrices = _unflatten_weights(params, self._shape)


def _get_active_neurons(active_probability, num_neurons):
    """0/1 vector of active neurons, at least one active (mlp.py:448-462); draws from ``random``."""
    if active_probability <= 0.0 or active_probability > 1.0:
        raise ValueError('0 < active_probability <= 1')
    active = [1.0 if random.uniform(0, 1) < active_probability else 0.0 for _ in range(num_neurons)]
    if 1.0 not in active:
        active[random.randint(0, len(active) - 1)] = 1.0
    return numpy.array(active)


class DropoutMLP(MLP):
    """MLP trained with dropout (mlp.py:330-427).

    Every train_step draws a fresh 0/1 mask for the inputs and for each hidden layer (same
    ``random`` call order as the reference).  On the device the input mask is folded into the rows
    of W0 / dW0 and the hidden masks multiply the transfer outputs in the forward GEMM epilogue;
    derivatives are evaluated on the masked outputs but not masked themselves (mlp.py:438-445).
    After training, the first ``activate`` scales every weight matrix by the hidden keep
    probability (mlp.py:374-384).
    """

    def __init__(self, shape, transfers=None, optimizer=None, error_func=None, input_active_probability=0.8,
                 hidden_active_probability=0.5):
        if optimizer is None:
            from ..optimize import SteepestDescent   # BFGS cannot track a problem that changes every step
            optimizer = SteepestDescent()
        super(DropoutMLP, self).__init__(shape, transfers, optimizer, error_func)
        self._inp_act_prob = input_active_probability
        self._hid_act_prob = hidden_active_probability
        self._during_training = False
        self._did_post_training = True

    def activate(self, input_tensor):
        if (not self._during_training) and (not self._did_post_training):
            self._post_training()
            self._did_post_training = True
        return super(DropoutMLP, self).activate(input_tensor)

    def _post_training(self):
        for i, _ in enumerate(self._weight_matrices):
            self._weight_matrices[i] = self._weight_matrices[i] * self._hid_act_prob

    def train_step(self, input_matrix, target_matrix):
        self._during_training = True
        self._did_post_training = False
        masks = [_get_active_neurons(self._inp_act_prob, self._shape[0])]
        for num_neurons in self._shape[1:-1]:
            masks.append(_get_active_neurons(self._hid_act_prob, num_neurons))
        flat = numpy.ascontiguousarray(numpy.hstack(masks), dtype=numpy.float64)
        with self.resident(input_matrix, target_matrix):
            ws = self._make_resident(input_matrix, target_matrix)
            nat.check(nat.lib.opl_mlp_set_masks_host(ws.pointer, flat.ctypes.data_as(ctypes.c_void_p), flat.size))
            ws.data_epoch += 1
            try:
                error = super(DropoutMLP, self).train_step(input_matrix, target_matrix)
            finally:
                nat.check(nat.lib.opl_mlp_set_masks_host(ws.pointer, None, 0))
                ws.data_epoch += 1
                self._during_training = False
        return error
