"""CPU oracle for the MLP objective / gradient hot path (NumPy float64).

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module; the product package (``optimal-py-learning_b200/``) never does and has
no CPU fallback.

This is a restatement (not a copy) of the reference algorithm, written as pure
functions over a *model spec* ``(shape, transfers, error)``:

  forward            learning/architecture/mlp.py:134-163
  flat layout        learning/architecture/mlp.py:313-327
  transfers          learning/transfer.py:47-130, learning/calculate.py:91-176
  errors             learning/error.py:46-116
  backprop           learning/architecture/mlp.py:213-276, 279-295

Pinned: ``oracle/gen_golden.py`` runs the real reference (Py3-shimmed, scratch
dir) beside these functions and the committed ``tests/golden/*.npz`` hold the
reference's outputs; ``tests/test_oracle.py`` re-checks the oracle against them.

Transfer ids / error ids are the same integers the C-ABI uses
(``include/opl_b200.h``).
"""
import numpy

# transfer ids (include/opl_b200.h: OPL_TRANSFER_*)
LINEAR, TANH, SOFTPLUS, GAUSSIAN, SOFTMAX = 0, 1, 2, 3, 4
# error ids (OPL_ERROR_*)
MSE, CROSS_ENTROPY = 0, 1


def param_count(shape):
    """n = d1 + sum_l d_l * d_{l+1}  (only the first layer has a bias; mlp.py:79,318-327)."""
    return shape[1] + sum(a * b for a, b in zip(shape[:-1], shape[1:]))


def split_params(flat, shape):
    """Views into the flat vector: bias first, then each W_l row-major (mlp.py:318-327)."""
    bias = flat[:shape[1]]
    mats, at = [], shape[1]
    for rows, cols in zip(shape[:-1], shape[1:]):
        mats.append(flat[at:at + rows * cols].reshape(rows, cols))
        at += rows * cols
    return bias, mats


def join_params(bias, mats):
    """Inverse of split_params (mlp.py:313-315)."""
    return numpy.hstack([bias] + [m.ravel() for m in mats])


# --------------------------------------------------------------------------
# transfers: (id, param) pairs.  param is the Gaussian variance, else unused.
# --------------------------------------------------------------------------
def transfer_apply(tid, param, z):
    if tid == LINEAR:          # transfer.py:47-49
        return z
    if tid == TANH:            # calculate.py:91-93
        return numpy.tanh(z)
    if tid == SOFTPLUS:        # calculate.py:128-135 ("ReluTransfer" is softplus)
        with numpy.errstate(over='ignore'):
            out = numpy.log(1.0 + numpy.exp(z))
        blown = out == numpy.inf
        out[blown] = z[blown]
        return out
    if tid == GAUSSIAN:        # calculate.py:101-102
        return numpy.exp(-(z ** 2 / param))
    if tid == SOFTMAX:         # calculate.py:152-153
        e = numpy.exp(z - numpy.max(z, axis=-1, keepdims=True))
        return e / numpy.sum(e, axis=-1, keepdims=True)
    raise ValueError('unknown transfer id %r' % (tid,))


def transfer_backprop(tid, param, upstream, z, a, literal_softmax=True):
    """upstream (N,d) times the transfer Jacobian (mlp.py:279-295 on transfer.py derivatives)."""
    if tid == LINEAR:          # transfer.py:58 -> ones
        return upstream * numpy.ones(z.shape)
    if tid == TANH:            # calculate.py:96-98, from the OUTPUT
        return upstream * (1.0 - a ** 2)
    if tid == SOFTPLUS:        # calculate.py:138-141, from the INPUT
        with numpy.errstate(over='ignore'):
            return upstream * (1.0 / (1.0 + numpy.exp(-z)))
    if tid == GAUSSIAN:        # calculate.py:105-106
        return upstream * (-2.0 * z * a / param)
    if tid == SOFTMAX:
        if literal_softmax:    # calculate.py:165-174 + mlp.py:295 (per-row Jacobian, einsum)
            jac = numpy.einsum('ij...,i...->ij...', -a, a)
            iy, ix = numpy.diag_indices(a.shape[-1])
            jac[:, iy, ix] = a * (1. - a)
            return numpy.einsum('ij,ijk->ik', upstream, jac)
        # The same contraction without the (N,C,C) tensor, term by term as the Jacobian form above:
        #   delta_k = g_k a_k (1 - a_k) - a_k sum_{j != k} g_j a_j = a_k (g_k (1 - a_k) - (t - g_k a_k)),  t = sum_j g_j a_j.
        # (The shorter a_k (g_k - t) is algebraically equal but cancels for confident one-hot rows, a_k -> 1: it loses
        # eps / (1 - a_k) of relative accuracy and does NOT track the reference there; 1 - a_k is exact.)
        p = upstream * a
        t = numpy.sum(p, axis=-1, keepdims=True)
        return a * (upstream * (1.0 - a) - (t - p))
    raise ValueError('unknown transfer id %r' % (tid,))


# --------------------------------------------------------------------------
# errors
# --------------------------------------------------------------------------
def error_value(eid, a, y):
    if eid == MSE:             # error.py:54
        return numpy.mean(numpy.subtract(a, y) ** 2)
    if eid == CROSS_ENTROPY:   # error.py:89-99
        with numpy.errstate(invalid='raise', divide='ignore'):
            log_a = numpy.log(a)
        log_a = numpy.nan_to_num(log_a)
        return -numpy.mean(numpy.sum(log_a * y, axis=-1))
    raise ValueError('unknown error id %r' % (eid,))


def error_value_grad(eid, a, y):
    if eid == MSE:             # error.py:58-64
        diff = numpy.subtract(a, y)
        val = numpy.mean(diff ** 2)
        diff *= 2.0 / float(numpy.prod(a.shape))
        return val, diff
    if eid == CROSS_ENTROPY:   # error.py:105-116
        with numpy.errstate(invalid='ignore', divide='ignore'):
            ratio = y / a
        ratio = numpy.nan_to_num(ratio)
        val = error_value(eid, a, y)
        if a.ndim == 1:
            return val, -ratio
        return val, ratio / (-float(numpy.prod(a.shape[:-1])))
    raise ValueError('unknown error id %r' % (eid,))


# --------------------------------------------------------------------------
# forward / objective / gradient
# --------------------------------------------------------------------------
def forward(flat, shape, transfers, x):
    """Return (activations [A_0..A_L], pre-activations [Z_0..Z_{L-1}])  (mlp.py:148-160)."""
    bias, mats = split_params(flat, shape)
    acts, pres = [x], []
    z = numpy.dot(x, mats[0]) + bias
    pres.append(z)
    acts.append(transfer_apply(transfers[0][0], transfers[0][1], z))
    for w, (tid, tp) in list(zip(mats, transfers))[1:]:
        z = numpy.dot(acts[-1], w)
        pres.append(z)
        acts.append(transfer_apply(tid, tp, z))
    return acts, pres


def activate(flat, shape, transfers, x):
    return numpy.copy(forward(flat, shape, transfers, numpy.asarray(x, dtype=float))[0][-1])


def objective(flat, shape, transfers, eid, x, y):
    """mlp.py:195-199."""
    return error_value(eid, forward(flat, shape, transfers, x)[0][-1], y)


def objective_gradient(flat, shape, transfers, eid, x, y, literal_softmax=True):
    """(objective, flat gradient)  (mlp.py:202-276)."""
    bias, mats = split_params(flat, shape)
    acts, pres = forward(flat, shape, transfers, x)
    val, upstream = error_value_grad(eid, acts[-1], y)

    last = len(mats) - 1
    deltas = [None] * len(mats)
    deltas[last] = transfer_backprop(transfers[last][0], transfers[last][1], upstream,
                                     pres[last], acts[last + 1], literal_softmax)
    for l in range(last - 1, -1, -1):      # mlp.py:254-260
        deltas[l] = transfer_backprop(transfers[l][0], transfers[l][1],
                                      deltas[l + 1].dot(mats[l + 1].T),
                                      pres[l], acts[l + 1], literal_softmax)
    grads = [acts[l].T.dot(deltas[l]) for l in range(len(mats))]   # mlp.py:269-273
    return val, join_params(numpy.sum(deltas[0], axis=0), grads)   # mlp.py:276,313-315


def algorithmic_flops(shape, rows):
    """F = 2 N (3 sum d_l d_{l+1} - d0 d1)   (SURVEY.md section 8d)."""
    s = sum(a * b for a, b in zip(shape[:-1], shape[1:]))
    return 2.0 * rows * (3.0 * s - shape[0] * shape[1])


# --------------------------------------------------------------------------
# synthetic data as the reference generates it (data/datasets.py:257-291)
# --------------------------------------------------------------------------
def random_classification(rows, attrs, classes, rng):
    x = rng.random((rows, attrs)) * 2.0 - 1.0
    y = numpy.zeros((rows, classes))
    y[numpy.arange(rows), rng.integers(0, classes, size=rows)] = 1.0
    return x, y


def random_regression(rows, attrs, outs, rng):
    return rng.random((rows, attrs)) * 2.0 - 1.0, rng.random((rows, outs)) * 2.0 - 1.0


def random_params(shape, rng, scale=0.25):
    """uniform(-0.25, 0.25) as mlp.py:37,123."""
    return (2.0 * rng.random(param_count(shape)) - 1.0) * scale
