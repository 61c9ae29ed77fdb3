"""CPU oracle for the optimizer side of the hot path (NumPy float64).

TEST INFRASTRUCTURE ONLY (same rules as ``oracle/mlp_oracle.py``): never
imported by the product package.

Restates, as small state machines over plain callables:

  BFGS update (literal + rank-2)     learning/optimize/optimizer.py:288-338
  BFGS step                          learning/optimize/optimizer.py:234-285
  L-BFGS two-loop / history          learning/optimize/optimizer.py:414-508
  steepest descent (+momentum)       learning/optimize/optimizer.py:89-145
  Wolfe bracketing + zoom            learning/optimize/linesearch.py:230-399
  backtracking / Armijo              learning/optimize/linesearch.py:179-224,430-447
  initial-step rules                 learning/optimize/initialstep.py:66-256

Pinned by ``oracle/gen_golden.py`` against the real (shimmed) reference: iterate
traces, step sizes, H matrices and directions are stored in ``tests/golden``.
"""
import numpy

BIG = 1.79769313e+308


# ----------------------------------------------------------------------------
# quasi-Newton algebra
# ----------------------------------------------------------------------------
def bfgs_update_literal(h, s, y):
    """The reference formula evaluated literally: O(n^3), two n x n products (optimizer.py:321-338)."""
    ys = y.dot(s)
    if ys == 0.0:
        return h
    rho = 1.0 / ys
    eye = numpy.identity(s.shape[0])
    rs = rho * s
    left = eye - rs[:, None] * y
    right = eye - (rho * y)[:, None] * s
    return left.dot(h).dot(right) + rs[:, None] * s


def bfgs_update_rank2(h, s, y):
    """Algebraically identical O(n^2) form the CUDA kernels implement.

    H+ = H - rho s v^T - rho u s^T + (rho^2 c + rho) s s^T,
    u = H y, v = H^T y, c = y.u   (H is NOT assumed symmetric).
    """
    ys = y.dot(s)
    if ys == 0.0:
        return h
    rho = 1.0 / ys
    u = h.dot(y)
    v = y.dot(h)
    c = y.dot(u)
    return (h - rho * numpy.outer(s, v) - rho * numpy.outer(u, s)
            + (rho * rho * c + rho) * numpy.outer(s, s))


def gamma_scalar(s, y):
    """(s.y)/(y.y) with the reference's guards (optimizer.py:346-362)."""
    yy = y.dot(y)
    if yy == 0:
        return 1.0
    with numpy.errstate(all='ignore'):
        return float(numpy.nan_to_num(s.dot(y) / yy))


def lbfgs_direction(s_hist, y_hist, grad, gamma_from_oldest=True):
    """Two-loop recursion exactly as optimizer.py:452-508 (rho stores y.s; gamma from pair [0])."""
    q = numpy.copy(grad)
    ys_list, alpha_list = [], []
    for s, y in reversed(list(zip(s_hist, y_hist))):
        ys = y.dot(s)
        if ys != 0:
            a = s.dot(q) / ys
            q -= a * y
        else:
            a = None
        ys_list.append(ys)
        alpha_list.append(a)
    ys_list.reverse()
    alpha_list.reverse()
    if len(s_hist) == 0:
        q *= 1.0
    else:
        q *= gamma_scalar(s_hist[0], y_hist[0]) if gamma_from_oldest else 1.0
    for s, y, ys, a in zip(s_hist, y_hist, ys_list, alpha_list):
        if ys != 0:
            q += s * (a - y.dot(q) / ys)
    return -q


# ----------------------------------------------------------------------------
# initial-step rules (host scalars)
# ----------------------------------------------------------------------------
def _sanitise_initial(step):
    """nan -> 1e-10, negative -> 1, inf -> BIG (initialstep.py:153-178, 225-251)."""
    if numpy.isnan(step):
        step = 1e-10
    if step < 0:
        step = 1.0
    if numpy.isinf(step):
        step = BIG
    assert step > 0
    return step


class FOChange(object):
    """initialstep.py:116-187."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.prev_step = None
        self.prev_slope = None

    def first(self, obj, slope):
        if self.prev_step is None or slope == 0.0:
            step = 1.0
        else:
            with numpy.errstate(all='ignore'):
                step = self.prev_step * (numpy.float64(self.prev_slope) / slope)
        step = _sanitise_initial(step)
        self.prev_slope = slope
        return step

    def update(self, step):
        self.prev_step = step


class Quadratic(object):
    """initialstep.py:190-256."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.prev_obj = None

    def first(self, obj, slope):
        if self.prev_obj is None or slope == 0.0:
            step = 1.0
        else:
            with numpy.errstate(all='ignore'):
                step = (2.0 * (obj - self.prev_obj)) / numpy.float64(slope)
        step = _sanitise_initial(step)
        self.prev_obj = obj
        return step

    def update(self, step):
        pass


class IncrPrev(object):
    """initialstep.py:66-113."""

    def __init__(self, incr_rate=1.05, lower=0, upper=1.0):
        self.incr_rate, self.lower = incr_rate, lower
        self.upper = float('inf') if upper is None else upper
        self.reset()

    def reset(self):
        self.prev = 1.0

    def first(self, obj, slope):
        return max(self.lower, min(self.upper, self.incr_rate * self.prev))

    def update(self, step):
        self.prev = step


# ----------------------------------------------------------------------------
# line searches over phi(alpha) -> (objective, slope)
# ----------------------------------------------------------------------------
GROWTH = 1.5      # linesearch.py:227


def wolfe_search(phi, obj0, slope0, c1, c2, first_step, trace=None):
    """linesearch.py:230-317.  NOTE the bracketing Armijo test uses the NEW point's slope (:281)."""
    if numpy.isnan(obj0):
        return 1e-10
    prev_step, prev_obj = 0.0, obj0
    step = first_step
    i = 0
    while True:
        i += 1
        if i >= 100:
            return step
        val, slope = phi(step)
        if trace is not None:
            trace.append(step)
        if (i > 1 and val >= prev_obj) or (val > obj0 + c1 * step * slope):
            return _zoom(phi, prev_step, prev_obj, step, obj0, slope0, c1, c2, trace)
        elif abs(slope) <= -c2 * slope0:
            return step
        elif slope >= 0:
            return _zoom(phi, step, val, prev_step, obj0, slope0, c1, c2, trace)
        prev_step, prev_obj = step, val
        step *= GROWTH


def _zoom(phi, lo, lo_obj, hi, obj0, slope0, c1, c2, trace):
    """Bisection zoom (linesearch.py:320-389)."""
    i = 0
    while True:
        i += 1
        a, b = min(lo, hi), max(lo, hi)
        step = a + 0.5 * (b - a)
        assert step >= 0
        if i >= 100:
            return step
        val, slope = phi(step)
        if trace is not None:
            trace.append(step)
        if val > obj0 + c1 * step * slope0 or val >= lo_obj:
            hi = step
        else:
            if abs(slope) <= -c2 * slope0:
                return step
            if slope * (hi - lo) >= 0:
                hi = lo
            lo, lo_obj = step, val


def backtracking_search(phi_obj, obj0, slope0, c1, first_step, decr_rate):
    """linesearch.py:179-224 with the Armijo rule of :430-447."""
    if numpy.isnan(obj0):
        return 1e-25
    step = first_step
    while True:
        if step < 1e-25:
            return step
        if phi_obj(step) <= obj0 + (c1 * step) * slope0:
            assert step > 0
            return step
        step *= decr_rate


class Wolfe(object):
    """WolfeLineSearch (linesearch.py:134-176)."""

    def __init__(self, c1=1e-4, c2=0.9, initial=None):
        self.c1, self.c2 = c1, c2
        self.initial = Quadratic() if initial is None else initial
        self.trace = None

    def reset(self):
        self.initial.reset()

    def __call__(self, x, obj, grad, direction, obj_jac, obj_only):
        slope0 = grad.dot(direction)
        first = self.initial.first(obj, slope0)

        def phi(alpha):
            val, g = obj_jac(x + alpha * direction)
            return val, g.dot(direction)
        step = wolfe_search(phi, obj, slope0, self.c1, self.c2, first, self.trace)
        self.initial.update(step)
        return step


class Backtracking(object):
    """BacktrackingLineSearch (linesearch.py:85-131)."""

    def __init__(self, c1=0.5, decr_rate=0.5, initial=None):
        self.c1, self.decr_rate = c1, decr_rate
        self.initial = (IncrPrev(2.0 / decr_rate - 1.0, 0.0, None)
                        if initial is None else initial)

    def reset(self):
        self.initial.reset()

    def __call__(self, x, obj, grad, direction, obj_jac, obj_only):
        slope0 = grad.dot(direction)
        first = self.initial.first(obj, slope0)
        step = backtracking_search(lambda a: obj_only(x + a * direction), obj, slope0,
                                   self.c1, first, self.decr_rate)
        self.initial.update(step)
        return step


class FixedStep(object):
    """SetStepSize (linesearch.py:62-82)."""

    def __init__(self, step):
        self.step = step

    def reset(self):
        pass

    def __call__(self, x, obj, grad, direction, obj_jac, obj_only):
        return self.step


def _default_search():
    return Wolfe(1e-4, 0.9, FOChange())


# ----------------------------------------------------------------------------
# optimizers: .step(obj_jac, obj_only, x) -> (objective at x, next x)
# ----------------------------------------------------------------------------
class SteepestDescent(object):
    """optimizer.py:73-97."""

    def __init__(self, search=None):
        self.search = search or _default_search()
        self.grad = None

    def reset(self):
        self.grad = None
        self.search.reset()

    def step(self, obj_jac, obj_only, x):
        val, self.grad = obj_jac(x)
        a = self.search(x, val, self.grad, -self.grad, obj_jac, obj_only)
        return val, x - a * self.grad


class SteepestDescentMomentum(object):
    """optimizer.py:100-145: momentum term added AFTER the line search."""

    def __init__(self, search=None, momentum=0.2):
        self.search = search or _default_search()
        self.momentum = momentum
        self.grad = None
        self.prev = None

    def reset(self):
        self.grad = None
        self.prev = None
        self.search.reset()

    def step(self, obj_jac, obj_only, x):
        val, self.grad = obj_jac(x)
        d = -self.grad
        move = self.search(x, val, self.grad, d, obj_jac, obj_only) * d
        nxt = x + move
        if self.prev is not None:
            nxt += self.momentum * self.prev
        self.prev = move
        return val, nxt


class BFGS(object):
    """optimizer.py:170-285.  ``update`` selects the literal or rank-2 form of the H update."""

    def __init__(self, search=None, scaled_identity=False, reset_every=100,
                 update=bfgs_update_literal):
        self.search = search or _default_search()
        self.scaled_identity = scaled_identity
        self.reset_every = reset_every
        self.update = update
        self._clear()

    def _clear(self):
        self.grad = None
        self.count = 0
        self.prev_move = None
        self.prev_grad = None
        self.h = None

    def reset(self):
        self._clear()
        self.search.reset()

    def step(self, obj_jac, obj_only, x):
        self.count += 1
        if self.count == self.reset_every:        # optimizer.py:236-238
            self.reset()
        val, self.grad = obj_jac(x)
        g = self.grad
        if self.prev_move is None:                # first call: identity, not stored
            h = numpy.identity(g.shape[0])
        else:
            y = g - self.prev_grad
            if self.h is None:                    # second call: seed
                if self.scaled_identity:
                    seed = numpy.diag(numpy.repeat(gamma_scalar(self.prev_move, y), g.shape[0]))
                else:
                    seed = numpy.identity(g.shape[0])
                h = self.update(seed, self.prev_move, y)
            else:
                h = self.update(self.h, self.prev_move, y)
            self.h = h
        self.prev_grad = g
        d = -(h.dot(g))
        a = self.search(x, val, g, d, obj_jac, obj_only)
        self.prev_move = a * d
        return val, x + self.prev_move


class LBFGS(object):
    """optimizer.py:365-508."""

    def __init__(self, search=None, memory=5, gamma_from_oldest=True):
        self.search = search or _default_search()
        self.memory = memory
        self.gamma_from_oldest = gamma_from_oldest
        self._clear()

    def _clear(self):
        self.grad = None
        self.prev_move = None
        self.prev_grad = None
        self.s_hist, self.y_hist = [], []

    def reset(self):
        self._clear()
        self.search.reset()

    def step(self, obj_jac, obj_only, x):
        val, self.grad = obj_jac(x)
        g = self.grad
        if len(self.s_hist) >= self.memory:       # pop BEFORE append (optimizer.py:437-439)
            self.s_hist.pop(0)
            self.y_hist.pop(0)
        if self.prev_move is not None:
            self.s_hist.append(self.prev_move)
            self.y_hist.append(g - self.prev_grad)
        self.prev_grad = g
        d = lbfgs_direction(self.s_hist, self.y_hist, g, self.gamma_from_oldest)
        a = self.search(x, val, g, d, obj_jac, obj_only)
        self.prev_move = a * d
        return val, x + self.prev_move
