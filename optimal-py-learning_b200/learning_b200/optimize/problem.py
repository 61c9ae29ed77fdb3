"""Problem: the bundle of callables an Optimizer works on (reference: optimize/problem.py:33-150).

Same constructor and ``get_*`` attributes as the reference.  Missing callables are derived
from the ones given, in the reference's priority order:

    obj          obj_func, obj_jac_func[0], obj_hess_func[0], obj_jac_hess_func[0]
    jac          jac_func, obj_jac_func[1], jac_hess_func[0], obj_jac_hess_func[1]
    hess         hess_func, obj_hess_func[1], jac_hess_func[1], obj_jac_hess_func[2]
    obj_jac      obj_jac_func, obj_jac_hess_func[0,1], (obj, jac)
    obj_hess     obj_hess_func, obj_jac_hess_func[0,2], (obj, hess)
    jac_hess     jac_hess_func, obj_jac_hess_func[1,2], (jac, hess)
    obj_jac_hess obj_jac_hess_func, obj_jac + (hess,), (obj_hess[0], jac, obj_hess[1]),
                 (obj,) + jac_hess, (obj, jac, hess)

A Problem may additionally carry ``ray_obj_jac(x, alpha, direction) -> (obj, slope)`` and
``ray_obj(x, alpha, direction) -> obj``: fused line-search probes evaluated at
``x + alpha*direction`` without materialising the trial point on the host side (set by
MLP.train_step).  Line searches use them when present and fall back to get_obj_jac / get_obj.
"""


def _nothing(*args, **kwargs):
    return None


def _item(func, index):
    return lambda *args, **kwargs: func(*args, **kwargs)[index]


def _items(func, indices):
    return lambda *args, **kwargs: [func(*args, **kwargs)[i] for i in indices]


def _each(*funcs):
    return lambda *args, **kwargs: [f(*args, **kwargs) for f in funcs]


def _first(candidates):
    """candidates: [(source callable or None, builder)] -> first available, else _nothing."""
    for source, make in candidates:
        if source is not None:
            return make(source)
    return _nothing


class Problem(object):
    """Problem instance for an optimizer."""

    def __init__(self, obj_func=None, jac_func=None, hess_func=None, obj_jac_func=None,
                 obj_hess_func=None, jac_hess_func=None, obj_jac_hess_func=None):
        same = lambda f: f
        self.get_obj = _first([(obj_func, same), (obj_jac_func, lambda f: _item(f, 0)),
                               (obj_hess_func, lambda f: _item(f, 0)),
                               (obj_jac_hess_func, lambda f: _item(f, 0))])
        self.get_jac = _first([(jac_func, same), (obj_jac_func, lambda f: _item(f, 1)),
                               (jac_hess_func, lambda f: _item(f, 0)),
                               (obj_jac_hess_func, lambda f: _item(f, 1))])
        self.get_hess = _first([(hess_func, same), (obj_hess_func, lambda f: _item(f, 1)),
                                (jac_hess_func, lambda f: _item(f, 1)),
                                (obj_jac_hess_func, lambda f: _item(f, 2))])

        self.get_obj_jac = _first([(obj_jac_func, same), (obj_jac_hess_func, lambda f: _items(f, (0, 1))),
                                   (True, lambda _: _each(self.get_obj, self.get_jac))])
        self.get_obj_hess = _first([(obj_hess_func, same), (obj_jac_hess_func, lambda f: _items(f, (0, 2))),
                                    (True, lambda _: _each(self.get_obj, self.get_hess))])
        self.get_jac_hess = _first([(jac_hess_func, same), (obj_jac_hess_func, lambda f: _items(f, (1, 2))),
                                    (True, lambda _: _each(self.get_jac, self.get_hess))])

        get_hess, get_jac, get_obj = self.get_hess, self.get_jac, self.get_obj

        def from_obj_jac(f):
            return lambda *a, **k: tuple(f(*a, **k)) + (get_hess(*a, **k),)

        def from_obj_hess(f):
            def call(*a, **k):
                obj, hess = f(*a, **k)
                return obj, get_jac(*a, **k), hess
            return call

        def from_jac_hess(f):
            return lambda *a, **k: (get_obj(*a, **k),) + tuple(f(*a, **k))

        self.get_obj_jac_hess = _first([(obj_jac_hess_func, same), (obj_jac_func, from_obj_jac),
                                        (obj_hess_func, from_obj_hess), (jac_hess_func, from_jac_hess),
                                        (True, lambda _: _each(get_obj, get_jac, get_hess))])

        # optional fused probes (see module docstring)
        self.ray_obj_jac = None
        self.ray_obj = None
