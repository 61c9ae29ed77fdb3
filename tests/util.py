"""Shared helpers for the test-suite."""
import numpy


def rel_err(actual, expected):
    """Relative l2 error; non-finite entries of `expected` must be reproduced exactly."""
    a = numpy.asarray(actual, dtype=float).ravel()
    b = numpy.asarray(expected, dtype=float).ravel()
    assert a.shape == b.shape, (a.shape, b.shape)
    bad = ~numpy.isfinite(b)
    if bad.any() or (~numpy.isfinite(a)).any():
        if not (numpy.array_equal(a[bad], b[bad], equal_nan=True) and numpy.isfinite(a[~bad]).all()):
            return float('inf')
        a, b = a[~bad], b[~bad]
    den = numpy.linalg.norm(b)
    return float(numpy.linalg.norm(a - b) / (den if den > 0 else 1.0))


def specs(meta_entry):
    shape = tuple(meta_entry["shape"])
    transfers = [(int(t), float(p)) for t, p in meta_entry["transfers"]]
    return shape, transfers, int(meta_entry["error"])


def make_transfers(L, transfers):
    table = {0: L.LinearTransfer, 1: L.TanhTransfer, 2: L.ReluTransfer, 4: L.SoftmaxTransfer}
    return [L.GaussianTransfer(p) if t == 3 else table[t]() for t, p in transfers]


def make_error(L, eid):
    return L.MeanSquaredError() if eid == 0 else L.CrossEntropyError()
