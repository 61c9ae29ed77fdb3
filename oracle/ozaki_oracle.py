"""NumPy restatement of the digit arithmetic behind the parity mode's INT8 engine (csrc/ozaki.cu) -- TEST INFRASTRUCTURE,
never imported by the product.

The reference has no counterpart (its GEMMs are numpy.dot, mlp.py:150-160,254-273); this file pins OUR error-free
transformation on the CPU so that `-m "not gpu"` covers it: every operand row / column is scaled by a power of two and
cut into 6 signed radix-256 digits, the 21 digit products with p + q <= 5 are exact integer GEMMs, and the result is
recombined in float64.  Same formulas as oz_scale_of / oz_digits4 / the epilogue's Horner recombination.
"""
import numpy

S = 6
MAX_K = 16384          # 6 * 16384 * 128^2 < 2^31: one INT32 accumulator per diagonal stays exact


def scale_of(amax):
    """power of two with |x| / scale <= 0.495 (oz_scale_of)."""
    amax = numpy.asarray(amax, dtype=numpy.float64)
    m, e = numpy.frexp(numpy.where(amax > 0, amax, 1.0))        # amax = m 2^e, m in [0.5, 1)
    e = e + numpy.where(m <= 0.99, 1, 2)
    e = numpy.clip(e, -1000, 1023)
    return numpy.where(amax > 0, numpy.ldexp(1.0, e), 1.0)


def digits(x, scale):
    """x [.., k] float64, scale broadcastable -> int64 digits [S, .., k] in [-128, 127] (oz_digits4)."""
    fx = numpy.rint(x / scale * 2.0 ** 48).astype(numpy.int64)
    u = fx + 0x808080808080
    assert (u >= 0).all() and (u < 2 ** 48).all()
    out = [((u >> (8 * (S - 1 - s))) & 0xff) - 128 for s in range(S)]
    return numpy.stack(out)


def reconstruct(d, scale):
    return scale * sum(d[s].astype(numpy.float64) * 256.0 ** -(s + 1) for s in range(S))


def gemm(a, b):
    """a [M,K] . b [K,N] through the digit arithmetic: returns (C float64, diagonal accumulators int64 [S, M, N])."""
    sa = scale_of(numpy.abs(a).max(axis=1, keepdims=True))       # per row of A
    sb = scale_of(numpy.abs(b).max(axis=0, keepdims=True))       # per column of B
    da, db = digits(a, sa), digits(b, sb)
    acc = numpy.zeros((S,) + (a.shape[0], b.shape[1]), dtype=numpy.int64)
    for p in range(S):
        for q in range(S - p):
            acc[p + q] += da[p] @ db[q]
    v = acc[S - 1].astype(numpy.float64)
    for d in range(S - 2, -1, -1):                               # Horner, smallest weight first
        v = v * (1.0 / 256.0) + acc[d].astype(numpy.float64)
    return v * (1.0 / 65536.0) * sa * sb, acc
