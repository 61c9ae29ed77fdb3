"""CPU: the NumPy oracle against the golden vectors produced by the real reference."""
import numpy
import pytest

import mlp_oracle as mo
import optimize_oracle as oo
from util import rel_err, specs


def test_objgrad_cases_match_reference(golden_meta, golden_objgrad):
    g = golden_objgrad
    assert len(golden_meta["objgrad"]) >= 14
    for entry in golden_meta["objgrad"]:
        k = entry["key"]
        shape, transfers, eid = specs(entry)
        with numpy.errstate(all='ignore'):
            obj = mo.objective(g[k + "_w"], shape, transfers, eid, g[k + "_x"], g[k + "_y"])
            obj2, grad = mo.objective_gradient(g[k + "_w"], shape, transfers, eid, g[k + "_x"], g[k + "_y"])
            closed = mo.objective_gradient(g[k + "_w"], shape, transfers, eid, g[k + "_x"], g[k + "_y"],
                                           literal_softmax=False)[1]
            out = mo.activate(g[k + "_w"], shape, transfers, g[k + "_x"])
        assert rel_err(obj, g[k + "_obj"]) <= 1e-13, entry["name"]
        assert rel_err(obj2, g[k + "_obj"]) <= 1e-13, entry["name"]
        assert rel_err(grad, g[k + "_grad"]) <= 1e-13, entry["name"]
        assert rel_err(closed, g[k + "_grad"]) <= 1e-11, entry["name"]
        assert rel_err(out, g[k + "_out"]) <= 1e-13, entry["name"]


def test_known_answers_from_reference_tests(golden_meta):
    ka = golden_meta["known_answers"]
    # tests/test_calculate.py:193, :223 and tests/test_error.py:79-128 of the reference
    assert ka["softmax_1000"] == [0.0, 1.0]
    assert numpy.allclose(ka["softplus_0_1"], [0.6931, 1.3133], atol=1e-4)
    assert ka["softplus_big"][0] == 1000.0 and ka["softplus_big"][1] == 0.0
    assert ka["ce_deriv_01"] == [0.0, -1.0]
    z = numpy.array([[-1000.0, 1000.0]])
    assert mo.transfer_apply(mo.SOFTMAX, 0, z).tolist() == [[0.0, 1.0]]
    assert numpy.allclose(mo.transfer_apply(mo.SOFTPLUS, 0, numpy.array([0.0, 1.0])), ka["softplus_0_1"], atol=0)
    with numpy.errstate(all='ignore'):
        assert mo.transfer_apply(mo.SOFTPLUS, 0, numpy.array([1000.0, -1000.0])).tolist() == ka["softplus_big"]
        assert mo.error_value(mo.CROSS_ENTROPY, numpy.array([0.0, 1.0]), numpy.array([0.0, 1.0])) == ka["ce_exact"][0]
    assert mo.error_value(mo.CROSS_ENTROPY, numpy.full((2, 2), 0.5), numpy.array([[0.0, 1.0], [1.0, 0.0]])) \
        == pytest.approx(ka["ce_exact"][1], abs=0)


def test_mlp_forward_known_answers():
    # reference tests/architecture/test_mlp.py:41-66: all-ones (2,2,2) linear network
    shape, tr = (2, 2, 2), [(mo.LINEAR, 0), (mo.LINEAR, 0)]
    w = numpy.ones(mo.param_count(shape))
    for x, want in [([0, 0], [2, 2]), ([0.5, 0.5], [4, 4]), ([1, 0], [4, 4]), ([0.5, 1], [5, 5]), ([1, 1], [6, 6])]:
        assert numpy.allclose(mo.activate(w, shape, tr, numpy.array(x, dtype=float)), want)
    # test_mlp.py:116-126 perceptron
    w = numpy.array([0.0, 0.5, -0.5])
    assert mo.activate(w, (2, 1), [(mo.LINEAR, 0)], numpy.array([1.0, 1.0])).tolist() == [0.0]


def test_log_of_negative_raises():
    with pytest.raises(FloatingPointError):   # error.py:89-91
        mo.error_value(mo.CROSS_ENTROPY, numpy.array([[-0.5, 1.5]]), numpy.array([[0.0, 1.0]]))


def test_gradient_against_finite_differences():
    # reference testing/helpers.py:228-264: central differences, eps = (1.1e-16)^(1/3), tol 10 eps
    rng = numpy.random.default_rng(5)
    eps = 1.1e-16 ** (1.0 / 3.0)
    for shape, tr, eid in [((4, 5, 3), [(mo.SOFTPLUS, 0), (mo.SOFTMAX, 0)], mo.CROSS_ENTROPY),
                           ((3, 6, 4, 2), [(mo.TANH, 0), (mo.GAUSSIAN, 1.3), (mo.LINEAR, 0)], mo.MSE)]:
        x, y = (mo.random_classification if eid else mo.random_regression)(9, shape[0], shape[-1], rng)
        w = mo.random_params(shape, rng, 1.0)
        grad = mo.objective_gradient(w, shape, tr, eid, x, y)[1]
        fd = numpy.empty_like(w)
        for i in range(w.size):
            e = numpy.zeros_like(w)
            e[i] = eps
            fd[i] = (mo.objective(w + e, shape, tr, eid, x, y) - mo.objective(w - e, shape, tr, eid, x, y)) / (2 * eps)
        assert numpy.max(numpy.abs(fd - grad)) < 10 * eps


def test_bfgs_update_forms_match_reference(golden_meta, golden_qn):
    g = golden_qn
    for case in golden_meta["bfgs"]:
        k = case["key"]
        lit = oo.bfgs_update_literal(g[k + "_h"].copy(), g[k + "_s"], g[k + "_y"])
        rk2 = oo.bfgs_update_rank2(g[k + "_h"].copy(), g[k + "_s"], g[k + "_y"])
        assert rel_err(lit, g[k + "_hnext"]) <= 1e-14
        assert rel_err(rk2, g[k + "_hnext"]) <= 1e-12
    # reference tests/optimize/test_optimizer.py:48-62: symmetric result and secant equation H+ y = s
    k = golden_meta["bfgs"][2]["key"]
    hn = g[k + "_hnext"]
    assert numpy.allclose(hn, hn.T, atol=1e-12)
    assert numpy.allclose(hn.dot(g[k + "_y"]), g[k + "_s"], atol=1e-10)
    assert abs(oo.gamma_scalar(g["seed_s"], g["seed_y"]) - float(g["seed_gamma"])) <= 1e-16


def test_lbfgs_direction_matches_reference(golden_meta, golden_qn):
    g = golden_qn
    for case in golden_meta["lbfgs"]:
        k = case["key"]
        S, Y = g[k + "_S"], g[k + "_Y"]
        d = oo.lbfgs_direction([S[i] for i in range(case["m"])], [Y[i] for i in range(case["m"])], g[k + "_g"])
        assert rel_err(d, g[k + "_dir"]) <= 1e-14


def test_optimizer_traces_match_reference(golden_meta, golden_traces):
    t = golden_traces
    factories = {
        "bfgs": lambda: oo.BFGS(), "bfgs_scaled": lambda: oo.BFGS(scaled_identity=True),
        "bfgs_reset5": lambda: oo.BFGS(reset_every=5),
        "lbfgs": lambda: oo.LBFGS(), "sd": lambda: oo.SteepestDescent(),
        "sdm": lambda: oo.SteepestDescentMomentum(),
        "sd_backtracking": lambda: oo.SteepestDescent(oo.Backtracking()),
        "bfgs_wolfe_quadratic": lambda: oo.BFGS(oo.Wolfe()),
    }
    for run in golden_meta["traces"]:
        shape, tr, eid = specs(run)
        x, y = t[run["data"] + "_x"], t[run["data"] + "_y"]
        opt = factories[run["optimizer"]]()
        cur = t[run["key"] + "_w0"].copy()
        f = lambda v: mo.objective_gradient(v, shape, tr, eid, x, y)
        fo = lambda v: mo.objective(v, shape, tr, eid, x, y)
        with numpy.errstate(all='ignore'):
            for step in range(run["steps"]):
                val, cur = opt.step(f, fo, cur)
                assert rel_err(val, t[run["key"] + "_objs"][step]) <= 1e-9, (run["key"], step)
                assert rel_err(cur, t[run["key"] + "_xs"][step + 1]) <= 1e-9, (run["key"], step)


def test_rank2_bfgs_iterates_track_literal():
    """The O(n^2) update the kernels implement produces the same iterates as the literal O(n^3)
    form on a well-conditioned problem (hard parts)."""
    rng = numpy.random.default_rng(11)
    shape, tr, eid = (12, 10, 4), [(mo.SOFTPLUS, 0), (mo.SOFTMAX, 0)], mo.CROSS_ENTROPY
    x, y = mo.random_classification(200, 12, 4, rng)
    w0 = mo.random_params(shape, rng)
    f = lambda v: mo.objective_gradient(v, shape, tr, eid, x, y)
    fo = lambda v: mo.objective(v, shape, tr, eid, x, y)
    a, b = oo.BFGS(), oo.BFGS(update=oo.bfgs_update_rank2)
    xa, xb = w0.copy(), w0.copy()
    for _ in range(25):
        _, xa = a.step(f, fo, xa)
        _, xb = b.step(f, fo, xb)
        assert rel_err(xb, xa) <= 1e-9


# ----------------------------------------------------------------------------------------
# the digit arithmetic of the parity mode's INT8 engine (oracle/ozaki_oracle.py restates csrc/ozaki.cu)
# ----------------------------------------------------------------------------------------
def test_radix256_digits_are_an_error_free_transformation():
    import ozaki_oracle as oz
    rng = numpy.random.default_rng(0)
    x = rng.standard_normal((7, 300)) * numpy.exp(rng.standard_normal((7, 1)) * 8)
    x[0, :5] = [0.0, 1.0, -1.0, 2.0 ** -40, -(2.0 ** 20)]
    x[3] = 0.0
    scale = oz.scale_of(numpy.abs(x).max(axis=1, keepdims=True))
    assert (numpy.log2(scale) % 1 == 0).all() and (numpy.abs(x) <= 0.495 * scale + 1e-300).all()
    d = oz.digits(x, scale)
    assert d.min() >= -128 and d.max() <= 127
    # the digits reproduce x to 2^-49 of the row scale, exactly the rounding of x / scale * 2^48 to an integer
    assert (numpy.abs(oz.reconstruct(d, scale) - x) <= 2.0 ** -49 * scale * (1 + 1e-12)).all()
    assert (d[:, 3] == 0).all()                                   # a zero row has zero digits (scale 1)


def test_digit_gemm_matches_float64_and_stays_inside_int32():
    import ozaki_oracle as oz
    rng = numpy.random.default_rng(1)
    for (m, n, k, heavy) in [(33, 17, 200, False), (20, 9, 1500, True), (5, 4, 3, False)]:
        a = rng.random((m, k)) * 2 - 1
        b = (rng.random((k, n)) * 2 - 1) * 0.25
        if heavy:
            b *= numpy.exp(rng.standard_normal((k, 1)) * 3) * 1e-7
            a[:, ::7] = 0.0
        c, acc = oz.gemm(a, b)
        ref = a @ b
        bound = numpy.abs(a) @ numpy.abs(b)
        assert (numpy.abs(c - ref) <= 1e-12 * bound + 1e-300).all()
        assert numpy.abs(acc).max() < 2 ** 31                     # what one INT32 TMEM accumulator has to hold
    # worst case for the accumulators: every digit at -128, the longest contraction a tile is allowed
    assert oz.S * oz.MAX_K * 128 * 128 < 2 ** 31
