"""GPU: the CUDA path (through the C-ABI / the reference-shaped Python API) against the golden
vectors of the real reference and against the NumPy oracle.

FP64 parity gate (BASELINE.json north_star): objective and gradient within 1e-10 relative.
"""
import copy
import ctypes
import pickle

import numpy
import pytest

import mlp_oracle as mo
import optimize_oracle as oo
from util import make_error, make_transfers, rel_err, specs

pytestmark = pytest.mark.gpu

TOL = 1e-10   # relative, FP64 parity mode


@pytest.fixture(scope="module")
def L():
    import learning_b200
    from learning_b200 import _native
    assert _native.device_count() >= 1, "GPU tests need a CUDA device"
    return learning_b200


def _model(L, shape, transfers, eid, optimizer=None, precision=None):
    return L.MLP(shape, transfers=make_transfers(L, transfers), error_func=make_error(L, eid),
                 optimizer=optimizer or L.optimize.SteepestDescent(), precision=precision)


# ----------------------------------------------------------------------------------------
# objective / gradient / activate
# ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("small_path", [True, False])   # one-launch small-network kernel / general kernel schedule
def test_objgrad_golden_cases_host_api(L, golden_meta, golden_objgrad, small_path):
    g = golden_objgrad
    for entry in golden_meta["objgrad"]:
        k = entry["key"]
        shape, transfers, eid = specs(entry)
        model = _model(L, shape, transfers, eid)
        model.small_path = small_path
        obj = model._get_obj(g[k + "_w"].copy(), g[k + "_x"], g[k + "_y"])
        obj2, grad = model._get_obj_jac(g[k + "_w"].copy(), g[k + "_x"], g[k + "_y"])
        assert rel_err(obj, g[k + "_obj"]) <= TOL, (entry["name"], obj, float(g[k + "_obj"]))
        assert obj2 == obj or (numpy.isinf(obj) and numpy.isinf(obj2)), entry["name"]
        assert rel_err(grad, g[k + "_grad"]) <= TOL, (entry["name"], rel_err(grad, g[k + "_grad"]))
        # activate with the weights the evaluation left in the model (mlp.py:197)
        out = model.activate(g[k + "_x"])
        assert rel_err(out, g[k + "_out"]) <= TOL, entry["name"]
        assert rel_err(model.activate(g[k + "_x"][0]), g[k + "_out"][0]) <= TOL, entry["name"]


def test_device_api_matches_host_api_and_probe(L, golden_meta, golden_objgrad):
    from learning_b200 import _device as dev
    g = golden_objgrad
    for entry in golden_meta["objgrad"][:12]:
        k = entry["key"]
        shape, transfers, eid = specs(entry)
        model = _model(L, shape, transfers, eid)
        x, y, w = g[k + "_x"], g[k + "_y"], g[k + "_w"]
        obj_h, grad_h = model._get_obj_jac(w.copy(), x, y)
        wd = dev.to_device(w)
        obj_d, grad_d = model._get_obj_jac(wd, x, y)
        assert obj_d == obj_h
        assert numpy.array_equal(dev.to_host(grad_d), grad_h)          # same kernels, same order
        assert model._get_obj(wd, x, y) == obj_h
        # fused probe at w + alpha d == evaluation at the explicit point; slope == grad . d
        rng = numpy.random.default_rng(7)
        d = rng.standard_normal(w.size)
        alpha = 0.37
        ws = model._make_resident(x, y)
        obj_p, slope_p, grad_p = model._probe(ws, wd, alpha, dev.to_device(d), True)
        obj_e, grad_e = model._get_obj_jac(w + alpha * d, x, y)
        assert obj_p == obj_e
        assert numpy.array_equal(dev.to_host(grad_p), grad_e)
        assert rel_err(slope_p, grad_e.dot(d)) <= 1e-12


@pytest.mark.parametrize("shape,transfers,eid,rows", [
    ((784, 256, 10), [(2, 0), (4, 0)], 1, 3000),           # config-2 layer shapes, split-K, 16-byte path
    ((785, 255, 11), [(2, 0), (4, 0)], 1, 1531),           # odd everything: 8-byte cp.async path, edge tiles
    ((512, 128, 128, 10), [(2, 0), (2, 0), (4, 0)], 1, 2048),   # config-3 shape
    ((256, 192, 160, 10), [(1, 0), (3, 1.5), (0, 0)], 0, 1100),  # tanh / gaussian / linear + MSE
    ((130, 140, 150, 65), [(4, 0), (2, 0), (2, 0)], 0, 1200),   # hidden softmax, wide output
    ((300, 64, 7), [(1, 0), (4, 0)], 1, 1500),              # tanh hidden feeding the fused narrow head
    ((96, 72, 40, 5), [(2, 0), (0, 0), (0, 0)], 0, 1111),   # linear hidden + linear output + MSE through the fused head
    ((40, 70, 33, 65), [(4, 0), (2, 0), (2, 0)], 0, 517),   # hidden softmax, wide output
    ((3, 1), [(0, 0)], 0, 5),
    # BASELINE config 4's shape: INT8 forward / dW / dA (OZ_EPI_MUL_AUX) GEMMs, 32 classes (no fused head), digits of hidden layers
    ((1024, 256, 256, 32), [(2, 0), (2, 0), (4, 0)], 1, 4096),
    # BASELINE config 5's shape: four 1024-wide softplus layers, linear output + MSE (MLP defaults), fused narrow head
    ((256, 1024, 1024, 1024, 1024, 10), [(2, 0), (2, 0), (2, 0), (2, 0), (0, 0)], 0, 4096),
])
@pytest.mark.parametrize("precision", ["f64", "f64-dmma"])   # INT8 (Ozaki) engine for the large GEMMs / DMMA everywhere
def test_objgrad_against_oracle_at_size(L, shape, transfers, eid, rows, precision):
    rng = numpy.random.default_rng(hash(shape) % 1000)
    gen = mo.random_classification if eid == 1 else mo.random_regression
    x, y = gen(rows, shape[0], shape[-1], rng)
    w = mo.random_params(shape, rng)
    want_obj, want_grad = mo.objective_gradient(w, shape, transfers, eid, x, y, literal_softmax=rows < 2000)
    model = _model(L, shape, transfers, eid, precision=precision)
    obj, grad = model._get_obj_jac(w.copy(), x, y)
    assert rel_err(obj, want_obj) <= TOL
    assert rel_err(grad, want_grad) <= TOL, rel_err(grad, want_grad)
    assert rel_err(model._get_obj(w.copy(), x, y), want_obj) <= TOL
    assert rel_err(model.activate(x), mo.activate(w, shape, transfers, x)) <= TOL
    # bit-reproducible: fixed-order split-K reduction, no atomics
    obj2, grad2 = model._get_obj_jac(w.copy(), x, y)
    assert obj2 == obj and numpy.array_equal(grad2, grad)


@pytest.mark.parametrize("M,N,K,heavy", [
    (128, 64, 128, False), (300, 100, 500, False), (1, 1, 1, False), (257, 65, 130, True), (785, 256, 20000, True),
    (4096, 256, 784, False),
])
def test_int8_digit_gemm_is_fp64_accurate(L, M, N, K, heavy):
    """ozaki_gemm_kernel (6 radix-256 digits per operand, 21 exact INT8 products on tcgen05, FP64 recombination) against
    torch's FP64 matmul: the engine behind the parity mode's large GEMMs has to be FP64-accurate on its own."""
    import torch
    from learning_b200 import _native as nat, _device as dev
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.rand(M, K, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    b = (torch.rand(K, N, dtype=torch.float64, device="cuda", generator=g) * 2 - 1) * 0.25
    if heavy:   # rows / columns spanning many orders of magnitude, exact zeros, an all-zero row, a row of ones
        b = b * torch.exp(torch.randn(K, 1, dtype=torch.float64, device="cuda", generator=g) * 3) * 1e-7
        a[:, ::7] = 0.0
        a[M // 2] = 0.0
        a[0] = 1.0
    c = torch.full((M, N), float("nan"), dtype=torch.float64, device="cuda")
    sm, gm = ctypes.c_float(0), ctypes.c_float(0)
    nat.check(nat.lib.opl_bench_gemm_ozaki(M, N, K, dev.ptr(a), dev.ptr(b), dev.ptr(c), 0, 1, ctypes.byref(sm), ctypes.byref(gm)))
    torch.cuda.synchronize()
    ref = a @ b
    # error model: each operand is cut at 2^-48 of its row / column scale, so the bound is relative to |A| |B|
    bound = (a.abs() @ b.abs()).clamp_min(1e-300)
    assert float(((c - ref).abs() / bound).max()) <= 1e-11
    assert float((c - ref).norm() / ref.norm().clamp_min(1e-300)) <= 1e-11
    if heavy:
        assert float(c[M // 2].abs().max()) == 0.0      # an all-zero row stays exactly zero
    if M * N * K <= 300 * 100 * 500:
        # the CPU restatement of the digit arithmetic (oracle/ozaki_oracle.py) holds the same integers; only the FP64
        # recombination may round differently (fused multiply-add on the device)
        import ozaki_oracle as oz
        want, _ = oz.gemm(a.cpu().numpy(), b.cpu().numpy())
        got = c.cpu().numpy()
        assert (numpy.abs(got - want) <= 1e-15 * bound.cpu().numpy()).all()


def test_engines_agree_and_fused_head_matches_plain_schedule(L):
    """The INT8 engine + fused narrow head against the plain DMMA schedule on the same model and data (config-2 layer
    shapes, dropout masks included): both are FP64-accurate, so they agree far below the 1e-10 gate."""
    rng = numpy.random.default_rng(5)
    shape, transfers, eid = (784, 256, 10), [(2, 0), (4, 0)], 1
    x, y = mo.random_classification(5000, 784, 10, rng)
    w = mo.random_params(shape, rng)
    a = _model(L, shape, transfers, eid, precision="f64")
    b = _model(L, shape, transfers, eid, precision="f64-dmma")
    oa, ga = a._get_obj_jac(w.copy(), x, y)
    ob, gb = b._get_obj_jac(w.copy(), x, y)
    assert rel_err(oa, ob) <= 1e-12 and rel_err(ga, gb) <= 1e-11, (rel_err(oa, ob), rel_err(ga, gb))
    assert rel_err(a._get_obj(w.copy(), x, y), ob) <= 1e-12


@pytest.mark.parametrize("hidden", [(2, 0), (1, 0), (0, 0)])          # softplus / tanh / linear hidden layer
def test_delta_digits_formed_on_the_fly_match_the_materialised_schedule(L, hidden, monkeypatch):
    """Two-layer networks on the INT8 engine: the fused head leaves delta_0 = (delta_L W_1^T) o f' to the digit kernel of the
    dW_0 GEMM (csrc/ozaki.cu ozaki_slice_delta_prev: digits cut straight from delta_L, W_1 and the slope matrix, scales from a
    rigorous bound instead of the column maxima).  Against the oracle, and against the schedule that materialises delta_0
    (OPL_DIRECT_DELTA=0), including weights 1e6 apart in scale between the classes (a loose bound must only cost digits)."""
    rng = numpy.random.default_rng(11)
    shape, transfers, eid = (200, 128, 10), [hidden, (4, 0)], 1
    x, y = mo.random_classification(3000, 200, 10, rng)
    w = mo.random_params(shape, rng)
    w[128 + 200 * 128:].reshape(128, 10)[:, ::3] *= 1e-6               # some classes' weights six orders smaller
    want_obj, want_grad = mo.objective_gradient(w, shape, transfers, eid, x, y, literal_softmax=False)
    direct = _model(L, shape, transfers, eid)
    obj_d, grad_d = direct._get_obj_jac(w.copy(), x, y)
    monkeypatch.setenv("OPL_DIRECT_DELTA", "0")
    plain = _model(L, shape, transfers, eid)
    obj_p, grad_p = plain._get_obj_jac(w.copy(), x, y)
    monkeypatch.delenv("OPL_DIRECT_DELTA")
    assert rel_err(obj_d, want_obj) <= TOL and rel_err(grad_d, want_grad) <= TOL, rel_err(grad_d, want_grad)
    assert rel_err(obj_p, want_obj) <= TOL and rel_err(grad_p, want_grad) <= TOL
    assert obj_d == obj_p and rel_err(grad_d, grad_p) <= 1e-11, rel_err(grad_d, grad_p)
    print("delta digits on the fly: grad vs oracle %.1e, vs materialised schedule %.1e" % (rel_err(grad_d, want_grad), rel_err(grad_d, grad_p)))


@pytest.mark.parametrize("rows,big", [(3000, False), (2999, True)])
def test_hidden_transfer_applied_by_the_head_is_bit_identical(L, rows, big, monkeypatch):
    """Two-layer softplus networks on the INT8 engine: the forward digit GEMM stores the pre-activations and the fused head
    applies softplus / logistic to every tile it stages (csrc/mlp.cu narrow_head_kernel, ZIN) instead of the GEMM epilogue.
    Same routine on the same arguments, so objective and gradient must be BIT-identical to the epilogue schedule
    (OPL_LATE_TRANSFER=0) -- also with a ragged last tile and with pre-activations beyond +-708 (the full-form fix-up), for the
    objective-only evaluation too -- and both within the gate of the oracle."""
    rng = numpy.random.default_rng(23)
    shape, transfers, eid = (200, 128, 10), [(2, 0), (4, 0)], 1
    x, y = mo.random_classification(rows, 200, 10, rng)
    w = mo.random_params(shape, rng)
    if big:
        x[::7] *= 900.0                                                  # |z| up to ~1e4: the irregular-argument branch
        w[128 + 200 * 128:] *= 1e-3
    with numpy.errstate(all="ignore"):
        want_obj, want_grad = mo.objective_gradient(w, shape, transfers, eid, x, y, literal_softmax=False)
    late = _model(L, shape, transfers, eid)
    obj_l, grad_l = late._get_obj_jac(w.copy(), x, y)
    only_l = late._get_obj(w.copy(), x, y)
    monkeypatch.setenv("OPL_LATE_TRANSFER", "0")
    epi = _model(L, shape, transfers, eid)
    obj_e, grad_e = epi._get_obj_jac(w.copy(), x, y)
    only_e = epi._get_obj(w.copy(), x, y)
    monkeypatch.delenv("OPL_LATE_TRANSFER")
    assert obj_l == obj_e and only_l == only_e and numpy.array_equal(grad_l, grad_e)
    assert rel_err(obj_l, want_obj) <= TOL and rel_err(grad_l, want_grad) <= TOL, (rel_err(obj_l, want_obj), rel_err(grad_l, want_grad))
    assert rel_err(only_l, want_obj) <= TOL


def test_two_layer_schedule_probe_with_direction_and_foreign_stream(L):
    """The two-layer INT8 schedule end to end through the fused probe: the trial point x + alpha d is formed by the same
    cooperative launch that cuts the digits of W_0 (csrc/ozaki.cu trial_w0_digits_kernel), the head's folds run on the workspace's
    side stream (OPL_SIDE_FOLD) and rejoin before the probe scalars are read.  The probe must equal the evaluation at the explicit
    point bit for bit, its slope must be grad . d, and issuing everything under a non-default torch stream must change nothing."""
    import torch
    from learning_b200 import _device as dev
    rng = numpy.random.default_rng(31)
    shape, transfers, eid = (200, 128, 10), [(2, 0), (4, 0)], 1
    x, y = mo.random_classification(3000, 200, 10, rng)
    w = mo.random_params(shape, rng)
    d = rng.standard_normal(w.size) * 0.05
    alpha = 0.37
    model = _model(L, shape, transfers, eid)
    obj_e, grad_e = model._get_obj_jac(w + alpha * d, x, y)
    want_obj, want_grad = mo.objective_gradient(w + alpha * d, shape, transfers, eid, x, y, literal_softmax=False)
    assert rel_err(obj_e, want_obj) <= TOL and rel_err(grad_e, want_grad) <= TOL
    ws = model._make_resident(x, y)
    wd, dd = dev.to_device(w), dev.to_device(d)
    obj_p, slope_p, grad_p = model._probe(ws, wd, alpha, dd, True)
    assert obj_p == obj_e and numpy.array_equal(dev.to_host(grad_p), grad_e)
    assert rel_err(slope_p, grad_e.dot(d)) <= 1e-12
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        other = _model(L, shape, transfers, eid)
        ws2 = other._make_resident(x, y)
        for _ in range(3):                       # back to back: the side-stream fold of one probe against the next probe's kernels
            obj_s, slope_s, grad_s = other._probe(ws2, wd, alpha, dd, True)
        host_s = dev.to_host(grad_s)
    side.synchronize()
    assert obj_s == obj_e and slope_s == slope_p and numpy.array_equal(host_s, grad_e)


@pytest.mark.parametrize("precision", ["f64", "f64-dmma"])
def test_saturated_head_keeps_reference_semantics(L, precision):
    """Degenerate numerics on the fused-head / INT8 path: logits so large that target-class probabilities underflow to
    exactly 0 (log(0) -> nan_to_num -> -1.797e308 in the loss, Y/A -> 1.797e308 and a zero softmax Jacobian row in the
    gradient, error.py:89-116), and a linear output + cross entropy whose negative prediction must raise
    FloatingPointError (error.py:89-91)."""
    rng = numpy.random.default_rng(9)
    shape, transfers = (48, 64, 6), [(1, 0), (4, 0)]
    x, y = mo.random_classification(1300, 48, 6, rng)
    w = mo.random_params(shape, rng)
    w[64 + 48 * 64:] *= 4000.0                       # saturate the softmax
    with numpy.errstate(all="ignore"):
        want_obj, want_grad = mo.objective_gradient(w, shape, transfers, 1, x, y, literal_softmax=True)
    assert want_obj > 1e300                          # the case really is degenerate
    model = _model(L, shape, transfers, 1, precision=precision)
    obj, grad = model._get_obj_jac(w.copy(), x, y)
    assert rel_err(obj, want_obj) <= TOL
    assert rel_err(grad, want_grad) <= 1e-9, rel_err(grad, want_grad)
    lin = _model(L, (48, 64, 6), [(1, 0), (0, 0)], 1, precision=precision)
    with pytest.raises(FloatingPointError):
        lin._get_obj_jac(mo.random_params(shape, rng), x, y)


def test_full_size_properties_config2(L):
    """BASELINE config 2 at full size (60000 x 784, (784,256,10) softmax/CE): size-independent
    properties -- obj == obj_jac[0], directional derivative vs central differences, row-shard
    additivity (the contract the multi-GPU path relies on)."""
    rng = numpy.random.default_rng(0)
    shape, transfers, eid = (784, 256, 10), [(2, 0), (4, 0)], 1
    x, y = mo.random_classification(60000, 784, 10, rng)
    w = mo.random_params(shape, rng)
    model = _model(L, shape, transfers, eid)
    obj, grad = model._get_obj_jac(w.copy(), x, y)
    assert model._get_obj(w.copy(), x, y) == obj
    assert numpy.isfinite(obj) and obj > numpy.log(10.0) * 0.5
    d = rng.standard_normal(w.size)
    d /= numpy.linalg.norm(d)
    eps = 1e-5
    fd = (model._get_obj(w + eps * d, x, y) - model._get_obj(w - eps * d, x, y)) / (2 * eps)
    assert abs(fd - grad.dot(d)) <= 1e-7 * max(1.0, abs(fd))
    # oracle on a 4096-row slice; the full objective is the row-weighted mean of the slices
    sl = slice(1000, 5096)
    m2 = _model(L, shape, transfers, eid)
    o_s, g_s = m2._get_obj_jac(w.copy(), x[sl], y[sl])
    want_o, want_g = mo.objective_gradient(w, shape, transfers, eid, x[sl], y[sl], literal_softmax=False)
    assert rel_err(o_s, want_o) <= TOL and rel_err(g_s, want_g) <= TOL


@pytest.mark.parametrize("shape,transfers,eid", [
    ((4, 2, 3), [(2, 0), (4, 0)], 1),                         # BASELINE config 1 (README iris MLP)
    ((20, 16, 5), [(2, 0), (4, 0)], 1),
    ((7, 9, 6, 4), [(1, 0), (3, 0.8), (0, 0)], 0),            # tanh / gaussian / linear + MSE
    ((5, 8, 6, 3), [(4, 0), (2, 0), (4, 0)], 1),              # hidden softmax
    ((12, 128, 10), [(2, 0), (0, 0)], 0),                     # widest layer the kernel takes, n = 2944
    ((3, 1), [(0, 0)], 0),
])
@pytest.mark.parametrize("rows", [1, 33, 129, 1000, 5000])    # one partial CTA .. 40 CTAs folded by the last one
def test_small_network_one_launch_kernel(L, shape, transfers, eid, rows):
    """csrc/mlp_small.inl: trial point, forward, error, back-propagation, flat gradient, objective, slope and gradient norm
    in ONE launch -- against the oracle (1e-10), against the general kernel schedule, and through the fused probe."""
    import torch
    from learning_b200 import _device as dev
    rng = numpy.random.default_rng(rows + len(shape))
    gen = mo.random_classification if eid == 1 else mo.random_regression
    x, y = gen(rows, shape[0], shape[-1], rng)
    w = mo.random_params(shape, rng)
    want_obj, want_grad = mo.objective_gradient(w, shape, transfers, eid, x, y)
    small = _model(L, shape, transfers, eid)
    plain = _model(L, shape, transfers, eid)
    plain.small_path = False
    with small.resident(x, y):
        l0 = nat_launches(small)
        obj, grad = small._get_obj_jac(w.copy(), x, y)
        assert nat_launches(small) - l0 == 1                   # ONE kernel per evaluation
        assert rel_err(obj, want_obj) <= TOL and rel_err(grad, want_grad) <= TOL, (rel_err(obj, want_obj), rel_err(grad, want_grad))
        assert rel_err(small._get_obj(w.copy(), x, y), want_obj) <= TOL
        obj2, grad2 = small._get_obj_jac(w.copy(), x, y)
        assert obj2 == obj and numpy.array_equal(grad2, grad)  # fixed-order fold of the CTA partials
        # fused probe: scalars formed inside the same launch
        d = rng.standard_normal(w.size)
        ws = small._make_resident(x, y)
        l0 = nat_launches(small)
        val, slope, g = small._probe(ws, dev.to_device(w), 0.3, dev.to_device(d), True)
        assert nat_launches(small) - l0 == 1
        e_obj, e_grad = mo.objective_gradient(w + 0.3 * d, shape, transfers, eid, x, y)
        assert rel_err(val, e_obj) <= TOL and rel_err(dev.to_host(g), e_grad) <= TOL
        assert rel_err(slope, e_grad.dot(d)) <= 1e-9 and rel_err(g.opl_norm, numpy.linalg.norm(e_grad)) <= 1e-10
    obj_p, grad_p = plain._get_obj_jac(w.copy(), x, y)
    assert rel_err(obj, obj_p) <= 1e-12 and rel_err(grad, grad_p) <= 1e-11


def nat_launches(model):
    from learning_b200 import _native as nat
    return nat.lib.opl_mlp_launch_count(model._workspace.pointer)


def test_small_network_kernel_keeps_reference_edge_semantics(L):
    """Saturated softmax (target probability underflows to 0: log(0) -> -1.797e308, Y/A -> 1.797e308 and a zero Jacobian row)
    and log(negative) -> FloatingPointError on the one-launch path (error.py:89-116)."""
    rng = numpy.random.default_rng(9)
    shape, transfers = (6, 8, 4), [(1, 0), (4, 0)]
    x, y = mo.random_classification(200, 6, 4, rng)
    w = mo.random_params(shape, rng)
    w[8 + 48:] *= 4000.0
    with numpy.errstate(all="ignore"):
        want_obj, want_grad = mo.objective_gradient(w, shape, transfers, 1, x, y, literal_softmax=True)
    assert want_obj > 1e300
    model = _model(L, shape, transfers, 1)
    obj, grad = model._get_obj_jac(w.copy(), x, y)
    assert rel_err(obj, want_obj) <= TOL and rel_err(grad, want_grad) <= 1e-9
    lin = _model(L, shape, [(1, 0), (0, 0)], 1)
    with pytest.raises(FloatingPointError):
        lin._get_obj_jac(mo.random_params(shape, rng), x, y)
    with pytest.raises(FloatingPointError):
        lin._get_obj(mo.random_params(shape, rng), x, y)


def test_full_size_config2_against_oracle(L):
    """BASELINE config 2 at its real size -- all 60000 x 784 patterns, (784,256,10) softplus / softmax + cross entropy --
    against the oracle's objective and gradient of the SAME 60000 rows (not a slice), in the parity mode and in the fast
    mode; plus residency: an in-place edit of X between two calls must be seen."""
    rng = numpy.random.default_rng(0)
    shape, transfers, eid = (784, 256, 10), [(2, 0), (4, 0)], 1
    x, y = mo.random_classification(60000, 784, 10, rng)
    w = mo.random_params(shape, rng)
    want_obj, want_grad = mo.objective_gradient(w, shape, transfers, eid, x, y, literal_softmax=False)
    model = _model(L, shape, transfers, eid)
    obj, grad = model._get_obj_jac(w.copy(), x, y)
    assert rel_err(obj, want_obj) <= TOL and rel_err(grad, want_grad) <= TOL, (rel_err(obj, want_obj), rel_err(grad, want_grad))
    fast = _fast_model(L, shape, transfers, eid)
    obj_f, grad_f = fast._get_obj_jac(w.copy(), x, y)
    assert rel_err(obj_f, want_obj) <= FAST_TOL and rel_err(grad_f, want_grad) <= FAST_TOL, rel_err(grad_f, want_grad)
    # the reference always reads the arrays it is given: edit X in place (same object, same buffer), evaluate again
    x[::3] *= -0.5
    want_obj2, want_grad2 = mo.objective_gradient(w, shape, transfers, eid, x, y, literal_softmax=False)
    assert rel_err(want_grad2, want_grad) > 1e-3
    obj2, grad2 = model._get_obj_jac(w.copy(), x, y)
    assert rel_err(obj2, want_obj2) <= TOL and rel_err(grad2, want_grad2) <= TOL, rel_err(grad2, want_grad2)


def test_residency_is_explicit_never_guessed(L):
    """Dataset residency (architecture/mlp.py _make_resident): inside a residency scope the upload happens once; outside
    every call reads the arrays it is given; device tensors are re-read when their version counter moves."""
    import torch
    from learning_b200 import _device as dev
    rng = numpy.random.default_rng(3)
    shape, transfers, eid = (128, 64, 5), [(2, 0), (4, 0)], 1
    x, y = mo.random_classification(1536, 128, 5, rng)          # INT8-engine size: the digits of X are cached per set_data
    w = mo.random_params(shape, rng)
    model = _model(L, shape, transfers, eid)
    ref = lambda xx: mo.objective_gradient(w, shape, transfers, eid, xx, y)
    _, g0 = model._get_obj_jac(w.copy(), x, y)
    assert rel_err(g0, ref(x)[1]) <= TOL
    x[5:50] += 0.25                                              # in place, same object
    _, g1 = model._get_obj_jac(w.copy(), x, y)
    assert rel_err(g1, ref(x)[1]) <= TOL and rel_err(g1, g0) > 1e-6
    # a NEW array that happens to reuse a freed id / address is a different object
    for _ in range(3):
        xb = x + rng.standard_normal(x.shape) * 0.1
        _, gb = model._get_obj_jac(w.copy(), xb, y)
        assert rel_err(gb, ref(xb)[1]) <= TOL
        del xb
    # inside a scope: one upload, evaluations on the resident copy
    with model.resident(x, y):
        token = model._workspace.resident
        assert token is not None
        _, ga = model._get_obj_jac(w.copy(), x, y)
        assert model._workspace.resident is token               # no second upload
        assert numpy.array_equal(ga, g1)
    # device tensors: version counter
    xd, yd = dev.to_device(x), dev.to_device(y)
    wd = dev.to_device(w)
    _, gd = model._get_obj_jac(wd, xd, yd)
    assert numpy.array_equal(dev.to_host(gd), g1)
    xd[5:50] -= 0.25                                             # in place on the device: _version moves
    x[5:50] -= 0.25
    _, gd2 = model._get_obj_jac(wd, xd, yd)
    assert rel_err(dev.to_host(gd2), ref(x)[1]) <= TOL
    model.invalidate()
    _, gd3 = model._get_obj_jac(wd, xd, yd)
    assert numpy.array_equal(dev.to_host(gd3), dev.to_host(gd2))


def test_evaluation_follows_the_current_stream(L):
    """Handles re-bind to the caller's current CUDA stream: an evaluation, a BFGS direction and an
    L-BFGS direction issued under torch.cuda.stream(s) are ordered with the torch work of that stream."""
    import torch
    rng = numpy.random.default_rng(5)
    shape, transfers, eid = (20, 16, 5), [(2, 0), (4, 0)], 1
    x, y = mo.random_classification(300, 20, 5, rng)
    w = mo.random_params(shape, rng)
    for make in (L.optimize.BFGS, L.optimize.LBFGS):
        a = _model(L, shape, transfers, eid, optimizer=make())
        b = _model(L, shape, transfers, eid, optimizer=make())
        for m in (a, b):
            m.logging = False
            m._bias_vec, m._weight_matrices = w[:16].copy(), [w[16:336].reshape(20, 16).copy(), w[336:].reshape(16, 5).copy()]
        side = torch.cuda.Stream()
        for step in range(4):
            ea = a.train_step(x, y)
            with torch.cuda.stream(side if step % 2 else torch.cuda.current_stream()):
                eb = b.train_step(x, y)
            assert ea == eb
        assert numpy.array_equal(a._weight_matrices[0], b._weight_matrices[0])


def test_reference_known_answers(L, golden_meta):
    # learning/tests/architecture/test_mlp.py:41-66,116-126
    model = L.MLP((2, 2, 2), transfers=[L.LinearTransfer(), L.LinearTransfer()])
    model._bias_vec = numpy.ones(model._bias_vec.shape)
    model._weight_matrices = [numpy.ones(m.shape) for m in model._weight_matrices]
    for x, want in [([0, 0], [2, 2]), ([0.5, 0.5], [4, 4]), ([1, 0], [4, 4]), ([0.5, 1], [5, 5]), ([1, 1], [6, 6])]:
        assert numpy.allclose(model.activate(x), want, atol=1e-12)
    assert numpy.allclose(model.activate([[0, 0], [0.5, 0.5]]), [[2, 2], [4, 4]], atol=1e-12)
    model = L.MLP((2, 1), transfers=L.LinearTransfer())
    model._bias_vec[0] = 0.0
    model._weight_matrices[0][0][0] = 0.5
    model._weight_matrices[0][1][0] = -0.5
    assert (model.activate([1, 1]) == [0.0]).all()
    model._weight_matrices[0][0][0] = 1.0
    model._weight_matrices[0][1][0] = 2.0
    assert (model.activate([1, 1]) == [3.0]).all()
    with pytest.raises(ValueError):
        model.activate([1, 2, 3])
    with pytest.raises(ValueError):
        L.MLP((2, 2, 2), transfers=[L.LinearTransfer()])

    ka = golden_meta["known_answers"]
    # learning/tests/test_calculate.py:162-246, test_error.py:79-138
    assert L.SoftmaxTransfer()(numpy.array([-1000.0, 1000.0])).tolist() == ka["softmax_1000"]
    assert numpy.allclose(L.ReluTransfer()(numpy.array([0.0, 1.0])), ka["softplus_0_1"], rtol=1e-15)
    assert L.ReluTransfer()(numpy.array([1000.0, -1000.0])).tolist() == ka["softplus_big"]
    ce = L.CrossEntropyError()
    assert ce(numpy.array([0.0, 1.0]), numpy.array([0.0, 1.0])) == ka["ce_exact"][0]
    assert ce(numpy.full((2, 2), 0.5), numpy.array([[0.0, 1.0], [1.0, 0.0]])) == pytest.approx(ka["ce_exact"][1], rel=1e-15)
    err, deriv = ce.derivative(numpy.array([0.0, 1.0]), numpy.array([0.0, 1.0]))
    assert deriv.tolist() == ka["ce_deriv_01"] and err == ka["ce_exact"][0]
    with pytest.raises(FloatingPointError):      # error.py:89-91
        ce(numpy.array([-0.5, 1.5]), numpy.array([0.0, 1.0]))
    mse = L.MeanSquaredError()
    a, b = numpy.array([[1.0, 2.0], [3.0, 5.0]]), numpy.array([[0.0, 2.0], [1.0, 1.0]])
    assert mse(a, b) == pytest.approx(numpy.mean((a - b) ** 2), rel=1e-15)
    assert numpy.allclose(mse.derivative(a, b)[1], (a - b) * 2.0 / 4.0, rtol=1e-15)
    # transfer derivatives against the oracle's Jacobian products
    rng = numpy.random.default_rng(1)
    z = rng.standard_normal((5, 4))
    ones = numpy.ones_like(z)
    for t, tid, p in [(L.TanhTransfer(), mo.TANH, 0), (L.ReluTransfer(), mo.SOFTPLUS, 0),
                      (L.GaussianTransfer(0.7), mo.GAUSSIAN, 0.7), (L.LinearTransfer(), mo.LINEAR, 0)]:
        a = mo.transfer_apply(tid, p, z.copy())
        assert rel_err(t(z), a) <= 1e-14
        assert rel_err(t.derivative(z, a), mo.transfer_backprop(tid, p, ones, z, a)) <= 1e-14
    y = mo.transfer_apply(mo.SOFTMAX, 0, z)
    jac = L.SoftmaxTransfer().derivative(z, y)
    assert jac.shape == (5, 4, 4)
    up = rng.standard_normal((5, 4))
    assert rel_err(numpy.einsum('ij,ijk->ik', up, jac), mo.transfer_backprop(mo.SOFTMAX, 0, up, z, y)) <= 1e-13


def test_custom_plugins_are_rejected_not_run_on_cpu(L):
    class MyTransfer(L.Transfer):
        def __call__(self, x):
            return x

    class MyError(L.ErrorFunc):
        pass
    with pytest.raises(TypeError):
        L.MLP((2, 2, 2), transfers=MyTransfer())
    with pytest.raises(TypeError):
        L.MLP((2, 2, 2), error_func=MyError())
    with pytest.raises(TypeError):
        L.optimize.BFGS(initial_hessian_func=lambda *a: None)


def test_reset_determinism_and_pickle(L):
    # learning/tests/architecture/test_mlp.py:225-246
    import random
    a, b = L.MLP((3, 4, 2)), L.MLP((3, 4, 2))
    for m in (a, b):
        random.seed(0)
        numpy.random.seed(0)
        m.reset()
    assert a.serialize() == b.serialize()
    # draw order: matrices first, then bias (mlp.py:125-130)
    numpy.random.seed(0)
    w0 = (2 * numpy.random.random((3, 4)) - 1) * 0.25
    assert numpy.array_equal(a._weight_matrices[0], w0)
    x = numpy.random.random((6, 3))
    out = a.activate(x)
    clone = L.MLP.unserialize(a.serialize())
    assert numpy.array_equal(clone.activate(x), out)
    assert numpy.array_equal(copy.deepcopy(a).activate(x), out)
    # optimizer with device state pickles (retries path, base.py:223)
    a.logging = False
    a.train(x, numpy.random.random((6, 2)), iterations=3)
    pickle.loads(pickle.dumps(a._optimizer, protocol=2))


# ----------------------------------------------------------------------------------------
# quasi-Newton kernels
# ----------------------------------------------------------------------------------------
def test_bfgs_update_golden(L, golden_meta, golden_qn):
    from learning_b200.optimize import optimizer as opt
    g = golden_qn
    for case in golden_meta["bfgs"]:
        k = case["key"]
        got = opt._bfgs_eq(g[k + "_h"], g[k + "_s"], g[k + "_y"])
        assert rel_err(got, g[k + "_hnext"]) <= 1e-12, (case, rel_err(got, g[k + "_hnext"]))
    # reference tests/optimize/test_optimizer.py:48-62 : symmetry + secant equation
    h = numpy.identity(2)
    s, y = numpy.array([1.0, 2.0]), numpy.array([0.5, 1.5])
    hn = opt._bfgs_eq(h, s, y)
    assert numpy.allclose(hn, hn.T, atol=1e-14) and numpy.allclose(hn.dot(y), s, atol=1e-12)


def _bfgs_handle(L, n, begin=None, end=None):
    from learning_b200 import _native as nat, _device as dev
    out = ctypes.c_void_p()
    nat.check(nat.lib.opl_bfgs_create(ctypes.byref(out), n, 0 if begin is None else begin,
                                      n if end is None else end, dev.stream_ptr()))
    return out


@pytest.mark.parametrize("n,seed_mode", [(19, 0), (19, 1), (130, 0), (257, 1), (1000, 0)])
def test_bfgs_direction_sequence_against_oracle(L, golden_qn, n, seed_mode):
    """Drive the C-ABI state machine through 8 iterations of synthetic (g, s) pairs and compare
    every direction and the materialised H with the literal reference formula."""
    import torch
    from learning_b200 import _native as nat, _device as dev
    rng = numpy.random.default_rng(n + seed_mode)
    handle = _bfgs_handle(L, n)
    h_ref = None
    prev_g = prev_s = None
    for it in range(8):
        g = rng.standard_normal(n)
        if it == 4:
            g = prev_g.copy()            # y = 0  => y.s == 0 => H unchanged (optimizer.py:328-330)
        gd = dev.to_device(g)
        out = torch.empty(n, dtype=torch.float64, device='cuda')
        nat.check(nat.lib.opl_bfgs_direction_device(handle, dev.ptr(gd), dev.ptr(prev_s), dev.ptr(prev_g_dev) if it else None,
                                                    seed_mode, dev.ptr(out)))
        if it == 0:
            want = -g
        else:
            s, y = prev_s_host, g - prev_g
            if h_ref is None:
                gamma = oo.gamma_scalar(s, y) if seed_mode == 1 else 1.0
                h_ref = oo.bfgs_update_literal(numpy.identity(n) * gamma, s, y)
            else:
                h_ref = oo.bfgs_update_literal(h_ref, s, y)
            want = -(h_ref.dot(g))
        assert rel_err(dev.to_host(out), want) <= 1e-11, (it, rel_err(dev.to_host(out), want))
        if it >= 1:
            hm = torch.empty((n, n), dtype=torch.float64, device='cuda')
            nat.check(nat.lib.opl_bfgs_materialize_device(handle, dev.ptr(hm)))
            assert rel_err(dev.to_host(hm), h_ref) <= 1e-11, it
        prev_g, prev_g_dev = g, gd
        prev_s_host = rng.standard_normal(n) * 0.1 + 0.05 * want
        prev_s = dev.to_device(prev_s_host)
    nat.check(nat.lib.opl_bfgs_reset(handle))
    assert nat.lib.opl_bfgs_state(handle) == 0
    nat.check(nat.lib.opl_bfgs_destroy(handle))


@pytest.mark.parametrize("n,cut", [(301, 160), (1500, 700), (3000, 1000)])   # the last two: short shards of a wide H take
def test_bfgs_row_sharded_pass_equals_whole(L, n, cut):                      # the 2-D (row panel x column range) pass
    """Two row shards driven on one GPU (pass -> exchange -> finish) reproduce the single-device
    direction: the contract of the row-sharded multi-GPU path."""
    import torch
    from learning_b200 import _native as nat, _device as dev
    rng = numpy.random.default_rng(4)
    whole = _bfgs_handle(L, n)
    parts = [_bfgs_handle(L, n, 0, cut), _bfgs_handle(L, n, cut, n)]
    prev_s = prev_g = None
    for it in range(6):
        g = dev.to_device(rng.standard_normal(n))
        want = torch.empty(n, dtype=torch.float64, device='cuda')
        nat.check(nat.lib.opl_bfgs_direction_device(whole, dev.ptr(g), dev.ptr(prev_s), dev.ptr(prev_g), 1, dev.ptr(want)))
        u = torch.zeros(n, dtype=torch.float64, device='cuda')
        w = torch.zeros_like(u)
        v = torch.zeros_like(u)
        for handle in parts:
            ui, vi, wi = torch.zeros_like(u), torch.zeros_like(u), torch.zeros_like(u)
            nat.check(nat.lib.opl_bfgs_pass_device(handle, dev.ptr(g), dev.ptr(prev_s), dev.ptr(prev_g), 1,
                                                   dev.ptr(ui), dev.ptr(vi), dev.ptr(wi)))
            u += ui          # all-gather of disjoint slices
            w += wi
            v += vi          # all-reduce
        for handle in parts:
            got = torch.empty(n, dtype=torch.float64, device='cuda')
            nat.check(nat.lib.opl_bfgs_finish_device(handle, dev.ptr(g), dev.ptr(prev_s), dev.ptr(prev_g), 1,
                                                     dev.ptr(u), dev.ptr(v), dev.ptr(w), dev.ptr(got)))
            assert rel_err(dev.to_host(got), dev.to_host(want)) <= 1e-12, it
        prev_g = g
        prev_s = dev.to_device(rng.standard_normal(n) * 0.1) + 0.05 * want
    for handle in parts + [whole]:
        nat.check(nat.lib.opl_bfgs_destroy(handle))


def test_bfgs_large_n_properties(L):
    """n = 6000 (288 MB): secant equation H+ y = s and preserved symmetry after fused passes."""
    import torch
    from learning_b200 import _native as nat, _device as dev
    n = 6000
    rng = numpy.random.default_rng(9)
    handle = _bfgs_handle(L, n)
    nat.check(nat.lib.opl_bfgs_fill_synthetic(handle, 1234))
    g0 = dev.to_device(rng.standard_normal(n))
    s = dev.to_device(rng.standard_normal(n))
    y = rng.standard_normal(n)
    g1 = g0 + dev.to_device(y)
    d = torch.empty(n, dtype=torch.float64, device='cuda')
    nat.check(nat.lib.opl_bfgs_direction_device(handle, dev.ptr(g1), dev.ptr(s), dev.ptr(g0), 0, dev.ptr(d)))
    h = torch.empty((n, n), dtype=torch.float64, device='cuda')
    nat.check(nat.lib.opl_bfgs_materialize_device(handle, dev.ptr(h)))
    yd = g1 - g0
    assert rel_err(dev.to_host(h @ yd), dev.to_host(s)) <= 1e-9          # secant
    assert float((h - h.T).abs().max()) <= 1e-10                          # symmetric seed stays symmetric
    assert rel_err(dev.to_host(d), dev.to_host(-(h @ g1))) <= 1e-10      # direction == -H+ g
    nat.check(nat.lib.opl_bfgs_destroy(handle))


def test_lbfgs_direction_golden(L, golden_meta, golden_qn):
    import torch
    from learning_b200 import _native as nat, _device as dev
    g = golden_qn
    for case in golden_meta["lbfgs"]:
        k, n, m = case["key"], case["n"], case["m"]
        S, Y = g[k + "_S"], g[k + "_Y"]
        out = ctypes.c_void_p()
        nat.check(nat.lib.opl_lbfgs_create(ctypes.byref(out), n, 8, dev.stream_ptr()))
        # feed the history through the public call: prev_step = s_i, grad - prev_grad = y_i
        base = numpy.zeros(n)
        d = torch.empty(n, dtype=torch.float64, device='cuda')
        zero = dev.to_device(base)
        for i in range(m):
            gi, si = dev.to_device(base + Y[i]), dev.to_device(S[i])     # keep the tensors alive over the call
            nat.check(nat.lib.opl_lbfgs_direction_device(out, dev.ptr(gi), dev.ptr(si), dev.ptr(zero), 1, dev.ptr(d)))
        assert nat.lib.opl_lbfgs_history_len(out) == m
        # final direction for the golden gradient, without pushing a new pair
        gfinal = dev.to_device(g[k + "_g"])
        nat.check(nat.lib.opl_lbfgs_direction_device(out, dev.ptr(gfinal), None, None, 1, dev.ptr(d)))
        assert rel_err(dev.to_host(d), g[k + "_dir"]) <= 1e-12, (case, rel_err(dev.to_host(d), g[k + "_dir"]))
        nat.check(nat.lib.opl_lbfgs_destroy(out))


def test_lbfgs_history_is_a_ring_of_m(L):
    import torch
    from learning_b200 import _native as nat, _device as dev
    n, m = 64, 3
    rng = numpy.random.default_rng(2)
    out = ctypes.c_void_p()
    nat.check(nat.lib.opl_lbfgs_create(ctypes.byref(out), n, m, dev.stream_ptr()))
    S, Y = [], []
    d = torch.empty(n, dtype=torch.float64, device='cuda')
    for i in range(7):
        s = rng.standard_normal(n) * 0.1
        y = s * rng.uniform(0.5, 2.0, n)
        gnew = rng.standard_normal(n)
        S.append(s)
        Y.append(y)
        S, Y = S[-m:], Y[-m:]
        gd, sd, pd = dev.to_device(gnew), dev.to_device(s), dev.to_device(gnew - y)   # keep alive over the call
        nat.check(nat.lib.opl_lbfgs_direction_device(out, dev.ptr(gd), dev.ptr(sd), dev.ptr(pd), 1, dev.ptr(d)))
        # y is recomputed on the device as gnew - (gnew - y): compare against the same rounding
        y_dev = gnew - (gnew - y)
        Y[-1] = y_dev
        want = oo.lbfgs_direction(S, Y, gnew)
        assert rel_err(dev.to_host(d), want) <= 1e-12, i
        assert nat.lib.opl_lbfgs_history_len(out) == min(i + 1, m)
    nat.check(nat.lib.opl_lbfgs_destroy(out))


# ----------------------------------------------------------------------------------------
# optimizer iterate sequences (50 steps) against the reference's traces
# ----------------------------------------------------------------------------------------
def _optimizer_for(L, name):
    o = L.optimize
    return {
        "bfgs": lambda: o.BFGS(),
        "bfgs_scaled": lambda: o.BFGS(initial_hessian_func=o.optimizer.initial_hessian_scaled_identity),
        "bfgs_reset5": lambda: o.BFGS(iterations_per_reset=5),
        "lbfgs": lambda: o.LBFGS(),
        "sd": lambda: o.SteepestDescent(),
        "sdm": lambda: o.SteepestDescentMomentum(),
        "sd_backtracking": lambda: o.SteepestDescent(step_size_getter=o.BacktrackingLineSearch()),
        "bfgs_wolfe_quadratic": lambda: o.BFGS(step_size_getter=o.WolfeLineSearch()),
    }[name]()


def _set_weights(model, flat, shape):
    bias, mats = numpy.split(numpy.array(flat, dtype=float, copy=True), [shape[1]])
    model._bias_vec = bias
    model._weight_matrices = [m.reshape(s) for m, s in zip(
        numpy.split(mats, numpy.cumsum([a * b for a, b in zip(shape[:-1], shape[1:])])[:-1]),
        zip(shape[:-1], shape[1:]))]


def _flat(model):
    return numpy.hstack([model._bias_vec] + [m.ravel() for m in model._weight_matrices])


def _teacher_forced_step(L, run, t, record):
    """One train_step of the product from the ORACLE's x_k and optimizer state; returns (rel err of x_{k+1},
    rel err of the step x_{k+1} - x_k, rel err of H_{k+1} or None)."""
    from oracle_helpers import install_state
    shape, transfers, eid = specs(run)
    x_k, state, val, x_next, h_next = record
    model = _model(L, shape, transfers, eid, optimizer=_optimizer_for(L, run["optimizer"]))
    model.logging = False
    _set_weights(model, x_k, shape)
    install_state(L, model._optimizer, state)
    err = model.train_step(t[run["data"] + "_x"], t[run["data"] + "_y"])
    got = _flat(model)
    e_obj = rel_err(err, val)
    e_x = rel_err(got, x_next)
    e_step = rel_err(got - x_k, x_next - x_k)
    e_h = None
    if h_next is not None and hasattr(model._optimizer, "inverse_hessian"):
        e_h = rel_err(model._optimizer.inverse_hessian(), h_next)
    return e_obj, e_x, e_step, e_h


STEP_TOL = 1e-9          # relative agreement of the step x_{k+1} - x_k where the reference's own step is well conditioned


def _step_gate(run, t, record, e_step):
    """A teacher-forced step passes when it reproduces the reference's step to STEP_TOL, or -- where the reference's own step
    is ill conditioned -- to within 100x what the reference's step itself moves under 1e-15 relative perturbations of its
    gradients (oracle_helpers.reference_step_sensitivity).  Returns (passed, sensitivity or None)."""
    if e_step <= STEP_TOL:
        return True, None
    from oracle_helpers import reference_step_sensitivity
    sens = reference_step_sensitivity(run, t, record)
    return e_step <= 100.0 * sens, sens


def test_teacher_forced_steps_with_optimizer_state_match_reference(L, golden_meta, golden_traces):
    """Stateful teacher forcing (SURVEY section 7 hard part (i); reference optimize/optimizer.py:234-285,414-508): for
    EVERY recorded step k of all ten reference traces, load the reference's x_k and its optimizer state before the step
    (H_{k-1} / (s, y) history, previous step and gradient, iteration counter, initial-step memory), take one step and
    compare the objective and H_k to 1e-10, x_{k+1} to 1e-10 and the step x_{k+1} - x_k itself to 1e-9 -- except on steps
    where the reference's own step is ill conditioned (late iris L-BFGS steps subtract nearly equal gradients), where the
    bound is 100x the reference's measured sensitivity to rounding-level gradient perturbations.  The oracle that supplies
    the states reproduces the golden traces bit for bit (tests/test_oracle.py), which is re-asserted here."""
    from oracle_helpers import oracle_trace_with_states
    t = golden_traces
    report = {}
    for run in golden_meta["traces"]:
        records = oracle_trace_with_states(run, t)
        xs = t[run["key"] + "_xs"]
        worst = dict(obj=0.0, x=0.0, step=0.0, H=0.0)
        conditioned = []
        for k, record in enumerate(records):
            assert rel_err(record[3], xs[k + 1]) <= 1e-9, (run["key"], k)       # the oracle IS the reference trace
            e_obj, e_x, e_step, e_h = _teacher_forced_step(L, run, t, record)
            assert e_obj <= TOL, (run["key"], k, e_obj)
            assert e_h is None or e_h <= TOL, (run["key"], k, e_h)
            passed, sens = _step_gate(run, t, record, e_step)
            assert passed, (run["key"], k, "step error %.2e, reference sensitivity %.2e" % (e_step, sens))
            if sens is None:
                assert e_x <= TOL, (run["key"], k, e_x)
                worst["step"] = max(worst["step"], e_step)
                worst["x"] = max(worst["x"], e_x)
            else:
                conditioned.append((k, float("%.2e" % e_step), float("%.2e" % sens)))
            worst["obj"] = max(worst["obj"], e_obj)
            worst["H"] = max(worst["H"], e_h or 0.0)
        report[run["key"]] = dict(steps=len(records), ill_conditioned_steps=conditioned, **worst)
    print("teacher-forced steps (worst relative errors per trace; ill-conditioned steps as (k, step error, reference sensitivity)):", report)
    total = sum(r["steps"] for r in report.values())
    ill = sum(len(r["ill_conditioned_steps"]) for r in report.values())
    assert ill <= 0.05 * total, report                 # the conditioning clause is the exception, not the rule


def test_optimizer_iterate_sequences_match_reference(L, golden_meta, golden_traces):
    """Free-running: the same iterate sequence as the reference over the recorded steps (up to 50).  Gate: every trace
    agrees to 1e-8 on ALL its steps, or up to a first divergence that must come after step 20 and must be conditioning, not
    an arithmetic defect: the teacher-forced step from the reference's own state at that very step passes the step gate
    (1e-9, or 100x the reference's sensitivity to rounding-level gradient perturbations).  SURVEY.md section 7."""
    from oracle_helpers import oracle_trace_with_states
    t = golden_traces
    report = {}
    for run in golden_meta["traces"]:
        shape, transfers, eid = specs(run)
        x, y = t[run["data"] + "_x"], t[run["data"] + "_y"]
        model = _model(L, shape, transfers, eid, optimizer=_optimizer_for(L, run["optimizer"]))
        model.logging = False
        _set_weights(model, t[run["key"] + "_w0"], shape)
        xs, objs = t[run["key"] + "_xs"], t[run["key"] + "_objs"]
        agree = 0
        for step in range(run["steps"]):
            err = model.train_step(x, y)
            if rel_err(err, objs[step]) <= 1e-8 and rel_err(_flat(model), xs[step + 1]) <= 1e-8:
                agree = step + 1
            else:
                break
        report[run["key"]] = (agree, run["steps"])
    print("iterate agreement (steps matching reference to 1e-8):", report)
    for run in golden_meta["traces"]:
        agree, steps = report[run["key"]]
        if agree == steps:
            continue
        assert agree >= 20 and "_mid_" not in run["key"] and "_reg_" not in run["key"], (run["key"], agree, report)
        record = oracle_trace_with_states(run, t)[agree]
        e_obj, e_x, e_step, e_h = _teacher_forced_step(L, run, t, record)
        passed, sens = _step_gate(run, t, record, e_step)
        print("first divergence of %s at step %d: teacher-forced from the reference's state there: obj %.1e, step %.1e, "
              "reference sensitivity %s" % (run["key"], agree, e_obj, e_step, sens))
        assert e_obj <= TOL and passed, (run["key"], agree, e_obj, e_step, sens)


def test_accepted_probe_memo_changes_no_bits(L, golden_meta, golden_traces):
    """The evaluation at x_{k+1} that BFGS / L-BFGS ask for right after the line search accepted that very point (the
    reference's TODO, optimizer.py:69-72) is answered from the accepted probe: same iterates bit for bit, one evaluation
    per step fewer; and the weights stay on the device between train_steps until somebody looks at them."""
    t = golden_traces
    for name in ("tr_mid_bfgs", "tr_mid_lbfgs", "tr_iris_bfgs"):
        run = [r for r in golden_meta["traces"] if r["key"] == name][0]
        shape, transfers, eid = specs(run)
        x, y = t[run["data"] + "_x"], t[run["data"] + "_y"]
        models = []
        for memo in (True, False):
            model = _model(L, shape, transfers, eid, optimizer=_optimizer_for(L, run["optimizer"]))
            model.logging = False
            model.accepted_probe_memo = memo
            _set_weights(model, t[name + "_w0"], shape)
            errors = []
            with model.resident(x, y):
                for _ in range(15):
                    errors.append(model.train_step(x, y))
                    assert model.__dict__.get('_flat_dev') is not None       # nobody read the weights: still on the device
            models.append((model, errors))
        (a, ea), (b, eb) = models
        assert ea == eb, name
        assert numpy.array_equal(_flat(a), _flat(b)), name
        assert a.__dict__.get('_flat_dev') is None                            # reading them made the host copy authoritative
        assert rel_err(_flat(a), t[name + "_xs"][15]) <= 1e-8
        assert a.obj_jac_memo_hits == 14 and b.obj_jac_memo_hits == 0, (a.obj_jac_memo_hits, b.obj_jac_memo_hits)
        assert a.obj_jac_evaluations == b.obj_jac_evaluations - 14


def test_checkpoint_mid_training_continues_the_same_sequence(L, golden_meta, golden_traces):
    """Model.serialize / unserialize (base.py:350-373) mid-training: the reference pickles the optimizer's H / history /
    previous step and gradient with the model, so a restored model continues the SAME iterate sequence."""
    t = golden_traces
    for name in ("tr_mid_bfgs", "tr_mid_lbfgs", "tr_reg_sdm"):
        run = [r for r in golden_meta["traces"] if r["key"] == name][0]
        shape, transfers, eid = specs(run)
        x, y = t[run["data"] + "_x"], t[run["data"] + "_y"]
        model = _model(L, shape, transfers, eid, optimizer=_optimizer_for(L, run["optimizer"]))
        model.logging = False
        _set_weights(model, t[name + "_w0"], shape)
        for _ in range(6):
            model.train_step(x, y)
        clone = type(model).unserialize(model.serialize())
        for step in range(6, 12):
            ea, eb = model.train_step(x, y), clone.train_step(x, y)
            assert ea == eb, (name, step)
            assert numpy.array_equal(_flat(model), _flat(clone)), (name, step)
        assert rel_err(_flat(clone), t[name + "_xs"][12]) <= 1e-8


def test_readme_example_trains_like_the_reference(L, golden_traces):
    """BASELINE config 1: README iris MLP (4,2,3), BFGS + Wolfe + FOChange (examples/readme_example.py)."""
    import random
    t = golden_traces
    random.seed(0)
    numpy.random.seed(0)
    model = L.MLP((4, 2, 3), transfers=L.SoftmaxTransfer(), error_func=L.CrossEntropyError(),
                  optimizer=L.optimize.BFGS(step_size_getter=L.optimize.WolfeLineSearch(
                      initial_step_getter=L.optimize.FOChangeInitialStep())))
    model.logging = False
    before = L.validation.get_error(model, t["iris_x"], t["iris_y"])
    model.train(t["iris_x"], t["iris_y"])
    assert L.validation.get_error(model, t["iris_x"], t["iris_y"]) < before
    assert L.validation.get_accuracy(model, t["iris_test_x"], t["iris_test_y"]) >= 0.85
    assert model.obj_jac_evaluations > model.iteration      # line-search probes went through the device


def test_optimizers_on_user_numpy_problems(L):
    """learning/tests/optimize/test_optimizer.py:39-169: sphere from (10,10); L-BFGS == BFGS on
    Rosenbrock with unlimited memory; |x| objective with y == 0 must not divide by zero."""
    o = L.optimize
    sphere = o.Problem(obj_func=lambda v: float(numpy.sum(v ** 2)),
                       obj_jac_func=lambda v: (float(numpy.sum(v ** 2)), 2.0 * v))
    for make in (o.BFGS, o.LBFGS, o.SteepestDescent, o.SteepestDescentMomentum,
                 lambda: o.SteepestDescent(step_size_getter=o.BacktrackingLineSearch())):
        opt = make()
        v = numpy.array([10.0, 10.0])
        for _ in range(1000):
            val, v = opt.next(sphere, v)
            if val <= 1e-10:
                break
        assert val <= 1e-10 and isinstance(v, numpy.ndarray)

    def rosen(v):
        f = float(100.0 * (v[1] - v[0] ** 2) ** 2 + (1 - v[0]) ** 2)
        return f, numpy.array([-400.0 * v[0] * (v[1] - v[0] ** 2) - 2 * (1 - v[0]), 200.0 * (v[1] - v[0] ** 2)])
    prob = o.Problem(obj_func=lambda v: rosen(v)[0], obj_jac_func=rosen)
    a = o.BFGS()
    b = o.LBFGS(num_remembered_iterations=float('inf'),
                initial_hessian_scalar_func=o.optimizer.initial_hessian_one_scalar)
    va = vb = numpy.array([-1.2, 1.0])
    for _ in range(10):
        _, va = a.next(prob, va)
        _, vb = b.next(prob, vb)
        assert numpy.allclose(va, vb, atol=1e-3)

    absval = o.Problem(obj_func=lambda v: float(numpy.sum(numpy.abs(v))),
                       obj_jac_func=lambda v: (float(numpy.sum(numpy.abs(v))), numpy.sign(v)))
    opt = o.LBFGS()
    v = numpy.array([5.0, -3.0])
    for _ in range(6):
        val, v = opt.next(absval, v)
        assert numpy.all(numpy.isfinite(v))


def test_mlp_training_reduces_error(L):
    # learning/tests/architecture/test_mlp.py:70-101 (XOR)
    xor_x = numpy.array([[-1.0, -1.0], [-1.0, 1.0], [1.0, -1.0], [1.0, 1.0]])
    xor_y = numpy.array([[0.0, 1.0], [1.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
    numpy.random.seed(3)
    for kwargs in ({}, dict(transfers=L.SoftmaxTransfer(), error_func=L.CrossEntropyError())):
        model = L.MLP((2, 2, 2), **kwargs)
        model.logging = False
        before = L.validation.get_error(model, xor_x, xor_y)
        model.train(xor_x, xor_y, iterations=20)
        assert L.validation.get_error(model, xor_x, xor_y) < before


# ----------------------------------------------------------------------------------------
# fast mode: tcgen05 TF32 tensor cores + TMA, FP32 accumulate -- gate 1e-3 relative (north_star)
# ----------------------------------------------------------------------------------------
FAST_TOL = 1e-3        # BASELINE-sized layers: TF32 operand rounding (2^-11) averages out over the contraction
FAST_TOL_TINY = 5e-3   # toy golden cases (contraction lengths 1..10): no averaging, bound = a few operand roundings


def _fast_model(L, shape, transfers, eid):
    return L.MLP(shape, transfers=make_transfers(L, transfers), error_func=make_error(L, eid),
                 optimizer=L.optimize.SteepestDescent(), precision='tf32')


def test_fast_mode_golden_cases(L, golden_meta, golden_objgrad):
    g = golden_objgrad
    for entry in golden_meta["objgrad"]:
        if entry["name"] in ("overflow", "underflow_some_rows", "big_softplus_lin"):
            continue   # FP32 range: the +-1e308 / 4000-scaled edge cases belong to the FP64 parity mode
        k = entry["key"]
        shape, transfers, eid = specs(entry)
        model = _fast_model(L, shape, transfers, eid)
        obj, grad = model._get_obj_jac(g[k + "_w"].copy(), g[k + "_x"], g[k + "_y"])
        tol = FAST_TOL if min(shape) >= 32 else FAST_TOL_TINY
        assert rel_err(obj, g[k + "_obj"]) <= tol, (entry["name"], rel_err(obj, g[k + "_obj"]))
        assert rel_err(grad, g[k + "_grad"]) <= tol, (entry["name"], rel_err(grad, g[k + "_grad"]))
        assert rel_err(model._get_obj(g[k + "_w"].copy(), g[k + "_x"], g[k + "_y"]), g[k + "_obj"]) <= tol
        assert rel_err(model.activate(g[k + "_x"]), g[k + "_out"]) <= tol, entry["name"]


@pytest.mark.parametrize("shape,transfers,eid,rows", [
    ((784, 256, 10), [(2, 0), (4, 0)], 1, 4000),
    ((785, 255, 11), [(2, 0), (4, 0)], 1, 1531),
    ((512, 128, 128, 10), [(2, 0), (2, 0), (4, 0)], 1, 2048),
    ((256, 192, 160, 10), [(1, 0), (3, 1.5), (0, 0)], 0, 1000),
    ((40, 70, 33, 65), [(4, 0), (2, 0), (2, 0)], 0, 517),
    # TF32 tensor-path head: last 16-row hidden tile half empty (H % 16 == 8), odd k-step count, <= 8 classes, tanh slope,
    # a partial last step; then 16 classes, tanh output + MSE through the generic row stage
    ((64, 40, 5), [(1, 0), (4, 0)], 1, 1003),
    ((96, 72, 16), [(2, 0), (1, 0)], 0, 333),
    # BASELINE config 4 / config 5 shapes: 128x256 TF32 tiles, five layers of TF32 error growth (gate 1e-3)
    ((1024, 256, 256, 32), [(2, 0), (2, 0), (4, 0)], 1, 4096),
    ((256, 1024, 1024, 1024, 1024, 10), [(2, 0), (2, 0), (2, 0), (2, 0), (0, 0)], 0, 4096),
])
def test_fast_mode_against_oracle_at_size(L, shape, transfers, eid, rows):
    rng = numpy.random.default_rng(hash(shape) % 1000 + 1)
    gen = mo.random_classification if eid == 1 else mo.random_regression
    x, y = gen(rows, shape[0], shape[-1], rng)
    w = mo.random_params(shape, rng)
    want_obj, want_grad = mo.objective_gradient(w, shape, transfers, eid, x, y, literal_softmax=False)
    model = _fast_model(L, shape, transfers, eid)
    obj, grad = model._get_obj_jac(w.copy(), x, y)
    print("fast mode", shape, "obj rel", rel_err(obj, want_obj), "grad rel", rel_err(grad, want_grad))
    assert rel_err(obj, want_obj) <= FAST_TOL
    assert rel_err(grad, want_grad) <= FAST_TOL, rel_err(grad, want_grad)
    obj2, grad2 = model._get_obj_jac(w.copy(), x, y)
    assert obj2 == obj and numpy.array_equal(grad2, grad)      # deterministic


def test_fast_mode_trains(L, golden_traces):
    import random
    t = golden_traces
    random.seed(1)
    numpy.random.seed(1)
    model = L.MLP((4, 8, 3), transfers=L.SoftmaxTransfer(), error_func=L.CrossEntropyError(),
                  optimizer=L.optimize.LBFGS(), precision='tf32')
    model.logging = False
    before = L.validation.get_error(model, t["iris_x"], t["iris_y"])
    model.train(t["iris_x"], t["iris_y"], iterations=40)
    assert L.validation.get_error(model, t["iris_x"], t["iris_y"]) < 0.5 * before
    assert L.validation.get_accuracy(model, t["iris_test_x"], t["iris_test_y"]) >= 0.85
    with pytest.raises(ValueError):
        L.MLP((2, 2, 2), precision='fp8')
